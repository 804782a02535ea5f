// mdbg_gpu_firstpass -- C++ host driver over the C ABI: the GPU form of
//   metaMDBG readSelection <tmp> ... ; metaMDBG graph <tmp> --firstpass --min-abundance n
// for the part of those stages that is on the hot path (sketch + side outputs -> purgePalindromes ->
// k-min-mer count).  Reads FASTA/FASTQ (plain or gzip, through zlib), writes read_data_init.txt,
// read_stats.txt, read_data_corrected.txt, kminmerData_min.txt and kminmerData_abundance.txt in the
// reference's formats.
#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "mdbg_host.hpp"

using namespace mdbg_host;

static bool getline_gz(gzFile f, std::string& line) {
    line.clear();
    char buf[1 << 16];
    while (gzgets(f, buf, sizeof buf)) {
        size_t n = strlen(buf);
        bool eol = n && buf[n - 1] == '\n';
        if (eol) n--;
        if (n && buf[n - 1] == '\r') n--;
        line.append(buf, n);
        if (eol) return true;
    }
    return !line.empty();
}

// ---- stage-compatible entry points (row (f)2): the two subcommands AssemblyPipeline runs through system()
//   <exe> readSelection <tmpDir> <tmpDir>/read_data_init.txt <input.txt> --threads N --min-read-quality Q
//                       [--output-quality] [--skip-correction]            (src/pipeline/AssemblyPipeline.hpp:733-737)
//   <exe> graph <tmpDir> --threads N [--min-abundance n] [--firstpass]    (AssemblyPipeline.hpp:773-783)
// Both read <tmpDir>/parameters.gz exactly as Parameters::load does (src/Commons.hpp:1475-1497; writer
// AssemblyPipeline.hpp:1479-1517) and leave the files the next stage expects.  Checkpoint files are created by the
// orchestrator after the child returns 0 (AssemblyPipeline.hpp:1019-1027), not by the stage.
struct StageParameters {
    size_t minimizerSize = 0, kminmerSize = 0;
    float densityAssembly = 0;
    size_t kminmerSizeFirst = 0;
    float minimizerSpacingMean = 0, kminmerLengthMean = 0, kminmerOverlapMean = 0;
    size_t kminmerSizePrev = 0, kminmerSizeLast = 0, meanReadLength = 0;
    float densityCorrection = 0;
    bool useHomopolymerCompression = false;
    int dataType = 0;
    size_t snpmerSize = 0;

    void load(const std::string& filename) {
        gzFile f = gzopen(filename.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + filename);
        auto get = [&](void* p, unsigned n) { if (gzread(f, p, n) != (int)n) { gzclose(f); throw std::runtime_error("short read: " + filename); } };
        get(&minimizerSize, sizeof minimizerSize); get(&kminmerSize, sizeof kminmerSize);
        get(&densityAssembly, sizeof densityAssembly); get(&kminmerSizeFirst, sizeof kminmerSizeFirst);
        get(&minimizerSpacingMean, sizeof minimizerSpacingMean); get(&kminmerLengthMean, sizeof kminmerLengthMean);
        get(&kminmerOverlapMean, sizeof kminmerOverlapMean); get(&kminmerSizePrev, sizeof kminmerSizePrev);
        get(&kminmerSizeLast, sizeof kminmerSizeLast); get(&meanReadLength, sizeof meanReadLength);
        get(&densityCorrection, sizeof densityCorrection); get(&useHomopolymerCompression, sizeof useHomopolymerCompression);
        get(&dataType, sizeof dataType); get(&snpmerSize, sizeof snpmerSize);
        gzclose(f);
    }
};

// every record of the files listed in input.txt, in order (ReadParserParallel::parse, Commons.hpp:5846-5911)
template <typename Fn>
static uint64_t forEachRead(const std::string& inputTxt, uint64_t maxReads, Fn&& fn) {
    std::vector<std::string> files;
    {
        File list(inputTxt, "rb");
        char buf[4096];
        while (fgets(buf, sizeof buf, list.get())) {
            std::string name(buf);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r' || name.back() == ' ')) name.pop_back();
            if (!name.empty()) files.push_back(name);
        }
    }
    uint64_t index = 0;
    Read read;
    for (size_t fi = 0; fi < files.size() && index < maxReads; fi++) {
        gzFile f = gzopen(files[fi].c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + files[fi]);
        std::string line;
        bool have = getline_gz(f, line);
        while (have && index < maxReads) {
            if (line.empty()) { have = getline_gz(f, line); continue; }
            read._datasetIndex = fi;
            read._qual.clear();
            if (line[0] == '>') {
                read._header = line.substr(1);
                read._seq.clear();
                while ((have = getline_gz(f, line)) && (line.empty() || line[0] != '>')) read._seq += line;
            } else if (line[0] == '@') {
                read._header = line.substr(1);
                getline_gz(f, read._seq);
                getline_gz(f, line);
                getline_gz(f, read._qual);
                have = getline_gz(f, line);
            } else {
                gzclose(f);
                throw std::runtime_error("unrecognised record in " + files[fi] + ": " + line.substr(0, 40));
            }
            read._index = index++;
            fn(read);
        }
        gzclose(f);
    }
    return index;
}

static int stageReadSelection(int argc, char** argv) {
    if (argc < 5) { std::cerr << "usage: readSelection <tmpDir> <outputFile> <input.txt> [--threads N] [--min-read-quality Q] [--output-quality] [--skip-correction]\n"; return 2; }
    const std::string tmpDir = argv[2], outFile = argv[3], inputTxt = argv[4];
    bool skipCorrection = false;
    size_t batchMbp = 1024;
    for (int i = 5; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--skip-correction") skipCorrection = true;
        else if (a == "--threads" || a == "--min-read-quality" || a == "--batch-mbp") { if (a == "--batch-mbp" && i + 1 < argc) batchMbp = (size_t)atol(argv[i + 1]); i++; }
    }
    StageParameters P;
    P.load(tmpDir + "/parameters.gz");
    const uint32_t l = (uint32_t)P.minimizerSize;
    // ReadSelection::determineRepetitiveMinimizers (ReadSelection.hpp:497-561): nothing for HiFi; for ONT the first
    // 1 M reads are sketched at the correction density and the most frequent minimizers become the blacklist
    std::vector<uint32_t> blacklist;
    if (!P.useHomopolymerCompression) {
        Context cctx(l, P.densityCorrection, false);
        GpuReadSelectionFunctor counter(cctx, [](const ReadMinimizers&) {}, batchMbp << 20, /*sideOutputs=*/false);   // appends to the store
        forEachRead(inputTxt, 1000000, [&](const Read& r) { counter(r); });
        counter.flush();
        mdbg_repeats_out rep{};
        check(cctx.get(), mdbg_store_repetitive_minimizers(cctx.get(), 0.00001f, &rep), "mdbg_store_repetitive_minimizers");
        blacklist.assign(rep.minimizers, rep.minimizers + rep.n_selected);
    }
    {
        File bl(tmpDir + "/repetitiveMinimizers.bin");
        bl.put(blacklist.data(), sizeof(uint32_t), blacklist.size());
        bl.close();
    }
    Context ctx(l, P.densityAssembly, P.useHomopolymerCompression, blacklist);
    ReadDataWriter writer(outFile, l);
    GpuReadSelectionFunctor functor(ctx, [&](const ReadMinimizers& r) { writer.write(r); }, batchMbp << 20, /*sideOutputs=*/true);
    forEachRead(inputTxt, UINT64_MAX, [&](const Read& r) { functor(r); });
    functor.flush();
    writer.close();
    writer.writeReadStats(tmpDir + "/read_stats.txt");
    uint64_t changed = 0;
    if (P.useHomopolymerCompression || skipCorrection) {       // ReadSelection.hpp:300-302: purgePalindromes -> read_data_corrected.txt
        const uint32_t lastK = computeLastK(P.densityAssembly, writer.n50(), 4);
        changed = purgePalindromesAndWrite(ctx, 4, lastK, tmpDir + "/read_data_corrected.txt");
    }
    std::cout << "reads " << functor.nbReads() << " bases " << functor.nbBases() << " minimizers " << functor.nbSelectedMinimizers()
              << " repetitive_minimizers " << blacklist.size() << " purged_reads " << changed << std::endl;
    return 0;
}

static int stageGraph(int argc, char** argv) {
    if (argc < 3) { std::cerr << "usage: graph <tmpDir> [--threads N] [--min-abundance n] [--firstpass]\n"; return 2; }
    const std::string tmpDir = argv[2];
    bool firstPass = false, unitigs = false;
    uint32_t minAb = 0;
    for (int i = 3; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--firstpass") firstPass = true;
        else if (a == "--unitigs") unitigs = true;
        else if (a == "--min-abundance" && i + 1 < argc) minAb = (uint32_t)atoi(argv[++i]);
        else if (a == "--threads") i++;
    }
    StageParameters P;
    P.load(tmpDir + "/parameters.gz");
    if (!firstPass)
        throw std::runtime_error("graph without --firstpass needs the contig stage's refined abundances; only the first pass "
                                 "(KminmerCounter + rescue) is on the GPU path -- run the reference binary for k > firstK");
    Context ctx((uint32_t)P.minimizerSize, P.densityAssembly, P.useHomopolymerCompression);
    const uint64_t nReads = loadReadData(ctx, tmpDir + "/read_data_corrected.txt");
    GpuKminmerCounter counter(ctx, (uint32_t)P.kminmerSize, minAb);
    counter.execute(tmpDir + "/kminmerData_min.txt", tmpDir + "/kminmerData_abundance.txt");
    std::cout << "reads " << nReads << " kminmers " << counter._nbKminmers << " distinct " << counter._nbDistinct << " solid "
              << counter._nbSolidKminmers << " rescued " << counter._nbRescuedKminmers << " checksum " << counter._checksum << std::endl;
    if (unitigs) {                                       // createGfa's node side (CreateMdbg.cpp:876-925) on the table just built
        GpuUnitigBuilder ub(ctx, minAb);
        ub.execute(tmpDir + "/unitigGraph.nodes.bin", tmpDir + "/unitigGraph.nodes.abundances.bin",
                   tmpDir + "/unitigGraph.edges.successors.bin");
        std::cout << "unitigs " << ub._nbUnitigs << " circular " << ub._nbCircular << " checksum_unitig_nodes " << ub._checksumNodes
                  << " checksum_unitig_abundance " << ub._checksumAbundances << " unitig_edges " << ub._nbUnitigEdges
                  << " checksum_unitig_edges " << ub._checksumEdges << std::endl;
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && (std::string(argv[1]) == "readSelection" || std::string(argv[1]) == "graph")) {
        try {
            return std::string(argv[1]) == "readSelection" ? stageReadSelection(argc, argv) : stageGraph(argc, argv);
        } catch (const std::exception& e) {
            std::cerr << "error: " << e.what() << std::endl;
            return 1;
        }
    }
    if (argc < 3) {
        std::cerr << "usage: mdbg_gpu_firstpass readSelection <tmpDir> <outputFile> <input.txt> [...]   (metaMDBG's stage command lines,\n"
                     "       mdbg_gpu_firstpass graph <tmpDir> --firstpass [--min-abundance n]            parameters.gz in <tmpDir>)\n"
                     "       mdbg_gpu_firstpass <reads.fa|fq[.gz]> <outDir> [--ont] [-l 15] [-d 0.005] [-k 4] "
                     "[--min-abundance 2] [--last-k N] [--batch-mbp 1024] [--max-k K] [--edges] [--unitigs]\n"
                     "       mdbg_gpu_firstpass --from-read-data <read_data_corrected.txt> <outDir> [-k 4] [--min-abundance 2]\n"
                     "         (the `graph --firstpass` seam alone: count the minimizer-space reads of an existing file)\n"
                     "       --max-k K: also derive k+1 .. K from the previous table on the device (the k > firstK `graph`\n"
                     "         passes without the contig stage's refinements): kminmerData_{min,abundance}_k<K>.txt\n";
        return 2;
    }
    bool fromReadData = std::string(argv[1]) == "--from-read-data";
    if (fromReadData && argc < 4) { std::cerr << "--from-read-data needs <file> <outDir>\n"; return 2; }
    std::string input = fromReadData ? argv[2] : argv[1], outDir = fromReadData ? argv[3] : argv[2];
    bool hpc = true;
    uint32_t l = 15, k = 4, minAb = 2, lastK = 0, maxK = 0;
    bool writeEdges = false, writeUnitigs = false;
    float density = 0.005f;
    size_t batchMbp = 1024;
    for (int i = fromReadData ? 4 : 3; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : "0"); };
        if (a == "--ont") hpc = false;
        else if (a == "-l") l = (uint32_t)atoi(next().c_str());
        else if (a == "-d") density = (float)atof(next().c_str());
        else if (a == "-k") k = (uint32_t)atoi(next().c_str());
        else if (a == "--min-abundance") minAb = (uint32_t)atoi(next().c_str());
        else if (a == "--last-k") lastK = (uint32_t)atoi(next().c_str());
        else if (a == "--max-k") maxK = (uint32_t)atoi(next().c_str());
        else if (a == "--edges") writeEdges = true;
        else if (a == "--unitigs") writeUnitigs = true;
        else if (a == "--batch-mbp") batchMbp = (size_t)atol(next().c_str());
    }
    try {
        Context ctx(l, density, hpc);
        auto edgeKeys = [&](const std::string& file) {   // EdgeIndexer on the table the context holds
            if (writeEdges) {
                GpuEdgeIndexer edges(ctx, minAb);
                edges.execute(file);
                std::cout << "edges " << edges._nbEdges << " edge_checksum " << edges._checksum << "\n";
            }
            if (!writeUnitigs) return;
            GpuUnitigBuilder ub(ctx, minAb);
            const std::string dir = file.substr(0, file.find_last_of('/'));
            ub.execute(dir + "/unitigGraph.nodes.bin", dir + "/unitigGraph.nodes.abundances.bin", dir + "/unitigGraph.edges.successors.bin");
            std::cout << "unitigs " << ub._nbUnitigs << " circular " << ub._nbCircular << " checksum_unitig_nodes " << ub._checksumNodes
                      << " checksum_unitig_abundance " << ub._checksumAbundances << " unitig_edges " << ub._nbUnitigEdges
                      << " checksum_unitig_edges " << ub._checksumEdges << "\n";
        };
        auto nextKPasses = [&]() {                       // k+1 .. maxK from the table the context holds
            GpuNextKCounter nextK(ctx, minAb);
            for (uint32_t kk = k + 1; kk <= maxK; kk++) {
                nextK.execute(kk, outDir + "/kminmerData_min_k" + std::to_string(kk) + ".txt",
                              outDir + "/kminmerData_abundance_k" + std::to_string(kk) + ".txt");
                std::cout << "k " << kk << " kminmers " << nextK._nbKminmers << " checksum " << nextK._checksum << "\n";
            }
        };
        if (fromReadData) {
            const uint64_t nReads = loadReadData(ctx, input);
            GpuKminmerCounter counter(ctx, k, minAb);
            counter.execute(outDir + "/kminmerData_min.txt", outDir + "/kminmerData_abundance.txt");
            edgeKeys(outDir + "/edges.bin");
            nextKPasses();
            std::cout << "reads " << nReads << " kminmers " << counter._nbKminmers << " distinct " << counter._nbDistinct
                      << " solid " << counter._nbSolidKminmers << " rescued " << counter._nbRescuedKminmers << " checksum "
                      << counter._checksum << std::endl;
            return 0;
        }
        ReadDataWriter writer(outDir + "/read_data_init.txt", l);     // readSelection's record file + read_stats.txt
        GpuReadSelectionFunctor functor(ctx, [&](const ReadMinimizers& r) { writer.write(r); }, batchMbp << 20,
                                        /*sideOutputs=*/true);
        struct GzIn {                                    // closes on every exit path
            gzFile f;
            explicit GzIn(const std::string& name) : f(gzopen(name.c_str(), "rb")) {
                if (!f) throw std::runtime_error("cannot open " + name);
            }
            ~GzIn() { gzclose(f); }
        } gz(input);
        gzFile f = gz.f;
        std::string line;
        Read read;
        uint64_t index = 0;
        bool have = getline_gz(f, line);
        while (have) {
            if (line.empty()) { have = getline_gz(f, line); continue; }
            if (line[0] == '>') {                        // FASTA, possibly multi-line
                read._header = line.substr(1);
                read._seq.clear();
                while ((have = getline_gz(f, line)) && (line.empty() || line[0] != '>')) read._seq += line;
                read._index = index++;
                functor(read);
            } else if (line[0] == '@') {                 // FASTQ, 4-line records
                read._header = line.substr(1);
                getline_gz(f, read._seq);
                getline_gz(f, line);
                getline_gz(f, read._qual);
                read._index = index++;
                functor(read);
                have = getline_gz(f, line);
            } else {
                throw std::runtime_error("unrecognised record: " + line.substr(0, 40));
            }
        }
        functor.flush();
        writer.close();
        writer.writeReadStats(outDir + "/read_stats.txt");
        if (lastK == 0) lastK = computeLastK(density, writer.n50(), 4);   // on read_stats' N50, as upstream
        uint64_t changed = purgePalindromesAndWrite(ctx, 4, lastK, outDir + "/read_data_corrected.txt");
        GpuKminmerCounter counter(ctx, k, minAb);
        counter.execute(outDir + "/kminmerData_min.txt", outDir + "/kminmerData_abundance.txt");
        edgeKeys(outDir + "/edges.bin");
        nextKPasses();
        std::cout << "reads " << functor.nbReads() << " bases " << functor.nbBases() << " minimizers "
                  << functor.nbSelectedMinimizers() << " purged_reads " << changed << " kminmers " << counter._nbKminmers
                  << " distinct " << counter._nbDistinct << " solid " << counter._nbSolidKminmers << " rescued "
                  << counter._nbRescuedKminmers << " checksum "
                  << counter._checksum << " lastK " << lastK << " kernel_launches "
                  << mdbg_ctx_kernel_launches(ctx.get()) << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
