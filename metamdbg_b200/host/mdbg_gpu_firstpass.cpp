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

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cerr << "usage: mdbg_gpu_firstpass <reads.fa|fq[.gz]> <outDir> [--ont] [-l 15] [-d 0.005] [-k 4] "
                     "[--min-abundance 2] [--last-k N] [--batch-mbp 1024] [--max-k K] [--edges]\n"
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
    bool writeEdges = false;
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
        else if (a == "--batch-mbp") batchMbp = (size_t)atol(next().c_str());
    }
    try {
        Context ctx(l, density, hpc);
        auto edgeKeys = [&](const std::string& file) {   // EdgeIndexer on the table the context holds
            if (!writeEdges) return;
            GpuEdgeIndexer edges(ctx, minAb);
            edges.execute(file);
            std::cout << "edges " << edges._nbEdges << " edge_checksum " << edges._checksum << "\n";
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
