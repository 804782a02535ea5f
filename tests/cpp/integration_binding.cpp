// Compile-only proof that the reference-side bindings shown in INTEGRATION.md are real C++ against metaMDBG's own
// headers: the functor is accepted by ReadParserParallel::parse, the sink feeds ReadSelection::writeRead, and the
// KminmerCounter replacement fills the members CreateMdbg expects.  Compiled (not linked, not run) by
// tests/test_integration_compile.py where /root/reference exists:
//   g++ -std=gnu++20 -fopenmp -w -c -I/root/reference/src -I<repo>/include -I<repo>/metamdbg_b200/host
#include "Commons.hpp"
#include "graph/CreateMdbg.hpp"
#include "readSelection/ReadSelection.hpp"

#include "mdbg_host.hpp"

// ---- Seam 1a: replacement of ReadSelectionFunctor (src/readSelection/ReadSelection.hpp:669-1161) -------------
class ReadSelectionFunctorGpu {
public:
    ReadSelection& _readSelection;
    mdbg_host::GpuReadSelectionFunctor& _gpu;           // shared by all OpenMP thread copies
    ReadSelectionFunctorGpu(ReadSelection& rs, mdbg_host::GpuReadSelectionFunctor& gpu) : _readSelection(rs), _gpu(gpu) {}
    ReadSelectionFunctorGpu(const ReadSelectionFunctorGpu& copy) : _readSelection(copy._readSelection), _gpu(copy._gpu) {}
    void operator()(const Read& read) {
        mdbg_host::Read r{read._index, read._header, read._seq, read._qual, read._datasetIndex};
#pragma omp critical(mdbg_gpu_feed)
        _gpu(r);
    }
};

void readSelectionOnGpu(ReadSelection& rs) {
    mdbg_host::Context gpuCtx((uint32_t)rs._params._minimizerSize, rs._params._minimizerDensity_assembly,
                              rs._params._useHomopolymerCompression,
                              std::vector<uint32_t>(rs._isRepetitiveMinimizer.begin(), rs._isRepetitiveMinimizer.end()));
    mdbg_host::GpuReadSelectionFunctor gpu(
        gpuCtx,
        [&](const mdbg_host::ReadMinimizers& r) {
            // exactly the arguments of ReadSelection::writeRead (ReadSelection.hpp:386)
            Read read;
            read._index = r.readIndex;
            read._seq.resize(r.readLength);
            vector<MinimizerType> minimizers(r.minimizers, r.minimizers + r.n);
            vector<u_int32_t> minimizerPos(r.positions, r.positions + r.n);
            vector<u_int8_t> minimizerDirections(r.directions, r.directions + r.n);
            vector<u_int8_t> minimizerQualities(r.qualities, r.qualities + r.n);
            rs.writeRead(read, minimizers, minimizerPos, minimizerDirections, minimizerQualities, r.meanReadQuality);
        },
        size_t(1) << 30, /*sideOutputs=*/true);
    ReadParserParallel readParser(rs._inputFilename, false, false, rs._nbCores);
    readParser.parse(ReadSelectionFunctorGpu(rs, gpu));                    // src/Commons.hpp:5827-5922
    gpu.flush();
    rs.computeReadStats();
    // Seam 1b: purgePalindromes (ReadSelection.hpp:1374-1385)
    int lastK = Commons::computeLastK(rs._params._minimizerDensity_assembly, rs._n50ReadLength, rs._params._kminmerSizeFirst, 0);
    mdbg_host::purgePalindromesAndWrite(gpuCtx, (uint32_t)rs._params._kminmerSizeFirst, (uint32_t)lastK,
                                        rs._inputDir + "/read_data_corrected.txt");
}

// ---- Seam 2: replacement of KminmerCounter::execute in CreateMdbg::createMDBG (src/graph/CreateMdbg.cpp:284-326) ----
void kminmerCounterOnGpu(CreateMdbg& g) {
    mdbg_host::Context gpuCtx((uint32_t)g._params._minimizerSize, g._params._minimizerDensity_assembly, false);
    mdbg_host::loadReadData(gpuCtx, g._outputDir + "/read_data_corrected.txt");
    mdbg_host::GpuKminmerCounter counter(gpuCtx, (uint32_t)g._kminmerSize, (uint32_t)g._minAbundance);
    counter.execute(g._outputDir + "/kminmerData_min.txt", g._outputDir + "/kminmerData_abundance.txt");
    g._nbKminmersTotal += counter._nbSolidKminmers + counter._nbRescuedKminmers;
    g._nbRescuedKminmers = counter._nbRescuedKminmers;
}
