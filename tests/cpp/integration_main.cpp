// The reference's own readSelection stage (kseq FASTQ parser, ordered record writer, read stats) driven by the GPU
// functor of INTEGRATION.md instead of ReadSelectionFunctor: metaMDBG sources + libmdbg_b200.so in one binary.
// Built by oracle/Makefile into oracle/_ref/mdbg_ref_integrated (test infrastructure; needs /root/reference to
// build, travels prebuilt to the GPU box).  usage: mdbg_ref_integrated <input.txt> <dir> <l> <density> <hpc> <threads>
#include "integration_binding.cpp"

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s <input.txt> <dir> <l> <density> <hpc> <threads>\n", argv[0]); return 2; }
    ReadSelection rs;
    rs._inputFilename = argv[1];
    rs._inputDir = argv[2];
    rs._outputFilename = std::string(argv[2]) + "/read_data_init.txt";
    rs._nbCores = atoi(argv[6]);
    rs._minReadQuality = 0;
    rs._outputQuality = true;
    rs._skipCorrection = true;
    rs._params._minimizerSize = atoi(argv[3]);
    rs._params._minimizerDensity_assembly = (float)atof(argv[4]);
    rs._params._minimizerDensity_correction = 0.025f;
    rs._params._useHomopolymerCompression = atoi(argv[5]) != 0;
    rs._params._kminmerSize = 4;
    rs._params._kminmerSizeFirst = 4;
    rs._params._kminmerSizePrev = 3;
    // what ReadSelection::execute / readSelection initialise before parsing (ReadSelection.hpp:92-111, 251-262)
    rs._nbKmers = 0; rs._nbBases = 0; rs._nbSelectedMinimizers = 0; rs._nbLowQualityReads = 0; rs._nbLowComplexityReads = 0;
    rs._readQualitySum = 0; rs._readQualityN = 0;
    rs._nextReadIndexWriter = 0; rs._debug_nbMinimizers = 0;
    rs._file_readData = ofstream(rs._outputFilename);
    try {
        readSelectionOnGpu(rs);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    rs._file_readData.close();
    return 0;
}
