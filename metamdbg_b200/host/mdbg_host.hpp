// mdbg_host.hpp -- C++ host-side mirror of the reference's functor interface for the
// sketch + count path, on top of the C ABI (include/mdbg_b200.h).  Header-only, C++17.
//
// It mirrors (metaMDBG source tree):
//   * ReadSelectionFunctor::operator()(const Read&)   src/readSelection/ReadSelection.hpp:669-1161
//     (the EncoderRLE + MinimizerParser::parse part, lines 682-690) -> GpuReadSelectionFunctor
//   * ReadSelection::purgePalindromes                  src/readSelection/ReadSelection.hpp:1374-1431
//   * CreateMdbg::KminmerCounter::execute              src/graph/CreateMdbg.hpp:3634-3883 -> GpuKminmerCounter
// and writes the reference's file formats:
//   read_data_corrected.txt   u32 n, u8 isCircular(0), u32 minimizers[n]   (Commons.hpp:7405-7440 reader)
//   kminmerData_min.txt       k * u32 per entry                            (Commons.hpp:4429-4446)
//   kminmerData_abundance.txt u128 hash (LE) + u32 abundance = 20 bytes    (Commons.hpp:4463-4471)
//
// Threading contract = the C ABI's: ReadParserParallel calls the functor from several OpenMP
// threads; feed() must be called under the parser's existing critical section (or from one thread).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mdbg_b200.h"

namespace mdbg_host {

// Same fields as the reference's Read (src/Commons.hpp:92-98).
struct Read {
    uint64_t _index = 0;
    std::string _header, _seq, _qual;
    uint64_t _datasetIndex = 0;
};

inline void check(mdbg_ctx* ctx, mdbg_status st, const char* what) {
    if (st != MDBG_OK) throw std::runtime_error(std::string(what) + ": " + mdbg_last_error(ctx));
}

// stdio handle that closes itself (read or write mode); put() throws on a short write (disk full, quota) instead of leaving a truncated
// read_data / kminmerData file behind for the next stage to misparse
class File {
public:
    File() = default;
    explicit File(const std::string& filename, const char* mode = "wb") : _name(filename) {
        _f = fopen(filename.c_str(), mode);
        if (!_f) throw std::runtime_error("cannot open " + filename);
    }
    ~File() { if (_f) fclose(_f); }
    File(const File&) = delete;
    File& operator=(const File&) = delete;
    FILE* get() const { return _f; }
    void put(const void* data, size_t size, size_t count) {
        if (count && fwrite(data, size, count, _f) != count) throw std::runtime_error("write failed: " + _name);
    }
    void close() {
        if (!_f) return;
        const int rc = fclose(_f);
        _f = nullptr;
        if (rc != 0) throw std::runtime_error("close failed: " + _name);
    }

private:
    FILE* _f = nullptr;
    std::string _name;
};

class Context {
public:
    Context(uint32_t minimizerSize, float density, bool useHomopolymerCompression,
            const std::vector<uint32_t>& repetitiveMinimizers = {}, int device = 0) {
        mdbg_params p{};
        p.minimizer_size = minimizerSize;
        p.density = density;
        p.use_hpc = useHomopolymerCompression ? 1 : 0;
        p.blacklist = repetitiveMinimizers.empty() ? nullptr : repetitiveMinimizers.data();
        p.n_blacklist = repetitiveMinimizers.size();
        mdbg_status st = mdbg_ctx_create(device, &p, &_ctx);
        if (st != MDBG_OK) throw std::runtime_error(std::string("mdbg_ctx_create: ") + mdbg_last_error(nullptr));
    }
    ~Context() { mdbg_ctx_destroy(_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mdbg_ctx* get() const { return _ctx; }

private:
    mdbg_ctx* _ctx = nullptr;
};

// Per-read result handed to the sink, in input order (what ReadSelection::writeRead receives).
struct ReadMinimizers {
    uint64_t readIndex;
    uint32_t readLength;
    uint32_t n;
    const uint32_t* minimizers;
    const uint32_t* positions;
    const uint8_t* directions;
    const uint8_t* qualities;      // per-minimizer min base quality (nullptr unless side outputs are on)
    float meanReadQuality;         // NaN without qualities
    bool lowComplexity;
};

// Batching replacement of ReadSelectionFunctor: reads are buffered until `batchBases` bases are
// pending, sketched on the GPU in one call, appended to the device-resident minimizer store and
// reported read by read, in input order, to `sink`.
class GpuReadSelectionFunctor {
public:
    using Sink = std::function<void(const ReadMinimizers&)>;

    // sideOutputs = also compute mean read quality, the low-complexity filter and per-minimizer qualities
    // (the rest of ReadSelectionFunctor::operator(), ReadSelection.hpp:870-920, 1047-1138)
    GpuReadSelectionFunctor(Context& ctx, Sink sink, size_t batchBases = size_t(1) << 30, bool sideOutputs = false)
        : _ctx(ctx), _sink(std::move(sink)), _batchBases(batchBases), _sideOutputs(sideOutputs) {
        _offsets.push_back(0);
        if (_sideOutputs) check(_ctx.get(), mdbg_ctx_set_read_filters(_ctx.get(), 1), "mdbg_ctx_set_read_filters");
    }

    void operator()(const Read& read) {                 // same call shape as the reference functor
        if (_sideOutputs) {
            if (!read._qual.empty() && read._qual.size() != read._seq.size())
                throw std::runtime_error("quality/sequence length mismatch in read " + read._header);
            if (_indices.empty()) _haveQual = !read._qual.empty();
            else if (_haveQual != !read._qual.empty()) flush();      // FASTA and FASTQ records never share a batch
            if (_indices.empty()) _haveQual = !read._qual.empty();
            if (_haveQual) _quals.insert(_quals.end(), read._qual.begin(), read._qual.end());
        }
        _bases.insert(_bases.end(), read._seq.begin(), read._seq.end());
        _offsets.push_back(_bases.size());
        _indices.push_back(read._index);
        _nbBases += read._seq.size();
        if (_bases.size() >= _batchBases) flush();
    }

    void flush() {
        const uint32_t n = (uint32_t)_indices.size();
        if (n == 0) return;
        mdbg_sketch_out out{};
        mdbg_aux_out aux{};
        if (_sideOutputs)
            check(_ctx.get(),
                  mdbg_sketch_batch_q(_ctx.get(), reinterpret_cast<const uint8_t*>(_bases.data()),
                                      _haveQual ? reinterpret_cast<const uint8_t*>(_quals.data()) : nullptr,
                                      _offsets.data(), n, /*append_to_store=*/1, &out, &aux),
                  "mdbg_sketch_batch_q");
        else
            check(_ctx.get(),
                  mdbg_sketch_batch(_ctx.get(), reinterpret_cast<const uint8_t*>(_bases.data()), _offsets.data(), n,
                                    /*append_to_store=*/1, &out),
                  "mdbg_sketch_batch");
        for (uint32_t r = 0; r < n; r++) {
            const uint64_t lo = out.min_offsets[r], hi = out.min_offsets[r + 1];
            if (_sink)
                _sink(ReadMinimizers{_indices[r], (uint32_t)(_offsets[r + 1] - _offsets[r]), (uint32_t)(hi - lo),
                                     out.minimizers + lo, out.positions + lo, out.directions + lo,
                                     _sideOutputs ? aux.qualities + lo : nullptr,
                                     _sideOutputs ? aux.mean_quality[r] : 0.0f,
                                     _sideOutputs ? aux.low_complexity[r] != 0 : false});
            _nbSelectedMinimizers += hi - lo;
        }
        _nbReads += n;
        _quals.clear();
        _bases.clear();
        _offsets.assign(1, 0);
        _indices.clear();
    }

    uint64_t nbReads() const { return _nbReads; }
    uint64_t nbBases() const { return _nbBases; }
    uint64_t nbSelectedMinimizers() const { return _nbSelectedMinimizers; }

private:
    Context& _ctx;
    Sink _sink;
    size_t _batchBases;
    bool _sideOutputs = false, _haveQual = false;
    std::vector<char> _quals;
    std::vector<char> _bases;
    std::vector<uint64_t> _offsets;
    std::vector<uint64_t> _indices;
    uint64_t _nbReads = 0, _nbBases = 0, _nbSelectedMinimizers = 0;
};

// ReadSelection::writeRead + computeReadStats (ReadSelection.hpp:305-491): read_data_init.txt records
//   u32 n, u8 isCircular, u32 min[n], u32 pos[n], u8 dir[n], u8 qual[n], f32 meanReadQuality, u32 readLength
// and read_stats.txt (u64 nbReads, u32 n50, f32 density, u64 nbBases, f32 avgQuality, u32 meanLength,
// u64 nbSelectedMinimizers).  Records arrive in input order, so no re-ordering queue is needed.
class ReadDataWriter {
public:
    ReadDataWriter(const std::string& filename, uint32_t minimizerSize) : _f(filename), _minimizerSize(minimizerSize) {}

    void write(const ReadMinimizers& r) {
        const uint32_t size = r.n;
        const uint8_t isCircular = 0;
        if (size && !r.qualities) throw std::runtime_error("ReadDataWriter needs the side outputs (sideOutputs = true)");
        _f.put(&size, 4, 1);
        _f.put(&isCircular, 1, 1);
        _f.put(r.minimizers, 4, size);
        _f.put(r.positions, 4, size);
        _f.put(r.directions, 1, size);
        _f.put(r.qualities, 1, size);
        _f.put(&r.meanReadQuality, 4, 1);
        _f.put(&r.readLength, 4, 1);
        _allReadSizes.push_back(r.readLength);
        _nbSelectedMinimizers += size;
        _nbKmers += (uint64_t)r.readLength - _minimizerSize + 1;     // ReadSelection.hpp:474 (unsigned wrap kept)
        _nbBases += r.readLength;
        if (!(r.meanReadQuality < 0.0f)) {                           // ReadSelection.hpp:906-917, minReadQuality = 0
            _readQualitySum += r.meanReadQuality;
            _readQualityN += 1;
        }
    }

    void close() { _f.close(); }

    static uint32_t computeN50(std::vector<uint32_t> lengths) {     // Utils::computeN50, Commons.hpp:2291-2322
        if (lengths.empty()) return 0;
        std::sort(lengths.begin(), lengths.end(), std::greater<uint32_t>());
        std::vector<uint64_t> cumuls;
        uint64_t cumul = 0;
        for (uint32_t x : lengths) { cumul += x; cumuls.push_back(cumul); }
        std::reverse(lengths.begin(), lengths.end());
        std::reverse(cumuls.begin(), cumuls.end());
        uint32_t n50 = lengths.back();
        const uint64_t halfsize = cumuls[0] / 2;
        for (size_t i = 0; i < lengths.size(); i++)
            if (cumuls[i] < halfsize) { n50 = lengths[i]; break; }
        return n50;
    }

    static uint64_t computeMeanLength(const std::vector<uint32_t>& lengths) {   // Utils::computeMeanLength, Commons.hpp:2324-2336
        long double sum = 0, cnt = 0;
        for (uint32_t x : lengths) { sum += x; cnt += 1; }
        return (uint64_t)(sum / cnt);                                 // 0/0 -> NaN -> conversion as upstream
    }

    uint32_t n50() const { return computeN50(_allReadSizes); }

    void writeReadStats(const std::string& filename) {
        const uint64_t nbReads = _allReadSizes.size();
        const uint32_t n50v = n50();
        const uint32_t meanLength = (uint32_t)computeMeanLength(_allReadSizes);
        const float minimizerDensity = (long double)_nbSelectedMinimizers / (long double)_nbKmers;
        const float averageQuality = _readQualitySum / _readQualityN;
        File f(filename);
        f.put(&nbReads, 8, 1);
        f.put(&n50v, 4, 1);
        f.put(&minimizerDensity, 4, 1);
        f.put(&_nbBases, 8, 1);
        f.put(&averageQuality, 4, 1);
        f.put(&meanLength, 4, 1);
        f.put(&_nbSelectedMinimizers, 8, 1);
        f.close();
    }

private:
    File _f;
    uint32_t _minimizerSize;
    std::vector<uint32_t> _allReadSizes;
    uint64_t _nbSelectedMinimizers = 0, _nbKmers = 0, _nbBases = 0;
    long double _readQualitySum = 0, _readQualityN = 0;
};

// Commons::computeLastK (Commons.hpp:1726-1741) as ReadSelection::purgePalindromes calls it (ReadSelection.hpp:1376,
// maxK = 0): size_t * float * 2.0f truncated, at least firstK + 2.
inline uint32_t computeLastK(float minimizerDensityAssembly, size_t n50ReadLength, size_t firstK) {
    const size_t lastK = n50ReadLength * minimizerDensityAssembly * 2.0f;
    return (uint32_t)std::max(lastK, firstK + 2);
}

// ReadSelection::purgePalindromes + the read_data_corrected.txt writer.
inline uint64_t purgePalindromesAndWrite(Context& ctx, uint32_t firstK, uint32_t lastK, const std::string& filename) {
    uint64_t changed = 0;
    check(ctx.get(), mdbg_purge_palindromes(ctx.get(), firstK, lastK, &changed), "mdbg_purge_palindromes");
    uint64_t nReads = 0, nMins = 0;
    check(ctx.get(), mdbg_store_size(ctx.get(), &nReads, &nMins), "mdbg_store_size");
    std::vector<uint64_t> offs(nReads + 1);
    std::vector<uint32_t> mins(nMins + 1);
    check(ctx.get(), mdbg_store_fetch(ctx.get(), offs.data(), mins.data()), "mdbg_store_fetch");
    File f(filename);
    for (uint64_t r = 0; r < nReads; r++) {
        const uint32_t size = (uint32_t)(offs[r + 1] - offs[r]);
        const uint8_t isCircular = 0;                    // CONTIG_LINEAR
        f.put(&size, sizeof size, 1);
        f.put(&isCircular, 1, 1);
        f.put(mins.data() + offs[r], sizeof(uint32_t), size);
    }
    f.close();
    return changed;
}

// KminmerParserParallel's input side (src/Commons.hpp:7367-7495, records u32 n, u8 isCircular, u32[n]): load a
// read_data_corrected.txt (or unitig_data.txt) into the context's device store, 256 MB of minimizers at a time.
inline uint64_t loadReadData(Context& ctx, const std::string& filename) {
    File in(filename, "rb");
    FILE* f = in.get();
    std::vector<uint32_t> mins;
    std::vector<uint64_t> offs{0};
    uint64_t nReads = 0;
    auto flush = [&]() {
        if (offs.size() > 1)
            check(ctx.get(), mdbg_store_append(ctx.get(), mins.data(), offs.data(), (uint32_t)(offs.size() - 1)),
                  "mdbg_store_append");
        mins.clear();
        offs.assign(1, 0);
    };
    for (;;) {
        uint32_t size;
        uint8_t isCircular;
        if (fread(&size, 4, 1, f) != 1) break;
        if (fread(&isCircular, 1, 1, f) != 1) throw std::runtime_error("truncated record in " + filename);
        const size_t at = mins.size();
        mins.resize(at + size);
        if (size && fread(mins.data() + at, 4, size, f) != size) throw std::runtime_error("truncated record in " + filename);
        offs.push_back(mins.size());
        nReads++;
        if (mins.size() > (size_t(64) << 20)) flush();
    }
    flush();
    return nReads;
}

// CreateMdbg::KminmerCounter (first pass): counts every k-min-mer of the device-resident reads and
// writes kminmerData_min.txt / kminmerData_abundance.txt.
class GpuKminmerCounter {
public:
    GpuKminmerCounter(Context& ctx, uint32_t kminmerSize, uint32_t minAbundance)
        : _ctx(ctx), _k(kminmerSize), _minAbundance(minAbundance) {}

    // minAbundance <= 1 is metaMDBG's default mode: the abundance >= 2 entries are followed by the rescued
    // abundance-1 entries (CreateMdbg.cpp:309-319, rescueKminmers)
    void execute(const std::string& kminmerFile, const std::string& abundanceFile) {
        check(_ctx.get(), mdbg_count_begin(_ctx.get(), _k, 0), "mdbg_count_begin");
        check(_ctx.get(), mdbg_count_add_store(_ctx.get(), 0, UINT64_MAX), "mdbg_count_add_store");
        if (_minAbundance <= 1) check(_ctx.get(), mdbg_count_rescue(_ctx.get(), &_nbReadsRescued), "mdbg_count_rescue");
        mdbg_table_out t{};
        check(_ctx.get(), mdbg_count_finalize(_ctx.get(), _minAbundance, &t), "mdbg_count_finalize");
        File fk(kminmerFile), fa(abundanceFile);
        fk.put(t.kminmers, sizeof(uint32_t), (size_t)t.n_entries * _k);
        std::vector<unsigned char> rec((size_t)t.n_entries * 20);     // 20-byte records, written in one go
        for (uint64_t i = 0; i < t.n_entries; i++) {
            memcpy(&rec[i * 20], t.hashes + 2 * i, 16);    // u128 little-endian: low = Murmur h2, high = h1
            memcpy(&rec[i * 20 + 16], t.abundances + i, 4);
        }
        fa.put(rec.data(), 20, (size_t)t.n_entries);
        fk.close();
        fa.close();
        _nbSolidKminmers = t.n_entries - t.n_rescued;
        _nbRescuedKminmers = t.n_rescued;
        _nbKminmers = t.n_instances;
        _nbDistinct = t.n_distinct;
        _checksum = t.checksum;
    }

    uint64_t _nbKminmers = 0, _nbSolidKminmers = 0, _nbRescuedKminmers = 0, _nbReadsRescued = 0, _nbDistinct = 0,
             _checksum = 0;

private:
    Context& _ctx;
    uint32_t _k, _minAbundance;
};

// `graph` at k > firstK (CreateMdbg.cpp:386-468): the abundance of a k-min-mer is the minimum over its two
// (k-1)-min-mers of the previous k's table (_kminmerAbundances; getRefinedAbundance CreateMdbg.hpp:3933-4005 for
// k = firstK+1, IndexKminmerFunctor CreateMdbg.hpp:951-1465 beyond), kept when > 1.  The previous table is the one
// the context holds from the k before (optionally patched with the contig stage's refined abundances through
// mdbg_prev_load between two calls); the reads stay in the device-resident store the whole time.
class GpuNextKCounter {
public:
    GpuNextKCounter(Context& ctx, uint32_t minAbundance) : _ctx(ctx), _minAbundance(minAbundance) {}

    // Derives the table of k from the context's current table (which must be the one of k-1) and writes
    // kminmerData_min / kminmerData_abundance for this k.  The table is sized from the previous table's entry
    // count; the pass is idempotent, so a table that fills up is redone at the worst-case size.
    void execute(uint32_t k, const std::string& kminmerFile, const std::string& abundanceFile) {
        uint64_t prevEntries = 0;
        check(_ctx.get(), mdbg_count_stats(_ctx.get(), _minAbundance, &prevEntries, nullptr, nullptr, nullptr), "mdbg_count_stats");
        check(_ctx.get(), mdbg_prev_from_current(_ctx.get(), _minAbundance), "mdbg_prev_from_current");
        const uint64_t expected = 2 * (prevEntries < 256 ? 256 : prevEntries);
        check(_ctx.get(), mdbg_count_begin(_ctx.get(), k, expected), "mdbg_count_begin");
        mdbg_status st = mdbg_count_add_store_next_k(_ctx.get(), 0, UINT64_MAX);
        if (st == MDBG_ERR_TABLE_FULL) {
            check(_ctx.get(), mdbg_count_begin(_ctx.get(), k, 0), "mdbg_count_begin");
            st = mdbg_count_add_store_next_k(_ctx.get(), 0, UINT64_MAX);
            _nbResized++;
        }
        check(_ctx.get(), st, "mdbg_count_add_store_next_k");
        mdbg_table_out t{};
        check(_ctx.get(), mdbg_count_finalize(_ctx.get(), _minAbundance, &t), "mdbg_count_finalize");
        File fk(kminmerFile), fa(abundanceFile);
        fk.put(t.kminmers, sizeof(uint32_t), (size_t)t.n_entries * k);
        std::vector<unsigned char> rec((size_t)t.n_entries * 20);
        for (uint64_t i = 0; i < t.n_entries; i++) {
            memcpy(&rec[i * 20], t.hashes + 2 * i, 16);
            memcpy(&rec[i * 20 + 16], t.abundances + i, 4);
        }
        fa.put(rec.data(), 20, (size_t)t.n_entries);
        fk.close();
        fa.close();
        _nbKminmers = t.n_entries;
        _checksum = t.checksum;
    }

    uint64_t _nbKminmers = 0, _checksum = 0, _nbResized = 0;

private:
    Context& _ctx;
    uint32_t _minAbundance;
};

// CreateMdbg::EdgeIndexer (CreateMdbg.hpp:4010-4232): the dereplicated hash128 keys of the (k-1)-prefixes / suffixes of
// the nodes of the context's current table, written as edges.bin is (16-byte little-endian u128 per key; the
// reference builds its edge MPHF over this file, so the key order is irrelevant).
class GpuEdgeIndexer {
public:
    GpuEdgeIndexer(Context& ctx, uint32_t minAbundance) : _ctx(ctx), _minAbundance(minAbundance) {}

    void execute(const std::string& edgeFile) {
        mdbg_edges_out e{};
        check(_ctx.get(), mdbg_edges_index(_ctx.get(), _minAbundance, &e), "mdbg_edges_index");
        File f(edgeFile);
        f.put(e.hashes, 16, (size_t)e.n_edges);            // {low, high} words = the u128 as it lies on disk
        f.close();
        _nbEdges = e.n_edges;
        _checksum = e.checksum;
    }

    uint64_t _nbEdges = 0, _checksum = 0;

private:
    Context& _ctx;
    uint32_t _minAbundance;
};

// CreateMdbg::computeUnitigNodes + computeDeterministicUnitigs + dumpUnitigAbundances (CreateMdbg.cpp:1521-1598,
// 1001-1043, 3335-3390): the unitig nodes of the context's current table as the reference's two files --
//   unitigGraph.nodes.bin             per unitig: u32 size, size x u32 minimizers, u32 unitigIndex (= 2 * record number)
//   unitigGraph.nodes.abundances.bin  per unitig: u32 unitigIndex, u32 count, count x u32 abundance of its k-min-mers
//   unitigGraph.edges.successors.bin  per unitig: u32 unitigIndex, u32 nbSuccessors, successors, u32 nbPredecessors,
//                                     predecessors (indexUnitigEdges + computeUnitigEdges, CreateMdbg.cpp:2915-3245)
// in the reference's deterministic order (ascending hash128 of the normalized minimizer sequence).
class GpuUnitigBuilder {
public:
    GpuUnitigBuilder(Context& ctx, uint32_t minAbundance) : _ctx(ctx), _minAbundance(minAbundance) {}

    void execute(const std::string& nodeFile, const std::string& abundanceFile, const std::string& edgeFile = "") {
        mdbg_unitigs_out u{};
        check(_ctx.get(), mdbg_unitigs_build(_ctx.get(), _minAbundance, &u), "mdbg_unitigs_build");
        if (!edgeFile.empty() && u.edge_offsets) {       // dumpUnitigEdge (CreateMdbg.cpp:2853-2912), record i = unitigIndex 2 i
            File fe(edgeFile);
            for (uint64_t i = 0; i < u.n_unitigs; i++) {
                const uint32_t from = (uint32_t)(2 * i);
                fe.put(&from, 4, 1);
                for (int o = 0; o < 2; o++) {            // successors, then predecessors
                    const uint64_t lo = u.edge_offsets[2 * i + o], hi = u.edge_offsets[2 * i + o + 1];
                    const uint32_t nb = (uint32_t)(hi - lo);
                    fe.put(&nb, 4, 1);
                    fe.put(u.edge_targets + lo, 4, nb);
                }
            }
            fe.close();
            _nbUnitigEdges = u.n_unitig_edges;
            _checksumEdges = u.checksum_edges;
        }
        File fn(nodeFile), fa(abundanceFile);
        const uint64_t km1 = u.k - 1;
        for (uint64_t i = 0; i < u.n_unitigs; i++) {
            const uint64_t j = u.order[i];
            const uint32_t size = (uint32_t)(u.offsets[j + 1] - u.offsets[j]), index = (uint32_t)(2 * i), nb = size - (uint32_t)km1;
            fn.put(&size, 4, 1);
            fn.put(u.minimizers + u.offsets[j], 4, size);
            fn.put(&index, 4, 1);
            fa.put(&index, 4, 1);
            fa.put(&nb, 4, 1);
            fa.put(u.node_abundances + (u.offsets[j] - j * km1), 4, nb);
        }
        fn.close();
        fa.close();
        _nbUnitigs = u.n_unitigs;
        _nbCircular = u.n_circular;
        _checksumNodes = u.checksum_nodes;
        _checksumAbundances = u.checksum_abundances;
    }

    uint64_t _nbUnitigs = 0, _nbCircular = 0, _checksumNodes = 0, _checksumAbundances = 0, _nbUnitigEdges = 0, _checksumEdges = 0;

private:
    Context& _ctx;
    uint32_t _minAbundance;
};

}  // namespace mdbg_host
