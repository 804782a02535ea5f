/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin extern "C" wrapper that #includes metaMDBG's OWN sources where they
 * lie under /root/reference (no reference code is copied into this repo) and
 * exposes the hot-path functions so that
 *   (1) oracle/mdbg_oracle.c can be validated against the real thing, and
 *   (2) bench.py --impl reference / cpu_baseline can time the reference's own
 *       CPU implementation (kind = "reference").
 * Built by oracle/Makefile into oracle/_ref/libmdbg_ref.so (git-ignored, but
 * shipped to the GPU box by gpurun).  Everything computational below is a
 * call into reference code:
 *   EncoderRLE::execute            src/Commons.hpp:4163-4203
 *   MinimizerParser::parse         src/utils/kmer/Kmer.hpp:1373-1456
 *   KmerModel::iterate             src/utils/kmer/Kmer.hpp:570-589
 *   MurmurHash3_x64_128[_original] src/utils/MurmurHash3.cpp:246-405
 *   Commons::purgePalindrome       src/Commons.hpp:1617-1723
 *   MDBG::getKminmers_complete     src/Commons.hpp:5150-5367
 *   KmerVec::{normalize,hash128,operator<}  src/Commons.hpp:740-1005
 *   Commons::sortParallel          src/Commons.hpp:1537-1574
 *   CreateMdbg::KminmerCounter     src/graph/CreateMdbg.hpp:3591-4007   (ref_graph_firstpass, ref_graph_next_k)
 *   CreateMdbg::rescueKminmers     src/graph/CreateMdbg.hpp:4517-4640   (ref_graph_firstpass)
 *   CreateMdbg::IndexKminmerFunctor src/graph/CreateMdbg.hpp:951-1465   (ref_graph_next_k)
 * ref_count (in-memory sort/count) is the only place where logic is written here: the run-length
 * count + ">=2" filter of KminmerCounter::dereplicatePartition/dumpKminmer without its file I/O;
 * ref_graph_* drive the reference's own classes through their file contract in a scratch directory.
 */
#include "Commons.hpp"
#include "graph/CreateMdbg.hpp"
#include "readSelection/ReadSelection.hpp"

#include <cstdint>
#include <cstdlib>
#include <cstring>

extern "C" {

uint64_t ref_murmur3_x64_128_h1(const void* key, int len, uint32_t seed) {
    return MurmurHash3_x64_128(key, len, seed);
}

void ref_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t out[2]) {
    MurmurHash3_x64_128_original(key, len, seed, out);
}

double ref_minimizer_bound(float density) {
    unordered_set<MinimizerType> none;
    MinimizerParser parser(15, density, none);
    return parser._minimizerBound;
}

size_t ref_hpc(const char* seq, size_t len, int hpc, char* out, uint64_t* rle_pos) {
    EncoderRLE enc;
    string rle;
    vector<u_int64_t> pos;
    string s(seq, len);
    enc.execute(s.c_str(), s.size(), rle, pos, hpc != 0);
    memcpy(out, rle.data(), rle.size());
    if (rle_pos) memcpy(rle_pos, pos.data(), pos.size() * sizeof(u_int64_t));
    return rle.size();
}

size_t ref_lmers(const char* seq, size_t len, int l, uint64_t* values, uint8_t* dirs) {
    KmerModel model(l);
    vector<u_int64_t> kmers;
    vector<u_int8_t> kdirs;
    string s(seq, len);
    if (!model.iterate(s.c_str(), s.size(), kmers, kdirs)) return 0;
    memcpy(values, kmers.data(), kmers.size() * sizeof(u_int64_t));
    memcpy(dirs, kdirs.data(), kdirs.size());
    return kmers.size();
}

/* ReadSelectionFunctor::operator() lines 682-690: EncoderRLE then parse. */
static size_t sketch_one(EncoderRLE& enc, MinimizerParser& parser, const char* seq, size_t len, int hpc,
                         vector<MinimizerType>& mins, vector<u_int32_t>& pos, vector<u_int8_t>& dirs) {
    string s(seq, len);
    string rle;
    vector<u_int64_t> rlePositions;
    enc.execute(s.c_str(), s.size(), rle, rlePositions, hpc != 0);
    parser.parse(rle, mins, pos, dirs);
    return mins.size();
}

size_t ref_sketch_read(const char* seq, size_t len, int l, float density, int hpc,
                       const uint32_t* blacklist, size_t n_blacklist,
                       uint32_t* minimizers, uint32_t* positions, uint8_t* directions, size_t cap) {
    unordered_set<MinimizerType> bl(blacklist, blacklist + n_blacklist);
    MinimizerParser parser(l, density, bl);
    EncoderRLE enc;
    vector<MinimizerType> mins;
    vector<u_int32_t> pos;
    vector<u_int8_t> dirs;
    size_t n = sketch_one(enc, parser, seq, len, hpc, mins, pos, dirs);
    for (size_t i = 0; i < n && i < cap; i++) {
        minimizers[i] = mins[i];
        positions[i] = pos[i];
        directions[i] = dirs[i];
    }
    return n;
}

size_t ref_purge_palindrome(const uint32_t* m, size_t n, size_t first_k, size_t last_k, uint32_t* out) {
    vector<MinimizerType> v(m, m + n);
    vector<MinimizerType> r = Commons::purgePalindrome(v, first_k, last_k);
    memcpy(out, r.data(), r.size() * sizeof(MinimizerType));
    return r.size();
}

size_t ref_apply_density(const uint32_t* m, size_t n, float density, uint32_t* out) {
    vector<MinimizerType> mins(m, m + n), f;
    vector<u_int32_t> pos, fp;
    vector<u_int8_t> dirs, quals, fd, fq;
    Utils::applyDensityThreshold(density, mins, pos, dirs, quals, f, fp, fd, fq);     // src/Commons.hpp:2507-2550
    memcpy(out, f.data(), f.size() * sizeof(MinimizerType));
    return f.size();
}

size_t ref_kminmers(const uint32_t* m, size_t n, int k, uint32_t* vecs, uint8_t* reversed) {
    vector<MinimizerType> mins(m, m + n);
    vector<u_int32_t> pos(n, 0);
    vector<u_int8_t> quals(n, 0);
    vector<ReadKminmerComplete> kms;
    MDBG::getKminmers_complete(k, mins, pos, kms, 0, quals);
    for (size_t i = 0; i < kms.size(); i++) {
        memcpy(vecs + i * (size_t)k, kms[i]._vec._kmers.data(), (size_t)k * sizeof(MinimizerType));
        if (reversed) reversed[i] = kms[i]._isReversed ? 1 : 0;
    }
    return kms.size();
}

void ref_hash128(const uint32_t* vec, int k, uint64_t out[2]) {
    KmerVec v;
    v._kmers.assign(vec, vec + k);
    u_int128_t h = v.hash128();
    out[0] = (uint64_t)(h >> 64);
    out[1] = (uint64_t)h;
}

/* First-pass count over an in-memory minimizer CSR using the reference's
 * types: getKminmers_complete -> KmerVec -> sortParallel -> run-length count
 * -> keep abundance >= max(2, min_abundance) -> hash128. */
size_t ref_count(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k, uint32_t min_abundance,
                 int n_threads, uint32_t** vecs_out, uint64_t** hashes_out, uint32_t** abundances_out,
                 uint64_t* n_instances, uint64_t* n_distinct) {
    vector<KmerVec> all;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        vector<KmerVec> local;
        vector<ReadKminmerComplete> kms;
#pragma omp for schedule(dynamic, 64)
        for (size_t r = 0; r < n_reads; r++) {
            size_t n = offs[r + 1] - offs[r];
            vector<MinimizerType> m(mins + offs[r], mins + offs[r] + n);
            vector<u_int32_t> pos(n, 0);
            vector<u_int8_t> quals(n, 0);
            MDBG::getKminmers_complete(k, m, pos, kms, (int)r, quals);
            for (auto& km : kms) local.push_back(km._vec);
        }
#pragma omp critical
        all.insert(all.end(), local.begin(), local.end());
    }
    Commons::sortParallel(all, all.size(), n_threads);

    vector<uint32_t> vecs;
    vector<uint64_t> hashes;
    vector<uint32_t> abs_;
    uint64_t distinct = 0;
    size_t i = 0;
    while (i < all.size()) {
        size_t j = i + 1;
        while (j < all.size() && all[j] == all[i]) j++;
        uint32_t ab = (uint32_t)(j - i);
        distinct++;
        if (ab > 1 && ab >= min_abundance) {
            vecs.insert(vecs.end(), all[i]._kmers.begin(), all[i]._kmers.end());
            u_int128_t h = all[i].hash128();
            hashes.push_back((uint64_t)(h >> 64));
            hashes.push_back((uint64_t)h);
            abs_.push_back(ab);
        }
        i = j;
    }
    if (n_instances) *n_instances = all.size();
    if (n_distinct) *n_distinct = distinct;
    size_t n = abs_.size();
    *vecs_out = (uint32_t*)malloc((vecs.size() + 1) * sizeof(uint32_t));
    *hashes_out = (uint64_t*)malloc((hashes.size() + 1) * sizeof(uint64_t));
    *abundances_out = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    memcpy(*vecs_out, vecs.data(), vecs.size() * sizeof(uint32_t));
    memcpy(*hashes_out, hashes.data(), hashes.size() * sizeof(uint64_t));
    memcpy(*abundances_out, abs_.data(), n * sizeof(uint32_t));
    return n;
}

/* The reference's CPU hot path end to end, OpenMP over reads as
 * ReadParserParallel does (src/Commons.hpp:5846-5911, one functor copy per
 * thread): HPC -> sketch -> [purgePalindrome] -> k-min-mers -> sort-count.
 * Returns the number of solid k-min-mers; *checksum = sum ab*low64(hash),
 * *n_minimizers = total selected minimizers.  Used ONLY as the timed CPU
 * baseline and as a cross-check. */
size_t ref_pipeline(const char* bases, const uint64_t* offsets, size_t n_reads, int l, float density, int hpc,
                    int k, int purge_last_k, uint32_t min_abundance, int n_threads, uint64_t* n_minimizers,
                    uint64_t* checksum, double* seconds_sketch, double* seconds_count) {
    if (n_threads < 1) n_threads = 1;
    vector<vector<MinimizerType>> per_read(n_reads);
    unordered_set<MinimizerType> bl;
    auto t0 = high_resolution_clock::now();
#pragma omp parallel num_threads(n_threads)
    {
        MinimizerParser parser(l, density, bl);
        EncoderRLE enc;
        vector<MinimizerType> mins;
        vector<u_int32_t> pos;
        vector<u_int8_t> dirs;
#pragma omp for schedule(dynamic, 16)
        for (size_t r = 0; r < n_reads; r++) {
            sketch_one(enc, parser, bases + offsets[r], offsets[r + 1] - offsets[r], hpc, mins, pos, dirs);
            if (purge_last_k > 0) per_read[r] = Commons::purgePalindrome(mins, 4, purge_last_k);
            else per_read[r] = mins;
        }
    }
    auto t1 = high_resolution_clock::now();
    vector<uint64_t> offs(n_reads + 1, 0);
    for (size_t r = 0; r < n_reads; r++) offs[r + 1] = offs[r] + per_read[r].size();
    vector<uint32_t> flat(offs[n_reads] + 1);
    for (size_t r = 0; r < n_reads; r++)
        if (!per_read[r].empty()) memcpy(flat.data() + offs[r], per_read[r].data(), per_read[r].size() * 4);
    uint32_t* v; uint64_t* h; uint32_t* a;
    uint64_t ni, nd;
    size_t n = ref_count(flat.data(), offs.data(), n_reads, k, min_abundance, n_threads, &v, &h, &a, &ni, &nd);
    auto t2 = high_resolution_clock::now();
    uint64_t cs = 0;
    for (size_t i = 0; i < n; i++) cs += (uint64_t)a[i] * h[2 * i + 1];
    free(v); free(h); free(a);
    if (n_minimizers) *n_minimizers = offs[n_reads];
    if (checksum) *checksum = cs;
    if (seconds_sketch) *seconds_sketch = duration<double>(t1 - t0).count();
    if (seconds_count) *seconds_count = duration<double>(t2 - t1).count();
    return n;
}

/* Same with the ONT-style density re-threshold between sketch and purge: reads are sketched at `density`
 * (--density-correction), Utils::applyDensityThreshold (src/Commons.hpp:2507-2550) keeps the minimizers that also
 * pass `assembly_density`, then purgePalindrome and the count. */
size_t ref_pipeline2(const char* bases, const uint64_t* offsets, size_t n_reads, int l, float density, int hpc,
                     float assembly_density, int k, int purge_last_k, uint32_t min_abundance, int n_threads,
                     uint64_t* n_minimizers_sketch, uint64_t* n_minimizers, uint64_t* checksum, double* seconds_sketch,
                     double* seconds_count) {
    if (n_threads < 1) n_threads = 1;
    vector<vector<MinimizerType>> per_read(n_reads);
    unordered_set<MinimizerType> bl;
    uint64_t n_sketch = 0;
    auto t0 = high_resolution_clock::now();
#pragma omp parallel num_threads(n_threads) reduction(+ : n_sketch)
    {
        MinimizerParser parser(l, density, bl);
        EncoderRLE enc;
        vector<MinimizerType> mins, minsF;
        vector<u_int32_t> pos, posF;
        vector<u_int8_t> dirs, dirsF, quals, qualsF;
#pragma omp for schedule(dynamic, 16)
        for (size_t r = 0; r < n_reads; r++) {
            sketch_one(enc, parser, bases + offsets[r], offsets[r + 1] - offsets[r], hpc, mins, pos, dirs);
            n_sketch += mins.size();
            const vector<MinimizerType>* use = &mins;
            if (assembly_density > 0) {
                Utils::applyDensityThreshold(assembly_density, mins, pos, dirs, quals, minsF, posF, dirsF, qualsF);
                use = &minsF;
            }
            if (purge_last_k > 0) per_read[r] = Commons::purgePalindrome(*use, 4, purge_last_k);
            else per_read[r] = *use;
        }
    }
    auto t1 = high_resolution_clock::now();
    vector<uint64_t> offs(n_reads + 1, 0);
    for (size_t r = 0; r < n_reads; r++) offs[r + 1] = offs[r] + per_read[r].size();
    vector<uint32_t> flat(offs[n_reads] + 1);
    for (size_t r = 0; r < n_reads; r++)
        if (!per_read[r].empty()) memcpy(flat.data() + offs[r], per_read[r].data(), per_read[r].size() * 4);
    uint32_t* v; uint64_t* h; uint32_t* a;
    uint64_t ni, nd;
    size_t n = ref_count(flat.data(), offs.data(), n_reads, k, min_abundance, n_threads, &v, &h, &a, &ni, &nd);
    auto t2 = high_resolution_clock::now();
    uint64_t cs = 0;
    for (size_t i = 0; i < n; i++) cs += (uint64_t)a[i] * h[2 * i + 1];
    free(v); free(h); free(a);
    if (n_minimizers_sketch) *n_minimizers_sketch = n_sketch;
    if (n_minimizers) *n_minimizers = offs[n_reads];
    if (checksum) *checksum = cs;
    if (seconds_sketch) *seconds_sketch = duration<double>(t1 - t0).count();
    if (seconds_count) *seconds_count = duration<double>(t2 - t1).count();
    return n;
}

// ---- the reference's own graph-stage classes, driven through their file contract -----------------

static void write_read_data(const string& filename, const uint32_t* mins, const uint64_t* offs, size_t n_reads) {
    // record format read by KminmerParserParallel (src/Commons.hpp:7405-7440): u32 n, u8 isCircular, u32[n]
    ofstream f(filename, std::ios::binary);
    for (size_t r = 0; r < n_reads; r++) {
        u_int32_t size = (u_int32_t)(offs[r + 1] - offs[r]);
        u_int8_t isCircular = 0;
        f.write((const char*)&size, sizeof(size));
        f.write((const char*)&isCircular, sizeof(isCircular));
        f.write((const char*)(mins + offs[r]), size * sizeof(u_int32_t));
    }
}

static size_t read_tables(const string& dir, int k, bool with_vecs, uint32_t** vecs_out, uint64_t** hashes_out,
                          uint32_t** abundances_out) {
    vector<uint64_t> hashes;
    vector<uint32_t> abs_;
    ifstream fa(dir + "/kminmerData_abundance.txt", std::ios::binary);
    while (true) {
        u_int128_t vecHash;
        AbundanceType abundance;
        bool iseof = MDBG::readKminmerAbundance(vecHash, abundance, fa);
        if (iseof) break;
        hashes.push_back((uint64_t)(vecHash >> 64));
        hashes.push_back((uint64_t)vecHash);
        abs_.push_back(abundance);
    }
    size_t n = abs_.size();
    *hashes_out = (uint64_t*)malloc((hashes.size() + 1) * 8);
    *abundances_out = (uint32_t*)malloc((n + 1) * 4);
    memcpy(*hashes_out, hashes.data(), hashes.size() * 8);
    memcpy(*abundances_out, abs_.data(), n * 4);
    *vecs_out = (uint32_t*)malloc((n * (size_t)k + 1) * 4);
    if (with_vecs) {
        ifstream fk(dir + "/kminmerData_min.txt", std::ios::binary);
        fk.read((char*)*vecs_out, n * (size_t)k * 4);
    }
    return n;
}

/* CreateMdbg::createMDBG first pass (src/graph/CreateMdbg.cpp:284-326): KminmerCounter::execute
 * (disk partitions, sortParallel, dump) then rescueKminmers when min_abundance <= 1.
 * Entries come back in file order: the n_solid counted entries first, rescued ones after. */
size_t ref_graph_firstpass(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k, uint32_t min_abundance,
                           int n_threads, const char* tmp_dir, uint32_t** vecs_out, uint64_t** hashes_out,
                           uint32_t** abundances_out, uint64_t* n_solid, uint64_t* n_rescued, double* seconds) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    write_read_data(dir + "/read_data_corrected.txt", mins, offs, n_reads);
    auto t0 = high_resolution_clock::now();
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;                       // CreateMdbg.cpp:223-226: max(nbBases/20e9, nbCores, 1)
    c._isFirstPass = true;
    c._minAbundance = min_abundance;
    c._kminmerFile = ofstream(dir + "/kminmerData_min.txt");
    c._kminmerAbundanceFile = ofstream(dir + "/kminmerData_abundance.txt");
    {
        CreateMdbg::KminmerCounter kminmerCounter(c);
        kminmerCounter.execute();
        if (n_solid) *n_solid = kminmerCounter._nbSolidKminmers;
    }
    c._kminmerFile.close();
    c._kminmerAbundanceFile.close();
    c._nbRescuedKminmers = 0;
    if (min_abundance <= 1) {                          // CreateMdbg.cpp:309-319
        c._kminmerFile.open(dir + "/kminmerData_min.txt", std::ios_base::app);
        c._kminmerAbundanceFile.open(dir + "/kminmerData_abundance.txt", std::ios_base::app);
        c.rescueKminmers();
        c._kminmerFile.close();
        c._kminmerAbundanceFile.close();
    }
    if (n_rescued) *n_rescued = c._nbRescuedKminmers;
    if (seconds) *seconds = duration<double>(high_resolution_clock::now() - t0).count();
    return read_tables(dir, k, true, vecs_out, hashes_out, abundances_out);
}

/* createMDBG for k > firstK (src/graph/CreateMdbg.cpp:386-468) with _kminmerAbundances given:
 * use_counter != 0 -> the KminmerCounter + getRefinedAbundance path (k == firstK+1),
 * else             -> the IndexKminmerFunctor path (k >= firstK+2).  unitig_data.txt is empty. */
size_t ref_graph_next_k(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k, const uint64_t* prev_hashes,
                        const uint32_t* prev_ab, size_t n_prev, int use_counter, int n_threads, const char* tmp_dir,
                        uint32_t** vecs_out, uint64_t** hashes_out, uint32_t** abundances_out) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    write_read_data(dir + "/read_data_corrected.txt", mins, offs, n_reads);
    { ofstream empty(dir + "/unitig_data.txt", std::ios::binary); }
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._kminmerSizePrev = k - 1;
    c._kminmerSizeFirst = 4;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;
    c._isFirstPass = false;
    c._minAbundance = 0;
    for (size_t i = 0; i < n_prev; i++) {
        u_int128_t h = ((u_int128_t)prev_hashes[2 * i] << 64) | prev_hashes[2 * i + 1];
        c._kminmerAbundances[h] = prev_ab[i];
    }
    if (use_counter) {
        c._kminmerFile.open(dir + "/kminmerData_min.txt");
        c._kminmerAbundanceFile.open(dir + "/kminmerData_abundance.txt");
        CreateMdbg::KminmerCounter kminmerCounter(c);
        kminmerCounter.execute();
        c._kminmerFile.close();
        c._kminmerAbundanceFile.close();
        return read_tables(dir, k, true, vecs_out, hashes_out, abundances_out);
    }
    KminmerParserParallel parser2(dir + "/read_data_corrected.txt", 0, k, false, false, n_threads);
    parser2.parse(CreateMdbg::IndexKminmerFunctor(c, false));
    c._kminmerAbundanceFile.open(dir + "/kminmerData_abundance.txt");
    for (const auto& it : c._mdbgNodesLight) {          // CreateMdbg.cpp:453-464
        u_int128_t vecHash = it.first;
        u_int32_t abundance = it.second;
        MDBG::writeKminmerAbundance(vecHash, abundance, c._kminmerAbundanceFile);
    }
    c._kminmerAbundanceFile.close();
    return read_tables(dir, k, false, vecs_out, hashes_out, abundances_out);
}

/* CreateMdbg::indexEdges' first step (src/graph/CreateMdbg.cpp:1177-1187): EdgeIndexer::execute on the node file
 * kminmerData_min.txt -- disk partitions by hash128 % P, sortParallel, dereplication into edges.bin.
 * Returns the keys of edges.bin in file order ({high, low} words), _nbEdges and _checksum. */
size_t ref_edge_index(const uint32_t* vecs, size_t n, int k, int n_threads, const char* tmp_dir, uint64_t** hashes_out,
                      uint64_t* nb_edges, uint64_t* checksum) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    {
        ofstream f(dir + "/kminmerData_min.txt", std::ios::binary);
        f.write((const char*)vecs, (std::streamsize)(n * (size_t)k * sizeof(uint32_t)));
    }
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;
    CreateMdbg::EdgeIndexer edgeIndexer(c);
    edgeIndexer.execute();
    if (nb_edges) *nb_edges = edgeIndexer._nbEdges;
    if (checksum) *checksum = edgeIndexer._checksum;
    vector<uint64_t> keys;
    ifstream fe(edgeIndexer.getOutputFilename(), std::ios::binary);
    while (true) {
        u_int128_t e;
        fe.read((char*)&e, sizeof e);
        if (fe.eof()) break;
        keys.push_back((uint64_t)(e >> 64));
        keys.push_back((uint64_t)e);
    }
    *hashes_out = (uint64_t*)malloc((keys.size() + 2) * 8);
    memcpy(*hashes_out, keys.data(), keys.size() * 8);
    return keys.size() / 2;
}

/* CreateMdbg::indexEdges (src/graph/CreateMdbg.cpp:1177-1275): EdgeIndexer, the BooPHF over edges.bin, then indexEdge
 * (:1277-1420) for every node of kminmerData_min.txt, with n_threads OpenMP threads.  Returns, per distinct edge key
 * (in first-seen order of the node file), the key {high, low}, whether the key's (k-1)-min-mer is a palindrome, and the
 * raw KminmerEdge33 slots as the reference left them: minimizer (0xFFFFFFFF = empty) and flags bit0 isReversed, bit1
 * isPrefix, bit2 hasMultipleSuccessors.  Which node is recorded in a slot depends on the thread arrival order; the
 * order-free content is derived from this by oracle/pyoracle.py::canonical_edge_values. */
size_t ref_edge_values(const uint32_t* vecs, size_t n, int k, int n_threads, const char* tmp_dir, uint64_t** keys_out,
                       uint8_t** palindrome_out, uint32_t** minimizers_out, uint8_t** flags_out) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    {
        ofstream f(dir + "/kminmerData_min.txt", std::ios::binary);
        f.write((const char*)vecs, (std::streamsize)(n * (size_t)k * sizeof(uint32_t)));
    }
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;
    c._mutexes.resize(1000);
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_init_lock(&c._mutexes[i]);
    c.indexEdges();
    vector<uint64_t> keys;
    vector<uint8_t> pal, flags;
    vector<uint32_t> mins;
    unordered_set<string> seen;
    for (size_t i = 0; i < n; i++) {
        KmerVec vec;
        vec._kmers.assign(vecs + i * (size_t)k, vecs + (i + 1) * (size_t)k);
        for (int side = 0; side < 2; side++) {
            bool rev;
            const KmerVec e = side ? vec.suffix().normalize(rev) : vec.prefix().normalize(rev);
            const u_int128_t h = e.hash128();
            const string hs((const char*)&h, sizeof h);
            if (!seen.insert(hs).second) continue;
            const KminmerEdge33& v = c._mdbgEdges10._values[c._mdbgEdges10._keys->lookup(h)];
            keys.push_back((uint64_t)(h >> 64));
            keys.push_back((uint64_t)h);
            pal.push_back(e.isPalindrome() ? 1 : 0);
            mins.push_back(v._minimizer1);
            flags.push_back(v._minimizer1 == (MinimizerType)-1 ? 0 : (uint8_t)((v._isReversed1 ? 1 : 0) | (v._isPrefix1 ? 2 : 0) | (v._hasMultipleSuccessors1 ? 4 : 0)));
            mins.push_back(v._minimizer2);
            flags.push_back(v._minimizer2 == (MinimizerType)-1 ? 0 : (uint8_t)((v._isReversed2 ? 1 : 0) | (v._isPrefix2 ? 2 : 0) | (v._hasMultipleSuccessors2 ? 4 : 0)));
        }
    }
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_destroy_lock(&c._mutexes[i]);
    c._mdbgEdges10.clear();
    const size_t ne = pal.size();
    *keys_out = (uint64_t*)malloc((2 * ne + 2) * 8);
    memcpy(*keys_out, keys.data(), 2 * ne * 8);
    *palindrome_out = (uint8_t*)malloc(ne + 1);
    memcpy(*palindrome_out, pal.data(), ne);
    *minimizers_out = (uint32_t*)malloc((2 * ne + 2) * 4);
    memcpy(*minimizers_out, mins.data(), 2 * ne * 4);
    *flags_out = (uint8_t*)malloc(2 * ne + 2);
    memcpy(*flags_out, flags.data(), 2 * ne);
    return ne;
}

/* CreateMdbg::indexEdges + computeUnitigNodes (src/graph/CreateMdbg.cpp:1177-1275, 1521-1598; walker
 * ComputeUnitigFunctor::computeUnitigNode2, CreateMdbg.hpp:2513-2916) and, when `deterministic`, computeDeterministicUnitigs
 * (CreateMdbg.cpp:1001-1043), on the node file kminmerData_min.txt with n_threads OpenMP threads -- the same call
 * sequence as createGfa (CreateMdbg.cpp:876-912).  Returns the records of unitigGraph.nodes.bin in file order as a CSR;
 * the unitigIndex field of record i is checked to be 2 * i. */
size_t ref_unitig_nodes(const uint32_t* vecs, size_t n, int k, int n_threads, int deterministic, const char* tmp_dir,
                        uint64_t** offs_out, uint32_t** mins_out) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    {
        ofstream f(dir + "/kminmerData_min.txt", std::ios::binary);
        f.write((const char*)vecs, (std::streamsize)(n * (size_t)k * sizeof(uint32_t)));
    }
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;
    c._mutexes.resize(1000);
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_init_lock(&c._mutexes[i]);
    c.indexEdges();
    c._unitigGraphFile_nodes = ofstream(dir + "/unitigGraph.nodes.bin");
    c.computeUnitigNodes();
    c._unitigGraphFile_nodes.close();
    c._mdbgEdges10.clear();
    if (deterministic) c.computeDeterministicUnitigs();
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_destroy_lock(&c._mutexes[i]);
    vector<uint64_t> offs;
    vector<uint32_t> mins;
    ifstream nf(dir + "/unitigGraph.nodes.bin", std::ios::binary);
    size_t idx = 0;
    while (true) {
        u_int32_t size;
        nf.read((char*)&size, sizeof size);
        if (nf.eof()) break;
        offs.push_back(mins.size());
        const size_t at = mins.size();
        mins.resize(at + size);
        nf.read((char*)&mins[at], (std::streamsize)size * sizeof(MinimizerType));
        UnitigType unitigIndex;
        nf.read((char*)&unitigIndex, sizeof unitigIndex);
        if (deterministic && (size_t)unitigIndex != 2 * idx) { fprintf(stderr, "ref_unitig_nodes: unitigIndex %llu at record %zu\n", (unsigned long long)unitigIndex, idx); abort(); }
        idx++;
    }
    offs.push_back(mins.size());
    *offs_out = (uint64_t*)malloc((offs.size() + 1) * 8);
    memcpy(*offs_out, offs.data(), offs.size() * 8);
    *mins_out = (uint32_t*)malloc((mins.size() + 1) * 4);
    memcpy(*mins_out, mins.data(), mins.size() * 4);
    return offs.size() - 1;
}

/* CreateMdbg::indexUnitigEdges + computeUnitigEdges (src/graph/CreateMdbg.cpp:2915-2994, 3088-3140; UnitigEdgeIndexer,
 * BooPHF, indexUnitigEdge, computeUnitigEdge -> getSuccessors_unitig / getPredecessors_unitig -> dumpUnitigEdge) on a
 * unitigGraph.nodes.bin written from the given records (record i: unitigIndex 2 i), with n_threads OpenMP threads -- the call
 * sequence of createGfa (CreateMdbg.cpp:913-925).  Returns the records of unitigGraph.edges.successors.bin re-ordered by
 * unitigIndex as a CSR over oriented unitigs (list 2 i = successors, 2 i + 1 = predecessors of record i; inside a list
 * the file's order), *nb_edges = _nbUnitigEdges, *checksum = _checksum_unitigEdges. */
size_t ref_unitig_edges(const uint32_t* mins, const uint64_t* offs, size_t n_unitigs, int k, int n_threads, const char* tmp_dir,
                        uint64_t** eoff_out, uint32_t** etgt_out, uint64_t* nb_edges, uint64_t* checksum) {
    if (n_threads < 1) n_threads = 1;
    const string dir(tmp_dir);
    {
        ofstream f(dir + "/unitigGraph.nodes.bin", std::ios::binary);
        for (size_t i = 0; i < n_unitigs; i++) {
            const u_int32_t size = (u_int32_t)(offs[i + 1] - offs[i]);
            const UnitigType index = (UnitigType)(2 * i);
            f.write((const char*)&size, sizeof size);
            f.write((const char*)(mins + offs[i]), (std::streamsize)size * sizeof(MinimizerType));
            f.write((const char*)&index, sizeof index);
        }
    }
    CreateMdbg c;
    c._outputDir = dir;
    c._kminmerSize = k;
    c._nbCores = n_threads;
    c._nbPartitions = n_threads;
    c._mutexes.resize(1000);
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_init_lock(&c._mutexes[i]);
    c._checksum_unitigEdges = 0;
    c.indexUnitigEdges();
    c._unitigGraphFile_edges_successors = ofstream(dir + "/unitigGraph.edges.successors.bin");
    c.computeUnitigEdges();
    c._unitigGraphFile_edges_successors.close();
    for (size_t i = 0; i < c._mutexes.size(); i++) omp_destroy_lock(&c._mutexes[i]);
    vector<vector<uint32_t>> lists(2 * n_unitigs);
    ifstream ef(dir + "/unitigGraph.edges.successors.bin", std::ios::binary);
    while (true) {
        UnitigType from;
        ef.read((char*)&from, sizeof from);
        if (ef.eof()) break;
        for (int o = 0; o < 2; o++) {
            u_int32_t nb;
            ef.read((char*)&nb, sizeof nb);
            vector<uint32_t>& l = lists[(size_t)from + (size_t)o];
            l.resize(nb);
            if (nb) ef.read((char*)l.data(), (std::streamsize)nb * sizeof(UnitigType));
        }
    }
    size_t tot = 0;
    for (auto& l : lists) tot += l.size();
    *eoff_out = (uint64_t*)malloc((2 * n_unitigs + 2) * 8);
    *etgt_out = (uint32_t*)malloc((tot + 1) * 4);
    size_t at = 0;
    for (size_t x = 0; x < 2 * n_unitigs; x++) {
        (*eoff_out)[x] = at;
        memcpy(*etgt_out + at, lists[x].data(), lists[x].size() * 4);
        at += lists[x].size();
    }
    (*eoff_out)[2 * n_unitigs] = at;
    if (nb_edges) *nb_edges = c._nbUnitigEdges;
    if (checksum) *checksum = c._checksum_unitigEdges;
    return tot;
}

/* The reference's whole readSelection stage (ReadSelection::execute, src/readSelection/ReadSelection.hpp:92-303):
 * kseq FASTA/FASTQ parsing, HPC, sketch, complexity / quality side outputs, ordered record writer, read stats,
 * purgePalindromes.  `input_list` is the text file listing the read files (what `metaMDBG asm` writes as input.txt).
 * Writes <dir>/read_data_init.txt, read_stats.txt, repetitiveMinimizers.bin, read_data_corrected.txt. */
int ref_read_selection(const char* input_list, const char* dir, int l, float density, int hpc, int n_threads,
                       int skip_correction, double* seconds) {
    auto t0 = high_resolution_clock::now();
    ReadSelection rs;
    rs._inputFilename = input_list;
    rs._inputDir = dir;
    rs._outputFilename = string(dir) + "/read_data_init.txt";
    rs._nbCores = n_threads < 1 ? 1 : n_threads;
    rs._minReadQuality = 0;                              // default of --min-read-quality (ReadSelection.hpp:163)
    rs._outputQuality = true;                            // ReadSelection.hpp:196
    rs._skipCorrection = skip_correction != 0;
    rs._params._minimizerSize = l;
    rs._params._minimizerDensity_assembly = density;
    rs._params._minimizerDensity_correction = 0.025f;
    rs._params._useHomopolymerCompression = hpc != 0;
    rs._params._kminmerSize = 4;
    rs._params._kminmerSizeFirst = 4;
    rs._params._kminmerSizePrev = 3;
    rs.execute();
    if (seconds) *seconds = duration<double>(high_resolution_clock::now() - t0).count();
    return 0;
}

// Utils::computeN50 / computeMeanLength (Commons.hpp:2291-2336) and Commons::computeLastK (Commons.hpp:1726-1741):
// the scalars ReadSelection derives from the read lengths (read_stats.txt, purgePalindromes' lastK)
uint64_t ref_compute_n50(const uint32_t* lengths, size_t n) {
    return Utils::computeN50(std::vector<u_int32_t>(lengths, lengths + n));
}
uint64_t ref_compute_mean_length(const uint32_t* lengths, size_t n) {
    return Utils::computeMeanLength(std::vector<u_int32_t>(lengths, lengths + n));
}
int ref_compute_last_k(float density, size_t n50, size_t first_k, size_t max_k) {
    return Commons::computeLastK(density, n50, first_k, max_k);
}

int ref_max_threads() { return omp_get_max_threads(); }

void ref_free(void* p) { free(p); }

}  // extern "C"
