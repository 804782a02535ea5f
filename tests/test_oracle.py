"""CPU tests: the C restatement (oracle/mdbg_oracle.c) against the known-answer
values of SURVEY.md section 8c and against the reference's own sources compiled
into oracle/_ref (tests marked `ref` skip where that library is absent)."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from metamdbg_b200 import synth


# ---- known-answer tests (reproduced with the compiled reference) ----

def test_kat_murmur_invalid_lmer(oracle):
    assert oracle.murmur_h1(struct.pack("<Q", 0xFFFFFFFFFFFFFFFF), 42) == 13161889351522953593


def test_kat_hash128_1234(oracle):
    h1, h2 = oracle.hash128(np.array([1, 2, 3, 4], dtype=np.uint32))
    assert (h1, h2) == (0x333C24B23227B2C3, 0x869BB9D3950B32E1)


def test_kat_bound(oracle):
    assert oracle.bound(0.005) == pytest.approx(9.223371830696346e16, rel=0, abs=16)
    assert oracle.bound(0.005) == float(np.float32(0.005)) * 2.0 ** 64


def test_threshold_is_exact_boundary(oracle):
    for d in (0.005, 0.025, 0.0005, 0.5):
        t, none = oracle.threshold(d)
        assert not none
        b = oracle.bound(d)
        assert float(np.float64(np.uint64(t))) < b
        assert not (float(np.float64(np.uint64(t + 1))) < b)


def test_hpc_edge_cases(oracle):
    assert oracle.hpc(b"")[0] == b"#"                   # upstream quirk: empty read -> "#"
    assert oracle.hpc(b"AAAA")[0] == b"A"
    assert oracle.hpc(b"AACCCGTT")[0] == b"ACGT"
    assert list(oracle.hpc(b"AACCCGTT")[1]) == [0, 2, 5, 6, 8]
    assert oracle.hpc(b"aA")[0] == b"aA"                # case-sensitive byte compare
    assert oracle.hpc(b"ANNNC")[0] == b"ANC"
    assert oracle.hpc(b"A#C")[0] == b"AC"               # '#' is the sentinel and vanishes
    assert oracle.hpc(b"AC#")[0] == b"AC#"              # ... except a run that ends the read (unconditional last push)
    assert oracle.hpc(b"AC###")[0] == b"AC#" and oracle.hpc(b"###")[0] == b"#" and oracle.hpc(b"#AC")[0] == b"AC"
    assert oracle.hpc(b"AACC", hpc=False)[0] == b"AACC"


def test_lmers_small(oracle):
    # ACGTA, l=3: code A0 C1 T2 G3; ACG fwd=0b000111=7, rc(CGT)=0b011110=30 -> 7 dir0
    v, d = oracle.lmers(b"ACGTA", 3)
    assert len(v) == 3
    assert v[0] == 7 and d[0] == 0
    v, d = oracle.lmers(b"ACNTA", 3)
    assert all(x == np.uint64(0xFFFFFFFFFFFFFFFF) for x in v)
    assert len(oracle.lmers(b"AC", 3)[0]) == 0


def test_sketch_trims_first_and_last(oracle):
    rng = np.random.default_rng(5)
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 4000))
    m, p, d = oracle.sketch_read(seq, 15, 0.9, False)     # d=0.9 selects ~everything valid
    n_lmers = len(seq) - 15 + 1
    assert p.min() >= 1 and p.max() <= n_lmers - 2
    assert np.all(np.diff(p.astype(np.int64)) > 0)


def test_kminmers_normalize_tie_and_order(oracle):
    v, r = oracle.kminmers(np.array([5, 1, 1, 5, 9], dtype=np.uint32), 4)
    assert list(v[0]) == [5, 1, 1, 5] and r[0] == 1      # palindromic window -> reversed flag
    assert list(v[1]) == [1, 1, 5, 9] and r[1] == 0
    v, r = oracle.kminmers(np.array([9, 2, 3, 4], dtype=np.uint32), 4)
    assert list(v[0]) == [4, 3, 2, 9] and r[0] == 1
    assert len(oracle.kminmers(np.array([1, 2, 3], dtype=np.uint32), 4)[0]) == 0


def test_purge_palindrome_basic(oracle):
    m = np.array([7, 5, 1, 1, 5, 9, 3], dtype=np.uint32)
    out, keep = oracle.purge_palindrome(m, 4, 6)
    # window [5,1,1,5] is a palindrome -> its first element (index 1) is banned
    assert keep[1] == 0
    nothing, keep2 = oracle.purge_palindrome(np.arange(20, dtype=np.uint32), 4, 10)
    assert np.all(keep2 == 1) and len(nothing) == 20


def test_count_matches_python_dict(oracle):
    rng = np.random.default_rng(0)
    reads = [rng.integers(0, 6, size=rng.integers(0, 30)).astype(np.uint32) for _ in range(200)]
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    mins = np.concatenate(reads).astype(np.uint32)
    for k in (4, 5, 7):
        res = oracle.count(mins, offs, k, 2)
        ref = {}
        for r in reads:
            for i in range(len(r) - k + 1):
                w = tuple(int(x) for x in r[i:i + k])
                wr = w[::-1]
                key = w if w < wr else wr
                ref[key] = ref.get(key, 0) + 1
        solid = {kk: c for kk, c in ref.items() if c >= 2}
        got = {tuple(int(x) for x in v): int(a) for v, a in zip(res["vecs"], res["abundances"])}
        assert got == solid
        assert res["n_distinct"] == len(ref)
        assert res["n_instances"] == sum(ref.values())


# ---- restatement vs the reference's own code (oracle/_ref) ----

def _random_reads(seed, n, mean, with_n=False, lower=False):
    rng = np.random.default_rng(seed)
    alpha = b"ACGT" + (b"N" if with_n else b"") + (b"acgt" if lower else b"")
    out = []
    for _ in range(n):
        ln = int(rng.integers(0, mean * 2))
        s = rng.choice(np.frombuffer(alpha, dtype=np.uint8), ln)
        # homopolymer runs so HPC has something to do
        rep = rng.integers(1, 4, size=ln)
        out.append(bytes(np.repeat(s, rep)[:ln]))
    return out


@pytest.mark.ref
def test_ref_murmur_and_hash(oracle, reference):
    rng = np.random.default_rng(1)
    for ln in list(range(0, 90)):
        key = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        assert oracle.murmur128(key, 0) == reference.murmur128(key, 0)
        assert oracle.murmur_h1(key, 42) == reference.murmur_h1(key, 42)
    for k in range(2, 22):
        v = rng.integers(0, 2 ** 32, k, dtype=np.uint64).astype(np.uint32)
        assert oracle.hash128(v) == reference.hash128(v)
    for d in (0.005, 0.025, 0.1):
        assert oracle.bound(d) == reference.bound(d)


@pytest.mark.ref
@pytest.mark.parametrize("hpc", [True, False])
def test_ref_hpc_lmers_sketch(oracle, reference, hpc):
    for seq in _random_reads(2, 60, 600, with_n=True, lower=True) + [b"", b"A", b"ACGT" * 4, b"A" * 50]:
        a, pa = oracle.hpc(seq, hpc)
        b, pb = reference.hpc(seq, hpc)
        assert a == b
        assert np.array_equal(pa, pb)
        for l in (3, 11, 15, 16):
            va, da = oracle.lmers(a, l)
            vb, db = reference.lmers(a, l)
            assert np.array_equal(va, vb) and np.array_equal(da, db)
        for dens in (0.005, 0.05, 0.3):
            ra = oracle.sketch_read(seq, 15, dens, hpc)
            rb = reference.sketch_read(seq, 15, dens, hpc)
            for x, y in zip(ra, rb):
                assert np.array_equal(x, y)


@pytest.mark.ref
def test_ref_sketch_blacklist(oracle, reference):
    seq = _random_reads(3, 1, 4000)[0]
    m, _, _ = oracle.sketch_read(seq, 15, 0.05, True)
    assert len(m) > 4
    bl = m[::3].copy()
    ra = oracle.sketch_read(seq, 15, 0.05, True, bl)
    rb = reference.sketch_read(seq, 15, 0.05, True, bl)
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    assert len(ra[0]) < len(m)


@pytest.mark.ref
def test_ref_synthetic_reads_sketch(oracle, reference):
    rs = synth.make_readset(40, 6000, seed=11, n_genomes=2, genome_len_range=(50_000, 80_000))
    bases, offs = synth.fill_reads(rs)
    raw = bases.tobytes()
    tot = 0
    for r in range(rs.n_reads):
        seq = raw[int(offs[r]):int(offs[r + 1])]
        ra = oracle.sketch_read(seq, 15, 0.005, True)
        rb = reference.sketch_read(seq, 15, 0.005, True)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        tot += len(ra[0])
    assert tot > 100
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    assert int(mo[-1]) == tot


@pytest.mark.ref
def test_ref_kminmers_and_purge(oracle, reference):
    rng = np.random.default_rng(4)
    for trial in range(300):
        n = int(rng.integers(0, 40))
        m = rng.integers(0, 4, n).astype(np.uint32)      # tiny alphabet -> many palindromes
        for k in (2, 4, 5, 8):
            va, ra = oracle.kminmers(m, k)
            vb, rb = reference.kminmers(m, k)
            assert np.array_equal(va, vb) and np.array_equal(ra, rb)
        last_k = int(rng.integers(5, 12))
        pa, _ = oracle.purge_palindrome(m, 4, last_k)
        pb = reference.purge_palindrome(m, 4, last_k)
        assert np.array_equal(pa, pb)


@pytest.mark.ref
def test_ref_count(oracle, reference):
    rng = np.random.default_rng(6)
    reads = [rng.integers(0, 9, size=rng.integers(0, 60)).astype(np.uint32) for _ in range(400)]
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    mins = np.concatenate(reads).astype(np.uint32)
    for k in (4, 6, 21):
        for min_ab in (0, 2, 3):
            a = oracle.count(mins, offs, k, min_ab)
            b = reference.count(mins, offs, k, min_ab, threads=3)
            assert np.array_equal(a["vecs"], b["vecs"])
            assert np.array_equal(a["hashes"], b["hashes"])
            assert np.array_equal(a["abundances"], b["abundances"])
            assert a["n_instances"] == b["n_instances"] and a["n_distinct"] == b["n_distinct"]


@pytest.mark.ref
def test_ref_pipeline_checksum(oracle, reference):
    rs = synth.make_readset(300, 5000, seed=21, n_genomes=1, genome_len_range=(60_000, 60_001))
    bases, offs = synth.fill_reads(rs)
    res = reference.pipeline(bases, offs, 15, 0.005, True, 4, purge_last_k=0, min_abundance=2, threads=2)
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    c = oracle.count(m, mo, 4, 2)
    assert res["n_minimizers"] == len(m)
    assert res["n_solid"] == len(c["abundances"]) and res["n_solid"] > 10
    assert res["checksum"] == oracle.checksum(c["hashes"], c["abundances"])


# ---- rows A7 (real KminmerCounter with disk partitions), A7b rescue, A8/A9 next-k: restatement vs the
# ---- reference's own CreateMdbg classes driven through their file contract

def _table(hashes, abund):
    return {(int(h[0]), int(h[1])): int(a) for h, a in zip(hashes, abund)}


def _minspace_reads(seed, n_reads=600, alphabet=400, max_len=80):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, alphabet, size=5000).astype(np.uint32)       # a "genome" in minimizer space
    reads = []
    for _ in range(n_reads):
        ln = int(rng.integers(0, max_len))
        st = int(rng.integers(0, len(base) - max_len))
        r = base[st:st + ln].copy()
        if ln and rng.random() < 0.3:
            r[rng.integers(0, ln)] = rng.integers(0, alphabet)           # an "error"
        if rng.random() < 0.5:
            r = r[::-1].copy()
        reads.append(r)
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return np.concatenate(reads).astype(np.uint32), offs


@pytest.mark.ref
@pytest.mark.parametrize("k", [4, 7])
def test_ref_graph_firstpass_counter_and_rescue(oracle, reference, k):
    mins, offs = _minspace_reads(41)
    # --min-abundance 2: pure count, no rescue
    g = reference.graph_firstpass(mins, offs, k, min_abundance=2, threads=3)
    c = oracle.count(mins, offs, k, 2)
    assert g["n_rescued"] == 0 and g["n_solid"] == len(c["abundances"]) > 50
    assert _table(g["hashes"], g["abundances"]) == _table(c["hashes"], c["abundances"])
    assert sorted(map(tuple, g["vecs"].tolist())) == sorted(map(tuple, c["vecs"].tolist()))
    # default mode (--min-abundance 0): solid entries, then the rescued abundance-1 entries
    g = reference.graph_firstpass(mins, offs, k, min_abundance=0, threads=2)
    ns = g["n_solid"]
    assert _table(g["hashes"][:ns], g["abundances"][:ns]) == _table(c["hashes"], c["abundances"])
    r = oracle.rescue(mins, offs, k, c["hashes"], c["abundances"])
    assert g["n_rescued"] == len(r["hashes"]) > 0
    assert np.all(g["abundances"][ns:] == 1)
    assert sorted(map(tuple, g["hashes"][ns:].tolist())) == sorted(map(tuple, r["hashes"].tolist()))
    assert sorted(map(tuple, g["vecs"][ns:].tolist())) == sorted(map(tuple, r["vecs"].tolist()))
    assert len(set(map(tuple, r["hashes"].tolist()))) == len(r["hashes"])       # rescued keys are unique


@pytest.mark.ref
def test_ref_graph_next_k_both_paths(oracle, reference):
    mins, offs = _minspace_reads(42)
    prev = oracle.count(mins, offs, 4, 2)
    rng = np.random.default_rng(0)
    prev_ab = prev["abundances"].copy()
    prev_ab[rng.random(len(prev_ab)) < 0.1] = 0           # "refined" abundances may be 0 (treated as <= 1)
    for k, use_counter in ((5, True), (5, False)):
        want = oracle.next_k(mins, offs, k, prev["hashes"], prev_ab)
        got = reference.graph_next_k(mins, offs, k, prev["hashes"], prev_ab, use_counter=use_counter, threads=2)
        assert _table(got["hashes"], got["abundances"]) == _table(want["hashes"], want["abundances"])
        assert len(want["abundances"]) > 50 and want["abundances"].min() >= 2
        if use_counter:
            assert sorted(map(tuple, got["vecs"].tolist())) == sorted(map(tuple, want["vecs"].tolist()))
    # chain: k=5 table feeds k=6 (IndexKminmerFunctor path)
    t5 = oracle.next_k(mins, offs, 5, prev["hashes"], prev["abundances"])
    want = oracle.next_k(mins, offs, 6, t5["hashes"], t5["abundances"])
    got = reference.graph_next_k(mins, offs, 6, t5["hashes"], t5["abundances"], use_counter=False, threads=3)
    assert _table(got["hashes"], got["abundances"]) == _table(want["hashes"], want["abundances"])
    assert len(want["abundances"]) > 20


# ---- rows A3b / A3c: side outputs and the record writer, against the reference's whole readSelection stage

def _write_fastq(path, reads, quals):
    with open(path, "wb") as f:
        for i, (s, q) in enumerate(zip(reads, quals)):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + q + b"\n")


def _aux_reads(seed):
    rng = np.random.default_rng(seed)
    rs = synth.make_readset(120, 5000, seed=seed, n_genomes=1, genome_len_range=(80_000, 80_001))
    bases, offs = synth.fill_reads(rs)
    raw = bases.tobytes()
    reads = [raw[int(offs[r]):int(offs[r + 1])] for r in range(rs.n_reads)]
    reads += [b"AC" * 1500, b"A" * 800 + b"ACGTTGCA" * 300, b"ACG" * 900, b"ACGT" * 10]   # low complexity / short
    quals = []
    for s in reads:
        q = rng.integers(2, 60, size=len(s)).astype(np.uint8) + 33
        q[rng.random(len(s)) < 0.01] = 33 + 93
        quals.append(q.tobytes())
    return reads, quals


@pytest.mark.ref
@pytest.mark.parametrize("hpc", [True, False])
def test_ref_read_selection_side_outputs(tmp_path, oracle, reference, hpc):
    reads, quals = _aux_reads(71)
    fq = tmp_path / "reads.fastq"
    _write_fastq(fq, reads, quals)
    dens = 0.005 if hpc else 0.025
    res = reference.read_selection([fq], 15, dens, hpc, threads=3, skip_correction=True, workdir=str(tmp_path))
    assert len(res["records"]) == len(reads)
    n_low = 0
    for r, (s, q) in enumerate(zip(reads, quals)):
        rec = res["records"][r]
        m, p, d = oracle.sketch_read(s, 15, dens, hpc, res["blacklist"])   # ONT: the stage's own blacklist
        mq, cx, mins_q = oracle.read_aux(s, q, 15, hpc, p)
        if cx > 5:                                            # low-complexity filter clears the record
            n_low += 1
            m, p, d, mins_q = m[:0], p[:0], d[:0], mins_q[:0]
        assert np.array_equal(rec["minimizers"], m) and np.array_equal(rec["positions"], p), r
        assert np.array_equal(rec["directions"], d) and np.array_equal(rec["qualities"], mins_q), r
        assert rec["read_length"] == len(s)
        assert np.float32(rec["mean_quality"]).tobytes() == np.float32(mq).tobytes(), r     # bit-exact float
    assert n_low >= 2 and (hpc or len(res["blacklist"]) >= 1)
    assert res["stats"]["n_reads"] == len(reads) and res["stats"]["n_bases"] == sum(len(s) for s in reads)
    # purged stream (thread-completion order upstream => compare as multisets)
    want = []
    for rec in res["records"]:
        q, _ = oracle.purge_palindrome(rec["minimizers"], 4, 50)
        want.append(tuple(int(x) for x in q))
    got = [tuple(int(x) for x in rec["minimizers"]) for rec in res["corrected"]]
    assert sorted(got) == sorted(want)


@pytest.mark.ref
def test_ref_apply_density(oracle, reference):
    rs = synth.make_readset(30, 6000, seed=13, n_genomes=1, genome_len_range=(60_000, 60_001), err=0.02)
    bases, offs = synth.fill_reads(rs)
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.025, False)
    for dens in (0.005, 0.01, 0.025, 0.5):
        a, b = oracle.apply_density(m, dens), reference.apply_density(m, dens)
        assert np.array_equal(a, b)
    low = oracle.apply_density(m, 0.005)
    assert 0 < len(low) < len(m)
    # sketching at 0.025 then thresholding at 0.005 == sketching at 0.005 (same hash, same bound)
    assert np.array_equal(low, oracle.sketch_batch(bases, offs, 15, 0.005, False)[1])


@pytest.mark.ref
def test_ref_reads_ending_in_the_hpc_sentinel(oracle, reference):
    """A '#' run at the END of a read survives EncoderRLE (Commons.hpp:4186), which moves the last selectable
    position by one; the restatement and the reference's own MinimizerParser agree on it."""
    rng = np.random.default_rng(5)
    for i in range(300):
        body = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(0, 600))))
        s = body + [b"#", b"##", b"N#", b"#A", b"a#"][i % 5]
        assert oracle.hpc(s)[0] == reference.hpc(s)[0]
        for l, d in ((15, 0.3), (11, 0.3), (15, 0.005)):
            for x, y in zip(oracle.sketch_read(s, l, d, True), reference.sketch_read(s, l, d, True)):
                assert np.array_equal(x, y)


@pytest.mark.ref
def test_ref_edge_indexer(oracle, reference):
    """Row F1: the restatement of CreateMdbg::EdgeIndexer against the reference's own class (disk partitions,
    sortParallel, dereplication into edges.bin): same key set, _nbEdges and _checksum, for the node set of a real
    count table and for small-alphabet vectors (palindromic prefixes / suffixes, shared keys)."""
    rng = np.random.default_rng(3)
    for k in (2, 3, 4, 5, 21):
        vecs = rng.integers(0, 40, (2500, k)).astype(np.uint32)
        a = oracle.edge_index(vecs, k)
        b = reference.edge_index(vecs, k, threads=3)
        assert b["nb_edges"] == len(b["hashes"]) == len(a["hashes"])
        assert set(map(tuple, a["hashes"].tolist())) == set(map(tuple, b["hashes"].tolist()))
        assert a["checksum"] == b["checksum"]
    reads, offs = _minspace_reads(11)
    nodes = oracle.count(reads, offs, 4, 2)["vecs"]
    a = oracle.edge_index(nodes, 4); b = reference.edge_index(nodes, 4, threads=2)
    assert len(nodes) > 500 and set(map(tuple, a["hashes"].tolist())) == set(map(tuple, b["hashes"].tolist()))
    assert a["checksum"] == b["checksum"] and b["nb_edges"] == len(a["hashes"])


@pytest.mark.ref
def test_ref_edge_values_order_free_form(oracle, reference):
    """Row F1, second step: CreateMdbg::indexEdges run for real (EdgeIndexer, BooPHF, indexEdge over every node, 1 and
    4 OpenMP threads); its KminmerEdge33 slots, reduced to their order-free content (pyoracle.canonical_edge_values),
    equal the oracle's restatement -- two orientation classes per key with count 0 / 1 / 2+ and the single offer's
    minimizer and flags.  Small alphabets give palindromic keys and many branching ones."""
    from oracle.pyoracle import canonical_edge_values
    rng = np.random.default_rng(3)
    for k, alpha in ((4, 30), (5, 30), (3, 12), (2, 8), (7, 50), (4, 6), (21, 40)):
        vecs = rng.integers(0, alpha, (2500, k)).astype(np.uint32)
        nodes = np.unique(np.array([oracle.kminmers(v, k)[0][0] for v in vecs], dtype=np.uint32), axis=0)
        a = oracle.edge_values(nodes, k)
        want = {(int(h[0]), int(h[1])): tuple(tuple(int(x) for x in c) for c in v) for h, v in zip(a["hashes"], a["values"])}
        assert len(want) == len(oracle.edge_index(nodes, k)["hashes"])
        for threads in (1, 4):
            assert canonical_edge_values(reference.edge_values(nodes, k, threads=threads)) == want, (k, alpha, threads)
    reads, offs = _minspace_reads(11)
    nodes = oracle.count(reads, offs, 4, 2)["vecs"]
    a = oracle.edge_values(nodes, 4)
    want = {(int(h[0]), int(h[1])): tuple(tuple(int(x) for x in c) for c in v) for h, v in zip(a["hashes"], a["values"])}
    assert canonical_edge_values(reference.edge_values(nodes, 4, threads=3)) == want and len(want) > 500


@pytest.mark.ref
def test_ref_unitig_nodes(oracle, reference):
    """Row F1, third step: CreateMdbg::indexEdges + computeUnitigNodes + computeDeterministicUnitigs run for real (1 and 4
    OpenMP threads) against the oracle's restatement: the records of unitigGraph.nodes.bin are identical -- same unitigs,
    same normalized minimizer sequences, same order -- for clean paths, small alphabets (palindromic keys, hairpins,
    branching everywhere) and circular genomes."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import unitig_cases
    n_circular_like = 0
    for i, (k, mins, offs) in enumerate(unitig_cases(seed=5)):
        nodes = oracle.count(mins, offs, k, 2)["vecs"]
        a = oracle.unitigs(nodes, k)
        for threads in (1, 4):
            b = reference.unitig_nodes(nodes, k, threads=threads)
            assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["minimizers"], b["minimizers"]), (i, k, threads)
        assert len(a["offsets"]) - 1 > 0
        n_circular_like += int(i % 4 == 3)
        # unitig graph edges: indexUnitigEdges + computeUnitigEdges for real; identical file content with one thread,
        # identical lists up to the arrival order inside a list with four
        e = oracle.unitig_edges(a["offsets"], a["minimizers"], k)
        r1 = reference.unitig_edges(a["offsets"], a["minimizers"], k, threads=1)
        assert np.array_equal(e["offsets"], r1["offsets"]) and np.array_equal(e["targets"], r1["targets"]), (i, k)
        assert (e["n_edges"], e["checksum"]) == (r1["n_edges"], r1["checksum"])
        r4 = reference.unitig_edges(a["offsets"], a["minimizers"], k, threads=4)
        as_lists = lambda d: [sorted(d["targets"][int(d["offsets"][x]):int(d["offsets"][x + 1])].tolist()) for x in range(len(d["offsets"]) - 1)]
        assert as_lists(e) == as_lists(r4) and e["checksum"] == r4["checksum"]
    # many unitigs with many edges between them: unitig indices beyond 2^16, so that the reference's 32-bit products in
    # _checksum_unitigEdges wrap
    rng = np.random.default_rng(9)
    vecs = rng.integers(0, 40, (60000, 4)).astype(np.uint32)
    nodes = np.unique(np.array([oracle.kminmers(v, 4)[0][0] for v in vecs], dtype=np.uint32), axis=0)
    a = oracle.unitigs(nodes, 4); b = reference.unitig_nodes(nodes, 4, threads=4)
    assert len(a["offsets"]) - 1 > 40000 and np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["minimizers"], b["minimizers"])
    e = oracle.unitig_edges(a["offsets"], a["minimizers"], 4); r1 = reference.unitig_edges(a["offsets"], a["minimizers"], 4, threads=1)
    assert np.array_equal(e["offsets"], r1["offsets"]) and np.array_equal(e["targets"], r1["targets"]) and e["n_edges"] > 100000
    assert e["checksum"] == r1["checksum"]
    reads, offs = _minspace_reads(11)
    nodes = oracle.count(reads, offs, 4, 2)["vecs"]
    a = oracle.unitigs(nodes, 4); b = reference.unitig_nodes(nodes, 4, threads=3)
    assert len(nodes) > 500 and np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["minimizers"], b["minimizers"])
    assert n_circular_like >= 4
