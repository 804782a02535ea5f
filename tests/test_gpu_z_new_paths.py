"""GPU parity tests of the paths added AFTER the last GPU run of round 1 (they have run on the CPU emulator, not yet
on a B200): sketch-kernel variants + autotune, the piece pipeline of host batches, the multi-k sweep, edge keys and
values, reads ending in the HPC sentinel, the driver's smoke() and the C++ driver's --max-k / --edges.  Kept in their
own file, which sorts last, so that the suites validated on hardware run first."""
import os
import subprocess

import numpy as np
import pytest

from metamdbg_b200 import synth
from tests.test_gpu_parity import (EMULATED, assert_sketch_equal, built, device_array, engine, table_dict)  # noqa: F401

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("hpc,dens", [(True, 0.005), (False, 0.025)])
def test_sketch_kernel_variants_and_autotune(built, oracle, hpc, dens):
    """Every arithmetic variant of the sketch kernel's unrolled block gives the oracle's sketch, and
    mdbg_ctx_autotune_sketch only ever activates a variant whose complete output matched variant 0 on the device."""
    import warnings
    rs = synth.make_readset(1500, 8000, seed=314, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.01)
    bases, offs = synth.fill_reads(rs)
    bases = bases.copy()
    rng = np.random.default_rng(9)
    for r in rng.choice(rs.n_reads, 40, replace=False):                       # some blocks leave the fast path
        lo, hi = int(offs[r]), int(offs[r + 1])
        if hi - lo > 10:
            bases[lo + int(rng.integers(0, hi - lo))] = ord("N")
    want = oracle.sketch_batch(bases, offs, 15, dens, hpc)
    eng = engine(15, dens, hpc)
    assert eng.sketch_variant == int(os.environ.get("MDBG_SKETCH_VARIANT", "2"))
    p_b, keep_b = device_array(bases, pad=64)
    p_o, keep_o = device_array(offs.astype(np.uint64))
    res = eng.autotune_sketch(p_b, p_o, rs.n_reads, int(offs[-1]))
    assert res["identical"][0] and res["n_minimizers"] == len(want[1]) and res["n_reads"] == rs.n_reads
    assert res["identical"][res["chosen"]] and eng.sketch_variant == res["chosen"]
    assert eng.store_size() == (0, 0)                                         # autotune appends nothing
    assert_sketch_equal(eng.sketch_batch(bases, offs), *want, tag=f"autotuned variant {res['chosen']}")
    for v, same in enumerate(res["identical"]):
        if not same:                                   # never activated by autotune; say so instead of hiding it
            warnings.warn(f"sketch variant {v} differs from variant 0 on this device (times {res['ms']})")
            continue
        eng.set_sketch_variant(v)
        assert_sketch_equal(eng.sketch_batch(bases, offs), *want, tag=f"forced variant {v}")
    with pytest.raises(Exception):
        eng.set_sketch_variant(len(res["identical"]))
    eng.close()
    del keep_b, keep_o


def _selected_lmer(oracle, l, dens):
    """(bases of) one l-mer the density threshold selects, found by sketching random sequence with the oracle."""
    rng = np.random.default_rng(4)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 20000)]
    mo, m, p, d = oracle.sketch_batch(seq, np.array([0, len(seq)], np.uint64), l, dens, False)
    assert len(p) > 3
    return seq[int(p[1]):int(p[1]) + l].copy()


@pytest.mark.parametrize("packing", [0, 1])
def test_piece_pipeline_many_pieces(built, oracle, packing, monkeypatch):
    """Host batches cut into many pieces (MDBG_PIECE_BYTES): per-piece scan / compaction / D2H behind each piece's
    sketch gives the same CSR and store as the whole-batch tail -- plain reads, reads whose minimizer count exceeds
    the up-front estimate (buffers grow mid-batch), a slot overflow in a late piece (pipeline cancelled, exact
    re-sketch), empty reads, and the next batch appended behind it."""
    monkeypatch.setenv("MDBG_PIECE_BYTES", "60000")
    rs = synth.make_readset(700, 6000, seed=91, n_genomes=2, genome_len_range=(150_000, 250_000))
    bases, offs = synth.fill_reads(rs)
    want = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    eng = engine(15, 0.005, True)
    eng.set_host_packing(packing)
    sk = eng.sketch_batch(bases, offs, append_to_store=True)
    assert_sketch_equal(sk, *want, tag="many pieces")
    info = eng.last_batch_info()
    assert info["n_pieces"] >= 40 and info["n_pieces_pipelined"] == info["n_pieces"] and not info["overflow_fallback"]
    assert info["packed"] == bool(packing)
    # a second batch with empty and tiny reads between the pieces, appended to the same store
    rs2 = synth.make_readset(300, 5000, seed=92, n_genomes=1, genome_len_range=(150_000, 250_000))
    b2, o2 = synth.fill_reads(rs2)
    lens = np.diff(o2.astype(np.int64))
    lens[::7] = 0
    lens[3::11] = 9
    o2b = np.zeros(len(lens) + 1, np.uint64)
    o2b[1:] = np.cumsum(lens)
    b2b = np.concatenate([b2[int(o2[r]):int(o2[r]) + int(lens[r])] for r in range(len(lens))])
    want2 = oracle.sketch_batch(b2b, o2b, 15, 0.005, True)
    assert_sketch_equal(eng.sketch_batch(b2b, o2b, append_to_store=True), *want2, tag="ragged second batch")
    so, sm = eng.store_fetch()
    assert np.array_equal(sm, np.concatenate([want[1], want2[1]]))
    assert np.array_equal(so, np.concatenate([want[0], want2[0][1:] + want[0][-1]]))
    # same batch without fetching the CSR (store only), then fetched afterwards
    eng.store_clear()
    assert eng.sketch_batch(bases, offs, append_to_store=True, fetch=False) is None
    assert_sketch_equal(eng.sketch_fetch(), *want, tag="deferred fetch")
    assert np.array_equal(eng.store_fetch()[1], want[1])
    eng.close()

    # minimizer-rich reads: 3-4 x the nominal density (tandem repeats of a selected l-mer), HPC off -> the
    # estimate-sized buffers must grow while copies are in flight; one read far above its slot -> overflow fallback
    lm = _selected_lmer(oracle, 15, 0.005)
    unit = np.concatenate([lm, np.frombuffer(b"ACGTA", np.uint8)])
    rng = np.random.default_rng(12)
    reads = []
    for r in range(400):
        rnd = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 4200)]
        reads.append(np.concatenate([rnd[:2100], np.tile(unit, 90), rnd[2100:]]))
    for with_overflow in (False, True):
        rr = list(reads)
        if with_overflow:
            rr[377] = np.tile(unit, 300)
        bb = np.concatenate(rr)
        oo = np.zeros(len(rr) + 1, np.uint64)
        oo[1:] = np.cumsum([len(x) for x in rr])
        want3 = oracle.sketch_batch(bb, oo, 15, 0.005, False)
        assert len(want3[1]) > 2.5 * 0.005 * len(bb)
        eng = engine(15, 0.005, False)
        eng.set_host_packing(packing)
        assert_sketch_equal(eng.sketch_batch(bb, oo, append_to_store=True), *want3, tag=f"rich reads overflow={with_overflow}")
        info = eng.last_batch_info()
        assert info["n_pieces"] >= 30 and info["overflow_fallback"] == with_overflow
        if not with_overflow:
            assert info["n_buffer_growths"] >= 1 and info["n_pieces_pipelined"] == info["n_pieces"]
        else:
            assert 0 < info["n_pieces_pipelined"] < info["n_pieces"]
        assert np.array_equal(eng.store_fetch()[1], want3[1])
        eng.close()


def test_multi_k_sweep_on_device_vs_oracle(built, oracle):
    """multi_k_sweep: k = 4 counted (+ rescue), k = 5..9 each derived from the previous table, all on the device-resident
    store; every k's table equals the oracle's chain (count -> rescue -> next_k -> next_k ...)."""
    from metamdbg_b200 import multi_k_sweep
    rs = synth.make_readset(1200, 7000, seed=23, n_genomes=2, genome_len_range=(120_000, 200_000))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    for rescue in (False, True):
        tables = {}
        res = multi_k_sweep(eng, 4, 9, min_abundance=0 if rescue else 2, rescue=rescue,
                            on_table=lambda k, e: tables.__setitem__(k, e.count_finalize(0 if rescue else 2)))
        assert [r["k"] for r in res] == list(range(4, 10))
        prev = oracle.count(sm, so, 4, 2)
        ph, pa = prev["hashes"], prev["abundances"]
        if rescue:
            rr = oracle.rescue(sm, so, 4, ph, pa)
            assert res[0]["n_reads_rescued"] == rr["n_reads_rescued"]
            ph = np.concatenate([ph, rr["hashes"]]) if len(rr["hashes"]) else ph
            pa = np.concatenate([pa, np.ones(len(rr["hashes"]), np.uint32)])
        assert tables[4].as_dict() == table_dict(ph, pa)
        for k in range(5, 10):
            nk = oracle.next_k(sm, so, k, ph, pa)
            assert tables[k].as_dict() == table_dict(nk["hashes"], nk["abundances"]), f"k={k} rescue={rescue}"
            assert res[k - 4]["n_entries"] == len(nk["abundances"]) > 500
            ph, pa = nk["hashes"], nk["abundances"]
    # right-sized next-k tables: a table sized far too small fills up, the k is redone at the worst-case size
    # (the pass is idempotent), and the tables do not change
    want_sums = [r["checksum"] for r in multi_k_sweep(eng, 4, 7, table_headroom=0)]
    assert [r["checksum"] for r in multi_k_sweep(eng, 4, 7, table_headroom=0.001)] == want_sums
    assert [r["checksum"] for r in multi_k_sweep(eng, 4, 7)] == want_sums
    eng.close()


def test_sketch_reads_ending_in_the_hpc_sentinel(built, oracle):
    """EncoderRLE (Commons.hpp:4172-4190) swallows every run of '#' except one that ends the read -- its final
    `rleSequence += lastChar` is unconditional -- so such a read is one HPC base longer and its last selectable
    position moves by one (found by scripts/fuzz_capi_emulated.py).  Dense selection makes that position count."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for i in range(300):
        body = rng.choice(acgt, int(rng.integers(0, 1800)))
        if i % 5 == 0 and len(body) > 10:
            body[rng.integers(0, len(body), 3)] = ord("#")              # interior sentinels too
        tail = np.frombuffer((b"#", b"##", b"", b"N#", b"#A")[i % 5], np.uint8)
        reads.append(np.concatenate([body, tail]).astype(np.uint8))
    reads += [np.frombuffer(x, np.uint8) for x in (b"#", b"###", b"ACGTACGTTGCATGCA#")]
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    bases = np.concatenate(reads)
    for l, dens in ((15, 0.3), (15, 0.005), (11, 0.3)):
        for packing in (0, 1):                                          # dirty reads of a packed batch stay ASCII
            eng = engine(l, dens, True)
            eng.set_host_packing(packing)
            assert_sketch_equal(eng.sketch_batch(bases, offs), *oracle.sketch_batch(bases, offs, l, dens, True),
                                tag=f"l={l} d={dens} packing={packing}")
            eng.close()
        eng = engine(l, dens, False)                                    # without HPC '#' is an ordinary character
        assert_sketch_equal(eng.sketch_batch(bases, offs), *oracle.sketch_batch(bases, offs, l, dens, False))
        eng.close()


def test_edge_index_vs_oracle(built, oracle):
    """Row F1 (CreateMdbg::EdgeIndexer): the dereplicated prefix / suffix keys of the node set, for the first-pass
    table, the default-mode table (rescued nodes included) and a next-k table; set, count and checksum as the
    oracle's restatement (pinned against the reference's own EdgeIndexer in tests/test_oracle.py)."""
    rs = synth.make_readset(1500, 7000, seed=29, n_genomes=2, genome_len_range=(120_000, 200_000), err=0.003)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)

    def check(k, min_ab):
        tab = eng.count_finalize(min_ab)
        got = eng.edges_index(min_ab)
        want = oracle.edge_index(tab.kminmers, k)
        assert got["n_nodes"] == len(tab.abundances) and got["n_edges"] == len(want["hashes"]) > 500
        assert got["checksum"] == want["checksum"]
        assert {(int(h[1]), int(h[0])) for h in got["hashes"]} == {(int(h[0]), int(h[1])) for h in want["hashes"]}
        assert len(got["hashes"]) == len({(int(h[0]), int(h[1])) for h in got["hashes"]})      # no duplicate key
        # edge values (indexEdge / successorExists, order-free): two orientation classes per key
        wv = oracle.edge_values(tab.kminmers, k)
        want_v = {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])}
        got_v = {(int(h[1]), int(h[0])): v.tolist() for h, v in zip(got["hashes"], got["values"])}
        assert got_v == want_v

    eng.count_begin(4)
    eng.count_add_store()
    check(4, 2)
    eng.count_rescue()
    check(4, 0)
    eng.prev_from_current(0)
    eng.count_begin(5)
    eng.count_add_store_next_k()
    check(5, 0)
    for k in (2, 3, 9):
        eng.count_begin(k)
        eng.count_add_store()
        check(k, 2)
    eng.close()


def test_graft_entry_smoke(built, capsys):
    """__graft_entry__.smoke(): the driver's one small invocation of the hot path, checked against the oracle."""
    import __graft_entry__ as g
    g.smoke()
    out = capsys.readouterr().out
    assert "smoke OK" in out                     # (the line also reports whether variant 1 matched: informational)


def test_cpp_driver_max_k_and_edges(tmp_path, oracle):
    """The C++ driver's k > firstK passes (GpuNextKCounter, --max-k) and edge keys (GpuEdgeIndexer, --edges) on a
    read_data_corrected.txt written by its own first stage, default mode (rescued entries feed k = 5)."""
    import __graft_entry__ as g
    g.build()
    EXE = os.path.join(ROOT, "metamdbg_b200", "host", "mdbg_gpu_firstpass")
    rs = synth.make_readset(1500, 7000, seed=56, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.004)
    bases, offs = synth.fill_reads(rs)
    fa = tmp_path / "reads.fasta"
    raw = bases.tobytes()
    with open(fa, "wb") as f:
        for r in range(rs.n_reads):
            f.write(b">r%d\n" % r + raw[int(offs[r]):int(offs[r + 1])] + b"\n")
    d1 = tmp_path / "one"
    d1.mkdir()
    out = subprocess.run([EXE, str(fa), str(d1), "--batch-mbp", "4"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    pm, po = [], [0]
    for rr in range(rs.n_reads):
        q, _ = oracle.purge_palindrome(m[int(mo[rr]):int(mo[rr + 1])], 4, 60)
        pm.append(q); po.append(po[-1] + len(q))
    m, mo = np.concatenate(pm).astype(np.uint32), np.array(po, np.uint64)
    c = oracle.count(m, mo, 4, 2)
    r = oracle.rescue(m, mo, 4, c["hashes"], c["abundances"])
    # --max-k: the k > firstK passes (GpuNextKCounter) from the table the context holds, default mode: the previous
    # table of k = 5 carries the rescued entries, as kminmerData_abundance.txt does upstream
    d3 = tmp_path / "three"
    d3.mkdir()
    out = subprocess.run([EXE, "--from-read-data", str(d1 / "read_data_corrected.txt"), str(d3), "--min-abundance", "0",
                          "--max-k", "7", "--edges", "--unitigs"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    # --edges: edges.bin of the k = 4 node set (GpuEdgeIndexer) = the oracle's EdgeIndexer key set
    nodes = np.frombuffer(open(d3 / "kminmerData_min.txt", "rb").read(), dtype=np.uint32).reshape(-1, 4)
    we = oracle.edge_index(nodes, 4)
    eb = np.frombuffer(open(d3 / "edges.bin", "rb").read(), dtype=np.uint64).reshape(-1, 2)       # {low, high}
    assert {(int(h[1]), int(h[0])) for h in eb} == {(int(h[0]), int(h[1])) for h in we["hashes"]} and len(eb) == len(we["hashes"]) > 500
    words = out.stdout.split()
    assert int(words[words.index("edges") + 1]) == len(eb) and int(words[words.index("edge_checksum") + 1]) == we["checksum"]
    # --unitigs: unitigGraph.nodes.bin of that node set (GpuUnitigBuilder) = the bytes computeUnitigNodes +
    # computeDeterministicUnitigs write (oracle restatement, pinned against the reference's own run), and the per-node
    # abundances of dumpUnitigAbundances
    wu = oracle.unitigs(nodes, 4)
    want_bytes, want_ab = b"", b""
    abf = np.frombuffer(open(d3 / "kminmerData_abundance.txt", "rb").read(), dtype=np.uint8).reshape(-1, 20)
    ab_of = {(int(h), int(l)): int(c_) for l, h, c_ in zip(abf[:, 0:8].copy().view(np.uint64)[:, 0], abf[:, 8:16].copy().view(np.uint64)[:, 0],
                                                          abf[:, 16:20].copy().view(np.uint32)[:, 0])}
    for i in range(len(wu["offsets"]) - 1):
        seq = wu["minimizers"][int(wu["offsets"][i]):int(wu["offsets"][i + 1])]
        want_bytes += np.uint32(len(seq)).tobytes() + seq.tobytes() + np.uint32(2 * i).tobytes()
        a = np.array([ab_of[oracle.hash128(v)] for v in oracle.kminmers(seq, 4)[0]], np.uint32)
        want_ab += np.uint32(2 * i).tobytes() + np.uint32(len(a)).tobytes() + a.tobytes()
    assert open(d3 / "unitigGraph.nodes.bin", "rb").read() == want_bytes and len(wu["offsets"]) > 50
    assert open(d3 / "unitigGraph.nodes.abundances.bin", "rb").read() == want_ab
    assert int(words[words.index("unitigs") + 1]) == len(wu["offsets"]) - 1
    # unitigGraph.edges.successors.bin: the records indexUnitigEdges + computeUnitigEdges write with one thread
    we = oracle.unitig_edges(wu["offsets"], wu["minimizers"], 4)
    want_e = b""
    for i in range(len(wu["offsets"]) - 1):
        want_e += np.uint32(2 * i).tobytes()
        for o in (0, 1):
            lst = we["targets"][int(we["offsets"][2 * i + o]):int(we["offsets"][2 * i + o + 1])]
            want_e += np.uint32(len(lst)).tobytes() + lst.tobytes()
    assert open(d3 / "unitigGraph.edges.successors.bin", "rb").read() == want_e and we["n_edges"] > 50
    assert int(words[words.index("unitig_edges") + 1]) == we["n_edges"] and int(words[words.index("checksum_unitig_edges") + 1]) == we["checksum"]
    ph = np.concatenate([c["hashes"], r["hashes"]]); pa = np.concatenate([c["abundances"], np.ones(len(r["hashes"]), np.uint32)])
    for kk in (5, 6, 7):
        nk = oracle.next_k(m, mo, kk, ph, pa)
        ab = np.frombuffer(open(d3 / f"kminmerData_abundance_k{kk}.txt", "rb").read(), dtype=np.uint8).reshape(-1, 20)
        lo = ab[:, 0:8].copy().view(np.uint64)[:, 0]; hi = ab[:, 8:16].copy().view(np.uint64)[:, 0]
        cnt = ab[:, 16:20].copy().view(np.uint32)[:, 0]
        assert {(int(h), int(l)): int(c_) for h, l, c_ in zip(hi, lo, cnt)} == \
            {(int(h[0]), int(h[1])): int(a) for h, a in zip(nk["hashes"], nk["abundances"])}, f"k={kk}"
        vec = np.frombuffer(open(d3 / f"kminmerData_min_k{kk}.txt", "rb").read(), dtype=np.uint32).reshape(-1, kk)
        assert sorted(map(tuple, vec.tolist())) == sorted(map(tuple, nk["vecs"].tolist())) and len(vec) > 500
        ph, pa = nk["hashes"], nk["abundances"]
