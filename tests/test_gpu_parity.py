"""GPU parity tests proper: libmdbg_b200.so (through the C ABI) against the CPU
oracle on the same seeded inputs, against the golden vectors minted from the
reference's code, and size-independent properties.  Bit-exact everywhere
(integer / byte / index work)."""
import os

import numpy as np
import pytest

from metamdbg_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


EMULATED = bool(os.environ.get("MDBG_EMU_LIB"))      # child run of tests/test_capi_emulated_cpu.py (see conftest.py)


@pytest.fixture(scope="module")
def built():
    if EMULATED:
        return True
    import __graft_entry__ as g
    g.build()
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return True


def real_gpu_only():
    if EMULATED:
        pytest.skip("uses torch device tensors: real GPU only")


def device_array(arr: np.ndarray, pad: int = 0):
    """(device pointer, owner) of a copy of `arr` (+ pad bytes): a torch CUDA tensor on the GPU box; under the CPU
    emulator device memory is host memory, so a 16-byte aligned numpy copy plays the part."""
    raw = np.ascontiguousarray(arr).view(np.uint8).ravel()
    if EMULATED:
        store = np.zeros(len(raw) + pad + 16, np.uint8)
        shift = (-store.ctypes.data) % 16
        view = store[shift:shift + len(raw) + pad]
        view[:len(raw)] = raw
        return view.ctypes.data, (store, view)
    import torch
    t = torch.zeros(len(raw) + pad, dtype=torch.uint8, device="cuda:0")
    t[:len(raw)] = torch.from_numpy(raw.copy()).to("cuda:0")
    return t.data_ptr(), t


def engine(l=15, d=0.005, hpc=True, bl=None):
    from metamdbg_b200 import Engine
    return Engine(l, d, hpc, bl)


def assert_sketch_equal(sk, mo, m, p, d, tag=""):
    assert np.array_equal(sk.min_offsets, mo), f"min_offsets {tag}"
    assert np.array_equal(sk.minimizers, m), f"minimizers {tag}"
    assert np.array_equal(sk.positions, p), f"positions {tag}"
    assert np.array_equal(sk.directions, d), f"directions {tag}"


def table_dict(hashes_h1h2, abund):
    return {(int(h[0]), int(h[1])): int(a) for h, a in zip(hashes_h1h2, abund)}


def check_table(tab, ref_hashes, ref_abund, ref_vecs, stats=None):
    assert tab.as_dict() == table_dict(ref_hashes, ref_abund)
    got_vecs = {(int(h[1]), int(h[0])): tuple(int(x) for x in v) for h, v in zip(tab.hashes, tab.kminmers)}
    want_vecs = {(int(h[0]), int(h[1])): tuple(int(x) for x in v) for h, v in zip(ref_hashes, ref_vecs)}
    assert got_vecs == want_vecs
    if stats is not None:
        assert [tab.n_instances, tab.n_distinct] == [int(stats[0]), int(stats[1])]


# ---------------------------------------------------------------- sketch

def test_sketch_hifi_vs_oracle(built, oracle):
    rs = synth.make_readset(2000, 9000, seed=5, n_genomes=3, genome_len_range=(200_000, 400_000))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    sk = eng.sketch_batch(bases, offs)
    assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.005, True))
    assert sk.min_offsets[-1] > 20_000
    eng.close()


def test_sketch_ont_vs_oracle(built, oracle):
    rs = synth.make_readset(1500, 8000, seed=6, n_genomes=2, genome_len_range=(200_000, 300_000), err=0.02)
    bases, offs = synth.fill_reads(rs)
    eng = engine(15, 0.025, False)
    sk = eng.sketch_batch(bases, offs)
    assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.025, False))
    eng.close()


def test_sketch_golden_hifi_and_ont(built):
    g = load("hifi_small.npz")
    eng = engine()
    sk = eng.sketch_batch(g["bases"], g["offsets"])
    assert_sketch_equal(sk, g["min_offsets"], g["minimizers"], g["positions"], g["directions"], "hifi golden")
    eng.close()
    g = load("ont_small.npz")
    eng = engine(15, 0.025, False, g["blacklist"])
    sk = eng.sketch_batch(g["bases"], g["offsets"])
    assert_sketch_equal(sk, g["min_offsets"], g["minimizers"], g["positions"], g["directions"], "ont golden")
    eng.close()


def test_sketch_edge_cases_golden(built):
    """N / lower case / '#' / empty / short / homopolymer reads, l in {5,11,15,16},
    densities up to 0.6 (exercises the slot-overflow re-run), HPC on and off."""
    g = load("edge_cases.npz")
    for tag in g["cases"]:
        tag = str(tag)
        l = int(tag.split("_")[0][1:]); dens = float(tag.split("_")[1][1:]); hpc = tag.endswith("h1")
        eng = engine(l, dens, hpc)
        sk = eng.sketch_batch(g["bases"], g["offsets"])
        assert_sketch_equal(sk, g[tag + "_mo"], g[tag + "_m"], g[tag + "_p"], g[tag + "_d"], tag)
        eng.close()


def test_sketch_long_and_ragged_reads(built, oracle):
    """One 300 kbp read, many tiny reads, every start alignment 0..15."""
    rng = np.random.default_rng(9)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 300_000))]
    for a in range(40):
        reads.append(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(0, 700)) + a)))
    reads.append(bytes(np.repeat(rng.choice(np.frombuffer(b"ACGT", np.uint8), 5000), rng.integers(1, 9, 5000))))
    bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    for hpc in (True, False):
        eng = engine(15, 0.01, hpc)
        sk = eng.sketch_batch(bases, offs)
        assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.01, hpc), tag=f"hpc={hpc}")
        eng.close()


def test_sketch_empty_batch(built):
    eng = engine()
    sk = eng.sketch_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert sk.n_reads == 0 and len(sk.minimizers) == 0
    eng.close()


def test_sketch_reverse_complement_symmetry(built):
    """Size-independent property: sketching the reverse complement yields the same
    canonical minimizers mirrored, with flipped strands."""
    rs = synth.make_readset(400, 12000, seed=12, n_genomes=1, genome_len_range=(500_000, 500_001))
    bases, offs = synth.fill_reads(rs)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = np.empty_like(bases)
    for r in range(len(offs) - 1):
        lo, hi = int(offs[r]), int(offs[r + 1])
        rc[lo:hi] = comp[bases[lo:hi]][::-1]
    eng = engine(15, 0.005, False)          # no HPC: positions map exactly
    a = eng.sketch_batch(bases, offs)
    b = eng.sketch_batch(rc, offs)
    n_checked = 0
    for r in range(a.n_reads):
        ma, pa, da = a.read(r)
        mb, pb, db = b.read(r)
        n_l = int(offs[r + 1] - offs[r]) - 15 + 1
        assert np.array_equal(ma, mb[::-1])
        assert np.array_equal(pa, (n_l - 1 - pb.astype(np.int64))[::-1])
        pal = ma == 0  # never true; palindromic l-mers impossible for odd l
        assert np.array_equal(da[~pal], (1 - db[::-1])[~pal])
        n_checked += len(ma)
    assert n_checked > 10_000
    eng.close()


# ---------------------------------------------------------------- store / purge / count

def test_purge_and_count_minspace_golden(built):
    g = load("minspace_palindromes.npz")
    eng = engine()
    eng.store_append(g["minimizers"], g["offsets"])
    changed = eng.purge_palindromes(4, 12)
    assert changed > 0
    so, sm = eng.store_fetch()
    assert np.array_equal(so, g["purged_offsets"]) and np.array_equal(sm, g["purged_minimizers"])
    for k in (4, 6, 9, 21):
        eng.count_begin(k)
        eng.count_add_store()
        tab = eng.count_finalize(3)
        check_table(tab, g[f"k{k}_hashes"], g[f"k{k}_abund"], g[f"k{k}_vecs"], g[f"k{k}_stats"])
    eng.close()


def test_full_path_hifi_golden(built):
    """Reads -> sketch -> store -> purge -> count(k=4,5,7) on the device, against
    the reference-minted table."""
    g = load("hifi_small.npz")
    eng = engine()
    eng.sketch_batch(g["bases"], g["offsets"], append_to_store=True, fetch=False)
    eng.purge_palindromes(4, 60)
    so, sm = eng.store_fetch()
    assert np.array_equal(so, g["purged_offsets"]) and np.array_equal(sm, g["purged_minimizers"])
    for k in (4, 5, 7):
        eng.count_begin(k)
        eng.count_add_store()
        tab = eng.count_finalize(2)
        check_table(tab, g[f"k{k}_hashes"], g[f"k{k}_abund"], g[f"k{k}_vecs"], g[f"k{k}_stats"])
        st = eng.count_stats(2)
        assert st["n_entries"] == len(tab.abundances) and st["checksum"] == tab.checksum
    eng.close()


def test_count_vs_oracle_multi_k_and_batches(built, oracle):
    rs = synth.make_readset(3000, 9000, seed=31, n_genomes=1, genome_len_range=(300_000, 300_001))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    half = 1500
    eng.sketch_batch(bases[:int(offs[half])], offs[:half + 1], append_to_store=True, fetch=False)
    eng.sketch_batch(bases[int(offs[half]):], offs[half:] - offs[half], append_to_store=True, fetch=False)
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    so, sm = eng.store_fetch()
    assert np.array_equal(so, mo) and np.array_equal(sm, m)
    for k, min_ab in ((4, 2), (4, 5), (8, 0), (13, 2), (21, 2)):
        eng.count_begin(k)
        eng.count_add_store(0, half)            # two insert launches into one table
        eng.count_add_store(half, 2 ** 63)
        tab = eng.count_finalize(min_ab)
        ref = oracle.count(m, mo, k, min_ab)
        check_table(tab, ref["hashes"], ref["abundances"], ref["vecs"], [ref["n_instances"], ref["n_distinct"]])
        assert tab.checksum == oracle.checksum(ref["hashes"], ref["abundances"])
    eng.close()


def test_count_add_host_reads(built, oracle):
    rng = np.random.default_rng(3)
    reads = [rng.integers(0, 50, size=int(rng.integers(0, 90))).astype(np.uint32) for _ in range(2000)]
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    mins = np.concatenate(reads).astype(np.uint32)
    from metamdbg_b200 import KminmerCounter
    eng = engine()
    kc = KminmerCounter(eng, 5, expected_distinct=200_000)
    kc.add_reads(mins[:int(offs[700])], offs[:701])
    kc.add_reads(mins, offs[700:])               # offsets not starting at 0
    tab = kc.execute(2)
    ref = oracle.count(mins, offs, 5, 2)
    check_table(tab, ref["hashes"], ref["abundances"], ref["vecs"])
    eng.close()


def test_table_full_is_reported(built, monkeypatch):
    """A table that is too small for its keys is rebuilt larger (round 2); with that switched off the pass reports
    MDBG_ERR_TABLE_FULL as before, and the context stays usable."""
    from metamdbg_b200 import MdbgError
    rng = np.random.default_rng(4)
    mins = rng.integers(0, 2 ** 32, size=200_000, dtype=np.uint64).astype(np.uint32)
    offs = np.array([0, len(mins)], np.uint64)
    eng = engine()
    eng.count_begin(4, expected_distinct=256)    # far too small: grows
    eng.count_add(mins, offs)
    assert eng.count_stats(2)["n_distinct"] == len(mins) - 3
    eng.close()
    monkeypatch.setenv("MDBG_TABLE_AUTOGROW", "0")
    eng = engine()
    eng.count_begin(4, expected_distinct=256)
    with pytest.raises(MdbgError) as e:
        eng.count_add(mins, offs)
    assert e.value.status == 4
    eng.count_begin(4, expected_distinct=len(mins))
    eng.count_add_store()
    assert eng.count_stats(2)["n_distinct"] == len(mins) - 3
    eng.close()


# ---------------------------------------------------------------- device-resident path + generator

def test_device_generator_matches_numpy_and_device_sketch(built, oracle):
    real_gpu_only()
    import torch
    rs = synth.make_readset(500, 7000, seed=77, n_genomes=2, genome_len_range=(100_000, 200_000))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    dev = torch.device("cuda:0")
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    d_vs = torch.from_numpy(rs.vstart.astype(np.int64)).to(dev)
    d_st = torch.from_numpy(rs.strand).to(dev)
    d_bases = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device=dev)
    eng.synth_fill_reads(d_bases.data_ptr(), d_off.data_ptr(), d_vs.data_ptr(), d_st.data_ptr(), rs.n_reads, 0,
                         rs.seed, rs.err_q24)
    eng.synchronize()
    assert np.array_equal(d_bases[:int(offs[-1])].cpu().numpy(), bases)
    out = eng.sketch_batch_device(d_bases.data_ptr(), d_off.data_ptr(), rs.n_reads, int(offs[-1]), True)
    sk = eng.sketch_fetch()
    assert out.n_minimizers == len(sk.minimizers)
    assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.005, True))
    eng.close()


def test_python_mirror_single_read(built, oracle):
    from metamdbg_b200 import MinimizerParser
    rs = synth.make_readset(3, 20000, seed=8, n_genomes=1, genome_len_range=(100_000, 100_001))
    bases, offs = synth.fill_reads(rs)
    mp = MinimizerParser(15, 0.005, None, True)
    seq = bases[:int(offs[1])].tobytes()
    m, p, d = mp.parse(seq)
    rm, rp, rd = oracle.sketch_read(seq, 15, 0.005, True)
    assert np.array_equal(m, rm) and np.array_equal(p, rp) and np.array_equal(d, rd)
    mp.engine.close()


# ---------------------------------------------------------------- rescue (A7b) and next-k passes (A8/A9)

def _h1h2(tab):
    """CountTable.hashes (low, high) -> list of (h1, h2) tuples."""
    return [(int(h[1]), int(h[0])) for h in tab.hashes]


def test_rescue_and_next_k_golden(built):
    g = load("minspace_multik.npz")
    mins, offs = g["minimizers"], g["offsets"]
    ns = int(g["k4_n_solid"])
    eng = engine()
    eng.store_append(mins, offs)
    # default mode: count, rescue, finalize -> solid + rescued entries
    eng.count_begin(4)
    eng.count_add_store()
    n_reads_rescued = eng.count_rescue()
    tab = eng.count_finalize(0)
    assert n_reads_rescued > 0 and tab.n_rescued == int(g["k4_n_rescued"])
    assert tab.as_dict() == table_dict(g["k4_hashes"], g["k4_abund"])
    want_vecs = {(int(h[0]), int(h[1])): tuple(int(x) for x in v) for h, v in zip(g["k4_hashes"], g["k4_vecs"])}
    got_vecs = {hh: tuple(int(x) for x in v) for hh, v in zip(_h1h2(tab), tab.kminmers)}
    assert got_vecs == want_vecs
    # k = 5 from the table just built (device side), k = 6 from the k = 5 table
    eng.prev_from_current(0)
    eng.count_begin(5)
    eng.count_add_store_next_k()
    t5 = eng.count_finalize(0)
    assert t5.as_dict() == table_dict(g["k5_hashes"], g["k5_abund"])
    want_vecs = {(int(h[0]), int(h[1])): tuple(int(x) for x in v) for h, v in zip(g["k5_hashes"], g["k5_vecs"])}
    assert {hh: tuple(int(x) for x in v) for hh, v in zip(_h1h2(t5), t5.kminmers)} == want_vecs
    eng.prev_from_current(0)
    eng.count_begin(6)
    eng.count_add_store_next_k()
    t6 = eng.count_finalize(0)
    assert t6.as_dict() == table_dict(g["k6_hashes"], g["k6_abund"]) and len(t6.abundances) > 20
    eng.close()


def test_next_k_with_host_loaded_and_patched_prev_table(built, oracle):
    """prev table from host pairs (kminmerData_abundance_prev.txt) + a patch (refined abundances incl. zeros)."""
    rs = synth.make_readset(2500, 9000, seed=61, n_genomes=1, genome_len_range=(250_000, 250_001))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    prev = oracle.count(sm, so, 4, 2)
    rng = np.random.default_rng(5)
    patch_idx = np.nonzero(rng.random(len(prev["abundances"])) < 0.2)[0]
    patched = prev["abundances"].copy()
    patched[patch_idx] = rng.integers(0, 4, size=len(patch_idx)).astype(np.uint32)
    lohi = np.stack([prev["hashes"][:, 1], prev["hashes"][:, 0]], axis=1)       # (low, high) as on disk
    eng.prev_load(lohi, prev["abundances"], clear=True)
    eng.prev_load(lohi[patch_idx], patched[patch_idx], clear=False)
    for k in (5,):
        eng.count_begin(k)
        eng.count_add_store_next_k()
        tab = eng.count_finalize(0)
        want = oracle.next_k(sm, so, k, prev["hashes"], patched)
        assert tab.as_dict() == table_dict(want["hashes"], want["abundances"]) and len(want["abundances"]) > 1000
    eng.close()


def test_rescue_on_sketched_reads_vs_oracle(built, oracle):
    rs = synth.make_readset(3000, 8000, seed=62, n_genomes=3, genome_len_range=(150_000, 600_000), err=0.004)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    eng.count_begin(4)
    eng.count_add_store()
    eng.count_rescue()
    tab = eng.count_finalize(0)
    c = oracle.count(sm, so, 4, 2)
    r = oracle.rescue(sm, so, 4, c["hashes"], c["abundances"])
    want = table_dict(c["hashes"], c["abundances"])
    for h in r["hashes"]:
        want[(int(h[0]), int(h[1]))] = 1
    assert tab.as_dict() == want and tab.n_rescued == len(r["hashes"]) > 0
    eng.close()


# ---------------------------------------------------------------- side outputs (A3b) against the real readSelection stage

@pytest.mark.parametrize("tag,hpc,dens", [("hifi", True, 0.005), ("ont", False, 0.025)])
def test_side_outputs_golden(built, tag, hpc, dens):
    g = load(f"readselection_{tag}.npz")
    eng = engine(15, dens, hpc, g["blacklist"])
    eng.set_read_filters(True)
    sk, aux = eng.sketch_batch_q(g["bases"], g["quals"], g["offsets"])
    assert_sketch_equal(sk, g["min_offsets"], g["minimizers"], g["positions"], g["directions"], tag)
    assert np.array_equal(aux["qualities"], g["qualities"])
    assert aux["mean_quality"].tobytes() == g["mean_quality"].tobytes()         # bit-exact floats
    assert aux["low_complexity"].sum() >= 2
    # filtered reads are exactly the empty records of long reads
    lens = np.diff(g["offsets"].astype(np.int64))
    empty = np.diff(g["min_offsets"].astype(np.int64)) == 0
    assert np.all(aux["low_complexity"][lens > 2000] == empty[lens > 2000])
    # without qualities (FASTA): per-minimizer quality is 1 and the mean quality is NaN
    sk2, aux2 = eng.sketch_batch_q(g["bases"], None, g["offsets"])
    assert np.array_equal(sk2.minimizers, g["minimizers"]) and np.all(aux2["qualities"] == 1)
    assert np.all(np.isnan(aux2["mean_quality"]))
    eng.close()


def test_side_outputs_vs_oracle_with_overflow(built, oracle):
    """density 0.5 overflows the padded slots: qualities must survive the exact re-sketch."""
    rng = np.random.default_rng(17)
    rs = synth.make_readset(60, 3000, seed=88, n_genomes=1, genome_len_range=(50_000, 50_001))
    bases, offs = synth.fill_reads(rs)
    quals = (rng.integers(1, 70, size=len(bases)).astype(np.uint8) + 33)
    eng = engine(15, 0.5, True)
    sk, aux = eng.sketch_batch_q(bases, quals, offs)
    raw, qraw = bases.tobytes(), quals.tobytes()
    for r in range(rs.n_reads):
        s, q = raw[int(offs[r]):int(offs[r + 1])], qraw[int(offs[r]):int(offs[r + 1])]
        m, p, d = oracle.sketch_read(s, 15, 0.5, True)
        mq, cx, mins_q = oracle.read_aux(s, q, 15, True, p)
        lo, hi = int(sk.min_offsets[r]), int(sk.min_offsets[r + 1])
        assert np.array_equal(sk.minimizers[lo:hi], m) and np.array_equal(aux["qualities"][lo:hi], mins_q)
        assert np.float32(aux["mean_quality"][r]).tobytes() == np.float32(mq).tobytes()
        assert aux["complexity"][r] == cx or (np.isnan(cx) and np.isnan(aux["complexity"][r]))
    eng.close()


# ---------------------------------------------------------------- host batches travel 2-bit packed

def test_packed_host_transfer_with_dirty_reads(built, oracle):
    """> 1 MB host batches are 2-bit packed before H2D; reads with N / lower case / '#' keep their ASCII form.
    Results must equal the oracle and the unpacked transfer, HPC on and off."""
    rng = np.random.default_rng(23)
    rs = synth.make_readset(700, 6000, seed=23, n_genomes=2, genome_len_range=(100_000, 200_000))
    bases, offs = synth.fill_reads(rs)
    reads = [bytearray(bases[int(offs[r]):int(offs[r + 1])].tobytes()) for r in range(rs.n_reads)]
    for r in range(0, rs.n_reads, 7):                      # every 7th read gets something outside "ACGT"
        rd = reads[r]
        kind = (r // 7) % 4
        pos = int(rng.integers(0, len(rd)))
        if kind == 0:
            rd[pos:pos + 3] = b"NNN"[:len(rd) - pos]
        elif kind == 1:
            rd[pos] = ord("acgt"[int(rng.integers(0, 4))])
        elif kind == 2:
            rd[pos] = ord("#")
        else:
            rd[0] = ord("n")
    reads += [bytearray(b""), bytearray(b"ACGT"), bytearray(b"A" * 5000), bytearray(b"N" * 100)]
    flat = np.frombuffer(b"".join(bytes(x) for x in reads), np.uint8).copy()
    o2 = np.zeros(len(reads) + 1, np.uint64)
    o2[1:] = np.cumsum([len(x) for x in reads])
    assert len(flat) > (1 << 21)
    for hpc, dens in ((True, 0.005), (False, 0.02)):
        want = oracle.sketch_batch(flat, o2, 15, dens, hpc)
        eng = engine(15, dens, hpc)
        sk_packed = eng.sketch_batch(flat, o2, append_to_store=True)
        assert_sketch_equal(sk_packed, *want, tag=f"packed hpc={hpc}")
        eng.set_host_packing(False)
        sk_ascii = eng.sketch_batch(flat, o2)
        assert_sketch_equal(sk_ascii, *want, tag=f"ascii hpc={hpc}")
        so, sm = eng.store_fetch()
        assert np.array_equal(so, want[0]) and np.array_equal(sm, want[1])
        eng.close()


def test_ont_density_rethreshold_and_count(built, oracle):
    """BASELINE config 3 shape: ONT reads (no HPC) sketched at the correction density 0.025, re-thresholded to the
    assembly density 0.005 (Utils::applyDensityThreshold), then counted at k=4."""
    rs = synth.make_readset(2500, 8000, seed=33, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.02)
    bases, offs = synth.fill_reads(rs)
    eng = engine(15, 0.025, False)
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.025, False)
    changed = eng.store_apply_density(0.005)
    so, sm = eng.store_fetch()
    want, wo = [], [0]
    for r in range(rs.n_reads):
        q = oracle.apply_density(m[int(mo[r]):int(mo[r + 1])], 0.005)
        want.append(q); wo.append(wo[-1] + len(q))
    assert changed > 0 and np.array_equal(sm, np.concatenate(want)) and np.array_equal(so, np.array(wo, np.uint64))
    eng.count_begin(4)
    eng.count_add_store()
    tab = eng.count_finalize(2)
    ref = oracle.count(sm, so, 4, 2)
    check_table(tab, ref["hashes"], ref["abundances"], ref["vecs"], [ref["n_instances"], ref["n_distinct"]])
    eng.close()


def test_hybrid_transfer_from_pinned_host_memory(built, oracle):
    """Pinned caller buffers: pieces travel either as ASCII by DMA (PCIe idle) or 2-bit packed (PCIe busy); both
    modes -- and dirty reads inside packed pieces -- must give the oracle's sketch."""
    real_gpu_only()
    import torch
    rs = synth.make_readset(40_000, 9000, seed=29, n_genomes=3, genome_len_range=(300_000, 600_000))
    sub = rs.subset(0, 40_000)
    dev = torch.device("cuda:0")
    eng = engine()
    d_off = torch.from_numpy(sub.offsets.astype(np.int64)).to(dev)
    d_vs = torch.from_numpy(sub.vstart.astype(np.int64)).to(dev)
    d_st = torch.from_numpy(sub.strand).to(dev)
    d_bases = torch.empty(sub.n_bases + 64, dtype=torch.uint8, device=dev)
    eng.synth_fill_reads(d_bases.data_ptr(), d_off.data_ptr(), d_vs.data_ptr(), d_st.data_ptr(), sub.n_reads, 0,
                         sub.seed, sub.err_q24)
    eng.synchronize()
    h = torch.empty(sub.n_bases, dtype=torch.uint8, pin_memory=True)
    h.copy_(d_bases[:sub.n_bases])
    torch.cuda.synchronize()
    hb = h.numpy()
    for r in range(0, sub.n_reads, 501):                     # a few dirty reads
        hb[int(sub.offsets[r]) + 7] = ord("N")
    assert sub.n_bases > 300 * (1 << 20)                      # several 128 MB pieces
    eng.set_host_packing(2)                                   # hybrid: ASCII-by-DMA and packed pieces mixed
    n = eng.sketch_batch_ptr(h.data_ptr(), sub.offsets, append_to_store=True)
    sk = eng.sketch_fetch()
    assert n == len(sk.minimizers)
    eng.set_host_packing(True)                                # packed pieces only
    n2 = eng.sketch_batch_ptr(h.data_ptr(), sub.offsets, append_to_store=False)
    sk2 = eng.sketch_fetch()
    assert n2 == n
    assert_sketch_equal(sk2, sk.min_offsets, sk.minimizers, sk.positions, sk.directions, "packed vs hybrid")
    # oracle on a sample of reads (incl. the dirty ones) + device-resident path on everything
    raw = hb.tobytes()
    for r in list(range(0, sub.n_reads, 501)) + list(range(1, sub.n_reads, 997)):
        m, p, d = oracle.sketch_read(raw[int(sub.offsets[r]):int(sub.offsets[r + 1])], 15, 0.005, True)
        gm, gp, gd = sk.read(r)
        assert np.array_equal(gm, m) and np.array_equal(gp, p) and np.array_equal(gd, d), r
    d_bases[:sub.n_bases].copy_(h)
    eng.sketch_batch_device(d_bases.data_ptr(), d_off.data_ptr(), sub.n_reads, sub.n_bases, False)
    ref = eng.sketch_fetch()
    assert_sketch_equal(sk, ref.min_offsets, ref.minimizers, ref.positions, ref.directions, "hybrid vs device-resident")
    eng.close()


def test_device_resident_packed_input(built, oracle):
    """Reads kept 2-bit packed in HBM (the layout the design brief names) sketch to the same CSR as ASCII reads."""
    real_gpu_only()
    import torch
    rs = synth.make_readset(1500, 7000, seed=44, n_genomes=2, genome_len_range=(100_000, 200_000))
    bases, offs = synth.fill_reads(rs)
    words, woff = synth.pack_2bit(bases, offs)
    dev = torch.device("cuda:0")
    d_words = torch.from_numpy(words.astype(np.int64).astype(np.uint32).view(np.int32)).to(dev)
    d_woff = torch.from_numpy(woff.astype(np.int64)).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    for hpc in (True, False):
        eng = engine(15, 0.005, hpc)
        out = eng.sketch_batch_device_packed(d_words.data_ptr(), d_woff.data_ptr(), d_off.data_ptr(), rs.n_reads,
                                             int(offs[-1]), True)
        sk = eng.sketch_fetch()
        assert out.n_minimizers == len(sk.minimizers)
        assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.005, hpc), tag=f"packed device hpc={hpc}")
        eng.close()


@pytest.mark.parametrize("dens", [0.0, 1.0, 0.999])
def test_sketch_extreme_densities(built, oracle, dens):
    """density 0 selects nothing; density >= 1 overflows the 32-bit high-word test and must fall back to the exact
    path (every position but the trimmed first/last is selected)."""
    rng = np.random.default_rng(3)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n)) for n in (0, 14, 15, 16, 17, 700, 2500)]
    reads.append(b"ACGTNACGT" * 60)
    bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    for hpc in (True, False):
        eng = engine(15, dens, hpc)
        sk = eng.sketch_batch(bases, offs)
        assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, dens, hpc), tag=f"d={dens} hpc={hpc}")
        if dens == 0.0:
            assert len(sk.minimizers) == 0
        eng.close()


def test_sketch_arbitrary_bytes(built, oracle):
    """Reads of arbitrary bytes 1..255 (>= 0x80 included): the keep-mask compares raw bytes, the base code is
    (c >> 1) & 3 with bit 3 = invalid for ANY byte (Kmer.hpp:505-556), host buffers take the ASCII route for such
    reads even when 2-bit packing is on.  Both transfer modes, both HPC modes, fast (l=15) and generic l.
    NUL is excluded: reads are NUL-free by contract (the reference's non-HPC branch builds a
    C string and would stop at it, its FASTQ parser never produces one)."""
    rng = np.random.default_rng(21)
    reads = []
    for i in range(60):
        n = int(rng.integers(0, 3000))
        kind = i % 4
        if kind == 0:
            s = rng.integers(1, 256, n).astype(np.uint8)
        elif kind == 1:                                         # runs of arbitrary bytes (HPC on raw bytes)
            runs = rng.geometric(0.4, max(1, n // 2))
            s = np.repeat(rng.integers(1, 256, len(runs)).astype(np.uint8), runs)[:n]
        elif kind == 2:                                         # ACGT with valid-coded non-ACGT bytes ('a', 'e', 0x02 ...)
            s = rng.choice(np.frombuffer(b"ACGTacgtBEDF\x02\x04\x06\xe1\xe3", np.uint8), n)
        else:
            s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
        reads.append(s.tobytes())
    bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    for l, dens in ((15, 0.05), (11, 0.02)):
        for hpc in (True, False):
            want = oracle.sketch_batch(bases, offs, l, dens, hpc)
            assert int(want[0][-1]) > 50
            for packing in (0, 1):
                eng = engine(l, dens, hpc)
                eng.set_host_packing(packing)
                sk = eng.sketch_batch(bases, offs)
                assert_sketch_equal(sk, *want, tag=f"l={l} hpc={hpc} packing={packing}")
                eng.close()


def test_bad_host_arrays_are_rejected(built):
    """Error behaviour of the boundary: malformed CSR arrays come back as MDBG_ERR_ARG (status 2) with a message,
    nothing is launched, and the context stays usable."""
    from metamdbg_b200.engine import MdbgError
    eng = engine(15, 0.05, True)
    bases = np.frombuffer(b"ACGTTGCAAGGCTTAACCGGTTAACG" * 8, np.uint8).copy()
    good = np.array([0, 100, len(bases)], np.uint64)
    launches0 = eng.kernel_launches
    for offs in (np.array([4, 100, 200], np.uint64),           # does not start at 0
                 np.array([0, 150, 100], np.uint64),           # decreasing
                 np.array([0, 1 << 31, (1 << 31) + 5], np.uint64)):   # a read of 2^31 bases
        with pytest.raises(MdbgError) as ei:
            eng.sketch_batch(bases, offs)
        assert ei.value.status == 2 and "offsets" in str(ei.value)
        with pytest.raises(MdbgError):
            eng.sketch_batch_q(bases, None, offs)
    with pytest.raises(MdbgError) as ei:
        eng.store_append(np.arange(10, dtype=np.uint32), np.array([0, 7, 3], np.uint64))
    assert ei.value.status == 2 and "non-decreasing" in str(ei.value)
    assert eng.kernel_launches == launches0
    sk = eng.sketch_batch(bases, good)                          # still works afterwards
    assert sk.n_reads == 2
    eng.close()


def test_side_outputs_refuse_the_hpc_sentinel_in_reads(built, oracle):
    """'#' is EncoderRLE's internal sentinel: with HPC on the side-output entry point refuses a base string that
    holds it (MDBG_ERR_ARG) instead of returning quality windows that differ from the reference's shifted
    rlePositions; the plain sketch of the same batch stays bit-exact, and with HPC off the batch is accepted."""
    from metamdbg_b200.engine import MdbgError
    rng = np.random.default_rng(8)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 900)) for _ in range(4)]
    reads[2] = reads[2][:400] + b"#" + reads[2][401:]
    bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    quals = rng.integers(40, 70, len(bases)).astype(np.uint8)
    eng = engine(15, 0.05, True)
    with pytest.raises(MdbgError) as ei:
        eng.sketch_batch_q(bases, quals, offs)
    assert ei.value.status == 2 and "'#'" in str(ei.value)
    assert_sketch_equal(eng.sketch_batch(bases, offs), *oracle.sketch_batch(bases, offs, 15, 0.05, True), tag="plain sketch")
    eng.close()
    eng = engine(15, 0.05, False)
    sk, aux = eng.sketch_batch_q(bases, quals, offs)
    assert_sketch_equal(sk, *oracle.sketch_batch(bases, offs, 15, 0.05, False), tag="hpc off")
    raw = bases.tobytes()
    for r in range(len(reads)):
        lo, hi = int(offs[r]), int(offs[r + 1])
        pos = sk.positions[int(sk.min_offsets[r]):int(sk.min_offsets[r + 1])]
        _, _, ql = oracle.read_aux(raw[lo:hi], quals[lo:hi].tobytes(), 15, False, pos)
        assert np.array_equal(aux["qualities"][int(sk.min_offsets[r]):int(sk.min_offsets[r + 1])], ql)
    eng.close()
