"""GPU parity tests of the round-2 paths: the packed sketch kernel (variant 2: bulk-copy staged 2-bit input, table
compaction, bit-packed ring), the device-side ASCII -> 2-bit packer, and what was rebuilt around the count table.
All through the C ABI, all bit-exact against the oracle.  They also run on the CPU emulator
(tests/test_capi_emulated_cpu.py) wherever they only need device_array()."""
import os

import numpy as np
import pytest

from metamdbg_b200 import synth
from tests.test_gpu_parity import (EMULATED, assert_sketch_equal, built, device_array, engine, table_dict)  # noqa: F401

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_ASCII = 1 << 63


def _edge_case_reads(rng):
    """Lengths around every boundary of the packed kernel: 16 (word), 64 (16-byte group), 512 (roll block), 1024 (fill
    step), 4096 (bulk-copy tile), plus homopolymer-rich, tandem and dirty reads."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for n in (0, 1, 2, 14, 15, 16, 17, 31, 32, 33, 63, 64, 65, 511, 512, 513, 526, 527, 528, 1023, 1024, 1025, 1039, 2047,
              2048, 2049, 4095, 4096, 4097, 8191, 8192, 8193, 12288, 20000):
        reads.append(acgt[rng.integers(0, 4, n)])
    for i in range(30):
        n = int(rng.integers(30, 9000))
        s = acgt[rng.integers(0, 4, n)].copy()
        if i % 3 == 0:                                      # homopolymer runs, some of them hundreds long
            j = 0
            while j < n:
                run = int(rng.integers(1, 300 if i % 6 == 0 else 6))
                s[j:j + run] = s[j]
                j += run
        if i % 5 == 1:                                      # dirty: N, lower case, IUPAC -> stays ASCII
            for p in rng.integers(0, n, 3):
                s[p] = rng.choice(np.frombuffer(b"NnacgtRY", np.uint8))
        if i % 7 == 2:
            u = int(rng.integers(1, 7))
            s[u:] = np.resize(s[:u], n - u)                 # tandem repeat
        reads.append(s)
    reads.append(np.full(5000, ord("A"), np.uint8))         # one HPC base
    reads.append(np.tile(np.frombuffer(b"AC", np.uint8), 3000))
    bases = np.concatenate(reads).astype(np.uint8)
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return bases, offs


@pytest.mark.parametrize("hpc,dens", [(True, 0.005), (True, 0.05), (False, 0.025), (True, 0.6)])
def test_packed_kernel_edge_cases_device_batch(built, oracle, hpc, dens):
    """ASCII batch in device memory -> pack pass -> packed kernel (+ byte-ring kernel for the dirty reads) equals the
    oracle on reads whose lengths sit on every internal boundary; density 0.6 forces the slot-overflow re-sketch."""
    bases, offs = _edge_case_reads(np.random.default_rng(77))
    want = oracle.sketch_batch(bases, offs, 15, dens, hpc)
    eng = engine(15, dens, hpc)
    assert eng.sketch_variant == int(os.environ.get("MDBG_SKETCH_VARIANT", "2"))
    p_b, keep_b = device_array(bases, pad=64)
    p_o, keep_o = device_array(offs.astype(np.uint64))
    for v in (2, 1, 0):
        eng.set_sketch_variant(v)
        out = eng.sketch_batch_device(p_b, p_o, len(offs) - 1, int(offs[-1]), False)
        sk = eng.sketch_fetch()
        assert out.n_minimizers == len(want[1])
        assert_sketch_equal(sk, *want, tag=f"device batch, variant {v}, hpc={hpc}, d={dens}")
    eng.close()
    del keep_b, keep_o


@pytest.mark.parametrize("hpc", [True, False])
def test_pack_device_and_packed2(built, oracle, hpc):
    """mdbg_pack_device writes the documented layout (word offsets in closed form, dirty reads flagged), and
    mdbg_sketch_batch_device_packed2 on it equals the oracle; the same words at UNALIGNED word offsets (the host
    style layout of round 1) go through the kernel's realigned first step."""
    bases, offs = _edge_case_reads(np.random.default_rng(5))
    n = len(offs) - 1
    want = oracle.sketch_batch(bases, offs, 15, 0.01, hpc)
    eng = engine(15, 0.01, hpc)
    p_b, keep_b = device_array(bases, pad=64)
    p_o, keep_o = device_array(offs.astype(np.uint64))
    n_words = eng.pack_device_words(int(offs[-1]), n)
    p_w, keep_w = device_array(np.full(n_words, 0xA5A5A5A5, np.uint32))
    p_s, keep_s = device_array(np.zeros(n, np.uint64))
    eng.pack_device(p_b, p_o, n, int(offs[-1]), p_w, p_s)
    eng.sketch_batch_device_packed2(p_w, p_s, p_b, p_o, n, int(offs[-1]), False)
    assert_sketch_equal(eng.sketch_fetch(), *want, tag=f"pack_device + packed2 hpc={hpc}")
    eng.synchronize()
    if EMULATED:                                            # device memory is host memory: look at the layout itself
        src = keep_s[1].view(np.uint64)
        words = keep_w[1].view(np.uint32)
        for r in range(n):
            lo, hi = int(offs[r]), int(offs[r + 1])
            s = bases[lo:hi]
            clean = bool(np.isin(s, np.frombuffer(b"ACGT", np.uint8)).all())
            if not clean:
                assert int(src[r]) == SRC_ASCII | lo
                continue
            w0 = ((lo >> 6) + r) << 2
            assert int(src[r]) == w0
            codes = ((s >> 1) & 3).astype(np.uint64)
            for j in range(0, hi - lo, 16):
                c = codes[j:j + 16]
                assert int(words[w0 + j // 16]) == int((c << (2 * np.arange(len(c), dtype=np.uint64))).sum()), (r, j)
    # host-style contiguous packing: arbitrary word alignment of the read starts
    clean_mask = np.array([bool(np.isin(bases[int(offs[r]):int(offs[r + 1])], np.frombuffer(b"ACGT", np.uint8)).all())
                           for r in range(n)])
    words, woff = synth.pack_2bit(bases, offs)
    src2 = woff[:n].astype(np.uint64).copy()
    src2[~clean_mask] = np.uint64(SRC_ASCII) | offs[:n][~clean_mask]
    p_w2, keep_w2 = device_array(np.concatenate([words.astype(np.uint32), np.zeros(8, np.uint32)]))
    p_s2, keep_s2 = device_array(src2)
    eng.sketch_batch_device_packed2(p_w2, p_s2, p_b, p_o, n, int(offs[-1]), False)
    assert_sketch_equal(eng.sketch_fetch(), *want, tag=f"unaligned packed layout hpc={hpc}")
    eng.close()
    del keep_b, keep_o, keep_w, keep_s, keep_w2, keep_s2


def test_packed_kernel_host_batches_with_dirty_reads(built, oracle, monkeypatch):
    """Host batches: packed by the host threads (16-byte aligned read starts, ASCII spill for dirty reads) or sent as
    ASCII and packed on the device, in many pieces -- the dirty list of every piece reaches the byte-ring kernel."""
    monkeypatch.setenv("MDBG_PIECE_BYTES", "60000")
    monkeypatch.setenv("MDBG_PACK_MIN_BYTES", "0")
    bases, offs = _edge_case_reads(np.random.default_rng(11))
    want = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    for packing in (1, 0):
        eng = engine(15, 0.005, True)
        eng.set_host_packing(packing)
        sk = eng.sketch_batch(bases, offs)
        info = eng.last_batch_info()
        assert info["n_pieces"] > 3
        assert_sketch_equal(sk, *want, tag=f"host batch packing={packing}")
        eng.close()


# ---------------------------------------------------------------- count table: sizing, rebuild, later passes

def test_table_grows_when_the_estimate_is_too_small(built, oracle):
    """The table is sized for the expected number of distinct k-min-mers; a store of (almost) all-distinct windows
    outgrows the default estimate (a quarter of the windows), the pass is abandoned and the table rebuilt larger --
    also when the store was inserted in several ranges, and for a next-k pass."""
    rng = np.random.default_rng(12)
    n_reads, per = 600, 300
    mins = rng.integers(0, 2**32 - 1, size=n_reads * per, dtype=np.uint64).astype(np.uint32)
    offs = (np.arange(n_reads + 1) * per).astype(np.uint64)
    mins[per * 10:per * 20] = mins[:per * 10]                     # some repeated reads: solid k-min-mers exist
    mins[per * 20:per * 30] = mins[:per * 10]
    want = oracle.count(mins, offs, 4, 2)
    eng = engine()
    eng.store_append(mins, offs)
    eng.count_begin(4, 0)
    eng.count_add_store(0, 200)                                   # three ranges: the rebuild re-inserts all of them
    eng.count_add_store(200, 450)
    eng.count_add_store(450, n_reads)
    tab = eng.count_finalize(2)
    assert tab.as_dict() == table_dict(want["hashes"], want["abundances"]) and len(want["abundances"]) > 2000
    assert tab.n_distinct > 0.8 * (n_reads * (per - 3))          # the estimate (1/4 of the windows) was far too small
    # explicit, much too small expectation
    eng.count_begin(4, 1000)
    eng.count_add_store()
    assert eng.count_finalize(2).as_dict() == table_dict(want["hashes"], want["abundances"])
    # next-k pass into a table sized for 1000 entries
    eng.prev_from_current(2)
    eng.count_begin(5, 600)
    eng.count_add_store_next_k()
    t5 = eng.count_finalize(2)
    w5 = oracle.next_k(mins, offs, 5, want["hashes"], want["abundances"])
    assert t5.as_dict() == table_dict(w5["hashes"], w5["abundances"]) and len(w5["abundances"]) > 1500
    eng.close()


def test_min_abundance_applies_to_the_first_pass_only(built, oracle):
    """--min-abundance >= 3: the first pass drops abundance-2 k-min-mers, later passes keep every k-min-mer with a
    derived abundance > 1 (CreateMdbg.hpp:3868-3869) -- visible once the previous-k table carries refined
    abundances of 2 (patches from the contig stage)."""
    rs = synth.make_readset(2500, 9000, seed=63, n_genomes=1, genome_len_range=(250_000, 250_001))
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    first = oracle.count(sm, so, 4, 3)                            # first pass, min abundance 3
    eng.count_begin(4)
    eng.count_add_store()
    assert eng.count_finalize(3).as_dict() == table_dict(first["hashes"], first["abundances"])
    eng.prev_from_current(3)                                      # the count table itself, filtered at lookup time
    rng = np.random.default_rng(6)
    patch_idx = np.nonzero(rng.random(len(first["abundances"])) < 0.3)[0]
    patched = first["abundances"].copy()
    patched[patch_idx] = rng.integers(0, 4, size=len(patch_idx)).astype(np.uint32)     # 0, 1, 2, 3
    lohi = np.stack([first["hashes"][:, 1], first["hashes"][:, 0]], axis=1)
    eng.prev_load(lohi[patch_idx], patched[patch_idx], clear=False)
    prev_h, prev_a = first["hashes"], patched
    for k in (5, 6):
        eng.count_begin(k)
        eng.count_add_store_next_k()
        want = oracle.next_k(sm, so, k, prev_h, prev_a)
        assert (want["abundances"] == 2).sum() > 50               # entries a min-abundance-3 filter would drop
        st = eng.count_stats(3)
        assert st["n_entries"] == len(want["abundances"])
        tab = eng.count_finalize(3)
        assert tab.as_dict() == table_dict(want["hashes"], want["abundances"]), k
        eng.prev_from_current(3)
        prev_h, prev_a = want["hashes"], want["abundances"]
    eng.close()


def test_store_rewrite_ends_the_table(built):
    """Slots reference store positions: clearing or purging the store ends the current table instead of leaving
    dangling references behind (finalize then reports a state error, nothing is read out of bounds)."""
    from metamdbg_b200.engine import MdbgError
    rng = np.random.default_rng(2)
    mins = rng.integers(0, 50, size=4000, dtype=np.uint64).astype(np.uint32)
    offs = (np.arange(41) * 100).astype(np.uint64)
    eng = engine()
    eng.store_append(mins, offs)
    eng.count_begin(4)
    eng.count_add_store()
    assert len(eng.count_finalize(2).abundances) > 0
    eng.store_clear()
    with pytest.raises(MdbgError) as e:
        eng.count_finalize(2)
    assert e.value.status == 3                                    # MDBG_ERR_STATE
    eng.store_append(mins, offs)                                  # and the context is still usable
    eng.count_begin(4)
    eng.count_add_store()
    assert len(eng.count_finalize(2).abundances) > 0
    eng.close()


def test_host_batch_handed_over_packed(built, oracle, monkeypatch):
    """mdbg_host_pack_read + mdbg_sketch_batch_packed: a reader that packs while parsing hands over 2-bit words (dirty
    reads in an ASCII spill); same CSR as the ASCII entry point, in many pieces and in one; bad layouts are refused."""
    from metamdbg_b200.engine import MdbgError
    bases, offs = _edge_case_reads(np.random.default_rng(21))
    want = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    eng = engine(15, 0.005, True)
    words, src, asc = eng.host_pack_reads(bases, offs)
    assert (src >> np.uint64(63)).sum() >= 5 and len(asc) > 0          # the dirty reads travel as ASCII
    for piece_bytes in ("50000", None):
        if piece_bytes:
            monkeypatch.setenv("MDBG_PIECE_BYTES", piece_bytes)
        else:
            monkeypatch.delenv("MDBG_PIECE_BYTES")
        sk = eng.sketch_batch_packed(words, src, asc, offs, append_to_store=True)
        assert_sketch_equal(sk, *want, tag=f"packed host batch, piece bytes {piece_bytes}")
    assert eng.store_size() == (2 * (len(offs) - 1), 2 * len(want[1]))
    bad = src.copy()
    clean = np.nonzero((src >> np.uint64(63)) == 0)[0]
    bad[clean[3]], bad[clean[4]] = bad[clean[4]], bad[clean[3]]          # decreasing word offsets
    with pytest.raises(MdbgError) as e:
        eng.sketch_batch_packed(words, bad, asc, offs)
    assert e.value.status == 2
    eng.close()


# ---------------------------------------------------------------- (f)3: FASTQ / FASTA text, record split on the device

def _fastq_text(reads, crlf=False, quals=None):
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for i, s in enumerate(reads):
        q = quals[i] if quals is not None else b"I" * len(s)
        out.append(b"@read_%d some description" % i + eol + bytes(s) + eol + b"+" + eol + q + eol)
    return b"".join(out)


@pytest.mark.parametrize("hpc", [True, False])
def test_fastx_text_ingest(built, oracle, hpc):
    """mdbg_sketch_fastx: the sketch of raw FASTQ / FASTA text (records split on the device) equals the sketch of the
    reads a host parser extracts from it -- '\\n' and '\\r\\n' line ends, '@' inside quality lines, dirty reads, empty
    reads, a block cut in the middle of a record (consumed_bytes), a last line without newline."""
    from metamdbg_b200.engine import MdbgError
    bases, offs = _edge_case_reads(np.random.default_rng(31))
    reads = [bases[int(offs[r]):int(offs[r + 1])].tobytes() for r in range(len(offs) - 1)]
    want = oracle.sketch_batch(bases, offs, 15, 0.01, hpc)
    rng = np.random.default_rng(8)
    quals = [bytes(rng.integers(33, 75, len(s), dtype=np.uint8)) for s in reads]
    quals[3] = b"@" * len(reads[3]); quals[7] = b"+" * len(reads[7])         # quality lines that look like headers
    eng = engine(15, 0.01, hpc)
    for crlf in (False, True):
        text = _fastq_text(reads, crlf, quals)
        sk, info = eng.sketch_fastx(text)
        assert info == dict(n_records=len(reads), consumed_bytes=len(text), n_bases=int(offs[-1]), format="fastq")
        assert_sketch_equal(sk, *want, tag=f"fastq crlf={crlf} hpc={hpc}")
    # FASTA, 2-line, last line without a newline
    fa = b"".join(b">r%d\n" % i + s + b"\n" for i, s in enumerate(reads))[:-1]
    sk, info = eng.sketch_fastx(fa, is_final=True)
    assert info["n_records"] == len(reads) and info["format"] == "fasta" and info["consumed_bytes"] == len(fa)
    assert_sketch_equal(sk, *want, tag="fasta")
    # a block that ends inside record 40: records 0..39 are sketched, the caller continues at consumed_bytes
    text = _fastq_text(reads, False, quals)
    cut = text.index(b"@read_40 ") + 25
    sk, info = eng.sketch_fastx(text[:cut], is_final=False, append_to_store=True)
    assert info["n_records"] == 40 and text[info["consumed_bytes"]:].startswith(b"@read_40 ")
    sk2, info2 = eng.sketch_fastx(text[info["consumed_bytes"]:], is_final=True, append_to_store=True)
    assert info2["n_records"] == len(reads) - 40
    assert np.array_equal(np.concatenate([sk.minimizers, sk2.minimizers]), want[1])
    assert np.array_equal(np.concatenate([sk.positions, sk2.positions]), want[2])
    so, sm = eng.store_fetch()
    assert np.array_equal(sm, want[1]) and np.array_equal(so, want[0])
    # multi-line FASTA is not handled on the device: reported, the caller falls back to its host parser
    with pytest.raises(MdbgError) as e:
        eng.sketch_fastx(b">a\nACGT\nACGT\n>b\nAC\n")
    assert e.value.status == 2
    with pytest.raises(MdbgError):
        eng.sketch_fastx(b"ACGT\n")
    sk, info = eng.sketch_fastx(b"")
    assert info["n_records"] == 0 and len(sk.minimizers) == 0
    eng.close()


# ---------------------------------------------------------------- lookup-free next-k passes, repetitive minimizers

def test_next_k_without_lookups_equals_lookups(built, oracle, monkeypatch):
    """k >= firstK + 2 on an unpatched previous table runs without lookups (per-position values of the previous pass);
    tables are identical to the lookup form (MDBG_NO_STREAM_NEXT_K=1) and to the oracle, and a host patch of the
    previous-k table switches the next pass back to lookups."""
    from metamdbg_b200 import multi_k_sweep
    rs = synth.make_readset(2500, 9000, seed=64, n_genomes=2, genome_len_range=(150_000, 300_000), err=0.003)
    bases, offs = synth.fill_reads(rs)
    runs = {}
    for tag, env in (("stream", None), ("lookup", "1")):
        if env:
            monkeypatch.setenv("MDBG_NO_STREAM_NEXT_K", env)
        eng = engine()
        eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
        tabs = {}
        multi_k_sweep(eng, 4, 12, 0, rescue=True, on_table=lambda k, e: tabs.__setitem__(k, e.count_finalize(0).as_dict()))
        runs[tag] = tabs
        if tag == "stream":
            so, sm = eng.store_fetch()
            # patched previous table: k = 13 from the k = 12 table with some abundances lowered
            t12 = eng.count_finalize(0)
            eng.prev_from_current(0)
            idx = np.arange(0, len(t12.abundances), 7)
            eng.prev_load(t12.hashes[idx], np.ones(len(idx), np.uint32), clear=False)
            eng.count_begin(13)
            eng.count_add_store_next_k()
            got13 = eng.count_finalize(0).as_dict()
            pa = t12.abundances.copy(); pa[idx] = 1
            ph = np.stack([t12.hashes[:, 1], t12.hashes[:, 0]], axis=1)          # (h1, h2) as the oracle takes them
            want13 = oracle.next_k(sm, so, 13, ph, pa)
            assert got13 == table_dict(want13["hashes"], want13["abundances"]) and len(got13) > 100
        eng.close()
    assert runs["stream"] == runs["lookup"] and all(len(runs["stream"][k]) > 200 for k in range(4, 13))
    solid = oracle.count(sm, so, 4, 2)
    resc = oracle.rescue(sm, so, 4, solid["hashes"], solid["abundances"])
    ph = np.concatenate([solid["hashes"], resc["hashes"]]) if len(resc["hashes"]) else solid["hashes"]
    pa = np.concatenate([solid["abundances"], np.ones(len(resc["hashes"]), np.uint32)])
    for k in range(5, 13):
        nk = oracle.next_k(sm, so, k, ph, pa)
        ph, pa = nk["hashes"], nk["abundances"]
        assert runs["stream"][k] == table_dict(ph, pa), k


def test_repetitive_minimizers(built, oracle):
    """determineRepetitiveMinimizers on the device: counts of every minimizer of the stored reads, the
    max(1, int(1e-5f * distinct)) most frequent ones; installed as blacklist they disappear from later sketches."""
    rng = np.random.default_rng(17)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    repeat = acgt[rng.integers(0, 4, 400)]
    reads = []
    for i in range(300):
        s = acgt[rng.integers(0, 4, int(rng.integers(3000, 6000)))].copy()
        for _ in range(int(rng.integers(0, 4))):              # a repeat element planted a few times per read
            p = int(rng.integers(0, len(s) - 400))
            s[p:p + 400] = repeat
        reads.append(s)
    bases = np.concatenate(reads)
    offs = np.zeros(len(reads) + 1, np.uint64); offs[1:] = np.cumsum([len(r) for r in reads])
    eng = engine(15, 0.025, False)
    sk = eng.sketch_batch(bases, offs, append_to_store=True)
    vals, counts = np.unique(sk.minimizers, return_counts=True)
    for frac in (0.00001, 0.001, 0.5):
        rep = eng.repetitive_minimizers(frac)
        want_n = max(1, int(np.float32(frac) * np.float32(len(vals))))
        assert rep["n_distinct"] == len(vals) and len(rep["minimizers"]) == min(want_n, len(vals))
        order = np.lexsort((vals, -counts.astype(np.int64)))          # count descending, value ascending
        assert np.array_equal(rep["minimizers"], vals[order][:want_n].astype(np.uint32))
        assert np.array_equal(rep["counts"], counts[order][:want_n].astype(np.uint32))
        assert rep["min_count_selected"] == counts[order][want_n - 1]
        assert rep["n_with_min_count"] == int((counts == rep["min_count_selected"]).sum())
    rep = eng.repetitive_minimizers(0.001)
    eng.set_blacklist(rep["minimizers"])
    sk2 = eng.sketch_batch(bases, offs)
    want = oracle.sketch_batch(bases, offs, 15, 0.025, False, blacklist=np.sort(rep["minimizers"]))
    assert_sketch_equal(sk2, *want, tag="sketch with the device-made blacklist")
    assert len(sk2.minimizers) < len(sk.minimizers) and not np.isin(sk2.minimizers, rep["minimizers"]).any()
    eng.set_blacklist(None)
    assert_sketch_equal(eng.sketch_batch(bases, offs), sk.min_offsets, sk.minimizers, sk.positions, sk.directions, "blacklist cleared")
    eng.close()


def test_postings_index(built, oracle):
    """mdbg_count_postings: for every solid k-min-mer the (read, window) pairs of its occurrences in the stored reads
    (ReadCorrection::IndexReadsFunctor) -- list lengths are the abundances, lists hold exactly the occurrences."""
    rs = synth.make_readset(400, 6000, seed=66, n_genomes=1, genome_len_range=(60_000, 60_001), err=0.002)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    k = 4
    eng.count_begin(k)
    eng.count_add_store()
    tab = eng.count_finalize(2)
    post = eng.count_postings(2)
    assert post["k"] == k and len(post["hashes"]) == len(tab.abundances) and len(post["reads"]) == int(tab.abundances.sum())
    want_ab = {(int(h[0]), int(h[1])): int(a) for h, a in zip(tab.hashes, tab.abundances)}
    # occurrences from the oracle's window enumeration: hash128 of every normalized window of every read
    occ = {}
    for r in range(rs.n_reads):
        m = sm[int(so[r]):int(so[r + 1])]
        if len(m) < k:
            continue
        vecs = oracle.kminmers(m, k)[0]
        for i, v in enumerate(vecs):
            h1, h2 = oracle.hash128(np.ascontiguousarray(v, np.uint32))
            occ.setdefault((h2, h1), set()).add((r, i))
    got_keys = set()
    for j, h in enumerate(post["hashes"]):
        key = (int(h[0]), int(h[1]))
        lo, hi = int(post["offsets"][j]), int(post["offsets"][j + 1])
        assert hi - lo == want_ab[key]
        pairs = set(zip(post["reads"][lo:hi].tolist(), post["windows"][lo:hi].tolist()))
        assert len(pairs) == hi - lo and pairs == occ[key], key
        got_keys.add(key)
    assert got_keys == set(want_ab) and len(got_keys) > 200
    eng.close()


def test_multi_k_run_in_the_library(built, oracle):
    """mdbg_multi_k_run = the host loop of multik.py inside the library: same per-k statistics, last table current."""
    from metamdbg_b200 import multi_k_sweep
    rs = synth.make_readset(2000, 8000, seed=67, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.003)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    want = multi_k_sweep(eng, 4, 10, 0, rescue=True)
    got = eng.multi_k_run(4, 10, 0, rescue=True)
    assert [(g["k"], g["n_entries"], g["checksum"]) for g in got] == [(w["k"], w["n_entries"], w["checksum"]) for w in want]
    assert got[0]["n_reads_rescued"] == want[0]["n_reads_rescued"] > 0
    assert eng.count_stats(0)["checksum"] == got[-1]["checksum"]          # the k = 10 table is current
    eng.close()


def test_a_loop_over_k_stops_allocating_after_its_first_sweep(built, oracle):
    """mdbg_prev_from_current swaps the buffers of the count table and the previous-k table with every k, so either
    buffer has to hold the largest table sooner or later.  A buffer of the pair is therefore never allocated smaller
    than its sibling (ensure_table_buf), and when the current buffer is still too small while the previous-k table sits
    in a large one, the previous table's live slots move and the buffers trade places (table_reset) -- a second sweep
    must not allocate device memory (a cudaMalloc / cudaFree pair of a table's size cost 5 - 70 ms on the B200 boxes),
    and the previous-k table must survive whichever of the two happened: every table equals the oracle's."""
    from metamdbg_b200 import multi_k_sweep
    rs = synth.make_readset(1500, 8000, seed=71, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.002)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    so, sm = eng.store_fetch()
    first = multi_k_sweep(eng, 4, 9)
    allocs = eng.allocations()[0]
    again = multi_k_sweep(eng, 4, 9)
    assert eng.allocations()[0] == allocs                                    # nothing (re)allocated by the second sweep
    assert [r["checksum"] for r in again] == [r["checksum"] for r in first]
    want = {}
    prev = oracle.count(sm, so, 4, 2)
    ph, pa = prev["hashes"], prev["abundances"]
    for k in range(5, 10):
        nk = oracle.next_k(sm, so, k, ph, pa)
        ph, pa = nk["hashes"], nk["abundances"]
        want[k] = table_dict(ph, pa)
    assert eng.count_finalize(2).as_dict() == want[9]
    # small buffers first, then a table that outgrows them while the previous-k table is small: whatever the library does
    # about the buffers, the previous table's content must still be there for the pass
    eng2 = engine()
    eng2.store_append(sm, so)
    eng2.count_begin(4, 1000)                        # a small table, grown on demand
    eng2.count_add_store()
    eng2.prev_from_current(2)
    eng2.count_begin(5, 1000)
    eng2.count_add_store_next_k()
    assert eng2.count_finalize(2).as_dict() == want[5]
    eng2.prev_from_current(2)                        # previous = k = 5
    eng2.count_begin(6, 200_000)                     # far more than either buffer holds
    eng2.count_add_store_next_k()
    assert eng2.count_finalize(2).as_dict() == want[6]
    eng2.prev_from_current(2)                        # previous = k = 6, a small part of the large buffer
    eng2.count_begin(7, 150_000)                     # more than the small buffer held
    eng2.count_add_store_next_k()
    assert eng2.count_finalize(2).as_dict() == want[7] and len(want[7]) > 300
    eng2.prev_from_current(2)
    eng2.count_begin(8, 1000)
    eng2.count_add_store_next_k()
    assert eng2.count_finalize(2).as_dict() == want[8]
    eng2.close()
    eng.close()


def test_unitig_nodes_vs_oracle_and_golden(built, oracle):
    """Row F1, third step (mdbg_unitigs_build): unitig links between oriented nodes, list ranking by pointer jumping,
    cycle cut at the smallest hash128, normalized sequences and their hash128 -- the records of unitigGraph.nodes.bin
    in the reference's deterministic order, against vectors minted from the reference's own computeUnitigNodes +
    computeDeterministicUnitigs (clean paths, small alphabets with palindromic keys / hairpins, circular genomes) and,
    on the node set of sketched reads (first pass, rescued nodes, a next-k table), against the oracle."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "minspace_unitigs.npz"))
    eng = engine()
    n_circ = 0
    for i in range(int(g["n_cases"])):
        k = int(g[f"c{i}_k"])
        eng.store_clear()
        eng.store_append(g[f"c{i}_minimizers"], g[f"c{i}_offsets"])
        eng.count_begin(k)
        eng.count_add_store()
        got = eng.unitig_records(2)
        assert got["n_nodes"] == len(g[f"c{i}_nodes"])
        assert np.array_equal(got["offsets"], g[f"c{i}_unitig_offsets"]), i
        assert np.array_equal(got["minimizers"], g[f"c{i}_unitig_minimizers"]), i
        # unitig graph edges: the records of unitigGraph.edges.successors.bin as one reference thread writes them
        assert np.array_equal(got["edge_offsets"], g[f"c{i}_edge_offsets"]), i
        assert np.array_equal(got["edge_targets"], g[f"c{i}_edge_targets"]), i
        assert [got["n_unitig_edges"], got["checksum_edges"]] == [int(x) for x in g[f"c{i}_edge_stats"]], i
        n_circ += got["n_circular"]
    assert n_circ >= 4
    eng.close()

    rs = synth.make_readset(1500, 7000, seed=31, n_genomes=2, genome_len_range=(120_000, 200_000), err=0.003)
    bases, offs = synth.fill_reads(rs)
    eng = engine()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)

    def check(k, min_ab, at_least):
        tab = eng.count_finalize(min_ab)
        want = oracle.unitigs(tab.kminmers, k)
        got = eng.unitigs_build(min_ab)
        assert got["n_nodes"] == len(tab.abundances) and got["n_unitigs"] == len(want["offsets"]) - 1 >= at_least
        rec = eng.unitig_records(min_ab)
        assert np.array_equal(rec["offsets"], want["offsets"]) and np.array_equal(rec["minimizers"], want["minimizers"])
        # hashes as the oracle's, order = ascending (high, low); every node lies in a unitig (a self-reverse-complementary
        # path holds its nodes in both orientations, so the windows can outnumber the nodes)
        hs = got["hashes"][got["order"]]
        assert np.array_equal(hs[:, 1], want["hashes"][:, 0]) and np.array_equal(hs[:, 0], want["hashes"][:, 1])
        # dumpUnitigAbundances: the table's abundance of every k-min-mer of every unitig; the two checksums the reference logs
        ab_of = tab.as_dict()
        offs_r, mins_r = rec["offsets"], rec["minimizers"]
        want_ab, cs_nodes, cs_ab = [], 0, 0
        for i in range(len(offs_r) - 1):
            seq = mins_r[int(offs_r[i]):int(offs_r[i + 1])]
            vecs, _ = oracle.kminmers(seq, k)
            a = [ab_of[oracle.hash128(v)] for v in vecs]
            want_ab += a
            cs_nodes = (cs_nodes + int(seq.astype(np.uint64).sum()) * len(seq) * (2 * i)) % 2 ** 64
            cs_ab = (cs_ab + sum(a) * len(a)) % 2 ** 64
        assert np.array_equal(rec["abundances"], np.array(want_ab, np.uint32))
        we = oracle.unitig_edges(want["offsets"], want["minimizers"], k)
        assert np.array_equal(rec["edge_offsets"], we["offsets"]) and np.array_equal(rec["edge_targets"], we["targets"])
        assert (rec["n_unitig_edges"], rec["checksum_edges"]) == (we["n_edges"], we["checksum"])
        assert rec["checksum_nodes"] == cs_nodes and rec["checksum_abundances"] == cs_ab
        assert int(np.sum(np.diff(got["offsets"]).astype(np.int64) - (k - 1))) >= got["n_nodes"]

    eng.count_begin(4)
    eng.count_add_store()
    check(4, 2, 20)
    eng.count_rescue()
    check(4, 0, 20)
    eng.prev_from_current(0)
    eng.count_begin(5)
    eng.count_add_store_next_k()
    check(5, 0, 20)
    for k in (2, 3, 9):
        eng.count_begin(k)
        eng.count_add_store()
        check(k, 2, 1)
    # an empty node set
    eng.store_clear()
    eng.count_begin(4)
    eng.count_add_store()
    e = eng.unitigs_build(2)
    assert e["n_unitigs"] == 0 and e["n_nodes"] == 0 and len(e["offsets"]) == 1
    eng.close()
