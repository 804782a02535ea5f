#!/usr/bin/env python
"""Torch-free GPU self-test of the newest host-side paths (seconds, not minutes: `import torch` alone costs a minute
on a fresh box).  Everything goes through the C ABI (ctypes) and is compared with the CPU oracle:

  pieces     host batches cut into ~50 pieces (MDBG_PIECE_BYTES): per-piece scan / compaction / D2H, packed and
             ASCII transfer, minimizer-rich reads (buffers grow mid-batch), a slot overflow (exact re-sketch)
  ranks N    N ranks = N threads on ONE GPU over tests/cpp/fake_nccl.cpp built with -DFAKE_NCCL_CUDA: owner merge,
             multi-rank rescue, replicated previous-k table and two next-k passes (tests/emu_multirank_child.py)

  variants   sketch variant 1 in many small launches, reads ending in '#', edge keys (late round-1 kernel changes)

usage: gpu_selftest.py pieces | variants | ranks N         (prints one JSON line, exit code 0 = identical to the oracle)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pieces():
    os.environ["MDBG_PIECE_BYTES"] = "40000"
    from metamdbg_b200 import Engine, synth
    from oracle.pyoracle import Oracle
    orc = Oracle()
    rs = synth.make_readset(500, 4200, seed=91, n_genomes=2, genome_len_range=(150_000, 250_000))
    bases, offs = synth.fill_reads(rs)
    want = orc.sketch_batch(bases, offs, 15, 0.005, True)
    out = {}

    def same(sk, w):
        return all(np.array_equal(a, b) for a, b in zip((sk.min_offsets, sk.minimizers, sk.positions, sk.directions), w))

    for packing in (1, 0):
        eng = Engine(15, 0.005, True)
        eng.set_host_packing(packing)
        ok = same(eng.sketch_batch(bases, offs, append_to_store=True), want)
        info = eng.last_batch_info()
        ok &= info["n_pieces"] >= 40 and info["n_pieces_pipelined"] == info["n_pieces"] and info["packed"] == bool(packing)
        ok &= bool(np.array_equal(eng.store_fetch()[1], want[1]))
        eng.store_clear()
        eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
        ok &= same(eng.sketch_fetch(), want)
        out[f"plain_packing{packing}"] = bool(ok)
        eng.close()
    # minimizer-rich reads (tandem repeats of a selected l-mer): growth with copies in flight, then one overflowing read
    rng = np.random.default_rng(4)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 20000)]
    p = orc.sketch_batch(seq, np.array([0, len(seq)], np.uint64), 15, 0.005, False)[2]
    unit = np.concatenate([seq[int(p[1]):int(p[1]) + 15], np.frombuffer(b"ACGTA", np.uint8)])
    reads = []
    for r in range(260):
        rnd = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 4200)]
        reads.append(np.concatenate([rnd[:2100], np.tile(unit, 90), rnd[2100:]]))
    for with_overflow in (False, True):
        rr = list(reads)
        if with_overflow:
            rr[237] = np.tile(unit, 300)
        bb = np.concatenate(rr)
        oo = np.zeros(len(rr) + 1, np.uint64)
        oo[1:] = np.cumsum([len(x) for x in rr])
        w3 = orc.sketch_batch(bb, oo, 15, 0.005, False)
        eng = Engine(15, 0.005, False)
        eng.set_host_packing(1)
        ok = same(eng.sketch_batch(bb, oo, append_to_store=True), w3)
        info = eng.last_batch_info()
        ok &= info["overflow_fallback"] == with_overflow and bool(np.array_equal(eng.store_fetch()[1], w3[1]))
        if not with_overflow:
            ok &= info["n_buffer_growths"] >= 1 and info["n_pieces_pipelined"] == info["n_pieces"]
        out[f"rich_overflow{int(with_overflow)}"] = bool(ok)
        out[f"rich_overflow{int(with_overflow)}_info"] = {k: info[k] for k in ("n_pieces", "n_pieces_pipelined", "n_buffer_growths")}
        eng.close()
    return out


def variants():
    """Sketch variant 1 in MANY small launches (the condition under which its first version went wrong), reads that
    end in EncoderRLE's '#' sentinel, and the edge keys of a count table -- the three kernel changes made after the
    last GPU run of round 1."""
    os.environ["MDBG_PIECE_BYTES"] = "20000"
    os.environ["MDBG_PACK_MIN_BYTES"] = "0"
    from metamdbg_b200 import Engine, synth
    from oracle.pyoracle import Oracle
    orc = Oracle()
    out = {}
    rs = synth.make_readset(900, 4000, seed=5, n_genomes=2, genome_len_range=(150_000, 250_000))
    bases, offs = synth.fill_reads(rs)
    want = orc.sketch_batch(bases, offs, 15, 0.005, True)

    def same(sk, w):
        return all(np.array_equal(a, b) for a, b in zip((sk.min_offsets, sk.minimizers, sk.positions, sk.directions), w))

    for v in (1, 0):
        for packing in (1, 0):
            eng = Engine(15, 0.005, True)
            eng.set_sketch_variant(v)
            eng.set_host_packing(packing)
            ok = all(same(eng.sketch_batch(bases, offs), want) for _ in range(3))      # repeated: fresh CTAs every time
            out[f"variant{v}_packing{packing}_{eng.last_batch_info()['n_pieces']}pieces"] = bool(ok)
            eng.close()
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = [np.concatenate([rng.choice(acgt, int(rng.integers(0, 1800))), np.frombuffer((b"#", b"##", b"", b"N#", b"#A")[i % 5], np.uint8)])
             for i in range(300)]
    o2 = np.zeros(len(reads) + 1, np.uint64)
    o2[1:] = np.cumsum([len(r) for r in reads])
    b2 = np.concatenate(reads).astype(np.uint8)
    for l, d in ((15, 0.3), (11, 0.3), (15, 0.005)):
        eng = Engine(l, d, True)
        out[f"sentinel_l{l}_d{d}"] = bool(same(eng.sketch_batch(b2, o2), orc.sketch_batch(b2, o2, l, d, True)))
        eng.close()
    eng = Engine(15, 0.005, True)
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    for k, rescue in ((4, False), (4, True), (6, False)):
        eng.count_begin(k)
        eng.count_add_store()
        if rescue:
            eng.count_rescue()
        tab = eng.count_finalize(0 if rescue else 2)
        got = eng.edges_index(0 if rescue else 2)
        we = orc.edge_index(tab.kminmers, k)
        out[f"edges_k{k}_rescue{int(rescue)}"] = bool(
            got["n_edges"] == len(we["hashes"]) and got["checksum"] == we["checksum"] and
            {(int(h[1]), int(h[0])) for h in got["hashes"]} == {(int(h[0]), int(h[1])) for h in we["hashes"]})
        wv = orc.edge_values(tab.kminmers, k)
        out[f"edge_values_k{k}_rescue{int(rescue)}"] = bool(
            {(int(h[1]), int(h[0])): v.tolist() for h, v in zip(got["hashes"], got["values"])} ==
            {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])})
    eng.close()
    return out


def ranks(n):
    os.environ.setdefault("MDBG_EMU_LIB", os.path.join(ROOT, "metamdbg_b200", "libmdbg_b200.so"))     # the REAL library
    sys.argv = ["emu_multirank_child.py", str(n), "4"]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import io
    from contextlib import redirect_stdout
    import emu_multirank_child as child
    buf = io.StringIO()
    with redirect_stdout(buf):
        child.main()
    lines = buf.getvalue().strip().splitlines()
    return {"ok": lines[-1] == "OK", "log": lines}


if __name__ == "__main__":
    t0 = time.time()
    what = sys.argv[1]
    try:
        res = pieces() if what == "pieces" else variants() if what == "variants" else ranks(int(sys.argv[2]))
        ok = all(v for k, v in res.items() if isinstance(v, bool))
    except Exception as e:  # noqa: BLE001
        res, ok = {"error": repr(e)}, False
    print(json.dumps({"test": " ".join(sys.argv[1:]), "ok": ok, "seconds": round(time.time() - t0, 2), **res}), flush=True)
    sys.exit(0 if ok else 1)
