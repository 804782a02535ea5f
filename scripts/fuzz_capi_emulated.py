#!/usr/bin/env python
"""Randomised scenarios through the WHOLE C ABI on the CPU emulator (tests/_emu.py builds libmdbg_b200_emu.so from
the product sources), compared with the oracle: random batches (ragged, dirty, homopolymer, tandem reads), random
l / density / HPC / blacklist / transfer mode, several batches appended to the store, device-pointer entry points
(host pointers are device pointers in the emulator), side outputs, purge, density re-threshold, count at random k
in several read ranges, rescue, next-k.  CPU only, no GPU involved; the emulated library is test infrastructure.

    python scripts/fuzz_capi_emulated.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rand_read(rng, max_len):
    kind = int(rng.integers(0, 7))
    n = int(rng.integers(0, max_len))
    acgt = np.frombuffer(b"ACGT", np.uint8)
    if kind == 0 or n == 0:
        s = rng.choice(acgt, n)
    elif kind == 1:
        s = rng.choice(np.frombuffer(b"ACGTNacgtn#RY", np.uint8), n)
    elif kind == 2:
        runs = rng.geometric(0.35, max(1, n // 2))
        s = np.repeat(rng.choice(acgt, len(runs)), runs)[:n]
    elif kind == 3:
        unit = rng.choice(acgt, int(rng.integers(1, 9)))
        s = np.tile(unit, n // len(unit) + 1)[:n]
    elif kind == 4:
        s = rng.choice(acgt, n)
        s[rng.integers(0, n, max(1, n // 60))] = ord("N")
    elif kind == 5:
        s = rng.integers(1, 256, n).astype(np.uint8)
    else:
        s = rng.choice(acgt, int(rng.integers(0, 40)))
    return s.astype(np.uint8)


def make_batch(rng, n_reads, max_len):
    reads = [rand_read(rng, max_len) for _ in range(n_reads)]
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    bases = np.concatenate(reads).astype(np.uint8) if offs[-1] else np.zeros(0, np.uint8)
    return bases, offs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--deferred", action="store_true", help="adversarial stream execution (-DEMU_DEFERRED)")
    a = ap.parse_args()
    if a.deferred:
        os.environ["MDBG_EMU_EXTRA_FLAGS"] = (os.environ.get("MDBG_EMU_EXTRA_FLAGS", "") + " -DEMU_DEFERRED").strip()
    import _emu
    from metamdbg_b200 import _capi
    tmp = tempfile.mkdtemp(prefix="mdbg_emu_")
    _capi.LIB_PATH = _emu.build_emulated_library(tmp)
    from metamdbg_b200 import Engine
    from oracle import pyoracle
    orc = pyoracle.Oracle()
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n = dict(scenarios=0, minimizers=0, table_entries=0)
    while time.time() < t_end:
        l = int(rng.choice([15, 15, 15, 11, 16, 7, 13]))
        dens = float(rng.choice([0.005, 0.005, 0.05, 0.0025, 0.3, 0.7]))
        hpc = bool(rng.integers(0, 2))
        bl = None
        # host-batch plumbing knobs: piece size (many pieces even for small batches), piece pipeline on/off, packed
        # transfer for small batches too, and the arithmetic variant of the sketch kernel
        for key, choices in (("MDBG_PIECE_BYTES", [None, "3000", "20000", "200000"]), ("MDBG_PIECE_PIPELINE", [None, "0", "1"]),
                             ("MDBG_PACK_MIN_BYTES", [None, "0", "5000"])):
            v = choices[int(rng.integers(0, len(choices)))]
            if v is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = v
        eng = Engine(l, dens, hpc)
        eng.set_host_packing(int(rng.choice([-1, 0, 1])))
        eng.set_sketch_variant(int(rng.integers(0, 2)))
        all_m, all_off = [], [0]
        scen = (l, dens, hpc, {k: os.environ.get(k) for k in ("MDBG_PIECE_BYTES", "MDBG_PIECE_PIPELINE", "MDBG_PACK_MIN_BYTES")},
                "variant", eng.sketch_variant)
        for b in range(int(rng.integers(1, 4))):
            big = rng.integers(0, 6) == 0
            bases, offs = make_batch(rng, int(rng.integers(1, 60)), 150_000 if big else 4000)
            want = orc.sketch_batch(bases, offs, l, dens, hpc)
            mode = int(rng.integers(0, 3))
            if mode == 0:
                sk = eng.sketch_batch(bases, offs, append_to_store=True)
            elif mode == 1:                                   # device-pointer entry (16-byte aligned buffer + slack)
                buf = np.zeros(len(bases) + 64 + 16, np.uint8)
                sh = (-buf.ctypes.data) % 16
                buf[sh:sh + len(bases)] = bases
                if len(offs) > 1 and rng.integers(0, 4) == 0:  # autotune on this batch first (appends nothing)
                    tune = eng.autotune_sketch(buf.ctypes.data + sh, offs.ctypes.data, len(offs) - 1, len(bases))
                    assert all(tune["identical"]), ("autotune", scen, tune)
                    eng.set_sketch_variant(int(rng.integers(0, 2)))
                    scen = scen + ("autotuned, variant now", eng.sketch_variant)
                eng.sketch_batch_device(buf.ctypes.data + sh, offs.ctypes.data, len(offs) - 1, len(bases), True)
                sk = eng.sketch_fetch()
            else:                                             # side outputs, with or without qualities
                if hpc and (bases == ord("#")).any():         # refused by design (EncoderRLE's sentinel)
                    try:
                        eng.sketch_batch_q(bases, None, offs)
                        raise AssertionError(("'#' accepted", scen))
                    except Exception as e:
                        assert "'#'" in str(e), e
                    bases = np.where(bases == ord("#"), ord("N"), bases).astype(np.uint8)
                    want = orc.sketch_batch(bases, offs, l, dens, hpc)
                q = None
                if rng.integers(0, 2):
                    q = rng.integers(33, 90, len(bases)).astype(np.uint8)
                eng.set_read_filters(False)
                sk, aux = eng.sketch_batch_q(bases, q, offs, append_to_store=True)
                raw = bases.tobytes()
                for r in rng.choice(len(offs) - 1, min(6, len(offs) - 1), replace=False):
                    lo, hi = int(offs[r]), int(offs[r + 1])
                    if b"N" in raw[lo:hi] or any(c not in b"ACGT" for c in raw[lo:hi][:0]):
                        pass
                    pos = want[2][int(want[0][r]):int(want[0][r + 1])]
                    mq, cx, ql = orc.read_aux(raw[lo:hi], q[lo:hi].tobytes() if q is not None else b"", l, hpc, pos)
                    got_q = aux["qualities"][int(sk.min_offsets[r]):int(sk.min_offsets[r + 1])]
                    assert np.array_equal(got_q, ql), ("min qualities", scen, r)
                    if q is not None and hi > lo:
                        assert np.float32(aux["mean_quality"][r]).tobytes() == np.float32(mq).tobytes(), ("mean q", scen, r)
                    c_got = aux["complexity"][r]
                    assert (np.isnan(c_got) and np.isnan(cx)) or c_got == cx, ("complexity", scen, r, c_got, cx)
            for x, y, nm in zip((sk.min_offsets, sk.minimizers, sk.positions, sk.directions), want, ("offsets", "minimizers", "positions", "directions")):
                if not np.array_equal(x, y):
                    np.savez(os.path.join(tempfile.gettempdir(), "fuzz_capi_failure.npz"), bases=bases, offs=offs)
                assert np.array_equal(x, y), ("sketch", scen, b, mode, nm, len(x), len(y))
            n["minimizers"] += len(want[1])
            all_m.append(want[1])
            for r in range(len(offs) - 1):
                all_off.append(all_off[-1] + int(want[0][r + 1] - want[0][r]))
        mins = np.concatenate(all_m).astype(np.uint32) if all_m else np.zeros(0, np.uint32)
        moff = np.array(all_off, np.uint64)
        so, sm = eng.store_fetch()
        assert np.array_equal(so, moff) and np.array_equal(sm, mins), ("store", scen)
        # ONT-style density re-threshold on the store (Utils::applyDensityThreshold), sometimes
        if rng.integers(0, 3) == 0 and len(mins):
            d2 = float(rng.choice([dens / 5, dens / 2, dens]))
            eng.store_apply_density(d2)
            pm, po = [], [0]
            for r in range(len(moff) - 1):
                qv = orc.apply_density(mins[int(moff[r]):int(moff[r + 1])], d2)
                pm.append(qv); po.append(po[-1] + len(qv))
            mins = np.concatenate(pm).astype(np.uint32) if pm else np.zeros(0, np.uint32)
            moff = np.array(po, np.uint64)
            so, sm = eng.store_fetch()
            assert np.array_equal(so, moff) and np.array_equal(sm, mins), ("apply_density", scen, d2)
        # purge (sometimes), then count / rescue / next-k at a random k
        if rng.integers(0, 2):
            lk = int(rng.integers(5, 40))
            eng.purge_palindromes(4, lk)
            pm, po = [], [0]
            for r in range(len(moff) - 1):
                qv, _ = orc.purge_palindrome(mins[int(moff[r]):int(moff[r + 1])], 4, lk)
                pm.append(qv); po.append(po[-1] + len(qv))
            mins = np.concatenate(pm).astype(np.uint32) if pm else np.zeros(0, np.uint32)
            moff = np.array(po, np.uint64)
            so, sm = eng.store_fetch()
            assert np.array_equal(so, moff) and np.array_equal(sm, mins), ("purge", scen, lk)
        k = int(rng.choice([4, 4, 2, 3, 5, 9, 21]))
        min_ab = int(rng.choice([0, 2, 3]))
        eng.count_begin(k, 0)
        cut = int(rng.integers(0, len(moff)))
        eng.count_add_store(0, cut)
        eng.count_add_store(cut, len(moff) - 1)
        ref = orc.count(mins, moff, k, min_ab)
        tab = eng.count_finalize(min_ab)
        want_t = {(int(h[0]), int(h[1])): int(x) for h, x in zip(ref["hashes"], ref["abundances"])}
        assert tab.as_dict() == want_t, ("count", scen, k, min_ab, len(want_t))
        assert (tab.n_instances, tab.n_distinct) == (ref["n_instances"], ref["n_distinct"]), ("count totals", scen, k)
        n["table_entries"] += len(want_t)
        if k >= 2 and rng.integers(0, 2):                     # edge keys of the node set (CreateMdbg::EdgeIndexer)
            ed = eng.edges_index(min_ab)
            we = orc.edge_index(tab.kminmers, k)
            assert ed["n_edges"] == len(we["hashes"]) and ed["checksum"] == we["checksum"], ("edges", scen, k)
            assert {(int(h[1]), int(h[0])) for h in ed["hashes"]} == {(int(h[0]), int(h[1])) for h in we["hashes"]}, ("edge set", scen, k)
            wv = orc.edge_values(tab.kminmers, k)
            assert {(int(h[1]), int(h[0])): v.tolist() for h, v in zip(ed["hashes"], ed["values"])} == \
                {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])}, ("edge values", scen, k)
        if rng.integers(0, 2):                                # default mode: rescue on top of the >= 2 table
            solid = orc.count(mins, moff, k, 2)
            resc = orc.rescue(mins, moff, k, solid["hashes"], solid["abundances"])
            nr = eng.count_rescue()
            t2 = eng.count_finalize(0)
            want_r = {(int(h[0]), int(h[1])): int(x) for h, x in zip(solid["hashes"], solid["abundances"])}
            for h in resc["hashes"]:
                want_r[(int(h[0]), int(h[1]))] = 1
            assert t2.as_dict() == want_r and nr == resc["n_reads_rescued"], ("rescue", scen, k)
        elif k < 21 and len(want_t):
            solid = orc.count(mins, moff, k, 2)
            eng.prev_from_current(2)
            eng.count_begin(k + 1, 0)
            eng.count_add_store_next_k()
            nk = orc.next_k(mins, moff, k + 1, solid["hashes"], solid["abundances"])
            t3 = eng.count_finalize(0)
            want_n = {(int(h[0]), int(h[1])): int(x) for h, x in zip(nk["hashes"], nk["abundances"])}
            assert t3.as_dict() == want_n, ("next_k", scen, k + 1)
        eng.close()
        n["scenarios"] += 1
    print("no difference:", n)


if __name__ == "__main__":
    main()
