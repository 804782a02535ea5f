#!/usr/bin/env python
"""Randomised comparison of the C restatement (oracle/mdbg_oracle.c) with the reference's own code compiled into
oracle/_ref/libmdbg_ref.so, on adversarial inputs the unit tests only sample: arbitrary bytes (0..255), tandem
repeats and homopolymer runs, reads around the l-mer length, every l in 2..16 (hash path) and k in 2..40.

    python scripts/fuzz_oracle_vs_ref.py [--seconds 120] [--seed 1]

Needs /root/reference (to build oracle/_ref); CPU only.  Exit code 0 = no difference found."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402


def rand_read(rng, max_len):
    kind = rng.integers(0, 8)
    n = int(rng.integers(0, max_len))
    if kind == 0:                                           # plain ACGT
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    elif kind == 1:                                         # IUPAC / lower case / sentinel characters mixed in
        s = rng.choice(np.frombuffer(b"ACGTNacgtn#RYKM-*", np.uint8), n)
    elif kind == 2:                                         # arbitrary bytes, 0 excluded (C strings upstream)
        s = rng.integers(1, 256, n).astype(np.uint8)
    elif kind == 3:                                         # homopolymer runs with geometric lengths
        runs = rng.geometric(0.3, max(1, n // 3))
        s = np.repeat(rng.choice(np.frombuffer(b"ACGT", np.uint8), len(runs)), runs)[:n]
    elif kind == 4:                                         # tandem repeat of a short unit
        unit = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(1, 9)))
        s = np.tile(unit, n // len(unit) + 1)[:n]
    elif kind == 5:                                         # palindromic (reverse-complement symmetric) sequence
        h = rng.choice(np.frombuffer(b"ACGT", np.uint8), n // 2)
        comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
        s = np.concatenate([h, comp[h[::-1]]])
    elif kind == 6:                                         # mostly ACGT with sparse N
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
        if n:
            s[rng.integers(0, n, max(1, n // 50))] = ord("N")
    else:                                                   # very short
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(0, 40)))
    return s.astype(np.uint8).tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    orc, ref = pyoracle.Oracle(), pyoracle.Reference()
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n_cases = {"sketch": 0, "minspace": 0, "graph": 0}
    densities = [0.005, 0.0025, 0.01, 0.05, 0.3, 0.77, 1e-4]
    while time.time() < t_end:
        # ---- base space: hpc, l-mers, sketch ------------------------------------------------------------
        for _ in range(40):
            seq = rand_read(rng, 1500)
            hpc = bool(rng.integers(0, 2))
            x, px = orc.hpc(seq, hpc)
            y, py = ref.hpc(seq, hpc)
            assert x == y and np.array_equal(px, py), ("hpc", seq, hpc)
            l = int(rng.integers(2, 17))
            va, da = orc.lmers(x, l)
            vb, db = ref.lmers(x, l)
            assert np.array_equal(va, vb) and np.array_equal(da, db), ("lmers", seq, l)
            d = float(densities[rng.integers(0, len(densities))])
            bl = None
            if rng.integers(0, 3) == 0:
                m0 = orc.sketch_read(seq, l, d, hpc)[0]
                if len(m0):
                    bl = np.unique(m0[rng.integers(0, len(m0), max(1, len(m0) // 4))])
            ra = orc.sketch_read(seq, l, d, hpc, bl)
            rb = ref.sketch_read(seq, l, d, hpc, bl)
            for u, v in zip(ra, rb):
                assert np.array_equal(u, v), ("sketch", seq, l, d, hpc, bl)
            n_cases["sketch"] += 1
        # ---- minimizer space: k-min-mers, purge, density re-threshold, count ------------------------------
        alpha = int(rng.choice([2, 3, 5, 17, 1 << 20]))
        reads = [rng.integers(0, alpha, size=int(rng.integers(0, 70))).astype(np.uint32) for _ in range(120)]
        if rng.integers(0, 2):                              # duplicate some reads (and their reverses) -> abundance
            reads += [r[::-1].copy() if rng.integers(0, 2) else r.copy() for r in reads[: len(reads) // 2]]
        for m in reads[:40]:
            k = int(rng.integers(2, 41))
            va, fa = orc.kminmers(m, k)
            vb, fb = ref.kminmers(m, k)
            assert np.array_equal(va, vb) and np.array_equal(fa, fb), ("kminmers", m, k)
            lk = int(rng.integers(5, 40))
            pa, _ = orc.purge_palindrome(m, 4, lk)
            pb = ref.purge_palindrome(m, 4, lk)
            assert np.array_equal(pa, pb), ("purge", m, lk)
            big = rng.integers(0, 1 << 32, size=len(m) + 5, dtype=np.uint64).astype(np.uint32)
            d = float(densities[rng.integers(0, len(densities))])
            assert np.array_equal(orc.apply_density(big, d), ref.apply_density(big, d)), ("density", big, d)
            n_cases["minspace"] += 1
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(r) for r in reads])
        mins = np.concatenate(reads).astype(np.uint32) if reads else np.zeros(0, np.uint32)
        k = int(rng.choice([2, 3, 4, 5, 7, 12, 21, 33]))
        for min_ab in (0, 2, 3):
            x = orc.count(mins, offs, k, min_ab)
            y = ref.count(mins, offs, k, min_ab, threads=2)
            for key in ("vecs", "hashes", "abundances"):
                assert np.array_equal(x[key], y[key]), ("count", k, min_ab, key)
            assert x["n_instances"] == y["n_instances"] and x["n_distinct"] == y["n_distinct"]
        # ---- the real graph stages: counter + rescue, next k ------------------------------------------
        # upstream quirk, not reproduced: KminmerCounter::dereplicatePartition (CreateMdbg.hpp:3839-3842) pushes
        # {default KmerVec, uninitialised abundance} for an EMPTY partition file, so inputs with fewer distinct
        # k-min-mers than partitions (= threads here) make the reference emit a garbage record
        solid = orc.count(mins, offs, k, 2)
        if solid["n_instances"] == 0:
            continue
        gthreads = 2 if solid["n_distinct"] >= 64 else 1
        g = ref.graph_firstpass(mins, offs, k, min_abundance=0, threads=gthreads)
        resc = orc.rescue(mins, offs, k, solid["hashes"], solid["abundances"])
        got = {(int(h[0]), int(h[1])): int(ab) for h, ab in zip(solid["hashes"].reshape(-1, 2), solid["abundances"])}
        for h in resc["hashes"].reshape(-1, 2):
            got[(int(h[0]), int(h[1]))] = 1
        want = {(int(h[0]), int(h[1])): int(ab) for h, ab in zip(g["hashes"].reshape(-1, 2), g["abundances"])}
        assert got == want, ("firstpass+rescue", k, len(got), len(want))
        if k < 33:
            prev_h, prev_ab = g["hashes"], g["abundances"]
            nx = orc.next_k(mins, offs, k + 1, prev_h, prev_ab)
            ny = ref.graph_next_k(mins, offs, k + 1, prev_h, prev_ab, use_counter=False, threads=gthreads)
            a1 = {(int(h[0]), int(h[1])): int(ab) for h, ab in zip(nx["hashes"].reshape(-1, 2), nx["abundances"])}
            b1 = {(int(h[0]), int(h[1])): int(ab) for h, ab in zip(ny["hashes"].reshape(-1, 2), ny["abundances"])}
            assert a1 == b1, ("next_k", k + 1, len(a1), len(b1))
        n_cases["graph"] += 1
    print("no difference:", n_cases)


if __name__ == "__main__":
    main()
