#!/usr/bin/env python
"""Randomised MULTI-RANK scenarios on the CPU emulator: N ranks = N threads, each with its own emulated context, over
the in-process fake NCCL (tests/cpp/fake_nccl.cpp).  Random rank count, uneven and EMPTY shards (fewer reads than
ranks), random k / density, count + owner merge, optional rescue, a next-k chain through the replicated previous-k
table -- the union of the ranks' tables must equal the oracle's table of the whole read set and every key must sit
on its owner.  CPU only; the emulated library and the fake NCCL are test infrastructure.

    python scripts/fuzz_multirank_emulated.py [--seconds 300] [--seed 1]
"""
import argparse
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    if not os.environ.get("MDBG_FUZZ_CHILD"):          # the fake libnccl.so.2 must be on LD_LIBRARY_PATH at start-up
        import _emu
        tmp = tempfile.mkdtemp(prefix="mdbg_emu_mr_")
        lib = _emu.build_emulated_library(tmp)
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", os.path.join(tmp, "libnccl.so.2"),
                        os.path.join(ROOT, "tests", "cpp", "fake_nccl.cpp"), "-lpthread", "-ldl"], check=True)
        env = dict(os.environ, MDBG_FUZZ_CHILD="1", MDBG_EMU_LIB=lib,
                   LD_LIBRARY_PATH=tmp + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
        sys.exit(subprocess.run([sys.executable] + sys.argv, env=env).returncode)

    from metamdbg_b200 import _capi
    _capi.LIB_PATH = os.environ["MDBG_EMU_LIB"]
    from metamdbg_b200 import Engine, synth
    from metamdbg_b200.parallel import owner_of
    from oracle import pyoracle
    orc = pyoracle.Oracle()
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n = dict(scenarios=0, entries=0, empty_ranks=0)
    while time.time() < t_end:
        world = int(rng.integers(2, 7))
        n_reads = int(rng.choice([0, 1, 3, 7, 40, 150]))
        k = int(rng.choice([4, 4, 3, 5, 7]))
        dens = float(rng.choice([0.05, 0.02, 0.2]))
        rescue = bool(rng.integers(0, 2))
        chain = int(rng.integers(0, 3))
        min_ab = 0 if rescue else 2
        rs = synth.make_readset(max(n_reads, 1), 4000, seed=int(rng.integers(1, 10 ** 6)), n_genomes=1,
                                genome_len_range=(40_000, 60_000), err=0.002)
        bases, offs = synth.fill_reads(rs)
        if n_reads == 0:
            bases, offs = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
        cuts = np.sort(rng.integers(0, n_reads + 1, world - 1)) if n_reads else np.zeros(world - 1, np.int64)
        bounds = [0] + [int(c) for c in cuts] + [n_reads]                 # uneven shards, some of them empty
        n["empty_ranks"] += sum(1 for r in range(world) if bounds[r] == bounds[r + 1])
        scen = dict(world=world, n_reads=n_reads, k=k, dens=dens, rescue=rescue, chain=chain, bounds=bounds)
        uid = Engine.nccl_unique_id()
        tables, errors = [None] * world, []
        edges = [None] * world

        def rank_main(rank):
            try:
                eng = Engine(15, dens, True)
                eng.comm_init(rank, world, uid)
                lo, hi = bounds[rank], bounds[rank + 1]
                sub = (offs[lo:hi + 1] - offs[lo]).astype(np.uint64)
                eng.sketch_batch(bases[int(offs[lo]):int(offs[hi])], sub, append_to_store=True, fetch=False)
                got = []
                eng.count_begin(k, 0)
                eng.count_add_store()
                eng.count_merge()
                if rescue:
                    eng.count_rescue()
                got.append(eng.count_finalize(min_ab))
                edges[rank] = eng.edges_index(min_ab)                    # collective: keys travel to their owners
                for kk in range(k + 1, k + 1 + chain):
                    eng.prev_from_current(min_ab)
                    eng.count_begin(kk, 0)
                    eng.count_add_store_next_k()
                    eng.count_merge()
                    got.append(eng.count_finalize(min_ab))
                tables[rank] = got
                eng.close()
            except Exception as e:                                        # noqa: BLE001
                errors.append((rank, repr(e)))

        th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(600)
        assert not errors and all(t is not None for t in tables), (scen, errors)
        mo, m, _, _ = orc.sketch_batch(bases, offs, 15, dens, True)
        ref = orc.count(m, mo, k, 2)
        ph, pa = ref["hashes"], ref["abundances"]
        if rescue:
            rr = orc.rescue(m, mo, k, ph, pa)
            if len(rr["hashes"]):
                ph = np.concatenate([ph, rr["hashes"]]); pa = np.concatenate([pa, np.ones(len(rr["hashes"]), np.uint32)])
        if k >= 2:                                                        # edge keys of the first table's nodes
            nodes = np.concatenate([tables[r][0].kminmers for r in range(world)]) if sum(len(tables[r][0].abundances) for r in range(world)) else np.zeros((0, k), np.uint32)
            we = orc.edge_index(nodes, k)
            got_e = set()
            for rank in range(world):
                for h in edges[rank]["hashes"]:
                    key = (int(h[1]), int(h[0]))
                    assert int(owner_of(np.uint64(key[0]), world)) == rank and key not in got_e, ("edge owner", scen)
                    got_e.add(key)
            assert got_e == {(int(h[0]), int(h[1])) for h in we["hashes"]}, ("edges", scen, len(got_e), len(we["hashes"]))
            assert sum(e["checksum"] for e in edges) % 2 ** 64 == we["checksum"], ("edge checksum", scen)
            wv = orc.edge_values(nodes, k)
            got_v = {}
            for e in edges:
                for h, v in zip(e["hashes"], e["values"]):
                    got_v[(int(h[1]), int(h[0]))] = v.tolist()
            assert got_v == {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])}, ("edge values", scen)
        for step in range(chain + 1):
            if step:
                nk = orc.next_k(m, mo, k + step, ph, pa)
                ph, pa = nk["hashes"], nk["abundances"]
            want = {(int(h[0]), int(h[1])): int(x) for h, x in zip(ph, pa)}
            got = {}
            for rank in range(world):
                for key, ab in tables[rank][step].as_dict().items():
                    assert int(owner_of(np.uint64(key[0]), world)) == rank and key not in got, ("owner", scen, step)
                    got[key] = ab
            assert got == want, ("table", scen, step, len(got), len(want))
            n["entries"] += len(want)
        n["scenarios"] += 1
    print("no difference:", n)


if __name__ == "__main__":
    main()
