"""Child process of tests/test_capi_emulated_cpu.py: the multi-k loop (multi_k_sweep, keys-only owner merge after every
k) on N ranks = N threads over the in-process fake NCCL, four sweeps in a row.  With several ranks three table buffers
rotate (count table, previous-k table, and the replication source), they trade places and grow to a common size: every
sweep must give the same tables, the owner-partitioned tables of every k must add up to the single-context table, and
the allocations must die out (three buffers rotate: at most one grows per sweep after the first, none in the fourth).  Environment: MDBG_EMU_LIB, LD_LIBRARY_PATH with the fake libnccl.so.2.
Arguments: n_ranks last_k.  Prints OK."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metamdbg_b200 import _capi  # noqa: E402

_capi.LIB_PATH = os.environ["MDBG_EMU_LIB"]
from metamdbg_b200 import Engine, multi_k_sweep, synth  # noqa: E402
from metamdbg_b200.parallel import shard_range  # noqa: E402

MASK = (1 << 64) - 1


def main():
    n_ranks, last_k = int(sys.argv[1]), int(sys.argv[2])
    rs = synth.make_readset(1500, 6000, seed=13, n_genomes=3, genome_len_range=(100_000, 150_000), err=0.002)
    bases, offs = synth.fill_reads(rs)
    uid = Engine.nccl_unique_id()
    out, errors = [None] * n_ranks, []

    def rank_main(rank):
        try:
            eng = Engine(15, 0.02, True)
            eng.comm_init(rank, n_ranks, uid)
            lo, hi = shard_range(rs.n_reads, rank, n_ranks)
            eng.sketch_batch(bases[int(offs[lo]):int(offs[hi])], (offs[lo:hi + 1] - offs[lo]).astype(np.uint64),
                             append_to_store=True, fetch=False)
            res = []
            for _ in range(4):
                a0 = eng.allocations()
                r = multi_k_sweep(eng, 4, last_k, 2, merge="hashes", world=n_ranks)
                a1 = eng.allocations()
                res.append((a1[0] - a0[0], a1[1] - a0[1], [(x["n_entries"], x["checksum"]) for x in r]))
            out[rank] = res
            eng.close()
        except Exception as e:                           # noqa: BLE001
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    one = Engine(15, 0.02, True)
    one.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    want = [(x["n_entries"], x["checksum"]) for x in multi_k_sweep(one, 4, last_k, 2)]
    one.close()
    for sweep in range(4):
        got = [(sum(out[r][sweep][2][i][0] for r in range(n_ranks)), sum(out[r][sweep][2][i][1] for r in range(n_ranks)) & MASK)
               for i in range(len(want))]
        assert got == want, (sweep, got, want)
    for r in range(n_ranks):
        assert out[r][3][0] == 0 and out[r][1][0] <= 2 and out[r][2][0] <= 1, (r, [x[:2] for x in out[r]])   # three buffers rotate: settled by sweep 4
    print("allocations / buffer trades per sweep, rank 0:", [x[:2] for x in out[0]])
    print("OK")


if __name__ == "__main__":
    main()
