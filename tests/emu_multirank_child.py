"""Child process of tests/test_capi_emulated_cpu.py: N ranks = N threads, each with its own context of the
CPU-emulated library, merged through mdbg_count_merge over the in-process fake NCCL (tests/cpp/fake_nccl.cpp).
Environment: MDBG_EMU_LIB (emulated library), LD_LIBRARY_PATH with the fake libnccl.so.2 first.  Prints OK."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metamdbg_b200 import _capi  # noqa: E402

_capi.LIB_PATH = os.environ["MDBG_EMU_LIB"]
from metamdbg_b200 import Engine, synth  # noqa: E402
from metamdbg_b200.parallel import owner_of, shard_range  # noqa: E402
from oracle import pyoracle  # noqa: E402


def main():
    n_ranks = int(sys.argv[1])
    k = int(sys.argv[2])
    orc = pyoracle.Oracle()
    rs = synth.make_readset(260, 5000, seed=13, n_genomes=2, genome_len_range=(60_000, 90_000))
    bases, offs = synth.fill_reads(rs)
    uid = Engine.nccl_unique_id()                       # loads the (fake) NCCL once, before the threads start
    results, errors = [None] * n_ranks, []

    def rank_main(rank):
        try:
            eng = Engine(15, 0.05, True)
            eng.comm_init(rank, n_ranks, uid)
            lo, hi = shard_range(rs.n_reads, rank, n_ranks)
            sub_offs = (offs[lo:hi + 1] - offs[lo]).astype(np.uint64)
            eng.sketch_batch(bases[int(offs[lo]):int(offs[hi])], sub_offs, append_to_store=True, fetch=False)
            eng.purge_palindromes(4, 50)
            eng.count_begin(k, 0)
            eng.count_add_store()
            eng.count_merge()
            results[rank] = (eng.count_finalize(2), eng.count_stats(2))
            eng.close()
        except Exception as e:                           # noqa: BLE001
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(n_ranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    assert not errors, errors
    assert all(r is not None for r in results), "a rank did not finish"
    # oracle on the whole read set
    mo, m, _, _ = orc.sketch_batch(bases, offs, 15, 0.05, True)
    pm, po = [], [0]
    for r in range(rs.n_reads):
        q, _ = orc.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, 50)
        pm.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(pm).astype(np.uint32); po = np.array(po, np.uint64)
    ref = orc.count(pm, po, k, 2)
    want = {(int(h[0]), int(h[1])): int(a) for h, a in zip(ref["hashes"], ref["abundances"])}
    merged, instances, distinct = {}, 0, 0
    for rank, (tab, st) in enumerate(results):
        for key, ab in tab.as_dict().items():
            assert owner_of(key[0], n_ranks) == rank, "key on the wrong owner"
            assert key not in merged, "key on two owners"
            merged[key] = ab
        instances += st["n_instances"]; distinct += st["n_distinct"]
    assert merged == want and len(want) > 3000, (len(merged), len(want))
    assert instances == ref["n_instances"] and distinct == ref["n_distinct"], "occurrences not conserved"
    print(f"{n_ranks} ranks, k={k}: {len(want)} solid k-min-mers, {instances} occurrences conserved")
    print("OK")


if __name__ == "__main__":
    main()
