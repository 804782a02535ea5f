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
    n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 260        # fewer reads than ranks: some ranks hold no read at all
    rs = synth.make_readset(n_reads, 5000, seed=13, n_genomes=2, genome_len_range=(60_000, 90_000))
    bases, offs = synth.fill_reads(rs)
    uid = Engine.nccl_unique_id()                       # loads the (fake) NCCL once, before the threads start
    results, errors = [None] * n_ranks, []
    next_results = [None] * n_ranks
    rescue_results = [None] * n_ranks
    edge_results = [None] * n_ranks
    light_results = [None] * n_ranks

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
            edge_results[rank] = eng.edges_index(2)          # collective: keys of the owned nodes -> their owner ranks
            # default mode (--min-abundance 0): rescue across ranks, then the table with the rescued abundance-1 entries
            n_resc = eng.count_rescue()
            rescue_results[rank] = (n_resc, eng.count_finalize(0))
            # multi-k on the device: previous-k table replicated from the owners, next-k pass on this rank's reads,
            # owner merge of the (key -> value) tables, twice
            chain = []
            for kk in (k + 1, k + 2):
                eng.prev_from_current(2)
                eng.count_begin(kk, 0)
                eng.count_add_store_next_k()
                eng.count_merge()
                chain.append(eng.count_finalize(0))
            next_results[rank] = chain
            # the same two passes with the keys-only merge (no vectors on the owners), from a fresh first pass
            eng.count_begin(k, 0)
            eng.count_add_store()
            eng.count_merge()
            eng.count_rescue()
            light = []
            for kk in (k + 1, k + 2):
                eng.prev_from_current(2)
                eng.count_begin(kk, 0)
                eng.count_add_store_next_k()
                eng.count_merge_hashes()
                t = eng.count_finalize(0)
                assert t.kminmers.shape[0] == 0
                light.append(t)
            light_results[rank] = light
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
    assert merged == want and (len(want) > 3000 or n_reads < 260), (len(merged), len(want))
    assert instances == ref["n_instances"] and distinct == ref["n_distinct"], "occurrences not conserved"
    # edge keys (CreateMdbg::EdgeIndexer) of the solid nodes: disjoint per owner, union = the oracle's key set
    we = orc.edge_index(ref["vecs"], k)
    got_e = set()
    for rank, ed in enumerate(edge_results):
        for h in ed["hashes"]:
            key = (int(h[1]), int(h[0]))
            assert owner_of(key[0], n_ranks) == rank and key not in got_e
            got_e.add(key)
    assert got_e == {(int(h[0]), int(h[1])) for h in we["hashes"]}, (len(got_e), len(we["hashes"]))
    assert sum(ed["checksum"] for ed in edge_results) % 2 ** 64 == we["checksum"]
    assert sum(ed["n_nodes"] for ed in edge_results) == len(ref["abundances"])
    wv = orc.edge_values(ref["vecs"], k)                  # order-free indexEdge content, folded on the key's owner
    want_v = {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])}
    got_v = {}
    for ed in edge_results:
        for h, v in zip(ed["hashes"], ed["values"]):
            got_v[(int(h[1]), int(h[0]))] = v.tolist()
    assert got_v == want_v, "edge values"
    print(f"  edge keys: {len(got_e)} distinct over {n_ranks} owners, values identical")
    # rescue: solid + rescued entries over all ranks = the oracle's table of the whole read set
    resc = orc.rescue(pm, po, k, ref["hashes"], ref["abundances"])
    want_r = dict(want)
    for h in resc["hashes"]:
        want_r[(int(h[0]), int(h[1]))] = 1
    got_r, reads_rescued, n_rescued = {}, 0, 0
    for rank, (n_resc, tab) in enumerate(rescue_results):
        reads_rescued += n_resc
        n_rescued += tab.n_rescued
        for key, ab in tab.as_dict().items():
            assert owner_of(key[0], n_ranks) == rank and key not in got_r
            got_r[key] = ab
    assert got_r == want_r and n_rescued == len(resc["hashes"]) and (n_rescued > 50 or n_reads < 260), (len(got_r), len(want_r), n_rescued)
    assert reads_rescued == resc["n_reads_rescued"], (reads_rescued, resc["n_reads_rescued"])
    print(f"  rescue: {n_rescued} abundance-1 k-min-mers of {reads_rescued} reads flagged on their owners")
    # the multi-k chain against the oracle's next-k restatement (IndexKminmerFunctor / getRefinedAbundance); the
    # previous-k table holds the solid AND the rescued entries, as kminmerData_abundance.txt does in default mode
    prev_h = np.concatenate([ref["hashes"], resc["hashes"]]) if len(resc["hashes"]) else ref["hashes"]
    prev_a = np.concatenate([ref["abundances"], np.ones(len(resc["hashes"]), np.uint32)])
    for step, kk in enumerate((k + 1, k + 2)):
        nk = orc.next_k(pm, po, kk, prev_h, prev_a)
        want_k = {(int(h[0]), int(h[1])): int(a) for h, a in zip(nk["hashes"], nk["abundances"])}
        got_k = {}
        for rank in range(n_ranks):
            for key, ab in next_results[rank][step].as_dict().items():
                assert owner_of(key[0], n_ranks) == rank and key not in got_k
                got_k[key] = ab
        assert got_k == want_k and (len(want_k) > 1000 or n_reads < 260), (kk, len(got_k), len(want_k))
        got_l = {}
        for rank in range(n_ranks):
            for key, ab in light_results[rank][step].as_dict().items():
                assert owner_of(key[0], n_ranks) == rank and key not in got_l
                got_l[key] = ab
        assert got_l == want_k, ("keys-only merge", kk, len(got_l), len(want_k))
        prev_h, prev_a = nk["hashes"], nk["abundances"]
        print(f"  next-k {kk}: {len(want_k)} entries identical over {n_ranks} ranks")
    print(f"{n_ranks} ranks, k={k}: {len(want)} solid k-min-mers, {instances} occurrences conserved")
    print("OK")


if __name__ == "__main__":
    main()
