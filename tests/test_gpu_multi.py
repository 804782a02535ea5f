"""N=2 GPU test of mdbg_count_merge (NCCL all-to-all inside the library): the union of the
ranks' finalised tables equals the oracle's table for all reads.  Skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest

from metamdbg_b200 import synth

pytestmark = pytest.mark.gpu
K = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from metamdbg_b200 import Engine
    from metamdbg_b200.parallel import init_engine_comm
    rs = synth.make_readset(3000, 8000, seed=91, n_genomes=2, genome_len_range=(200_000, 300_000))
    bases, offs = synth.fill_reads(rs.shard(rank, world))
    eng = Engine(15, 0.005, True, device=rank)
    init_engine_comm(eng, rank, world, torch.device("cuda", rank))
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    eng.purge_palindromes(4, 80)
    for k in (K, 9):
        eng.count_begin(k)
        eng.count_add_store()
        eng.count_merge()
        tab = eng.count_finalize(2)
        ed = eng.edges_index(2)                          # collective: edge keys of the owned nodes -> their owners
        np.savez(os.path.join(out_dir, f"rank{rank}_k{k}.npz"), hashes=tab.hashes, abund=tab.abundances,
                 vecs=tab.kminmers, edge_keys=ed["hashes"], edge_values=ed["values"],
                 edge_checksum=np.array([ed["checksum"]], np.uint64))
    # default mode + multi-k over NCCL: rescue across ranks, then k = 5, 6 from the replicated previous-k table
    from metamdbg_b200 import multi_k_sweep

    def keep(k, e):
        tab = e.count_finalize(0)
        np.savez(os.path.join(out_dir, f"chain_rank{rank}_k{k}.npz"), hashes=tab.hashes, abund=tab.abundances)

    res = multi_k_sweep(eng, K, K + 2, min_abundance=0, rescue=True, merge=True, on_table=keep)
    np.save(os.path.join(out_dir, f"chain_rank{rank}_rescued.npy"), np.array([res[0]["n_reads_rescued"]]))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_merge_matches_oracle(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    g.build()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = synth.make_readset(3000, 8000, seed=91, n_genomes=2, genome_len_range=(200_000, 300_000))
    bases, offs = synth.fill_reads(rs)
    mo, m, _, _ = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    for k in (K, 9):
        got = {}
        for r in range(world):
            z = np.load(tmp_path / f"rank{r}_k{k}.npz")
            for h, a, v in zip(z["hashes"], z["abund"], z["vecs"]):
                key = (int(h[1]), int(h[0]))
                assert key not in got
                got[key] = (int(a), tuple(int(x) for x in v))
        ref = oracle.count(m, mo, k, 2)
        want = {(int(h[0]), int(h[1])): (int(a), tuple(int(x) for x in v))
                for h, a, v in zip(ref["hashes"], ref["abundances"], ref["vecs"])}
        assert got == want and len(want) > 100
        we = oracle.edge_index(ref["vecs"], k)            # CreateMdbg::EdgeIndexer key set of this node set
        wv = oracle.edge_values(ref["vecs"], k)
        keys, cs, vals = set(), 0, {}
        for r in range(world):
            z = np.load(tmp_path / f"rank{r}_k{k}.npz")
            for h, v in zip(z["edge_keys"], z["edge_values"]):
                assert (int(h[1]), int(h[0])) not in keys
                keys.add((int(h[1]), int(h[0])))
                vals[(int(h[1]), int(h[0]))] = v.tolist()
            cs = (cs + int(z["edge_checksum"][0])) % 2 ** 64
        assert keys == {(int(h[0]), int(h[1])) for h in we["hashes"]} and cs == we["checksum"]
        assert vals == {(int(h[0]), int(h[1])): v.tolist() for h, v in zip(wv["hashes"], wv["values"])}
    # the chain: solid + rescued at k = 4, then two next-k passes (oracle restatement of the whole read set)
    pm, po = [], [0]
    for r in range(rs.n_reads):
        q, _ = oracle.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, 80)
        pm.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(pm).astype(np.uint32); po = np.array(po, np.uint64)
    solid = oracle.count(pm, po, K, 2)
    resc = oracle.rescue(pm, po, K, solid["hashes"], solid["abundances"])
    assert sum(int(np.load(tmp_path / f"chain_rank{r}_rescued.npy")[0]) for r in range(world)) == resc["n_reads_rescued"]
    ph = np.concatenate([solid["hashes"], resc["hashes"]]) if len(resc["hashes"]) else solid["hashes"]
    pa = np.concatenate([solid["abundances"], np.ones(len(resc["hashes"]), np.uint32)])
    for k in (K, K + 1, K + 2):
        if k > K:
            nk = oracle.next_k(pm, po, k, ph, pa)
            ph, pa = nk["hashes"], nk["abundances"]
        want = {(int(h[0]), int(h[1])): int(a) for h, a in zip(ph, pa)}
        got = {}
        for r in range(world):
            z = np.load(tmp_path / f"chain_rank{r}_k{k}.npz")
            for h, a in zip(z["hashes"], z["abund"]):
                key = (int(h[1]), int(h[0]))
                assert key not in got
                got[key] = int(a)
        assert got == want and len(want) > 100, f"chain k={k}"
