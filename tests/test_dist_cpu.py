"""world_size-2 gloo test (CPU) of the N>1 host logic: record sharding + the owner-partitioned
table merge protocol (SURVEY.md section 8e / DESIGN.md section 6), restated with the CPU oracle
and torch.distributed all_to_all: the union of the per-owner merged tables must equal the
single-process table, and the >=2 filter must run after the merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metamdbg_b200 import synth
from metamdbg_b200.parallel import owner_of, shard_range

K = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    orc = Oracle()
    rs = synth.make_readset(400, 5000, seed=77, n_genomes=1, genome_len_range=(60_000, 60_001))
    lo, hi = shard_range(rs.n_reads, rank, world)
    bases, offs = synth.fill_reads(rs.shard(rank, world))
    assert rs.shard(rank, world).index_base == lo and len(offs) - 1 == hi - lo
    mo, m, _, _ = orc.sketch_batch(bases, offs, 15, 0.005, True)
    local = orc.count(m, mo, K, keep_all=True)               # local table incl. abundance 1
    own = owner_of(local["hashes"][:, 0], world)
    # all-to-all of (vector, count) records, bucketed by owner
    send = []
    for d in range(world):
        sel = own == d
        rec = np.concatenate([local["vecs"][sel].astype(np.int64), local["abundances"][sel, None].astype(np.int64)], 1)
        send.append(torch.from_numpy(rec.reshape(-1, K + 1)))
    counts = torch.tensor([len(x) for x in send], dtype=torch.int64)
    all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    recv = [torch.zeros((int(all_counts[s][rank]), K + 1), dtype=torch.int64) for s in range(world)]
    # gloo has no all_to_all: pairwise isend/irecv
    reqs = []
    for d in range(world):
        if d == rank:
            recv[d].copy_(send[d])
        else:
            reqs.append(dist.isend(send[d], d))
            reqs.append(dist.irecv(recv[d], d))
    for r in reqs:
        r.wait()
    merged = {}
    for t in recv:
        for row in t.numpy():
            key = tuple(int(x) for x in row[:K])
            merged[key] = merged.get(key, 0) + int(row[K])
    # every key this rank received is owned by it
    for key in merged:
        h1, _ = orc.hash128(np.array(key, dtype=np.uint32))
        assert int(owner_of(np.array([h1], dtype=np.uint64), world)[0]) == rank
    solid = {k: v for k, v in merged.items() if v >= 2}         # filter AFTER the merge
    np.save(os.path.join(out_dir, f"rank{rank}.npy"),
            np.array([list(k) + [v] for k, v in solid.items()], dtype=np.int64).reshape(-1, K + 1))

    # ---- the multi-k step of the protocol: the previous-k table is REPLICATED (all-gather of
    # the owners' solid (hash, abundance) pairs), the next-k pass runs on the local reads, and the resulting
    # (vector -> value) entries go to their owners where equal keys are kept once, never summed
    pairs = np.array([list(orc.hash128(np.array(k, dtype=np.uint32))) + [v] for k, v in solid.items()],
                     dtype=np.uint64).reshape(-1, 3)
    n_pairs = torch.tensor([len(pairs)], dtype=torch.int64)
    all_n = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_n, n_pairs)
    cap = int(max(int(x) for x in all_n))
    padded = torch.zeros((cap, 3), dtype=torch.int64)
    padded[:len(pairs)] = torch.from_numpy(pairs.view(np.int64))
    gathered = [torch.zeros((cap, 3), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, padded)
    prev = np.concatenate([g.numpy()[:int(n)].view(np.uint64) for g, n in zip(gathered, all_n)])
    nk = orc.next_k(m, mo, K + 1, prev[:, :2].copy(), prev[:, 2].astype(np.uint32))      # local reads, replicated table
    own = owner_of(nk["hashes"][:, 0], world)
    send = []
    for d in range(world):
        sel = own == d
        rec = np.concatenate([nk["vecs"][sel].astype(np.int64), nk["abundances"][sel, None].astype(np.int64)], 1)
        send.append(torch.from_numpy(rec.reshape(-1, K + 2)))
    counts = torch.tensor([len(x) for x in send], dtype=torch.int64)
    all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    recv = [torch.zeros((int(all_counts[s_][rank]), K + 2), dtype=torch.int64) for s_ in range(world)]
    reqs = []
    for d in range(world):
        if d == rank:
            recv[d].copy_(send[d])
        else:
            reqs.append(dist.isend(send[d], d))
            reqs.append(dist.irecv(recv[d], d))
    for r_ in reqs:
        r_.wait()
    table = {}
    for t in recv:
        for row in t.numpy():
            key = tuple(int(x) for x in row[:K + 1])
            assert table.setdefault(key, int(row[K + 1])) == int(row[K + 1])      # same key, same value on every rank
    np.save(os.path.join(out_dir, f"next_rank{rank}.npy"),
            np.array([list(k) + [v] for k, v in table.items()], dtype=np.int64).reshape(-1, K + 2))
    dist.barrier()
    dist.destroy_process_group()


def test_owner_partitioned_merge_gloo(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        for row in np.load(tmp_path / f"rank{r}.npy"):
            key = tuple(int(x) for x in row[:K])
            assert key not in got                       # owners are disjoint
            got[key] = int(row[K])
    rs = synth.make_readset(400, 5000, seed=77, n_genomes=1, genome_len_range=(60_000, 60_001))
    bases, offs = synth.fill_reads(rs)
    mo, m, _, _ = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    ref = oracle.count(m, mo, K, 2)
    want = {tuple(int(x) for x in v): int(a) for v, a in zip(ref["vecs"], ref["abundances"])}
    assert got == want and len(want) > 50
    # the next-k tables of the two owners = the oracle's next-k table of the whole read set
    nk = oracle.next_k(m, mo, K + 1, ref["hashes"], ref["abundances"])
    want_n = {tuple(int(x) for x in v): int(a) for v, a in zip(nk["vecs"], nk["abundances"])}
    got_n = {}
    for r in range(world):
        for row in np.load(tmp_path / f"next_rank{r}.npy"):
            key = tuple(int(x) for x in row[:K + 1])
            assert key not in got_n
            got_n[key] = int(row[K + 1])
    assert got_n == want_n and len(want_n) > 30


def test_owner_of_matches_header_formula():
    rng = np.random.default_rng(1)
    h = rng.integers(0, 2 ** 64, size=1000, dtype=np.uint64)
    for world in (1, 2, 3, 8):
        o = owner_of(h, world)
        assert o.min() >= 0 and o.max() < world
        assert np.array_equal(o, np.array([((int(x) >> 32) * world) >> 32 for x in h]))


def test_shards_cover_all_reads():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8)):
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
