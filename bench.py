#!/usr/bin/env python
"""bench.py -- Gbp/s through minimizer-sketch + k-min-mer count (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU stages (oracle/_ref)

Headline workload (config.workload): BASELINE.json configs[1] per GPU -- 1 M synthetic HiFi reads x 15 kbp, l=15,
d=0.005, HPC on, k=4, abundance >= 2.  A "step" is one pass of the hot path over that batch:

    sketch (2-bit packed reads resident in HBM) -> minimizer store -> purgePalindromes -> k-min-mer insert
    -> [NCCL owner merge for N>1] -> statistics + on-device emit of (hash128, abundance, k-min-mer) arrays

`value` times it with the reads resident in HBM in the packed layout the design brief names (`ascii_resident` is the
same step on ASCII bytes, i.e. with the device-side pack pass in it); `e2e` times the same work through the
host-buffer C ABI (ASCII in pinned host memory -> 2-bit packing by the library's host threads -> H2D; D2H of the
minimizer CSR and of the finalised table inside the timed region).  For N>1 every rank processes its own shard of
N x 1 M reads (weak scaling) and the only data-path collective of the step is the owner-partitioned table merge.

Extras in the same JSON line (each with its own timing, none of them part of `value`): the multi-k loop k = 4..21 on
the resident store (collective for N>1), edge indexing, and the other BASELINE.json configs -- cfg3 (ONT, N=1),
cfg4 (5 M reads TOTAL, k = 4..21, strong scaling) and cfg5 (20 M reads TOTAL in 3 samples, strong scaling).  For N>1 a
parity leg runs first: count -> merge -> rescue -> two next-k passes -> edge index over real NCCL on a shared read
set, every stage compared with the CPU oracle of the whole set; a mismatch ends the run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L, K, MIN_AB = 15, 4, 2
SEED = 20260924
METRIC = "Gbp/s through minimizer-sketch + k-min-mer count"

# BASELINE.json configs.  cfg2 is the configuration the metric is quoted on (the default headline); cfg3 is the ONT
# shape: no homopolymer compression, sketch at the correction density 0.025 ("nanoMDBG density"), then
# Utils::applyDensityThreshold(0.005) on the minimizer store, purge, k = 4 count.  cfg4 / cfg5 fix the
# TOTAL number of reads (strong scaling): the 5 M-read multi-k sweep and the 20 M-read 3-sample co-assembly.
WORKLOADS = {
    "cfg2": dict(reads=1_000_000, per_gpu=True, read_len=15_000, density=0.005, hpc=True, asm_density=0.0, err=0.001,
                 samples=1, last_k=4,
                 label="cfg2: 1M synthetic HiFi reads x 15 kbp per GPU, l=15 d=0.005 HPC k=4 min-abundance 2"),
    "cfg3": dict(reads=2_000_000, per_gpu=True, read_len=8_000, density=0.025, hpc=False, asm_density=0.005, err=0.02,
                 samples=1, last_k=4,
                 label="cfg3: 2M synthetic ONT reads x 8 kbp per GPU, l=15 sketch d=0.025 (no HPC) -> density 0.005, "
                       "k=4 min-abundance 2"),
    "cfg4": dict(reads=5_000_000, per_gpu=False, read_len=15_000, density=0.005, hpc=True, asm_density=0.0, err=0.001,
                 samples=1, last_k=21,
                 label="cfg4: 5M synthetic HiFi reads x 15 kbp in TOTAL (strong scaling), l=15 d=0.005 HPC, sketch + purge + "
                       "k=4 count + next-k passes k=5..21 (previous-k tables, collectives for N>1), min-abundance 2"),
    "cfg5": dict(reads=20_000_000, per_gpu=False, read_len=15_000, density=0.005, hpc=True, asm_density=0.0, err=0.001,
                 samples=3, last_k=4,
                 label="cfg5: 20M synthetic HiFi reads x 15 kbp in TOTAL, 3 samples (datasetIndex 0-2, different abundance "
                       "profiles of one genome set) pooled as in the reference (strong scaling), k=4 min-abundance 2, "
                       "NCCL owner merge of the count tables"),
}


def usable_cpus() -> int:
    """CPUs this process can really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def host_bytes_available() -> int:
    """Host memory this process may still take: MemAvailable capped by the cgroup's remaining allowance."""
    avail = 1 << 62
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
                break
    except OSError:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            avail = min(avail, int(lim) - int(open("/sys/fs/cgroup/memory.current").read()))
    except (OSError, ValueError):
        pass
    return max(0, avail)


def purge_last_k(read_len: int, w: dict) -> int:
    # Commons::computeLastK (src/Commons.hpp:1726-1741): n50 * density * 2, at least firstK+2
    return max(int(read_len * np.float32(w["asm_density"] or w["density"]) * np.float32(2.0)), 6)


def workload_config(name: str, reads: int, read_len: int, genomes: int, world: int) -> dict:
    w = WORKLOADS[name]
    default = reads == w["reads"] and read_len == w["read_len"]
    return {
        "workload": w["label"] if default else
        f"{reads} synthetic reads x {read_len} bp {'per GPU' if w['per_gpu'] else 'in total'}, parameters of {name}: l=15 "
        f"d={w['density']} hpc={w['hpc']} assembly-density={w['asm_density'] or w['density']} k=4..{w['last_k']} min-abundance 2",
        "reads_per_gpu": reads if w["per_gpu"] else reads // max(1, world), "reads_total": reads * world if w["per_gpu"] else reads,
        "read_len_mean": read_len, "minimizer_size": L, "density": w["density"],
        "assembly_density": w["asm_density"] or w["density"], "hpc": w["hpc"], "k": K, "last_k": w["last_k"],
        "min_abundance": MIN_AB, "purge_last_k": purge_last_k(read_len, w), "genomes": genomes, "samples": w["samples"],
        "substitution_rate": w["err"],
        "input_format": "2-bit packed bases resident in HBM (16 per u32, 16-byte aligned reads); e2e: ASCII bases (Read::_seq) in pinned host memory",
        "l2_policy": "inputs (>= 1 GB per step) are far larger than the 126 MB L2; no explicit flush",
    }


def make_readsets(name: str, reads_total: int, read_len: int, genomes: int):
    """The read sets of a workload: one per sample (cfg5: three abundance profiles over the same genome set)."""
    from metamdbg_b200 import synth
    w = WORKLOADS[name]
    if w["samples"] == 1:
        return [synth.make_readset(reads_total, read_len, seed=SEED, n_genomes=genomes, err=w["err"])]
    out, base = [], 0
    for s in range(w["samples"]):
        n = reads_total * (s + 1) // w["samples"] - reads_total * s // w["samples"]
        rs = synth.make_readset(n, read_len, seed=SEED, n_genomes=genomes, err=w["err"], sample=s, index_base=base)
        out.append(rs)
        base += n
    return out


# ---------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.proc = index, [], set(), None
        self.max_mhz = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------- reference arm
def write_fastq(path: str, bases: np.ndarray, offs: np.ndarray):
    raw = bases.tobytes()
    with open(path, "wb", buffering=1 << 24) as f:
        for r in range(len(offs) - 1):
            s = raw[int(offs[r]):int(offs[r + 1])]
            f.write(b"@r%d\n" % r + s + b"\n+\n" + b"I" * len(s) + b"\n")


def run_reference(args, rank: int):
    """The reference's own CPU implementation of the path, all host threads, on a bounded sample of the workload.

    Headline (kind "reference"): the STOCK stages -- ReadSelection::execute on a FASTQ in tmpfs (kseq parsing, sketch,
    side outputs, ordered record writer, purgePalindromes; src/readSelection/ReadSelection.hpp:92-111) followed by
    CreateMdbg::KminmerCounter with its disk partitions + parallel sort (src/graph/CreateMdbg.hpp:3634-3642), i.e. what
    `metaMDBG readSelection` + `graph --firstpass` spend on this path; value = sample bases / (stage seconds).
    Extra: the same reference primitives driven in memory (no FASTQ, no disk), and the reference's multi-k loop.
    The thread count does not depend on OMP_NUM_THREADS (torch.distributed.run exports OMP_NUM_THREADS=1): every
    parallel region of the shim and of the stages takes its thread count as an argument."""
    if rank != 0:
        return
    from metamdbg_b200 import synth
    from oracle import pyoracle
    name = args.workload
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    try:
        ref = pyoracle.Reference()
        kind = "reference"
    except (FileNotFoundError, OSError):
        ref, kind = None, "port"
    world = max(1, args.gpus)
    reads_total = args.reads * world if w["per_gpu"] else args.reads
    n_sample = int(min(reads_total, args.ref_reads or 100_000))
    rs = make_readsets(name, reads_total, args.read_len, args.genomes)[0]
    n_sample = min(n_sample, rs.n_reads)
    bases, offs = synth.fill_reads(rs.subset(0, n_sample))
    n_bases = int(offs[-1])
    lk = purge_last_k(args.read_len, w)
    stock = ref is not None and not args.no_stages and name != "cfg3"      # ONT: the correction stage is out of scope
    tmp_base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    import tempfile
    tmp = tempfile.TemporaryDirectory(dir=tmp_base)
    fq = os.path.join(tmp.name, "reads.fastq")
    if stock:
        write_fastq(fq, bases, offs)
    threads = max(1, usable_cpus())
    state = {}

    def one(t: int):
        if stock:
            with tempfile.TemporaryDirectory(dir=tmp.name) as d:
                res = ref.read_selection([fq], L, w["density"], w["hpc"], threads=t, skip_correction=False, workdir=d)
            corr = res["corrected"]
            mins = np.concatenate([r["minimizers"] for r in corr]) if corr else np.zeros(0, np.uint32)
            mo = np.zeros(len(corr) + 1, np.uint64)
            mo[1:] = np.cumsum([len(r["minimizers"]) for r in corr])
            g = ref.graph_firstpass(mins, mo, K, min_abundance=MIN_AB, threads=t)
            state.update(mins=mins, mo=mo, table=g)
            return res["seconds"] + g["seconds"], dict(n_solid=g["n_solid"], n_minimizers=int(len(mins)),
                                                       readSelection_s=res["seconds"], graph_firstpass_count_s=g["seconds"])
        t0 = time.perf_counter()
        if ref is not None:
            res = ref.pipeline(bases, offs, L, w["density"], w["hpc"], K, purge_last_k=lk, min_abundance=MIN_AB, threads=t,
                               assembly_density=w["asm_density"])
        else:
            res = oracle_port_pipeline(pyoracle.Oracle(), bases, offs, lk, w)
        return time.perf_counter() - t0, res

    # the warm-up steps also pick the thread count: candidates from the usable CPUs (cgroup quota) up to every
    # logical CPU; the fastest is used for the timed steps ("all the host threads it can use")
    if ref is not None and args.warmup >= 1:
        eff = usable_cpus()
        cands = sorted({max(1, min(cores, c)) for c in (eff, 2 * eff, cores)})
        cands = cands[:max(1, args.warmup)]
        best = None
        for c in cands:
            t, _ = one(c)
            if best is None or t < best[0]:
                best = (t, c)
        threads = best[1]
        for _ in range(args.warmup - len(cands)):
            one(threads)
    else:
        threads = 1 if ref is None else threads
        for _ in range(args.warmup):
            one(threads)
    ts = []
    for _ in range(args.steps):
        t, res = one(threads)
        ts.append(t)
    total = sum(ts)
    gbps = n_bases * args.steps / total / 1e9
    sample = (f"first {n_sample} reads of the workload ({n_bases / 1e9:.3f} Gbp) per step; "
              + ("stock stages: ReadSelection::execute on a tmpfs FASTQ + KminmerCounter (disk partitions + sort), stage-internal seconds"
                 if stock else "reference primitives driven in memory (no FASTQ, no disk)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps,
        "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak" if w["per_gpu"] else "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": dict(workload_config(name, args.reads, args.read_len, args.genomes, world), sample=sample),
        "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "check": {"n_minimizers": res.get("n_minimizers"), "n_solid": res.get("n_solid")},
        "host": {"logical_cpus": cores, "usable_cpus": usable_cpus(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
    }
    if stock:
        line["reference_stages"] = {k: res[k] for k in ("readSelection_s", "graph_firstpass_count_s")}
        # extra: the same sample through the reference's primitives in memory (round 1's headline)
        t0 = time.perf_counter()
        mem = ref.pipeline(bases, offs, L, w["density"], w["hpc"], K, purge_last_k=lk, min_abundance=MIN_AB, threads=threads,
                           assembly_density=w["asm_density"])
        dt = time.perf_counter() - t0
        line["reference_in_memory"] = {"value": n_bases / dt / 1e9, "unit": "Gbp/s", "threads": threads,
                                       "same_table_as_stock_stages": mem["n_solid"] == res["n_solid"]}
    if ref is not None and args.multi_k > K and state.get("table") is not None and not args.no_ref_multi_k:
        # extra: the reference's multi-k loop on the same sample (CreateMdbg.cpp:386-468 driven per k by
        # AssemblyPipeline.hpp:606-671): k = 4 first pass above, then every k from the previous k's table --
        # getRefinedAbundance (k = 5, through KminmerCounter) and IndexKminmerFunctor (k >= 6)
        prev_h, prev_a = state["table"]["hashes"], state["table"]["abundances"]
        per_k = []
        for k in range(K + 1, args.multi_k + 1):
            t0 = time.perf_counter()
            nk = ref.graph_next_k(state["mins"], state["mo"], k, prev_h, prev_a, use_counter=(k == K + 1), threads=threads)
            per_k.append(round(time.perf_counter() - t0, 4))
            prev_h, prev_a = nk["hashes"], nk["abundances"]
        tot = res["graph_firstpass_count_s"] + sum(per_k)
        line["reference_multi_k"] = {"k_first": K, "k_last": args.multi_k, "seconds_per_k": [res["graph_firstpass_count_s"]] + per_k,
                                     "seconds_total": tot, "value": n_bases / tot / 1e9,
                                     "unit": "Gbp/s (sample bases / time of the k-loop alone)", "threads": threads,
                                     "n_entries_last_k": int(len(prev_a)),
                                     "timer": "host wall clock per k around the reference's graph stage code (incl. its file I/O)"}
    print(json.dumps(line), flush=True)
    tmp.cleanup()


def oracle_port_pipeline(orc, bases, offs, lk, w):
    """The path on the C restatement (single thread): sketch, [density re-threshold], purge, count."""
    mo, m, p, d = orc.sketch_batch(bases, offs, L, w["density"], w["hpc"])
    pm, po = [], [0]
    for r in range(len(offs) - 1):
        q = m[int(mo[r]):int(mo[r + 1])]
        if w["asm_density"]:
            q = orc.apply_density(q, w["asm_density"])
        q, _ = orc.purge_palindrome(q, 4, lk)
        pm.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(pm).astype(np.uint32) if pm else np.zeros(0, np.uint32)
    c = orc.count(pm, np.array(po, np.uint64), K, MIN_AB)
    return dict(n_solid=len(c["abundances"]), n_minimizers=len(pm), checksum=orc.checksum(c["hashes"], c["abundances"]),
                mins=pm, offs=np.array(po, np.uint64), table=c)


# ---------------------------------------------------------------- this engine
class DeviceReads:
    """A rank's shard of a workload, resident in HBM in the packed 2-bit layout (chunks of <= chunk_reads reads; each
    chunk optionally keeps its ASCII bytes).  Generated on the device, chunk by chunk, through one ASCII scratch."""

    def __init__(self, torch, dev, eng, readsets, rank, world, chunk_reads, keep_ascii):
        self.chunks, self.n_reads, self.n_bases = [], 0, 0
        scratch = None
        for rs_all in readsets:
            rs = rs_all.shard(rank, world)
            for lo in range(0, rs.n_reads, chunk_reads):
                sub = rs.subset(lo, min(rs.n_reads, lo + chunk_reads))
                n, nb = sub.n_reads, sub.n_bases
                d_off = torch.from_numpy(sub.offsets.astype(np.int64)).to(dev)
                d_vs = torch.from_numpy(sub.vstart.astype(np.int64)).to(dev)
                d_st = torch.from_numpy(sub.strand).to(dev)
                if keep_ascii:
                    d_bases = torch.empty(nb + 64, dtype=torch.uint8, device=dev)
                else:
                    if scratch is None or scratch.numel() < nb + 64:
                        scratch = None
                        scratch = torch.empty(int(1.05 * nb) + 64, dtype=torch.uint8, device=dev)
                    d_bases = scratch
                eng.synth_fill_reads(d_bases.data_ptr(), d_off.data_ptr(), d_vs.data_ptr(), d_st.data_ptr(), n,
                                     sub.index_base, sub.seed, sub.err_q24)
                d_words = torch.empty(eng.pack_device_words(nb, n) * 4, dtype=torch.uint8, device=dev)
                d_src = torch.empty(n * 8, dtype=torch.uint8, device=dev)
                eng.pack_device(d_bases.data_ptr(), d_off.data_ptr(), n, nb, d_words.data_ptr(), d_src.data_ptr())
                eng.synchronize()
                self.chunks.append(dict(n=n, n_bases=nb, off=d_off, words=d_words, src=d_src, offsets=sub.offsets,
                                        bases=d_bases if keep_ascii else None))
                self.n_reads += n
                self.n_bases += nb
        del scratch

    def sketch_all(self, eng, ascii_input=False):
        n_min = 0
        for c in self.chunks:
            if ascii_input:
                n_min += int(eng.sketch_batch_device(c["bases"].data_ptr(), c["off"].data_ptr(), c["n"], c["n_bases"], True).n_minimizers)
            else:
                # synthetic reads hold only A, C, G, T: no read is flagged for the ASCII buffer
                n_min += int(eng.sketch_batch_device_packed2(c["words"].data_ptr(), c["src"].data_ptr(),
                                                             c["bases"].data_ptr() if c["bases"] is not None else 0,
                                                             c["off"].data_ptr(), c["n"], c["n_bases"], True).n_minimizers)
        return n_min


def hot_path(eng, dr, w, lk, world, ascii_input=False, last_k=4, merge_every_k=True):
    """One step.  Returns (statistics of the LAST table incl. the on-device emit, minimizers sketched, per-k entries).

    merge_every_k=False (N > 1, multi-k): the tables of k = 5 .. last_k - 1 stay rank-local -- every rank holds the
    k-min-mers of ITS reads with their (global) abundances, which is all the next pass needs (it reads per-position
    values, no table) -- and only the first and the last table are merged to their owners."""
    eng.store_clear()
    n_min = dr.sketch_all(eng, ascii_input)
    if w["asm_density"]:
        eng.store_apply_density(w["asm_density"])
    eng.purge_palindromes(4, lk)
    eng.count_begin(K, 0)
    eng.count_add_store()
    if world > 1:
        eng.count_merge()
    tab = eng.count_finalize_device(MIN_AB)            # statistics + (hash128, abundance, k-min-mer) arrays left in HBM
    per_k = [tab["n_entries"]]
    n_keys = tab["n_entries"] * world                  # keys a rank's table of the next k will hold (about as many as now)
    for k in range(K + 1, last_k + 1):                 # multi-k: every further k from the previous k's table, on the device
        eng.prev_from_current(MIN_AB)
        eng.count_begin(k, int(1.3 * max(256, n_keys)))     # grows on demand
        eng.count_add_store_next_k()
        merged = world > 1 and (merge_every_k or k == last_k)
        if merged:
            eng.count_merge_hashes()                   # (hash128, abundance) records: the k > 4 tables need no vectors on the owner
        tab = eng.count_finalize_device(MIN_AB)
        per_k.append(tab["n_entries"])
        n_keys = tab["n_entries"] * (world if merged else 1)
    return tab, n_min, per_k


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import __graft_entry__ as ge

    ge.build()
    from metamdbg_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: int) -> int:
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def sum_u64_over_ranks(x: int) -> int:
        """sum mod 2^64 (checksums): the two 32-bit halves travel separately, so nothing overflows on the way"""
        if world == 1:
            return int(x) % (1 << 64)
        t = torch.tensor([int(x) & 0xFFFFFFFF, (int(x) >> 32) & 0xFFFFFFFF], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        lo, hi = int(t[0].item()), int(t[1].item())
        return (lo + (hi << 32)) % (1 << 64)

    def new_engine(w):
        e = Engine(L, w["density"], w["hpc"], device=local_rank)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.enable_timing(True)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            e.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        return e

    def timed(fn, warmup: int, steps: int):
        """W warm-up calls, then K calls timed with CUDA events on the launching stream between two barriers; max over ranks."""
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, out

    name = args.workload
    w = WORKLOADS[name]
    reads_total = args.reads * world if w["per_gpu"] else args.reads
    lk = purge_last_k(args.read_len, w)
    eng = new_engine(w)

    # ---- N > 1: parity of every collective stage over real NCCL, before anything is timed --------------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(args, eng, torch, dev, rank, world, sum_over_ranks, sum_u64_over_ranks, barrier)

    # ---- synthetic reads, generated straight into HBM; 2-bit packed copy = the resident input of `value` ------------
    readsets = make_readsets(name, reads_total, args.read_len, args.genomes)
    dr = DeviceReads(torch, dev, eng, readsets, rank, world, args.chunk_reads, keep_ascii=True)
    n_reads, n_bases = dr.n_reads, dr.n_bases
    c0 = dr.chunks[0]
    # engine set-up, outside every timed region: every sketch-kernel variant on this rank's first chunk (ASCII), complete
    # outputs compared on the device; reported, and a variant that differs from variant 0 ends the run
    if args.no_autotune:
        tune = {"chosen": 2, "identical": [True], "ms": [], "skipped": True}
    else:
        tune = eng.autotune_sketch(c0["bases"].data_ptr(), c0["off"].data_ptr(), c0["n"], c0["n_bases"])
    if not all(tune["identical"]):
        raise SystemExit(f"bench.py: a sketch kernel variant does not reproduce variant 0 on this device: {tune}")
    eng.set_sketch_variant(args.sketch_variant if args.sketch_variant >= 0 else 2)
    active_variant = eng.sketch_variant

    n_sketched = [0]

    def step_device():
        tab, n_sketched[0], _ = hot_path(eng, dr, w, lk, world, last_k=w["last_k"])
        return tab

    # ---- device-resident timing ---------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        stats = step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches
    allocs0 = eng.allocations()[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sketch_ms, insert_ms = [], []
    step_checksums = {stats["checksum"]} if args.warmup else set()
    e0.record()
    for _ in range(args.steps):
        stats = step_device()
        sketch_ms.append(eng.kernel_time_ms(0))
        insert_ms.append(eng.kernel_time_ms(1))
        step_checksums.add(stats["checksum"])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches - launches0
    allocs_timed = eng.allocations()[0] - allocs0            # steady state: the library must not allocate inside a step
    total_bases = sum_over_ranks(n_bases)
    value = total_bases * args.steps / (ms_total * 1e-3) / 1e9
    n_min_store = eng.store_size()[1]
    solid_total = sum_over_ranks(stats["n_entries"])
    checksum_local = stats["checksum"]
    checksum_total = sum_u64_over_ranks(checksum_local)
    # size-independent property at full size: every k-min-mer occurrence of every read is in exactly one table
    # (after the owner merge): sum of all abundances over all ranks == sum over reads of max(0, n_minimizers - k + 1)
    occ = None
    local_windows = 0
    if w["last_k"] == K:
        so, _ = eng.store_fetch()
        per_read = np.diff(so.astype(np.int64))
        local_windows = int(np.maximum(per_read - K + 1, 0).sum())         # k-min-mer instances of this rank's insert pass
        expect_instances = sum_over_ranks(local_windows)
        got_instances = sum_over_ranks(stats["n_instances"])
        if expect_instances != got_instances:
            raise SystemExit(f"bench.py: occurrence conservation violated: {got_instances} != {expect_instances}")
        occ = int(got_instances)

    # exclusive phase times of one more step (diagnostic, not part of any timed region)
    eng.phase_profile(True)
    step_device()
    step_phases = eng.phase_times()
    eng.phase_profile(False)

    # ---- extra: the same step with the reads resident as ASCII (device-side pack pass inside the step) ---------------
    ascii_leg = None
    if not args.no_ascii_leg and w["last_k"] == K:
        def step_ascii():
            return hot_path(eng, dr, w, lk, world, ascii_input=True)[0]
        ms_a, st_a = timed(step_ascii, 2, max(1, min(args.steps, 3)))
        ascii_leg = {"value": total_bases / (ms_a * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms_a,
                     "same_table_as_packed_leg": st_a["checksum"] == checksum_local,
                     "what": "identical step, reads resident as ASCII bytes: ASCII -> 2-bit pack kernel + packed sketch kernel"}
        if st_a["checksum"] != checksum_local:
            raise SystemExit("bench.py: the ASCII-resident step gives a different table than the packed-resident step")
        step_device()                                   # leave the packed leg's table current for the extras below

    # ---- extra: edge keys + order-free edge values of the k = 4 node set (CreateMdbg::EdgeIndexer / indexEdge), the
    # first step beyond the count table; the table of the last step is still current here
    edges_extra = None
    if not args.no_edges and w["last_k"] == K:
        try:
            t_e = []
            for _ in range(2):                              # 2nd = warm
                barrier()
                t0 = time.perf_counter()
                ed = eng.edges_index(MIN_AB, decode=False)
                t_e.append(max_over_ranks(time.perf_counter() - t0))
            edges_extra = {"k": K, "n_nodes": sum_over_ranks(ed["n_nodes"]), "n_edges": sum_over_ranks(ed["n_edges"]),
                           "checksum": sum_u64_over_ranks(ed["checksum"]),
                           "ms": round(1e3 * t_e[1], 3), "d2h_bytes": ed["n_edges"] * 32,
                           "timer": "host wall clock around mdbg_edges_index incl. the D2H of keys and values, max over ranks"}
            if ed.get("raw_values") is not None:
                edges_extra["branching_keys"] = int((((ed["raw_values"] >> np.uint64(34)) & np.uint64(1)) == 1).any(axis=1).sum())
        except Exception as e:                                # noqa: BLE001
            edges_extra = {"error": repr(e)}

    # ---- extra: unitig nodes of the same node set (computeUnitigNodes + computeDeterministicUnitigs): edge set, links,
    # list ranking, sequences, hashes on the device; the CSR's D2H and the host's sort of the unitig hashes are inside
    unitigs_extra = None
    if not args.no_edges and w["last_k"] == K and world == 1:
        try:
            t_u = []
            for _ in range(2):
                t0 = time.perf_counter()
                un = eng.unitigs_build(MIN_AB, copy=False)
                t_u.append(time.perf_counter() - t0)
            lens = np.diff(un["offsets"]).astype(np.int64)
            unitigs_extra = {"k": K, "n_nodes": un["n_nodes"], "n_unitigs": un["n_unitigs"], "n_circular": un["n_circular"],
                             "n_minimizers": int(un["offsets"][-1]), "longest_unitig_nodes": int(lens.max() - (K - 1)) if len(lens) else 0,
                             "ms": round(1e3 * t_u[1], 3), "d2h_bytes": int(un["offsets"][-1]) * 4 + un["n_unitigs"] * 25,
                             "checksum_of_hashes": int(np.sum(un["hashes"][:, 0], dtype=np.uint64)) if un["n_unitigs"] else 0,
                             "n_unitig_edges": un["n_unitig_edges"], "checksum_unitig_edges": un["checksum_edges"],
                             "checksum_unitig_nodes": un["checksum_nodes"], "checksum_unitig_abundances": un["checksum_abundances"],
                             "timer": "host wall clock around mdbg_unitigs_build (edge set + links + list ranking + sequences + hash128 + "
                                      "deterministic order + unitig graph edges on the device, D2H of the CSRs)"}
        except Exception as e:                                # noqa: BLE001
            unitigs_extra = {"error": repr(e)}

    # ---- extra (not the headline): the multi-k loop k = 4 .. 21 on the resident store (BASELINE config 4's shape on
    # this workload): k = 4 counted, every further k derived from the previous table on the device; collective
    # previous-k replication + value merge for N > 1
    multi_k = None
    if args.multi_k > K and w["last_k"] == K:
        try:
            from metamdbg_b200 import multi_k_sweep
            mk_merge = "hashes" if world > 1 else False
            sweeps, sweep_allocs = [], []
            for _ in range(3):                                  # the last one is timed: buffers have reached their sizes
                a0 = eng.allocations()
                sweeps.append(multi_k_sweep(eng, K, args.multi_k, MIN_AB, merge=mk_merge, world=world))
                a1 = eng.allocations()
                sweep_allocs.append({"device_allocations": a1[0] - a0[0], "table_buffer_trades": a1[1] - a0[1]})
            per_k = [round(1e3 * max_over_ranks(r["seconds"]), 3) for r in sweeps[-1]]
            total_s = sum(per_k) * 1e-3
            multi_k = {"k_first": K, "k_last": args.multi_k, "ms_per_k": per_k, "ms_total": round(sum(per_k), 3),
                       "value": total_bases / total_s / 1e9, "unit": "Gbp/s (input bases / time of the k-loop alone; the "
                       "sketch is not repeated, as in the reference)",
                       "n_entries_total": [sum_over_ranks(r["n_entries"]) for r in sweeps[-1]],
                       "timer": "host wall clock per k around device work ending in a D2H of the table statistics, max over ranks",
                       "ms_per_k_first_sweep": [round(1e3 * max_over_ranks(r["seconds"]), 3) for r in sweeps[0]],
                       "allocations_rank0": {"first_sweep": sweep_allocs[0], "second_sweep": sweep_allocs[1], "timed_sweep": sweep_allocs[-1]},
                       "same_tables_both_sweeps": all([r["checksum"] for r in sw] == [r["checksum"] for r in sweeps[0]] for sw in sweeps[1:])}
            # where the loop's time goes: one more sweep with the library's phase profile on (exclusive phase times, every
            # phase boundary synchronises the stream -- a diagnostic, slower than the timed sweep above)
            eng.phase_profile(True)
            multi_k_sweep(eng, K, args.multi_k, MIN_AB, merge=mk_merge, world=world)
            multi_k["phase_ms_profiled_sweep_rank0"] = eng.phase_times()
            multi_k["merge"] = "k = 4: owner merge with vectors; k > 4: keys-only owner merge (mdbg_count_merge_hashes)" if world > 1 else None
            eng.phase_profile(False)
        except Exception as e:                                # noqa: BLE001  -- an extra must not cost the headline line
            multi_k = {"error": repr(e)}

    # ---- end to end through the host-buffer C ABI ---------------------------------------------
    e2e = None
    if not args.no_e2e and len(dr.chunks) == 1:
        rs_off = c0["offsets"]
        e_reads = min(n_reads, args.e2e_reads or n_reads)
        # the pinned host copy of this rank's reads must fit beside the other ranks' (one node): use at most half of
        # what the host still has, split over the local ranks; fewer reads in the e2e leg is reported, not hidden
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        share = host_bytes_available() // (2 * max(1, local_world))
        while e_reads > args.e2e_batch and int(rs_off[e_reads]) > share:
            e_reads = max(args.e2e_batch, e_reads // 2)
        h_bases = None
        while h_bases is None:
            e_bases = int(rs_off[e_reads])
            try:
                h_bases = torch.empty(e_bases, dtype=torch.uint8, pin_memory=True)
            except RuntimeError:
                if e_reads <= 1024:
                    raise
                e_reads //= 2
        h_bases.copy_(c0["bases"][:e_bases])
        h_offs = rs_off[:e_reads + 1].copy()
        torch.cuda.synchronize()
        batch = args.e2e_batch
        d2h = [0]

        def step_e2e():
            d2h[0] = 0
            eng.store_clear()
            for lo in range(0, e_reads, batch):
                hi = min(e_reads, lo + batch)
                offs = h_offs[lo:hi + 1] - h_offs[lo]
                sk = eng.sketch_batch_ptr(h_bases.data_ptr() + int(h_offs[lo]), offs, True)
                d2h[0] += 8 * (hi - lo + 1) + 9 * sk
            if w["asm_density"]:
                eng.store_apply_density(w["asm_density"])
            eng.purge_palindromes(4, lk)
            eng.count_begin(K, 0)
            eng.count_add_store()
            if world > 1:
                eng.count_merge()
            tab = eng.count_finalize(MIN_AB)
            d2h[0] += len(tab.abundances) * (16 + 4 + 4 * K)
            return tab

        # parity of every e2e step (warm-up steps included): the finalised table of the host-buffer path must carry
        # the checksum of the device-resident leg (same reads, same merge); reported, never silently skipped
        comparable = e_reads == n_reads and w["last_k"] == K
        e2e_checks = []
        for _ in range(max(1, min(args.warmup, 2))):
            tab = step_e2e()
            e2e_checks.append(tab.checksum == checksum_local)
        e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        moved0 = eng.bytes_moved()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            tab = step_e2e()
            e2e_checks.append(tab.checksum == checksum_local)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        bad_steps = sum_over_ranks(sum(1 for ok in e2e_checks if not ok)) if comparable else None
        moved1 = eng.bytes_moved()
        e_total = sum_over_ranks(e_bases)
        e2e = {"value": e_total * e_steps / dt / 1e9, "unit": "Gbp/s",
               "h2d_bytes_per_step": int((moved1[0] - moved0[0]) // e_steps),
               "d2h_bytes_per_step": int((moved1[1] - moved0[1]) // e_steps),
               "host_input_bytes_per_step": int(e_bases + 8 * (e_reads + 1)),
               "transfer": "ASCII reads are 2-bit packed by the library's host threads before H2D (ASCII + device-side pack "
                           "pass when the process has too few CPUs); scan, compaction and the CSR's D2H run piece by piece "
                           "behind each piece's sketch",
               "steps": e_steps, "reads_per_gpu": e_reads, "host_batch_reads": batch,
               "last_host_batch": eng.last_batch_info(),
               "timer": "host wall clock around synchronous C-ABI calls, max over ranks",
               "table_checks": {"steps_checked": len(e2e_checks) * world if comparable else 0,
                                "steps_differing_from_device_leg": bad_steps,
                                "what": "checksum (sum abundance*hash) of the finalised table of every e2e step, warm-up "
                                        "included, against the device-resident leg on the same reads"}}
        if comparable and bad_steps:
            print(f"bench.py: WARNING: {bad_steps} e2e step(s) produced a table that differs from the device-resident leg",
                  file=sys.stderr, flush=True)
        del h_bases
        # extra: the same leg with a host batch that is ALREADY 2-bit packed (mdbg_sketch_batch_packed: what a reader
        # that packs while it parses hands over) -- PCIe carries a quarter of the bytes and no library thread packs
        try:
            src_all = np.frombuffer(c0["src"].cpu().numpy().tobytes(), dtype=np.uint64)[:e_reads + 1 if e_reads < n_reads else e_reads]
            n_w = int(eng.pack_device_words(int(rs_off[e_reads]), e_reads))
            h_words = torch.empty(n_w * 4, dtype=torch.uint8, pin_memory=True)
            h_words.copy_(c0["words"][:n_w * 4])
            torch.cuda.synchronize()

            def step_e2e_packed():
                eng.store_clear()
                for lo in range(0, e_reads, batch):
                    hi = min(e_reads, lo + batch)
                    offs = h_offs[lo:hi + 1] - h_offs[lo]
                    w_lo = int(src_all[lo])
                    w_hi = int(src_all[hi]) if hi < len(src_all) else n_w
                    eng.sketch_batch_packed_ptr(h_words.data_ptr() + 4 * w_lo, w_hi - w_lo, src_all[lo:hi] - np.uint64(w_lo), offs, True)
                if w["asm_density"]:
                    eng.store_apply_density(w["asm_density"])
                eng.purge_palindromes(4, lk)
                eng.count_begin(K, 0)
                eng.count_add_store()
                if world > 1:
                    eng.count_merge()
                return eng.count_finalize(MIN_AB)

            ok_p = [step_e2e_packed().checksum == checksum_local for _ in range(2)]
            barrier()
            moved0 = eng.bytes_moved()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                ok_p.append(step_e2e_packed().checksum == checksum_local)
            barrier()
            dtp = max_over_ranks(time.perf_counter() - t0)
            moved1 = eng.bytes_moved()
            e2e["packed_host_input"] = {"value": e_total * e_steps / dtp / 1e9, "unit": "Gbp/s",
                                        "h2d_bytes_per_step": int((moved1[0] - moved0[0]) // e_steps),
                                        "d2h_bytes_per_step": int((moved1[1] - moved0[1]) // e_steps),
                                        "same_table_as_device_leg": (all(ok_p) if comparable else None),
                                        "what": "mdbg_sketch_batch_packed on pinned 2-bit words (16-byte aligned reads), otherwise the same leg"}
            del h_words
        except Exception as e:                                # noqa: BLE001
            e2e["packed_host_input"] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N=1 only) + parity spot check against it ------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle
        cores = os.cpu_count() or 1
        try:
            ref = pyoracle.Reference()
            kind, threads = "reference", min(cores, 2 * usable_cpus())
        except (FileNotFoundError, OSError):
            ref, kind, threads = None, "port", 1
        n_sample = int(min(c0["n"], args.cpu_reads or (200_000 if ref is not None else 2_000)))
        s_offs = c0["offsets"][:n_sample + 1].copy()
        s_bases = c0["bases"][:int(s_offs[-1])].cpu().numpy()
        t0 = time.perf_counter()
        if ref is not None:
            res = ref.pipeline(s_bases, s_offs, L, w["density"], w["hpc"], K, purge_last_k=lk, min_abundance=MIN_AB,
                               threads=threads, assembly_density=w["asm_density"])
        else:
            res = oracle_port_pipeline(pyoracle.Oracle(), s_bases, s_offs, lk, w)
        ref_solid, ref_cs, ref_nmin = res["n_solid"], res["checksum"], res["n_minimizers"]
        dt = time.perf_counter() - t0
        # same sample through the GPU engine (packed-resident input): bit-exact fingerprint must agree
        eng.store_clear()
        eng.sketch_batch_device_packed2(c0["words"].data_ptr(), c0["src"].data_ptr(), c0["bases"].data_ptr(), c0["off"].data_ptr(),
                                        n_sample, int(s_offs[-1]), True)
        if w["asm_density"]:
            eng.store_apply_density(w["asm_density"])
        eng.purge_palindromes(4, lk)
        eng.count_begin(K, 0)
        eng.count_add_store()
        g = eng.count_stats(MIN_AB)
        ok = (g["n_entries"] == ref_solid and g["checksum"] == ref_cs and eng.store_size()[1] == ref_nmin)
        if not ok:
            raise SystemExit(f"bench.py: GPU result differs from the CPU {kind} on the sample: {g} vs "
                             f"{ref_solid}/{ref_cs}/{ref_nmin}")
        cpu_baseline = {"value": int(s_offs[-1]) / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": kind,
                        "sample": f"first {n_sample} reads ({int(s_offs[-1]) / 1e9:.3f} Gbp), one pass of the reference's "
                                  f"sketch + purge + count code driven in memory, "
                                  f"GPU fingerprint (n_minimizers, n_solid, checksum) identical"}
        # the graph side of the same sample: the reference's indexEdges + computeUnitigNodes + computeDeterministicUnitigs
        # on the sample's node set, against mdbg_unitigs_build -- the records of unitigGraph.nodes.bin must be identical
        if ref is not None and unitigs_extra is not None and "error" not in unitigs_extra:
            try:
                nodes = eng.count_finalize(MIN_AB).kminmers
                t0 = time.perf_counter()
                gu = eng.unitig_records(MIN_AB)
                t_gpu = time.perf_counter() - t0
                t0 = time.perf_counter()
                ru = ref.unitig_nodes(nodes, K, threads=threads)
                t_ref = time.perf_counter() - t0
                same = bool(np.array_equal(gu["offsets"], ru["offsets"]) and np.array_equal(gu["minimizers"], ru["minimizers"]))
                t0 = time.perf_counter()
                re_ = ref.unitig_edges(ru["offsets"], ru["minimizers"], K, threads=threads)
                t_ref_e = time.perf_counter() - t0
                lists = lambda d: [sorted(d["targets"][int(d["offsets"][x]):int(d["offsets"][x + 1])].tolist())
                                   for x in range(len(d["offsets"]) - 1)]
                same_e = bool(gu["n_unitig_edges"] == re_["n_edges"] and gu["checksum_edges"] == re_["checksum"] and
                              lists(dict(offsets=gu["edge_offsets"], targets=gu["edge_targets"])) == lists(re_))
                same = same and same_e
                if not same:
                    raise SystemExit("bench.py: the unitigs / unitig edges of the GPU differ from the reference's on the sample")
                unitigs_extra["cpu_reference_on_sample"] = {
                    "n_nodes": int(len(nodes)), "n_unitigs": int(len(ru["offsets"]) - 1), "seconds": round(t_ref, 3), "threads": threads,
                    "n_unitig_edges": int(re_["n_edges"]), "seconds_unitig_edges": round(t_ref_e, 3), "identical_edge_lists": same_e,
                    "gpu_seconds_same_sample": round(t_gpu, 4), "identical_records": same,
                    "what": "CreateMdbg::indexEdges + computeUnitigNodes + computeDeterministicUnitigs of oracle/_ref through their "
                            "file contract in a scratch directory, node set of the cpu_baseline sample; then indexUnitigEdges + "
                            "computeUnitigEdges on those records (lists compared as multisets: the reference's order inside a list "
                            "is its threads' arrival order)"}
            except SystemExit:
                raise
            except Exception as e:                            # noqa: BLE001
                unitigs_extra["cpu_reference_on_sample"] = {"error": repr(e)}

    # ---- extras: the other BASELINE.json configs ---------------------------------------------------------------------
    extras = {}
    want_extras = [x for x in args.extras.split(",") if x] if args.extras != "auto" else (
        (["cfg3"] if world == 1 else []) + ["cfg4", "cfg5"]
        if (name == "cfg2" and args.reads == w["reads"] and args.read_len == w["read_len"]) else [])
    if want_extras:
        del dr, c0
        if hasattr(torch.cuda, "empty_cache"):
            torch.cuda.empty_cache()
    for xn in want_extras:
        if xn == name or xn not in WORKLOADS:
            continue
        try:
            extras[xn] = run_extra(args, xn, torch, dev, rank, world, eng if WORKLOADS[xn]["density"] == w["density"] and
                                   WORKLOADS[xn]["hpc"] == w["hpc"] else None, new_engine, timed, sum_over_ranks,
                                   sum_u64_over_ranks)
        except Exception as e:                                # noqa: BLE001  -- an extra must not cost the headline line
            extras[xn] = {"error": repr(e)}
        if hasattr(torch.cuda, "empty_cache"):
            torch.cuda.empty_cache()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        sk_ms = float(np.mean(sketch_ms))
        kernel = {2: "sketch_packed_kernel<15>", 1: "sketch_kernel<15, 1>", 0: "sketch_kernel<15, 0>"}[active_variant]
        # SURVEY 8d / DESIGN.md: 0.25 B per input base read (2-bit packed) + 9 B per selected minimizer written
        algo_bytes = n_bases * 0.25 + n_sketched[0] * 9.0
        achieved = algo_bytes / (sk_ms * 1e-3) / 1e9
        traffic, ncu_pipes, int_issue = None, None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # only a capture of THIS kernel at THIS launch shape may annotate the line
            if tj.get("kernel") == kernel and tj.get("reads") == dr_reads_of(args, w, world) and len(sketch_ms):
                traffic = tj.get("dram_bytes_per_launch")
                ncu_pipes = tj.get("ncu")
                inst = tj.get("warp_instructions_per_launch")
                if inst and clocks and clocks.get("sm_mhz"):
                    peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6          # warp instructions per second: 4 schedulers per SM
                    ach = inst / (sk_ms * 1e-3)
                    int_issue = {"what": "integer-issue roofline (SURVEY 8d): warp instructions executed per launch (ncu "
                                         "smsp__inst_executed.sum of the committed capture) / live launch duration, against one "
                                         "instruction per scheduler per cycle at the sampled SM clock",
                                 "achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": peak_issue, "frac": ach / peak_issue,
                                 "thread_instructions_per_lmer": tj.get("thread_instructions_per_lmer")}
        # the second kernel of the step (K3, the k-min-mer insert pass): SURVEY 8d counts one 32-byte sector read and one
        # 32-byte write per k-min-mer instance; the local (pre-merge) pass of this rank
        table_roofline = None
        ins_ms = float(np.mean(insert_ms)) if len(insert_ms) else 0.0
        if ins_ms > 0 and local_windows:
            n_windows = local_windows
            t_ach = n_windows * 64.0 / (ins_ms * 1e-3) / 1e9
            table_roofline = {"bound": "hbm", "kernel": "insert_warp_kernel<4>", "achieved": t_ach, "peak": peak, "unit": "GB/s",
                              "frac": t_ach / peak, "ms_per_launch": ins_ms, "algorithmic_bytes_per_launch": n_windows * 64.0,
                              "windows_per_launch": n_windows, "share_of_step": ins_ms / (ms_total / args.steps),
                              "note": "every access is a random 32-byte sector of a table larger than L2: the pass is bound by "
                                      "DRAM's random-access rate, not its byte rate (rate against table size: DESIGN.md K3, "
                                      "profiles/r02_passes_ncu.txt)"}
        line = {
            "metric": METRIC, "value": value, "unit": "Gbp/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak" if w["per_gpu"] else "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": workload_config(name, args.reads, args.read_len, args.genomes, world), "e2e": e2e,
            "gpu_launches": int(launches), "device_allocations_in_timed_region": int(allocs_timed), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "ms_per_launch": sk_ms, "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "the kernel is bound by integer issue (one MurmurHash3_x64_128 per l-mer), not by HBM: see "
                                 "int_issue and DESIGN.md",
                         "binding_pipes_ncu": ncu_pipes, "int_issue": int_issue,
                         "share_of_step": sk_ms / (ms_total / args.steps)},
            "roofline_table_pass": table_roofline,
            "kernels_ms": {"sketch": sk_ms, "insert": float(np.mean(insert_ms))},
            "table_phase_ms_profiled_step_rank0": step_phases,
            "ascii_resident": ascii_leg, "multi_k": multi_k, "edges": edges_extra, "unitigs": unitigs_extra,
            "extras": extras,
            "sketch_autotune": dict(tune, active=active_variant,
                                    note="ms = sketch (+ pack pass for variant 2) + scan + compaction of an ASCII-resident batch, "
                                         "best of 2; every variant's whole output must equal variant 0's on the device"),
            "cpu_baseline": cpu_baseline,
            "check": {"n_minimizers_rank0": int(n_min_store), "n_solid_total": int(solid_total),
                      "checksum_rank0": int(checksum_local), "checksum_total": int(checksum_total),
                      "kminmer_occurrences_total": occ, "occurrences_conserved": occ is not None,
                      "device_steps_same_checksum": len(step_checksums) == 1,
                      "multi_gpu_parity": parity},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def dr_reads_of(args, w, world):
    return args.reads if w["per_gpu"] else args.reads // max(1, world)


def run_extra(args, xn, torch, dev, rank, world, eng, new_engine, timed, sum_over_ranks, sum_u64_over_ranks):
    """Another BASELINE.json config through the same step (strong scaling for cfg4 / cfg5: the read count is the total)."""
    w = WORKLOADS[xn]
    scale = args.extra_scale
    reads = max(64 * world, int(w["reads"] * scale))
    reads_total = reads * world if w["per_gpu"] else reads
    read_len = w["read_len"] if scale >= 0.01 else max(1500, w["read_len"] // 4)
    lk = purge_last_k(read_len, w)
    own = eng is None
    if own:
        eng = new_engine(w)
    t0 = time.perf_counter()
    dr = DeviceReads(torch, dev, eng, make_readsets(xn, reads_total, read_len, args.genomes), rank, world, args.chunk_reads,
                     keep_ascii=False)
    t_gen = time.perf_counter() - t0
    last_k = min(w["last_k"], args.multi_k) if w["last_k"] > K else K
    out = {}

    def step():
        tab, n_min, per_k = hot_path(eng, dr, w, lk, world, last_k=last_k)
        out.update(tab=tab, n_min=n_min, per_k=per_k)
        return tab

    ms, tab = timed(step, 3, 2)
    total_bases = sum_over_ranks(dr.n_bases)
    variant = None
    if world > 1 and last_k > K + 1:
        # the same loop with the tables of the intermediate k's left rank-local (merge at k = 4 and at the last k only)
        def step_local():
            return hot_path(eng, dr, w, lk, world, last_k=last_k, merge_every_k=False)[0]
        ms_l, tab_l = timed(step_local, 2, 2)
        variant = {"value": total_bases / (ms_l * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms_l,
                   "checksum_total_last_k": sum_u64_over_ranks(tab_l["checksum"]),
                   "same_last_table_as_per_k_merge": sum_u64_over_ranks(tab_l["checksum"]) == sum_u64_over_ranks(tab["checksum"]),
                   "what": "tables of k = 5 .. last_k - 1 stay rank-local (k-min-mers of the rank's own reads with their global "
                           "abundances; the next pass reads per-position values, no table); owner merge at k = 4 and at the last k"}
    res = {"config": workload_config(xn, reads, read_len, args.genomes, world), "scaling": "weak" if w["per_gpu"] else "strong",
           "value": total_bases / (ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms, "steps": 2, "warmup": 3,
           "n_bases_total": int(total_bases), "n_minimizers_total": sum_over_ranks(out["n_min"]),
           "n_entries_per_k_total": [sum_over_ranks(x) for x in out["per_k"]], "k_last": last_k,
           "checksum_total_last_k": sum_u64_over_ranks(tab["checksum"]), "resident_chunks_rank0": len(dr.chunks),
           "generate_and_pack_s": round(t_gen, 2), "merge_at_first_and_last_k_only": variant,
           "merge": ("k = 4: owner merge with vectors; every k > 4: keys-only owner merge (hash128, abundance)" if world > 1 and last_k > K else
                     "owner merge with vectors" if world > 1 else None),
           "timer": "CUDA events on the launching stream around 2 steps after 3 warm-up steps, barriers on both sides, max over ranks"}
    del dr
    if own:
        eng.close()
    return res


def multi_gpu_parity(args, eng, torch, dev, rank, world, sum_over_ranks, sum_u64_over_ranks, barrier):
    """All ranks shard ONE read set and run count -> merge -> rescue -> previous-k -> two next-k passes -> edge index
    over real NCCL; per stage the all-reduced (entries, checksum) pair is compared on rank 0 with the CPU oracle of the
    WHOLE set (oracle/: the C restatement, itself pinned against the reference's own code in tests/test_oracle.py).
    Any mismatch ends the run."""
    from metamdbg_b200 import synth
    n_reads, read_len, last = args.parity_reads, args.parity_read_len, 80
    rs = synth.make_readset(n_reads, read_len, seed=4242, n_genomes=3, genome_len_range=(300_000, 700_000), err=0.004)
    bases, offs = synth.fill_reads(rs.shard(rank, world))
    eng.store_clear()
    eng.sketch_batch(bases, offs, append_to_store=True, fetch=False)
    eng.purge_palindromes(4, last)
    got = {}

    def stage(tag, extra=0):
        st = eng.count_stats(0)
        got[tag] = (sum_over_ranks(st["n_entries"]), sum_u64_over_ranks(st["checksum"]), sum_over_ranks(extra))

    eng.count_begin(K, 0)
    eng.count_add_store()
    eng.count_merge()
    st2 = eng.count_stats(2)
    got["count"] = (sum_over_ranks(st2["n_entries"]), sum_u64_over_ranks(st2["checksum"]), sum_over_ranks(st2["n_distinct"]))
    n_resc = eng.count_rescue()
    stage("rescue", n_resc)
    ed = eng.edges_index(0)
    got["edges"] = (sum_over_ranks(ed["n_edges"]), sum_u64_over_ranks(ed["checksum"]), sum_over_ranks(ed["n_nodes"]))
    for k in (K + 1, K + 2):
        eng.prev_from_current(0)
        eng.count_begin(k, 0)
        eng.count_add_store_next_k()
        eng.count_merge()
        stage(f"k{k}")
    # the same chain once more with the keys-only merge of the multi-k loop (k > 4)
    eng.count_begin(K, 0)
    eng.count_add_store()
    eng.count_merge()
    eng.count_rescue()
    for k in (K + 1, K + 2, K + 3):
        eng.prev_from_current(0)
        eng.count_begin(k, 0)
        eng.count_add_store_next_k()
        eng.count_merge_hashes()
        stage(f"k{k}_keys_only_merge")
    barrier()
    verdict = None
    if rank == 0:
        from oracle.pyoracle import Oracle
        orc = Oracle()
        ab, ao = synth.fill_reads(rs)
        mo, m, _, _ = orc.sketch_batch(ab, ao, L, WORKLOADS[args.workload]["density"], WORKLOADS[args.workload]["hpc"])
        pm, po = [], [0]
        for r in range(rs.n_reads):
            q, _ = orc.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, last)
            pm.append(q); po.append(po[-1] + len(q))
        pm = np.concatenate(pm).astype(np.uint32); po = np.array(po, np.uint64)
        allk = orc.count(pm, po, K, keep_all=True)
        solid = orc.count(pm, po, K, 2)
        resc = orc.rescue(pm, po, K, solid["hashes"], solid["abundances"])
        ph = np.concatenate([solid["hashes"], resc["hashes"]]) if len(resc["hashes"]) else solid["hashes"]
        pa = np.concatenate([solid["abundances"], np.ones(len(resc["hashes"]), np.uint32)])
        pv = np.concatenate([solid["vecs"], resc["vecs"]]) if len(resc["hashes"]) else solid["vecs"]
        want = {"count": (len(solid["abundances"]), orc.checksum(solid["hashes"], solid["abundances"]), len(allk["abundances"])),
                "rescue": (len(pa), orc.checksum(ph, pa), resc["n_reads_rescued"])}
        we = orc.edge_index(pv, K)
        want["edges"] = (len(we["hashes"]), we["checksum"], len(pa))
        for k in (K + 1, K + 2, K + 3):
            nk = orc.next_k(pm, po, k, ph, pa)
            ph, pa = nk["hashes"], nk["abundances"]
            if k <= K + 2:
                want[f"k{k}"] = (len(pa), orc.checksum(ph, pa), 0)
            want[f"k{k}_keys_only_merge"] = (len(pa), orc.checksum(ph, pa), 0)
        verdict = {t: tuple(int(x) for x in got[t]) == tuple(int(x) for x in want[t]) for t in want}
        verdict["reads"] = n_reads
        verdict["what"] = ("(entries, sum abundance*hash, third figure: distinct keys / rescued reads / nodes) of every stage, "
                           "all-reduced over the ranks, against the CPU oracle of the whole read set")
        if not all(v for t, v in verdict.items() if t in want):
            raise SystemExit(f"bench.py: multi-GPU parity FAILED: got {got} want {want}")
    return verdict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS), help="BASELINE.json config of the headline (default: the one the metric is quoted on)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (cfg2/cfg3) or in total (cfg4/cfg5); 0 = the workload's")
    ap.add_argument("--read-len", type=int, default=0, help="mean read length (0 = the workload's)")
    ap.add_argument("--genomes", type=int, default=100)
    ap.add_argument("--err", type=float, default=-1.0, help="substitution rate of the synthetic reads (-1 = the workload's; experiments)")
    ap.add_argument("--chunk-reads", type=int, default=1_000_000, help="reads per resident chunk (device generation / sketch call)")
    ap.add_argument("--extras", default="auto", help="comma list of other configs to run as extras (auto: cfg3 at N=1, cfg4, cfg5 with the default headline; '' = none)")
    ap.add_argument("--extra-scale", type=float, default=1.0, help="shrink the extras' read counts (tests)")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per GPU in the e2e leg (0 = all)")
    ap.add_argument("--e2e-batch", type=int, default=262_144, help="reads per host-buffer C-ABI call")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-reads", type=int, default=0, help="sample size of the reference arm (0 = 100000)")
    ap.add_argument("--cpu-reads", type=int, default=0, help="sample size of the cpu_baseline leg (0 = 200000)")
    ap.add_argument("--parity-reads", type=int, default=20_000, help="N>1 parity leg: reads of the shared set")
    ap.add_argument("--parity-read-len", type=int, default=6_000)
    ap.add_argument("--sketch-variant", type=int, default=-1, help="force a sketch-kernel variant (-1 = the default, 2)")
    ap.add_argument("--multi-k", type=int, default=21, help="multi-k loops run up to this k (0 = off)")
    ap.add_argument("--no-autotune", action="store_true", help="skip the variant comparison of the sketch kernel (profiling runs)")
    ap.add_argument("--no-edges", action="store_true", help="skip the edge-key extra")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ascii-leg", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the parity leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="reference arm: time the in-memory harness instead of the stock stages")
    ap.add_argument("--no-ref-multi-k", action="store_true", help="reference arm: skip the multi-k extra")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.err >= 0:
        for wl in WORKLOADS.values():
            wl["err"] = args.err
    w = WORKLOADS[args.workload]
    if args.reads <= 0:
        args.reads = w["reads"]
    if args.read_len <= 0:
        args.read_len = w["read_len"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)       # timing rule: at least 3 warm-up steps
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
