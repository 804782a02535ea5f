"""Mint golden input/output vectors for the sketch+count path from the
REFERENCE'S OWN CODE (oracle/_ref/libmdbg_ref.so = metaMDBG sources compiled
where they lie under /root/reference; see oracle/ref_shim.cpp).

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
The resulting tests/golden/*.npz files are committed; /root/reference is never
read at test time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from metamdbg_b200 import synth  # noqa: E402
from oracle.pyoracle import Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sketch_all(ref, bases, offs, l, d, hpc, bl=None):
    raw = bases.tobytes()
    ms, ps, ds, mo = [], [], [], [0]
    for r in range(len(offs) - 1):
        m, p, dd = ref.sketch_read(raw[int(offs[r]):int(offs[r + 1])], l, d, hpc, bl)
        ms.append(m); ps.append(p); ds.append(dd); mo.append(mo[-1] + len(m))
    cat = lambda xs, t: np.concatenate(xs).astype(t) if xs else np.zeros(0, t)
    return np.array(mo, np.uint64), cat(ms, np.uint32), cat(ps, np.uint32), cat(ds, np.uint8)


def unitig_cases(seed=77):
    """Minimizer-space read sets whose node sets exercise the unitig walk: long clean unitigs, small alphabets (palindromic
    keys, hairpins, branching), circular genomes.  Every read is taken twice so that every window is solid at k."""
    rng = np.random.default_rng(seed)
    cases = []
    for it in range(16):
        k = int([3, 4, 5, 6, 8, 4, 21, 4][it % 8])
        mode = it % 4
        reads = []
        if mode == 0:
            g = rng.integers(1, 1 << 30, size=1200, dtype=np.uint32)
            for _ in range(120):
                st = int(rng.integers(0, 1100)); r = g[st:st + 90].copy()
                if rng.random() < 0.5: r = r[::-1]
                if rng.random() < 0.3: r[int(rng.integers(0, 90))] = rng.integers(1, 1 << 30)
                reads.append(r)
        elif mode == 1:
            g = rng.integers(1, 12, size=300, dtype=np.uint32)
            for _ in range(80):
                st = int(rng.integers(0, 260)); r = g[st:st + 40].copy()
                if rng.random() < 0.5: r = r[::-1]
                reads.append(r)
        elif mode == 2:
            for _ in range(40): reads.append(rng.integers(1, 4, size=k + 8, dtype=np.uint32))
        else:
            for _ in range(3):
                L = int(rng.integers(k, 200))
                g = rng.integers(1, 1 << 30, size=L, dtype=np.uint32)
                reads.append(np.concatenate([g, g[:k - 1]]))
                if rng.random() < 0.5: reads.append(np.concatenate([g, g, g[:k - 1]])[::-1])
            reads.append(rng.integers(1, 1 << 30, size=150, dtype=np.uint32))
        reads = [np.ascontiguousarray(r, dtype=np.uint32) for r in reads + reads]
        offs = np.zeros(len(reads) + 1, np.uint64)
        offs[1:] = np.cumsum([len(r) for r in reads])
        cases.append((k, np.concatenate(reads).astype(np.uint32), offs))
    return cases


def mint_unitigs(ref):
    """7. unitig nodes and unitig graph edges: the reference's own indexEdges + computeUnitigNodes +
    computeDeterministicUnitigs (oracle/ref_shim.cpp: ref_unitig_nodes) on the node set the reference's count gives for
    each case, then its indexUnitigEdges + computeUnitigEdges (ref_unitig_edges) on those records."""
    out = {}
    cases = unitig_cases()
    out["n_cases"] = np.array(len(cases))
    for i, (k, mins, offs) in enumerate(cases):
        nodes = ref.count(mins, offs, k, 2, threads=2)["vecs"]
        u = ref.unitig_nodes(nodes, k, threads=1 + (i % 2) * 3)
        out[f"c{i}_k"] = np.array(k); out[f"c{i}_minimizers"] = mins; out[f"c{i}_offsets"] = offs
        out[f"c{i}_nodes"] = nodes
        out[f"c{i}_unitig_offsets"] = u["offsets"]; out[f"c{i}_unitig_minimizers"] = u["minimizers"]
        e = ref.unitig_edges(u["offsets"], u["minimizers"], k, threads=1)      # one thread: the file's list order is deterministic
        out[f"c{i}_edge_offsets"] = e["offsets"]; out[f"c{i}_edge_targets"] = e["targets"]
        out[f"c{i}_edge_stats"] = np.array([e["n_edges"], e["checksum"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "minspace_unitigs.npz"), **out)


def main():
    ref = Reference()
    if "--only-unitigs" in sys.argv:
        mint_unitigs(ref)
        return

    # 1. HiFi-like: HPC on, l=15, d=0.005, k=4 count, abundance >= 2 (BASELINE config 0, scaled down)
    rs = synth.make_readset(600, 6000, seed=101, n_genomes=2, genome_len_range=(90_000, 140_000), err=0.001)
    bases, offs = synth.fill_reads(rs)
    mo, m, p, d = sketch_all(ref, bases, offs, 15, 0.005, True)
    purged, po = [], [0]
    for r in range(len(mo) - 1):
        q = ref.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, 60)
        purged.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(purged).astype(np.uint32)
    po = np.array(po, np.uint64)
    tabs = {}
    for k in (4, 5, 7):
        c = ref.count(pm, po, k, 2, threads=2)
        tabs[f"k{k}_hashes"] = c["hashes"]; tabs[f"k{k}_abund"] = c["abundances"]; tabs[f"k{k}_vecs"] = c["vecs"]
        tabs[f"k{k}_stats"] = np.array([c["n_instances"], c["n_distinct"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "hifi_small.npz"), seed=101, bases=bases, offsets=offs,
                        min_offsets=mo, minimizers=m, positions=p, directions=d,
                        purged_offsets=po, purged_minimizers=pm, **tabs)

    # 2. ONT-like: HPC off, d=0.025, 2 % substitutions, with a repetitive-minimizer blacklist
    rs = synth.make_readset(300, 4000, seed=202, n_genomes=1, genome_len_range=(150_000, 150_001), err=0.02)
    bases, offs = synth.fill_reads(rs)
    mo0, m0, _, _ = sketch_all(ref, bases, offs, 15, 0.025, False)
    vals, cnts = np.unique(m0, return_counts=True)
    bl = vals[np.argsort(-cnts, kind="stable")[:5]].astype(np.uint32)
    mo, m, p, d = sketch_all(ref, bases, offs, 15, 0.025, False, bl)
    c = ref.count(m, mo, 4, 2, threads=2)
    np.savez_compressed(os.path.join(HERE, "ont_small.npz"), seed=202, bases=bases, offsets=offs, blacklist=bl,
                        min_offsets=mo, minimizers=m, positions=p, directions=d,
                        k4_hashes=c["hashes"], k4_abund=c["abundances"], k4_vecs=c["vecs"],
                        k4_stats=np.array([c["n_instances"], c["n_distinct"]], np.uint64))

    # 3. edge cases: N, lower case, '#', empty / short reads, homopolymers, several l and densities
    rng = np.random.default_rng(303)
    reads = [b"", b"A", b"ACGT", b"A" * 300, b"AC" * 200, b"ACGTN" * 80, b"acgtACGT" * 60, b"#" * 10 + b"ACGGT" * 40,
             b"ACG#TTGCA" * 50, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 15)),
             bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 16)),
             bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 17))]
    for _ in range(40):
        ln = int(rng.integers(1, 3000))
        s = rng.choice(np.frombuffer(b"ACGTACGTACGTNacgt", np.uint8), ln)
        rep = rng.integers(1, 5, size=ln)
        reads.append(bytes(np.repeat(s, rep)[:ln]))
    bases = np.frombuffer(b"".join(reads), np.uint8).copy()
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    out = dict(bases=bases, offsets=offs)
    cases = []
    for l in (5, 11, 15, 16):
        for dens in (0.005, 0.05, 0.6):
            for hpc in (0, 1):
                mo, m, p, d = sketch_all(ref, bases, offs, l, dens, bool(hpc))
                tag = f"l{l}_d{dens}_h{hpc}"
                cases.append(tag)
                out[tag + "_mo"] = mo; out[tag + "_m"] = m; out[tag + "_p"] = p; out[tag + "_d"] = d
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)

    # 4. minimizer-space streams with many palindromes (tiny alphabet): purge + count for several k
    reads = [rng.integers(0, 5, size=int(rng.integers(0, 70))).astype(np.uint32) for _ in range(500)]
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    mins = np.concatenate(reads).astype(np.uint32)
    purged, po = [], [0]
    for r in reads:
        q = ref.purge_palindrome(r, 4, 12)
        purged.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(purged).astype(np.uint32)
    po = np.array(po, np.uint64)
    out = dict(minimizers=mins, offsets=offs, purged_minimizers=pm, purged_offsets=po)
    for k in (4, 6, 9, 21):
        c = ref.count(pm, po, k, 3, threads=2)
        out[f"k{k}_hashes"] = c["hashes"]; out[f"k{k}_abund"] = c["abundances"]; out[f"k{k}_vecs"] = c["vecs"]
        out[f"k{k}_stats"] = np.array([c["n_instances"], c["n_distinct"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "minspace_palindromes.npz"), **out)
    # 5. default-mode first pass (KminmerCounter + rescueKminmers) and the k=5,6 passes, from the reference's
    #    own CreateMdbg classes (oracle/ref_shim.cpp: ref_graph_firstpass / ref_graph_next_k)
    base = rng.integers(0, 400, size=6000).astype(np.uint32)
    reads = []
    for _ in range(900):
        ln = int(rng.integers(0, 90)); st = int(rng.integers(0, len(base) - 90))
        r = base[st:st + ln].copy()
        if ln and rng.random() < 0.3:
            r[rng.integers(0, ln)] = rng.integers(0, 400)
        if rng.random() < 0.5:
            r = r[::-1].copy()
        reads.append(r)
    offs = np.zeros(len(reads) + 1, np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    mins = np.concatenate(reads).astype(np.uint32)
    g4 = ref.graph_firstpass(mins, offs, 4, min_abundance=0, threads=2)
    g5 = ref.graph_next_k(mins, offs, 5, g4["hashes"], g4["abundances"], use_counter=True, threads=2)
    g6 = ref.graph_next_k(mins, offs, 6, g5["hashes"], g5["abundances"], use_counter=False, threads=2)
    np.savez_compressed(os.path.join(HERE, "minspace_multik.npz"), minimizers=mins, offsets=offs,
                        k4_hashes=g4["hashes"], k4_abund=g4["abundances"], k4_vecs=g4["vecs"],
                        k4_n_solid=g4["n_solid"], k4_n_rescued=g4["n_rescued"],
                        k5_hashes=g5["hashes"], k5_abund=g5["abundances"], k5_vecs=g5["vecs"],
                        k6_hashes=g6["hashes"], k6_abund=g6["abundances"])
    # 6. the whole readSelection stage (real ReadSelection::execute on a FASTQ): side outputs + record fields
    import tempfile
    rs = synth.make_readset(150, 5000, seed=404, n_genomes=1, genome_len_range=(90_000, 90_001))
    b6, o6 = synth.fill_reads(rs)
    raw = b6.tobytes()
    rd = [raw[int(o6[r]):int(o6[r + 1])] for r in range(rs.n_reads)]
    # (no read with N here: computeSequenceComplexity indexes kmerCounts[-1] for invalid 3-mers upstream -- heap corruption)
    rd += [b"AC" * 1500, b"A" * 800 + b"ACGTTGCA" * 300, b"ACG" * 900, b"ACGT" * 10]
    ql = []
    for sq in rd:
        q = rng.integers(2, 60, size=len(sq)).astype(np.uint8) + 33
        q[rng.random(len(sq)) < 0.01] = 33 + 93
        ql.append(q.tobytes())
    for tag, hpc, dens in (("hifi", True, 0.005), ("ont", False, 0.025)):
        with tempfile.TemporaryDirectory() as d:
            fq = os.path.join(d, "reads.fastq")
            with open(fq, "wb") as f:
                for i, (sq, q) in enumerate(zip(rd, ql)):
                    f.write(b"@r%d\n" % i + sq + b"\n+\n" + q + b"\n")
            res = ref.read_selection([fq], 15, dens, hpc, threads=2, skip_correction=True, workdir=d)
        recs = res["records"]
        mo = np.zeros(len(recs) + 1, np.uint64)
        mo[1:] = np.cumsum([len(r["minimizers"]) for r in recs])
        cat = lambda key, t: np.concatenate([r[key] for r in recs]).astype(t)
        offs6 = np.zeros(len(rd) + 1, np.uint64)
        offs6[1:] = np.cumsum([len(x) for x in rd])
        np.savez_compressed(os.path.join(HERE, f"readselection_{tag}.npz"),
                            bases=np.frombuffer(b"".join(rd), np.uint8), quals=np.frombuffer(b"".join(ql), np.uint8),
                            offsets=offs6, blacklist=res["blacklist"], min_offsets=mo,
                            minimizers=cat("minimizers", np.uint32), positions=cat("positions", np.uint32),
                            directions=cat("directions", np.uint8), qualities=cat("qualities", np.uint8),
                            mean_quality=np.array([r["mean_quality"] for r in recs], np.float32),
                            read_length=np.array([r["read_length"] for r in recs], np.uint32),
                            stats=np.array([res["stats"]["n_reads"], res["stats"]["n50"], res["stats"]["n_bases"],
                                            res["stats"]["mean_length"], res["stats"]["n_minimizers"]], np.uint64),
                            stats_f=np.array([res["stats"]["density"], res["stats"]["avg_quality"]], np.float32))
    mint_unitigs(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
