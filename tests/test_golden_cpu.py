"""The C restatement (oracle/) against the committed golden vectors minted from
the reference's own code (tests/golden/make_golden.py).  Runs anywhere."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def table_dict(hashes, abund):
    return {(int(h[0]), int(h[1])): int(a) for h, a in zip(hashes, abund)}


def test_hifi_small_sketch_purge_count(oracle):
    g = load("hifi_small.npz")
    mo, m, p, d = oracle.sketch_batch(g["bases"], g["offsets"], 15, 0.005, True)
    assert np.array_equal(mo, g["min_offsets"]) and np.array_equal(m, g["minimizers"])
    assert np.array_equal(p, g["positions"]) and np.array_equal(d, g["directions"])
    po = [0]; pm = []
    for r in range(len(mo) - 1):
        q, _ = oracle.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, 60)
        pm.append(q); po.append(po[-1] + len(q))
    assert np.array_equal(np.concatenate(pm), g["purged_minimizers"])
    assert np.array_equal(np.array(po, np.uint64), g["purged_offsets"])
    for k in (4, 5, 7):
        c = oracle.count(g["purged_minimizers"], g["purged_offsets"], k, 2)
        assert np.array_equal(c["vecs"], g[f"k{k}_vecs"])
        assert np.array_equal(c["hashes"], g[f"k{k}_hashes"]) and np.array_equal(c["abundances"], g[f"k{k}_abund"])
        assert [c["n_instances"], c["n_distinct"]] == list(g[f"k{k}_stats"])


def test_ont_small(oracle):
    g = load("ont_small.npz")
    mo, m, p, d = oracle.sketch_batch(g["bases"], g["offsets"], 15, 0.025, False, g["blacklist"])
    assert np.array_equal(mo, g["min_offsets"]) and np.array_equal(m, g["minimizers"])
    assert np.array_equal(p, g["positions"]) and np.array_equal(d, g["directions"])
    c = oracle.count(m, mo, 4, 2)
    assert table_dict(c["hashes"], c["abundances"]) == table_dict(g["k4_hashes"], g["k4_abund"])


def test_edge_cases(oracle):
    g = load("edge_cases.npz")
    for tag in g["cases"]:
        tag = str(tag)
        l = int(tag.split("_")[0][1:]); dens = float(tag.split("_")[1][1:]); hpc = tag.endswith("h1")
        mo, m, p, d = oracle.sketch_batch(g["bases"], g["offsets"], l, dens, hpc)
        assert np.array_equal(mo, g[tag + "_mo"]), tag
        assert np.array_equal(m, g[tag + "_m"]) and np.array_equal(p, g[tag + "_p"]) and np.array_equal(d, g[tag + "_d"]), tag


def test_minspace_palindromes(oracle):
    g = load("minspace_palindromes.npz")
    offs, mins = g["offsets"], g["minimizers"]
    pm = []
    for r in range(len(offs) - 1):
        q, _ = oracle.purge_palindrome(mins[int(offs[r]):int(offs[r + 1])], 4, 12)
        pm.append(q)
    assert np.array_equal(np.concatenate(pm), g["purged_minimizers"])
    for k in (4, 6, 9, 21):
        c = oracle.count(g["purged_minimizers"], g["purged_offsets"], k, 3)
        assert np.array_equal(c["hashes"], g[f"k{k}_hashes"]) and np.array_equal(c["abundances"], g[f"k{k}_abund"])
        assert np.array_equal(c["vecs"], g[f"k{k}_vecs"])


def test_minspace_multik(oracle):
    """Default-mode first pass (count + rescue) and the k=5 / k=6 passes against vectors minted from the
    reference's own KminmerCounter / rescueKminmers / IndexKminmerFunctor."""
    g = load("minspace_multik.npz")
    mins, offs = g["minimizers"], g["offsets"]
    ns = int(g["k4_n_solid"])
    c = oracle.count(mins, offs, 4, 2)
    assert table_dict(c["hashes"], c["abundances"]) == table_dict(g["k4_hashes"][:ns], g["k4_abund"][:ns])
    r = oracle.rescue(mins, offs, 4, c["hashes"], c["abundances"])
    assert len(r["hashes"]) == int(g["k4_n_rescued"]) > 0
    assert sorted(map(tuple, r["hashes"].tolist())) == sorted(map(tuple, g["k4_hashes"][ns:].tolist()))
    t5 = oracle.next_k(mins, offs, 5, g["k4_hashes"], g["k4_abund"])
    assert table_dict(t5["hashes"], t5["abundances"]) == table_dict(g["k5_hashes"], g["k5_abund"])
    t6 = oracle.next_k(mins, offs, 6, g["k5_hashes"], g["k5_abund"])
    assert table_dict(t6["hashes"], t6["abundances"]) == table_dict(g["k6_hashes"], g["k6_abund"])
    assert len(t6["abundances"]) > 20


@pytest.mark.parametrize("tag,hpc,dens", [("hifi", True, 0.005), ("ont", False, 0.025)])
def test_readselection_stage_golden(oracle, tag, hpc, dens):
    """Rows A3b/A3c: record fields of read_data_init.txt minted by the reference's whole readSelection stage."""
    g = load(f"readselection_{tag}.npz")
    raw, qraw, offs, mo = g["bases"].tobytes(), g["quals"].tobytes(), g["offsets"], g["min_offsets"]
    n_low = 0
    for r in range(len(offs) - 1):
        s, q = raw[int(offs[r]):int(offs[r + 1])], qraw[int(offs[r]):int(offs[r + 1])]
        m, p, d = oracle.sketch_read(s, 15, dens, hpc, g["blacklist"])
        mq, cx, mins_q = oracle.read_aux(s, q, 15, hpc, p)
        if cx > 5:
            n_low += 1
            m, p, d, mins_q = m[:0], p[:0], d[:0], mins_q[:0]
        lo, hi = int(mo[r]), int(mo[r + 1])
        assert np.array_equal(g["minimizers"][lo:hi], m) and np.array_equal(g["positions"][lo:hi], p)
        assert np.array_equal(g["directions"][lo:hi], d) and np.array_equal(g["qualities"][lo:hi], mins_q)
        assert np.float32(g["mean_quality"][r]).tobytes() == np.float32(mq).tobytes()
        assert int(g["read_length"][r]) == len(s)
    assert n_low >= 2


def test_unitig_nodes_golden(oracle):
    """Row F1, third step: the oracle's unitigs against unitigGraph.nodes.bin contents minted from the reference's own
    indexEdges + computeUnitigNodes + computeDeterministicUnitigs (tests/golden/make_golden.py::mint_unitigs)."""
    g = load("minspace_unitigs.npz")
    n_multi = n_edges = 0
    for i in range(int(g["n_cases"])):
        k = int(g[f"c{i}_k"])
        nodes = oracle.count(g[f"c{i}_minimizers"], g[f"c{i}_offsets"], k, 2)["vecs"]
        assert sorted(map(tuple, nodes.tolist())) == sorted(map(tuple, g[f"c{i}_nodes"].tolist()))
        u = oracle.unitigs(g[f"c{i}_nodes"], k)
        assert np.array_equal(u["offsets"], g[f"c{i}_unitig_offsets"]), i
        assert np.array_equal(u["minimizers"], g[f"c{i}_unitig_minimizers"]), i
        n_multi += int(np.any(np.diff(u["offsets"]) > k))
        e = oracle.unitig_edges(u["offsets"], u["minimizers"], k)           # unitigGraph.edges.successors.bin, 1-thread order
        assert np.array_equal(e["offsets"], g[f"c{i}_edge_offsets"]) and np.array_equal(e["targets"], g[f"c{i}_edge_targets"]), i
        assert [e["n_edges"], e["checksum"]] == [int(x) for x in g[f"c{i}_edge_stats"]]
        n_edges += e["n_edges"]
    assert n_multi >= 8 and n_edges > 500
