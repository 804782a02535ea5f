"""The C++ host driver (metamdbg_b200/host, reference-language host side over the C ABI):
FASTQ in, reference-format files out, compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from metamdbg_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "metamdbg_b200", "host", "mdbg_gpu_firstpass")


def test_cpp_driver_files_match_oracle(tmp_path, oracle):
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(EXE), "C++ host driver not built"
    rs = synth.make_readset(1200, 7000, seed=55, n_genomes=2, genome_len_range=(150_000, 250_000))
    bases, offs = synth.fill_reads(rs)
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        raw = bases.tobytes()
        for r in range(rs.n_reads):
            s = raw[int(offs[r]):int(offs[r + 1])]
            f.write(b"@r%d\n" % r + s + b"\n+\n" + b"I" * len(s) + b"\n")
    out = subprocess.run([EXE, str(fq), str(tmp_path), "-k", "4", "--min-abundance", "2", "--batch-mbp", "3"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    stats = dict(zip(out.stdout.split()[::2], out.stdout.split()[1::2]))
    # expected, from the oracle
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    pm, po = [], [0]
    for r in range(rs.n_reads):
        q, _ = oracle.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, 60)
        pm.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(pm).astype(np.uint32)
    po = np.array(po, np.uint64)
    ref = oracle.count(pm, po, 4, 2)
    # read_data_corrected.txt: u32 n, u8 0, u32[n]
    buf = open(tmp_path / "read_data_corrected.txt", "rb").read()
    pos, got = 0, []
    for r in range(rs.n_reads):
        n = int(np.frombuffer(buf, np.uint32, 1, pos)[0]); pos += 4
        assert buf[pos] == 0; pos += 1
        got.append(np.frombuffer(buf, np.uint32, n, pos)); pos += 4 * n
    assert pos == len(buf)
    assert np.array_equal(np.concatenate(got), pm)
    # kminmerData_abundance.txt: 20-byte records (u128 LE hash, u32 abundance)
    ab = np.frombuffer(open(tmp_path / "kminmerData_abundance.txt", "rb").read(), dtype=np.uint8).reshape(-1, 20)
    lo = ab[:, 0:8].copy().view(np.uint64)[:, 0]; hi = ab[:, 8:16].copy().view(np.uint64)[:, 0]
    cnt = ab[:, 16:20].copy().view(np.uint32)[:, 0]
    got_tab = {(int(h), int(l)): int(c) for h, l, c in zip(hi, lo, cnt)}
    want = {(int(h[0]), int(h[1])): int(a) for h, a in zip(ref["hashes"], ref["abundances"])}
    assert got_tab == want and len(want) > 100
    vec = np.frombuffer(open(tmp_path / "kminmerData_min.txt", "rb").read(), dtype=np.uint32).reshape(-1, 4)
    assert sorted(map(tuple, vec.tolist())) == sorted(map(tuple, ref["vecs"].tolist()))
    assert int(stats["solid"]) == len(want) and int(stats["reads"]) == rs.n_reads
    assert int(stats["checksum"]) == oracle.checksum(ref["hashes"], ref["abundances"])


def test_cpp_driver_read_data_init_is_byte_identical(tmp_path):
    """Row A3c: read_data_init.txt written by the C++ host driver equals, byte for byte, the file the reference's
    readSelection stage wrote for the same FASTQ (golden minted from ReadSelection::execute); read_stats.txt
    fields match (the average quality is a racy long double sum upstream: compared to 1e-5)."""
    import __graft_entry__ as g
    g.build()
    z = np.load(os.path.join(ROOT, "tests", "golden", "readselection_hifi.npz"))
    raw, qraw, offs = z["bases"].tobytes(), z["quals"].tobytes(), z["offsets"]
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        for r in range(len(offs) - 1):
            lo, hi = int(offs[r]), int(offs[r + 1])
            f.write(b"@r%d\n" % r + raw[lo:hi] + b"\n+\n" + qraw[lo:hi] + b"\n")
    out = subprocess.run([EXE, str(fq), str(tmp_path), "--batch-mbp", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    want = bytearray()
    mo = z["min_offsets"]
    for r in range(len(offs) - 1):
        lo, hi = int(mo[r]), int(mo[r + 1])
        want += np.uint32(hi - lo).tobytes() + b"\x00"
        want += z["minimizers"][lo:hi].tobytes() + z["positions"][lo:hi].tobytes()
        want += z["directions"][lo:hi].tobytes() + z["qualities"][lo:hi].tobytes()
        want += np.float32(z["mean_quality"][r]).tobytes() + np.uint32(z["read_length"][r]).tobytes()
    got = open(tmp_path / "read_data_init.txt", "rb").read()
    assert got == bytes(want)
    b = open(tmp_path / "read_stats.txt", "rb").read()
    assert len(b) == 40
    n_reads = int(np.frombuffer(b, np.uint64, 1, 0)[0]); n50 = int(np.frombuffer(b, np.uint32, 1, 8)[0])
    dens = np.frombuffer(b, np.float32, 1, 12)[0]; n_bases = int(np.frombuffer(b, np.uint64, 1, 16)[0])
    avgq = np.frombuffer(b, np.float32, 1, 24)[0]; mean_len = int(np.frombuffer(b, np.uint32, 1, 28)[0])
    n_min = int(np.frombuffer(b, np.uint64, 1, 32)[0])
    assert [n_reads, n50, n_bases, mean_len, n_min] == [int(x) for x in z["stats"]]
    assert np.float32(dens).tobytes() == np.float32(z["stats_f"][0]).tobytes()
    assert abs(float(avgq) - float(z["stats_f"][1])) <= 1e-5 * abs(float(z["stats_f"][1]))


def test_cpp_driver_graph_seam_and_default_mode(tmp_path, oracle):
    """`--from-read-data`: the graph --firstpass seam alone, on a read_data_corrected.txt written by stage one; with
    --min-abundance 0 (metaMDBG's default mode) the table carries the rescued abundance-1 entries."""
    import __graft_entry__ as g
    g.build()
    rs = synth.make_readset(1500, 7000, seed=56, n_genomes=2, genome_len_range=(150_000, 250_000), err=0.004)
    bases, offs = synth.fill_reads(rs)
    fa = tmp_path / "reads.fasta"
    raw = bases.tobytes()
    with open(fa, "wb") as f:
        for r in range(rs.n_reads):
            s = raw[int(offs[r]):int(offs[r + 1])]
            f.write(b">r%d\n" % r)
            for i in range(0, len(s), 80):                 # multi-line FASTA
                f.write(s[i:i + 80] + b"\n")
    d1, d2 = tmp_path / "one", tmp_path / "two"
    d1.mkdir(); d2.mkdir()
    out = subprocess.run([EXE, str(fa), str(d1), "--batch-mbp", "4"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    out = subprocess.run([EXE, "--from-read-data", str(d1 / "read_data_corrected.txt"), str(d2), "--min-abundance", "0"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    stats = dict(zip(out.stdout.split()[::2], out.stdout.split()[1::2]))
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, 0.005, True)
    pm, po = [], [0]
    for rr in range(rs.n_reads):
        q, _ = oracle.purge_palindrome(m[int(mo[rr]):int(mo[rr + 1])], 4, 60)
        pm.append(q); po.append(po[-1] + len(q))
    m, mo = np.concatenate(pm).astype(np.uint32), np.array(po, np.uint64)
    c = oracle.count(m, mo, 4, 2)
    r = oracle.rescue(m, mo, 4, c["hashes"], c["abundances"])
    want = {(int(h[0]), int(h[1])): int(a) for h, a in zip(c["hashes"], c["abundances"])}
    for h in r["hashes"]:
        want[(int(h[0]), int(h[1]))] = 1
    ab = np.frombuffer(open(d2 / "kminmerData_abundance.txt", "rb").read(), dtype=np.uint8).reshape(-1, 20)
    lo = ab[:, 0:8].copy().view(np.uint64)[:, 0]; hi = ab[:, 8:16].copy().view(np.uint64)[:, 0]
    cnt = ab[:, 16:20].copy().view(np.uint32)[:, 0]
    got = {(int(h), int(l)): int(c_) for h, l, c_ in zip(hi, lo, cnt)}
    assert got == want
    assert int(stats["solid"]) == len(c["abundances"]) and int(stats["rescued"]) == len(r["hashes"]) > 0
    # FASTA input: no qualities -> per-minimizer quality 1 and NaN mean quality in read_data_init.txt
    from oracle.pyoracle import parse_read_data
    recs = parse_read_data(str(d1 / "read_data_init.txt"), True)
    assert len(recs) == rs.n_reads and all(np.all(x["qualities"] == 1) for x in recs[:50])
    assert all(np.isnan(x["mean_quality"]) for x in recs[:50])


def test_reference_stage_driven_by_gpu_functor(tmp_path):
    """metaMDBG's own readSelection stage code (kseq FASTQ parser, ReadSelection::writeRead with its re-ordering queue,
    computeReadStats) with the GPU functor of INTEGRATION.md plugged in (oracle/_ref/mdbg_ref_integrated = reference
    sources + libmdbg_b200.so).  Its read_data_init.txt must be byte-identical to the file the unmodified stage
    wrote (golden), with the reads arriving from several parser threads."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mdbg_ref_integrated")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mdbg_ref_integrated not built (needs the reference sources)")
    z = np.load(os.path.join(ROOT, "tests", "golden", "readselection_hifi.npz"))
    raw, qraw, offs = z["bases"].tobytes(), z["quals"].tobytes(), z["offsets"]
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        for r in range(len(offs) - 1):
            lo, hi = int(offs[r]), int(offs[r + 1])
            f.write(b"@r%d\n" % r + raw[lo:hi] + b"\n+\n" + qraw[lo:hi] + b"\n")
    with open(tmp_path / "input.txt", "w") as f:
        f.write(str(fq) + "\n")
    out = subprocess.run([exe, str(tmp_path / "input.txt"), str(tmp_path), "15", "0.005", "1", "4"],
                         capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr[-2000:]
    want = bytearray()
    mo = z["min_offsets"]
    for r in range(len(offs) - 1):
        lo, hi = int(mo[r]), int(mo[r + 1])
        want += np.uint32(hi - lo).tobytes() + b"\x00"
        want += z["minimizers"][lo:hi].tobytes() + z["positions"][lo:hi].tobytes()
        want += z["directions"][lo:hi].tobytes() + z["qualities"][lo:hi].tobytes()
        want += np.float32(z["mean_quality"][r]).tobytes() + np.uint32(z["read_length"][r]).tobytes()
    assert open(tmp_path / "read_data_init.txt", "rb").read() == bytes(want)
    b = open(tmp_path / "read_stats.txt", "rb").read()
    got = [int(np.frombuffer(b, np.uint64, 1, 0)[0]), int(np.frombuffer(b, np.uint32, 1, 8)[0]),
           int(np.frombuffer(b, np.uint64, 1, 16)[0]), int(np.frombuffer(b, np.uint32, 1, 28)[0]),
           int(np.frombuffer(b, np.uint64, 1, 32)[0])]
    assert got == [int(x) for x in z["stats"]]
    assert os.path.getsize(tmp_path / "read_data_corrected.txt") > 0


def _write_parameters_gz(path, l=15, k=4, dens_asm=0.005, first_k=4, last_k=21, mean_len=0, dens_corr=0.025, hpc=True,
                         data_type=0, snpmer=0):
    """parameters.gz as AssemblyPipeline::writeParameters emits it (src/pipeline/AssemblyPipeline.hpp:1479-1517)."""
    import gzip
    import struct
    spacing = np.float32(1) / np.float32(dens_asm)
    len_mean = np.float32(spacing * np.float32(k - 1))
    rec = struct.pack("<QQfQfffQQQf?iQ", l, k, dens_asm, first_k, float(spacing), float(len_mean),
                      float(np.float32(len_mean - spacing)), first_k, last_k, mean_len, dens_corr, hpc, data_type, snpmer)
    with gzip.open(path, "wb") as f:
        f.write(rec)


@pytest.mark.parametrize("hifi", [True, False])
def test_cpp_driver_stage_compatible_subcommands(tmp_path, oracle, hifi):
    """Row (f)2: `readSelection <tmpDir> <out> <input.txt>` and `graph <tmpDir> --firstpass --min-abundance n` with
    metaMDBG's own positional command lines and parameters.gz; the files in <tmpDir> are the ones the next reference
    stage reads: read_data_init.txt / read_stats.txt / repetitiveMinimizers.bin / read_data_corrected.txt, then
    kminmerData_min.txt / kminmerData_abundance.txt (multiset-equal to the oracle's table)."""
    import __graft_entry__ as g
    g.build()
    rs = synth.make_readset(900, 6000, seed=58, n_genomes=2, genome_len_range=(120_000, 200_000), err=0.002)
    bases, offs = synth.fill_reads(rs)
    raw = bases.tobytes()
    files = []
    for part, (lo, hi) in enumerate(((0, 500), (500, rs.n_reads))):      # two input files = two datasets
        fq = tmp_path / f"reads{part}.fastq"
        with open(fq, "wb") as f:
            for r in range(lo, hi):
                s = raw[int(offs[r]):int(offs[r + 1])]
                f.write(b"@r%d\n" % r + s + b"\n+\n" + b"I" * len(s) + b"\n")
        files.append(str(fq))
    tmp = tmp_path / "tmp"
    tmp.mkdir()
    (tmp / "input.txt").write_text("\n".join(files) + "\n")
    dens, corr = 0.005, 0.025
    _write_parameters_gz(tmp / "parameters.gz", hpc=hifi, data_type=0 if hifi else 1, dens_asm=dens, dens_corr=corr)
    out = subprocess.run([EXE, "readSelection", str(tmp), str(tmp / "read_data_init.txt"), str(tmp / "input.txt"), "--threads", "4",
                          "--min-read-quality", "0", "--output-quality", "--skip-correction", "--batch-mbp", "2"],
                         capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr
    for name in ("read_data_init.txt", "read_stats.txt", "repetitiveMinimizers.bin", "read_data_corrected.txt"):
        assert (tmp / name).exists(), name
    bl = np.fromfile(tmp / "repetitiveMinimizers.bin", dtype=np.uint32)
    if hifi:
        assert len(bl) == 0                                  # ReadSelection.hpp:501: nothing for HiFi
    else:
        # ONT: the most frequent minimizers of the sketch at the correction density (no blacklist, no HPC)
        _, mc, _, _ = oracle.sketch_batch(bases, offs, 15, corr, False)
        vals, counts = np.unique(mc, return_counts=True)
        want_n = max(1, int(np.float32(0.00001) * np.float32(len(vals))))
        assert len(bl) == want_n and counts[np.isin(vals, bl)].min() >= np.sort(counts)[-want_n]
    mo, m, p, d = oracle.sketch_batch(bases, offs, 15, dens, hifi, blacklist=np.sort(bl) if len(bl) else None)
    b = open(tmp / "read_stats.txt", "rb").read()
    n50 = int(np.frombuffer(b, np.uint32, 1, 8)[0])
    assert int(np.frombuffer(b, np.uint64, 1, 0)[0]) == rs.n_reads and int(np.frombuffer(b, np.uint64, 1, 32)[0]) == len(m)
    last_k = max(int(np.float32(n50) * np.float32(dens) * np.float32(2.0)), 6)      # Commons::computeLastK
    pm, po = [], [0]
    for r in range(rs.n_reads):
        q, _ = oracle.purge_palindrome(m[int(mo[r]):int(mo[r + 1])], 4, last_k)
        pm.append(q); po.append(po[-1] + len(q))
    pm = np.concatenate(pm).astype(np.uint32); po = np.array(po, np.uint64)
    buf = open(tmp / "read_data_corrected.txt", "rb").read()
    pos, got = 0, []
    for r in range(rs.n_reads):
        n = int(np.frombuffer(buf, np.uint32, 1, pos)[0]); pos += 5
        got.append(np.frombuffer(buf, np.uint32, n, pos)); pos += 4 * n
    assert pos == len(buf) and np.array_equal(np.concatenate(got), pm)
    for min_ab in (2, 0):
        out = subprocess.run([EXE, "graph", str(tmp), "--threads", "4", "--min-abundance", str(min_ab), "--firstpass", "--unitigs"],
                             capture_output=True, text=True, timeout=180)
        assert out.returncode == 0, out.stderr
        ref = oracle.count(pm, po, 4, 2)
        want = {(int(h[0]), int(h[1])): int(a) for h, a in zip(ref["hashes"], ref["abundances"])}
        if min_ab == 0:                                     # default mode: rescued abundance-1 entries ride along
            for h in oracle.rescue(pm, po, 4, ref["hashes"], ref["abundances"])["hashes"]:
                want[(int(h[0]), int(h[1]))] = 1
        ab = np.frombuffer(open(tmp / "kminmerData_abundance.txt", "rb").read(), dtype=np.uint8).reshape(-1, 20)
        lo = ab[:, 0:8].copy().view(np.uint64)[:, 0]; hi = ab[:, 8:16].copy().view(np.uint64)[:, 0]
        cnt = ab[:, 16:20].copy().view(np.uint32)[:, 0]
        assert {(int(h), int(l)): int(c) for h, l, c in zip(hi, lo, cnt)} == want and len(want) > 100
        vec = np.frombuffer(open(tmp / "kminmerData_min.txt", "rb").read(), dtype=np.uint32).reshape(-1, 4)
        assert len(vec) == len(want)
        # --unitigs: createGfa's node side on the table just built -- unitigGraph.nodes.bin as the reference writes it
        wu = oracle.unitigs(vec, 4)
        want_bytes = b"".join(np.uint32(int(wu["offsets"][i + 1] - wu["offsets"][i])).tobytes() +
                              wu["minimizers"][int(wu["offsets"][i]):int(wu["offsets"][i + 1])].tobytes() + np.uint32(2 * i).tobytes()
                              for i in range(len(wu["offsets"]) - 1))
        assert open(tmp / "unitigGraph.nodes.bin", "rb").read() == want_bytes and len(wu["offsets"]) > 10
    # a later pass needs the contig stage: refused with a message, not silently wrong
    out = subprocess.run([EXE, "graph", str(tmp), "--threads", "4"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "firstpass" in out.stderr
