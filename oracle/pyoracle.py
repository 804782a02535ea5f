"""ctypes binding of the TEST-ONLY CPU oracle (oracle/mdbg_oracle.c) and, when
built, of the reference's own sources (oracle/_ref/libmdbg_ref.so).

TEST INFRASTRUCTURE: may be imported only from tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke().  Nothing in
metamdbg_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmdbg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmdbg_ref.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "mdbg_oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE, os.path.join(HERE, "libmdbg_oracle.so")], check=True)
    if os.path.exists("/root/reference/src/Commons.hpp"):
        integrated = os.path.join(HERE, "_ref", "mdbg_ref_integrated")
        engine = os.path.join(HERE, "..", "metamdbg_b200", "libmdbg_b200.so")
        if os.path.exists(REF_SO) and os.path.exists(engine) and (
                not os.path.exists(integrated) or os.path.getmtime(integrated) < os.path.getmtime(engine)):
            subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        if force or not os.path.exists(REF_SO) or \
                os.path.getmtime(REF_SO) < max(os.path.getmtime(os.path.join(HERE, "ref_shim.cpp")),
                                               os.path.getmtime(os.path.join(HERE, "Makefile"))):
            subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def parse_read_data(path: str, has_quality: bool):
    """read_data_init.txt (has_quality) / read_data_corrected.txt records
    (ReadSelection.hpp:415-467): u32 n, u8 isCircular, u32 min[n] [, u32 pos[n], u8 dir[n], u8 qual[n], f32 meanQ, u32 len]."""
    buf = open(path, "rb").read()
    pos, recs = 0, []
    while pos < len(buf):
        n = int(np.frombuffer(buf, np.uint32, 1, pos)[0]); pos += 4
        circ = buf[pos]; pos += 1
        rec = dict(circular=circ, minimizers=np.frombuffer(buf, np.uint32, n, pos).copy()); pos += 4 * n
        if has_quality:
            rec["positions"] = np.frombuffer(buf, np.uint32, n, pos).copy(); pos += 4 * n
            rec["directions"] = np.frombuffer(buf, np.uint8, n, pos).copy(); pos += n
            rec["qualities"] = np.frombuffer(buf, np.uint8, n, pos).copy(); pos += n
            rec["mean_quality"] = np.frombuffer(buf, np.float32, 1, pos)[0]; pos += 4
            rec["read_length"] = int(np.frombuffer(buf, np.uint32, 1, pos)[0]); pos += 4
        recs.append(rec)
    return recs


def parse_read_stats(path: str):
    """read_stats.txt (ReadSelection.hpp:305-384): u64 nbReads, u32 n50, f32 density, u64 nbBases, f32 avgQ,
    u32 meanLen, u64 nbSelectedMinimizers."""
    b = open(path, "rb").read()
    return dict(n_reads=int(np.frombuffer(b, np.uint64, 1, 0)[0]), n50=int(np.frombuffer(b, np.uint32, 1, 8)[0]),
                density=float(np.frombuffer(b, np.float32, 1, 12)[0]), n_bases=int(np.frombuffer(b, np.uint64, 1, 16)[0]),
                avg_quality=float(np.frombuffer(b, np.float32, 1, 24)[0]), mean_length=int(np.frombuffer(b, np.uint32, 1, 28)[0]),
                n_minimizers=int(np.frombuffer(b, np.uint64, 1, 32)[0]))


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class _Lib:
    """Common numpy front-end over the `orc_*` / `ref_*` entry points."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L, p = self.lib, prefix
        f = getattr(L, p + "murmur3_x64_128_h1"); f.restype = C.c_uint64; f.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        f = getattr(L, p + "murmur3_x64_128"); f.restype = None; f.argtypes = [C.c_void_p, C.c_int, C.c_uint32, _u64p]
        f = getattr(L, p + "minimizer_bound"); f.restype = C.c_double; f.argtypes = [C.c_float]
        f = getattr(L, p + "hpc"); f.restype = C.c_size_t; f.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        f = getattr(L, p + "lmers"); f.restype = C.c_size_t; f.argtypes = [C.c_char_p, C.c_size_t, C.c_int, _u64p, _u8p]
        f = getattr(L, p + "sketch_read"); f.restype = C.c_size_t
        f.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_float, C.c_int, _u32p, C.c_size_t, _u32p, _u32p, _u8p, C.c_size_t]
        f = getattr(L, p + "apply_density"); f.restype = C.c_size_t; f.argtypes = [_u32p, C.c_size_t, C.c_float, _u32p]
        f = getattr(L, p + "kminmers"); f.restype = C.c_size_t; f.argtypes = [_u32p, C.c_size_t, C.c_int, _u32p, _u8p]
        f = getattr(L, p + "hash128"); f.restype = None; f.argtypes = [_u32p, C.c_int, _u64p]
        getattr(L, p + "free").argtypes = [C.c_void_p]

    # -- scalar helpers ----------------------------------------------------
    def murmur_h1(self, key: bytes, seed: int) -> int:
        return int(getattr(self.lib, self.prefix + "murmur3_x64_128_h1")(key, len(key), seed))

    def murmur128(self, key: bytes, seed: int) -> tuple[int, int]:
        out = (C.c_uint64 * 2)()
        getattr(self.lib, self.prefix + "murmur3_x64_128")(key, len(key), seed, out)
        return int(out[0]), int(out[1])

    def bound(self, density: float) -> float:
        return float(getattr(self.lib, self.prefix + "minimizer_bound")(density))

    def hpc(self, seq: bytes, hpc: bool = True) -> tuple[bytes, np.ndarray]:
        out = np.zeros(len(seq) + 2, dtype=np.uint8)
        pos = np.zeros(len(seq) + 2, dtype=np.uint64)
        n = getattr(self.lib, self.prefix + "hpc")(seq, len(seq), int(hpc), out.ctypes.data, pos.ctypes.data)
        return out[:n].tobytes(), pos[:n + (1 if hpc else 0)].copy()

    def lmers(self, seq: bytes, l: int) -> tuple[np.ndarray, np.ndarray]:
        n = max(0, len(seq) - l + 1)
        v = np.zeros(n + 1, dtype=np.uint64)
        d = np.zeros(n + 1, dtype=np.uint8)
        k = getattr(self.lib, self.prefix + "lmers")(seq, len(seq), l, _p(v, _u64p), _p(d, _u8p))
        return v[:k].copy(), d[:k].copy()

    def sketch_read(self, seq: bytes, l: int, density: float, hpc: bool, blacklist: np.ndarray | None = None):
        cap = len(seq) + 1
        m = np.zeros(cap, dtype=np.uint32)
        p = np.zeros(cap, dtype=np.uint32)
        d = np.zeros(cap, dtype=np.uint8)
        bl = np.ascontiguousarray(np.sort(blacklist).astype(np.uint32)) if blacklist is not None and len(blacklist) else None
        n = getattr(self.lib, self.prefix + "sketch_read")(
            seq, len(seq), l, density, int(hpc), _p(bl, _u32p) if bl is not None else None,
            0 if bl is None else len(bl), _p(m, _u32p), _p(p, _u32p), _p(d, _u8p), cap)
        return m[:n].copy(), p[:n].copy(), d[:n].copy()

    def sketch_batch(self, bases: np.ndarray, offsets: np.ndarray, l: int, density: float, hpc: bool,
                     blacklist: np.ndarray | None = None):
        """Per-read loop -> CSR (min_offsets u64[n+1], minimizers, positions, directions)."""
        ms, ps, ds, offs = [], [], [], [0]
        raw = bases.tobytes()
        for r in range(len(offsets) - 1):
            m, p, d = self.sketch_read(raw[int(offsets[r]):int(offsets[r + 1])], l, density, hpc, blacklist)
            ms.append(m); ps.append(p); ds.append(d); offs.append(offs[-1] + len(m))
        cat = lambda xs, t: np.concatenate(xs).astype(t) if xs else np.zeros(0, t)
        return (np.array(offs, dtype=np.uint64), cat(ms, np.uint32), cat(ps, np.uint32), cat(ds, np.uint8))

    def apply_density(self, m: np.ndarray, density: float) -> np.ndarray:
        m = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.zeros(len(m) + 1, dtype=np.uint32)
        n = getattr(self.lib, self.prefix + "apply_density")(_p(m, _u32p), len(m), density, _p(out, _u32p))
        return out[:n].copy()

    def kminmers(self, m: np.ndarray, k: int):
        m = np.ascontiguousarray(m, dtype=np.uint32)
        nw = max(0, len(m) - k + 1)
        v = np.zeros((nw + 1) * k, dtype=np.uint32)
        rv = np.zeros(nw + 1, dtype=np.uint8)
        n = getattr(self.lib, self.prefix + "kminmers")(_p(m, _u32p), len(m), k, _p(v, _u32p), _p(rv, _u8p))
        return v[:n * k].reshape(n, k).copy(), rv[:n].copy()

    def hash128(self, vec: np.ndarray) -> tuple[int, int]:
        vec = np.ascontiguousarray(vec, dtype=np.uint32)
        out = (C.c_uint64 * 2)()
        getattr(self.lib, self.prefix + "hash128")(_p(vec, _u32p), len(vec), out)
        return int(out[0]), int(out[1])

    def _take(self, ptr, n, dtype):
        if n == 0:
            arr = np.zeros(0, dtype=dtype)
        else:
            arr = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
        getattr(self.lib, self.prefix + "free")(C.cast(ptr, C.c_void_p))
        return arr


class Oracle(_Lib):
    def __init__(self):
        build()
        super().__init__(ORACLE_SO, "orc_")
        L = self.lib
        L.orc_minimizer_threshold.restype = C.c_uint64
        L.orc_minimizer_threshold.argtypes = [C.c_float, C.POINTER(C.c_int)]
        L.orc_purge_palindrome.restype = C.c_size_t
        L.orc_purge_palindrome.argtypes = [_u32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p, _u8p]
        L.orc_count.restype = C.c_size_t
        L.orc_count.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(_u32p), C.POINTER(_u64p),
                                C.POINTER(_u32p), _u64p, _u64p]
        L.orc_sketch_batch.restype = C.c_size_t
        L.orc_sketch_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_int, C.c_float, C.c_int, _u32p, C.c_size_t,
                                       _u64p, _u32p, _u32p, _u8p, C.c_size_t]
        L.orc_read_aux.restype = None
        L.orc_read_aux.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, _u32p, C.c_size_t,
                                   C.POINTER(C.c_float), C.POINTER(C.c_double), _u8p]
        L.orc_rescue.restype = C.c_size_t
        L.orc_rescue.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, _u64p, _u32p, C.c_size_t, C.POINTER(_u32p),
                                 C.POINTER(_u64p), _u64p]
        L.orc_next_k.restype = C.c_size_t
        L.orc_next_k.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, _u64p, _u32p, C.c_size_t, C.POINTER(_u32p),
                                 C.POINTER(_u64p), C.POINTER(_u32p)]
        L.orc_edge_index.restype = C.c_size_t
        L.orc_edge_index.argtypes = [_u32p, C.c_size_t, C.c_int, C.POINTER(_u64p), C.POINTER(C.c_uint64)]
        L.orc_unitig_edges.restype = C.c_size_t
        L.orc_unitig_edges.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, C.POINTER(_u64p), C.POINTER(_u32p), C.POINTER(C.c_uint64)]
        L.orc_unitigs.restype = C.c_size_t
        L.orc_unitigs.argtypes = [_u32p, C.c_size_t, C.c_int, C.POINTER(_u64p), C.POINTER(_u32p), C.POINTER(_u64p)]
        L.orc_edge_values.restype = C.c_size_t
        L.orc_edge_values.argtypes = [_u32p, C.c_size_t, C.c_int, C.POINTER(_u64p), C.POINTER(_u32p)]
        L.orc_table_checksum.restype = C.c_uint64
        L.orc_table_checksum.argtypes = [_u64p, _u32p, C.c_size_t]

    def threshold(self, density: float) -> tuple[int, bool]:
        none = C.c_int(0)
        t = self.lib.orc_minimizer_threshold(density, C.byref(none))
        return int(t), bool(none.value)

    def sketch_batch(self, bases, offsets, l, density, hpc, blacklist=None, cap=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if cap is None:
            cap = max(1024, int(len(bases) * max(density, 0.001) * 4) + 64 * n)
        bl = np.ascontiguousarray(np.sort(blacklist).astype(np.uint32)) if blacklist is not None and len(blacklist) else None
        while True:
            mo = np.zeros(n + 1, dtype=np.uint64)
            m = np.zeros(cap, dtype=np.uint32); p = np.zeros(cap, dtype=np.uint32); d = np.zeros(cap, dtype=np.uint8)
            tot = self.lib.orc_sketch_batch(bases.ctypes.data, _p(offsets, _u64p), n, l, density, int(hpc),
                                            _p(bl, _u32p) if bl is not None else None, 0 if bl is None else len(bl),
                                            _p(mo, _u64p), _p(m, _u32p), _p(p, _u32p), _p(d, _u8p), cap)
            if tot <= cap:
                return mo, m[:tot].copy(), p[:tot].copy(), d[:tot].copy()
            cap = int(tot)

    def purge_palindrome(self, m: np.ndarray, first_k: int, last_k: int):
        m = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.zeros(len(m) + 1, dtype=np.uint32)
        keep = np.zeros(len(m) + 1, dtype=np.uint8)
        n = self.lib.orc_purge_palindrome(_p(m, _u32p), len(m), first_k, last_k, _p(out, _u32p), _p(keep, _u8p))
        return out[:n].copy(), keep[:len(m)].copy()

    def count(self, mins: np.ndarray, offs: np.ndarray, k: int, min_abundance: int = 2, keep_all: bool = False):
        """-> dict(vecs [n,k] u32, hashes [n,2] u64 (h1,h2), abundances [n] u32, n_instances, n_distinct).
        keep_all=True returns every distinct k-min-mer (abundance 1 included; multi-GPU merge tests)."""
        if keep_all:
            min_abundance = 0xFFFFFFFF
        mins = np.ascontiguousarray(mins, dtype=np.uint32)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        v = _u32p(); h = _u64p(); a = _u32p()
        ni = C.c_uint64(0); nd = C.c_uint64(0)
        n = self.lib.orc_count(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, min_abundance,
                               C.byref(v), C.byref(h), C.byref(a), C.byref(ni), C.byref(nd))
        vecs = self._take(v, n * k, np.uint32).reshape(n, k)
        hashes = self._take(h, n * 2, np.uint64).reshape(n, 2)
        abund = self._take(a, n, np.uint32)
        return dict(vecs=vecs, hashes=hashes, abundances=abund, n_instances=int(ni.value), n_distinct=int(nd.value))

    def read_aux(self, seq: bytes, qual: bytes, l: int, hpc: bool, positions: np.ndarray):
        """-> (mean_quality f32, complexity f64, qualities u8[n])."""
        positions = np.ascontiguousarray(positions, dtype=np.uint32)
        mq = C.c_float(0); cx = C.c_double(0)
        q = np.zeros(len(positions) + 1, dtype=np.uint8)
        self.lib.orc_read_aux(seq, qual, len(seq), len(qual), l, int(hpc), _p(positions, _u32p), len(positions),
                              C.byref(mq), C.byref(cx), _p(q, _u8p))
        return np.float32(mq.value), float(cx.value), q[:len(positions)].copy()

    def rescue(self, mins, offs, k, solid_hashes, solid_ab):
        """-> dict(vecs [n,k], hashes [n,2] (h1,h2), n_reads_rescued); rescued abundance is always 1."""
        mins = np.ascontiguousarray(mins, dtype=np.uint32); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        sh = np.ascontiguousarray(solid_hashes, dtype=np.uint64); sa = np.ascontiguousarray(solid_ab, dtype=np.uint32)
        v = _u32p(); h = _u64p(); nr = C.c_uint64(0)
        n = self.lib.orc_rescue(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, _p(sh, _u64p), _p(sa, _u32p), len(sa),
                                C.byref(v), C.byref(h), C.byref(nr))
        return dict(vecs=self._take(v, n * k, np.uint32).reshape(n, k), hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2),
                    n_reads_rescued=int(nr.value))

    def next_k(self, mins, offs, k, prev_hashes, prev_ab):
        """-> dict(vecs, hashes (h1,h2), abundances), sorted by hash."""
        mins = np.ascontiguousarray(mins, dtype=np.uint32); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ph = np.ascontiguousarray(prev_hashes, dtype=np.uint64); pa = np.ascontiguousarray(prev_ab, dtype=np.uint32)
        v = _u32p(); h = _u64p(); a = _u32p()
        n = self.lib.orc_next_k(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, _p(ph, _u64p), _p(pa, _u32p), len(pa),
                                C.byref(v), C.byref(h), C.byref(a))
        return dict(vecs=self._take(v, n * k, np.uint32).reshape(n, k), hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2),
                    abundances=self._take(a, n, np.uint32))

    def edge_index(self, vecs, k):
        """EdgeIndexer: distinct hash128 of the normalized (k-1)-prefix / suffix of every node.
        -> dict(hashes [n,2] (h1,h2) sorted, checksum)."""
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        h = _u64p(); cs = C.c_uint64(0)
        n = self.lib.orc_edge_index(_p(vecs, _u32p), len(vecs), k, C.byref(h), C.byref(cs))
        return dict(hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2), checksum=int(cs.value))

    def edge_values(self, vecs, k):
        """indexEdge in order-free form -> dict(hashes [n,2] (h1,h2) sorted, values [n,2,4]: per orientation class
        (count 0/1/2+, minimizer, isReversed, isPrefix))."""
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        h = _u64p(); v = _u32p()
        n = self.lib.orc_edge_values(_p(vecs, _u32p), len(vecs), k, C.byref(h), C.byref(v))
        return dict(hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2), values=self._take(v, 8 * n, np.uint32).reshape(n, 2, 4))

    def unitigs(self, vecs, k):
        """computeUnitigNodes + computeDeterministicUnitigs -> dict(offsets [n+1], minimizers, hashes [n,2]): the records
        of unitigGraph.nodes.bin in file order (unitigIndex = 2 * position)."""
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        o = _u64p(); m = _u32p(); h = _u64p()
        n = self.lib.orc_unitigs(_p(vecs, _u32p), len(vecs), k, C.byref(o), C.byref(m), C.byref(h))
        offs = self._take(o, n + 1, np.uint64)
        return dict(offsets=offs, minimizers=self._take(m, int(offs[-1]), np.uint32), hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2))

    def unitig_edges(self, unitig_offsets, unitig_minimizers, k):
        """indexUnitigEdges + computeUnitigEdges on the records of unitigGraph.nodes.bin -> dict(offsets [2n+1], targets,
        n_edges, checksum): list 2i = successors of record i (unitigIndex 2i), list 2i+1 = its predecessors."""
        offs = np.ascontiguousarray(unitig_offsets, dtype=np.uint64)
        mins = np.ascontiguousarray(unitig_minimizers, dtype=np.uint32)
        n = len(offs) - 1
        eo = _u64p(); et = _u32p(); cs = C.c_uint64(0)
        ne = self.lib.orc_unitig_edges(_p(mins, _u32p), _p(offs, _u64p), n, k, C.byref(eo), C.byref(et), C.byref(cs))
        return dict(offsets=self._take(eo, 2 * n + 1, np.uint64), targets=self._take(et, ne, np.uint32), n_edges=int(ne),
                    checksum=int(cs.value))

    def checksum(self, hashes: np.ndarray, abundances: np.ndarray) -> int:
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        abundances = np.ascontiguousarray(abundances, dtype=np.uint32)
        return int(self.lib.orc_table_checksum(_p(hashes, _u64p), _p(abundances, _u32p), len(abundances)))


class Reference(_Lib):
    """The reference's own code (oracle/_ref).  Raises FileNotFoundError when
    the library has not been built (it cannot be built on the GPU box)."""

    def __init__(self):
        build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO, "ref_")
        L = self.lib
        L.ref_purge_palindrome.restype = C.c_size_t
        L.ref_purge_palindrome.argtypes = [_u32p, C.c_size_t, C.c_size_t, C.c_size_t, _u32p]
        L.ref_count.restype = C.c_size_t
        L.ref_count.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, C.c_uint32, C.c_int, C.POINTER(_u32p),
                                C.POINTER(_u64p), C.POINTER(_u32p), _u64p, _u64p]
        L.ref_pipeline.restype = C.c_size_t
        L.ref_pipeline.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                   C.c_uint32, C.c_int, _u64p, _u64p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_max_threads.restype = C.c_int
        L.ref_read_selection.restype = C.c_int
        L.ref_read_selection.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_double)]
        L.ref_graph_firstpass.restype = C.c_size_t
        L.ref_graph_firstpass.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, C.c_uint32, C.c_int, C.c_char_p,
                                          C.POINTER(_u32p), C.POINTER(_u64p), C.POINTER(_u32p), _u64p, _u64p,
                                          C.POINTER(C.c_double)]
        L.ref_edge_index.restype = C.c_size_t
        L.ref_edge_index.argtypes = [_u32p, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.POINTER(_u64p),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_unitig_edges.restype = C.c_size_t
        L.ref_unitig_edges.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.POINTER(_u64p), C.POINTER(_u32p),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_unitig_nodes.restype = C.c_size_t
        L.ref_unitig_nodes.argtypes = [_u32p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_u64p), C.POINTER(_u32p)]
        L.ref_edge_values.restype = C.c_size_t
        L.ref_edge_values.argtypes = [_u32p, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.POINTER(_u64p), C.POINTER(_u8p),
                                      C.POINTER(_u32p), C.POINTER(_u8p)]
        L.ref_graph_next_k.restype = C.c_size_t
        L.ref_graph_next_k.argtypes = [_u32p, _u64p, C.c_size_t, C.c_int, _u64p, _u32p, C.c_size_t, C.c_int, C.c_int,
                                       C.c_char_p, C.POINTER(_u32p), C.POINTER(_u64p), C.POINTER(_u32p)]

    def purge_palindrome(self, m: np.ndarray, first_k: int, last_k: int):
        m = np.ascontiguousarray(m, dtype=np.uint32)
        out = np.zeros(len(m) + 1, dtype=np.uint32)
        n = self.lib.ref_purge_palindrome(_p(m, _u32p), len(m), first_k, last_k, _p(out, _u32p))
        return out[:n].copy()

    def count(self, mins, offs, k, min_abundance=2, threads=1):
        mins = np.ascontiguousarray(mins, dtype=np.uint32)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        v = _u32p(); h = _u64p(); a = _u32p()
        ni = C.c_uint64(0); nd = C.c_uint64(0)
        n = self.lib.ref_count(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, min_abundance, threads,
                               C.byref(v), C.byref(h), C.byref(a), C.byref(ni), C.byref(nd))
        vecs = self._take(v, n * k, np.uint32).reshape(n, k)
        hashes = self._take(h, n * 2, np.uint64).reshape(n, 2)
        abund = self._take(a, n, np.uint32)
        return dict(vecs=vecs, hashes=hashes, abundances=abund, n_instances=int(ni.value), n_distinct=int(nd.value))

    def graph_firstpass(self, mins, offs, k, min_abundance=0, threads=1):
        """The reference's own KminmerCounter (+ rescueKminmers when min_abundance <= 1) in a scratch dir."""
        import tempfile
        mins = np.ascontiguousarray(mins, dtype=np.uint32); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        v = _u32p(); h = _u64p(); a = _u32p()
        ns = C.c_uint64(0); nr = C.c_uint64(0); sec = C.c_double(0)
        with tempfile.TemporaryDirectory() as d:
            n = self.lib.ref_graph_firstpass(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, min_abundance, threads,
                                             d.encode(), C.byref(v), C.byref(h), C.byref(a), C.byref(ns), C.byref(nr),
                                             C.byref(sec))
        return dict(vecs=self._take(v, n * k, np.uint32).reshape(n, k), hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2),
                    abundances=self._take(a, n, np.uint32), n_solid=int(ns.value), n_rescued=int(nr.value),
                    seconds=float(sec.value))

    def edge_index(self, vecs, k, threads=1):
        """The reference's own CreateMdbg::EdgeIndexer on a node file (disk partitions + sortParallel) in a scratch
        dir -> dict(hashes [n,2] (h1,h2) in edges.bin order, nb_edges, checksum)."""
        import tempfile
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        h = _u64p(); ne = C.c_uint64(0); cs = C.c_uint64(0)
        with tempfile.TemporaryDirectory() as d:
            n = self.lib.ref_edge_index(_p(vecs, _u32p), len(vecs), k, threads, d.encode(), C.byref(h), C.byref(ne),
                                        C.byref(cs))
        return dict(hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2), nb_edges=int(ne.value), checksum=int(cs.value))

    def edge_values(self, vecs, k, threads=1):
        """The reference's own CreateMdbg::indexEdges (EdgeIndexer, BooPHF, indexEdge over all nodes) in a scratch
        dir -> dict(hashes [n,2] (h1,h2), palindrome [n], minimizers [n,2], flags [n,2]): the raw KminmerEdge33
        slots (flags: 1 isReversed, 2 isPrefix, 4 hasMultipleSuccessors; minimizer 0xFFFFFFFF = empty slot)."""
        import tempfile
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        h = _u64p(); pal = _u8p(); mn = _u32p(); fl = _u8p()
        with tempfile.TemporaryDirectory() as d:
            n = self.lib.ref_edge_values(_p(vecs, _u32p), len(vecs), k, threads, d.encode(), C.byref(h), C.byref(pal),
                                         C.byref(mn), C.byref(fl))
        return dict(hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2), palindrome=self._take(pal, n, np.uint8),
                    minimizers=self._take(mn, 2 * n, np.uint32).reshape(n, 2), flags=self._take(fl, 2 * n, np.uint8).reshape(n, 2))

    def unitig_nodes(self, vecs, k, threads=1, deterministic=True):
        """The reference's own indexEdges + computeUnitigNodes (+ computeDeterministicUnitigs) on a node file in a
        scratch dir -> dict(offsets [n+1], minimizers): the records of unitigGraph.nodes.bin in file order."""
        import tempfile
        vecs = np.ascontiguousarray(vecs, dtype=np.uint32).reshape(-1, k)
        o = _u64p(); m = _u32p()
        with tempfile.TemporaryDirectory() as d:
            n = self.lib.ref_unitig_nodes(_p(vecs, _u32p), len(vecs), k, threads, 1 if deterministic else 0, d.encode(),
                                          C.byref(o), C.byref(m))
        offs = self._take(o, n + 1, np.uint64)
        return dict(offsets=offs, minimizers=self._take(m, int(offs[-1]), np.uint32))

    def unitig_edges(self, unitig_offsets, unitig_minimizers, k, threads=1):
        """The reference's own indexUnitigEdges + computeUnitigEdges on a unitigGraph.nodes.bin written from the given
        records, in a scratch dir -> dict(offsets [2n+1], targets, n_edges, checksum) with the records of
        unitigGraph.edges.successors.bin re-ordered by unitigIndex (list 2i successors, 2i+1 predecessors; the order inside
        a list is the reference's arrival order: deterministic with one thread)."""
        import tempfile
        offs = np.ascontiguousarray(unitig_offsets, dtype=np.uint64)
        mins = np.ascontiguousarray(unitig_minimizers, dtype=np.uint32)
        n = len(offs) - 1
        eo = _u64p(); et = _u32p(); ne = C.c_uint64(0); cs = C.c_uint64(0)
        with tempfile.TemporaryDirectory() as d:
            tot = self.lib.ref_unitig_edges(_p(mins, _u32p), _p(offs, _u64p), n, k, threads, d.encode(), C.byref(eo), C.byref(et),
                                            C.byref(ne), C.byref(cs))
        return dict(offsets=self._take(eo, 2 * n + 1, np.uint64), targets=self._take(et, tot, np.uint32), n_edges=int(ne.value),
                    checksum=int(cs.value))

    def graph_next_k(self, mins, offs, k, prev_hashes, prev_ab, use_counter=False, threads=1):
        import tempfile
        mins = np.ascontiguousarray(mins, dtype=np.uint32); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ph = np.ascontiguousarray(prev_hashes, dtype=np.uint64); pa = np.ascontiguousarray(prev_ab, dtype=np.uint32)
        v = _u32p(); h = _u64p(); a = _u32p()
        with tempfile.TemporaryDirectory() as d:
            n = self.lib.ref_graph_next_k(_p(mins, _u32p), _p(offs, _u64p), len(offs) - 1, k, _p(ph, _u64p), _p(pa, _u32p),
                                          len(pa), int(use_counter), threads, d.encode(), C.byref(v), C.byref(h),
                                          C.byref(a))
        vecs = self._take(v, n * k, np.uint32).reshape(n, k)
        return dict(vecs=vecs if use_counter else None, hashes=self._take(h, 2 * n, np.uint64).reshape(n, 2),
                    abundances=self._take(a, n, np.uint32))

    def read_selection(self, fastq_paths, l=15, density=0.005, hpc=True, threads=1, skip_correction=False,
                       workdir=None):
        """The reference's whole readSelection stage on FASTA/FASTQ files.  Returns the parsed
        read_data_init.txt records, read_stats.txt and read_data_corrected.txt."""
        import tempfile
        ctx = tempfile.TemporaryDirectory() if workdir is None else None
        d = ctx.name if ctx else workdir
        with open(os.path.join(d, "input.txt"), "w") as f:
            for pth in fastq_paths:
                f.write(str(pth) + "\n")
        sec = C.c_double(0)
        self.lib.ref_read_selection(os.path.join(d, "input.txt").encode(), d.encode(), l, density, int(hpc), threads,
                                    int(skip_correction), C.byref(sec))
        out = dict(seconds=float(sec.value), records=parse_read_data(os.path.join(d, "read_data_init.txt"), True),
                   stats=parse_read_stats(os.path.join(d, "read_stats.txt")))
        bl = os.path.join(d, "repetitiveMinimizers.bin")            # u32 list (ReadSelection.hpp:553-556)
        out["blacklist"] = np.fromfile(bl, dtype=np.uint32) if os.path.exists(bl) else np.zeros(0, np.uint32)
        cp = os.path.join(d, "read_data_corrected.txt")
        out["corrected"] = parse_read_data(cp, False) if os.path.exists(cp) else None
        if ctx:
            ctx.cleanup()
        return out

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def pipeline(self, bases: np.ndarray, offsets: np.ndarray, l: int, density: float, hpc: bool, k: int,
                 purge_last_k: int = 0, min_abundance: int = 2, threads: int = 1, assembly_density: float = 0.0):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nm = C.c_uint64(0); cs = C.c_uint64(0); ts = C.c_double(0); tc = C.c_double(0)
        if assembly_density > 0:                  # ONT: sketch at `density`, Utils::applyDensityThreshold, purge, count
            f = self.lib.ref_pipeline2
            f.restype = C.c_size_t
            f.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_uint32,
                          C.c_int] + [C.POINTER(C.c_uint64)] * 3 + [C.POINTER(C.c_double)] * 2
            ns = C.c_uint64(0)
            n = f(bases.ctypes.data, _p(offsets, _u64p), len(offsets) - 1, l, density, int(hpc), assembly_density, k,
                  purge_last_k, min_abundance, threads, C.byref(ns), C.byref(nm), C.byref(cs), C.byref(ts), C.byref(tc))
            return dict(n_solid=int(n), n_minimizers=int(nm.value), n_minimizers_sketch=int(ns.value), checksum=int(cs.value),
                        seconds_sketch=float(ts.value), seconds_count=float(tc.value))
        n = self.lib.ref_pipeline(bases.ctypes.data, _p(offsets, _u64p), len(offsets) - 1, l, density, int(hpc), k,
                                  purge_last_k, min_abundance, threads, C.byref(nm), C.byref(cs), C.byref(ts), C.byref(tc))
        return dict(n_solid=int(n), n_minimizers=int(nm.value), checksum=int(cs.value),
                    seconds_sketch=float(ts.value), seconds_count=float(tc.value))


def canonical_edge_values(raw: dict) -> dict:
    """Order-free content of the reference's KminmerEdge33 slots (Reference.edge_values): {(h1, h2): ((count, minimizer,
    isReversed, isPrefix) of class A, ... of class B)} with class A = {(r,p): r == p} (and every slot of a
    palindromic key), class B = {r != p}; count 2 = hasMultipleSuccessors, whose recorded minimizer is arbitrary."""
    out = {}
    for h, pal, mins, flags in zip(raw["hashes"], raw["palindrome"], raw["minimizers"], raw["flags"]):
        cls = [(0, 0, 0, 0), (0, 0, 0, 0)]
        for m, f in zip(mins, flags):
            if int(m) == 0xFFFFFFFF:
                continue
            r, p, multi = int(f) & 1, (int(f) >> 1) & 1, (int(f) >> 2) & 1
            c = 0 if (pal or r == p) else 1
            assert cls[c][0] == 0, "two slots of one orientation class"
            cls[c] = (2, 0, 0, 0) if multi else (1, int(m), r, p)
        out[(int(h[0]), int(h[1]))] = tuple(cls)
    return out
