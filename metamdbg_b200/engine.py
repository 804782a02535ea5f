"""Host-side mirror of the reference's operator interface for the hot path,
implemented on top of the C ABI (include/mdbg_b200.h) -- every call below is a
call into libmdbg_b200.so; no computation happens in Python.

Naming follows metaMDBG (paths relative to its source tree):
  * ``MinimizerParser.parse``   <- src/utils/kmer/Kmer.hpp:1339-1456 (with the
    EncoderRLE step of src/Commons.hpp:4163-4203 folded in, as
    ReadSelectionFunctor chains them, src/readSelection/ReadSelection.hpp:682-690)
  * ``KminmerCounter``          <- src/graph/CreateMdbg.hpp:3591-3883
  * ``purge_palindromes``       <- src/Commons.hpp:1617-1723
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import AutotuneOut, AuxOut, BatchInfo, EdgesOut, FastxInfo, KStats, MdbgParams, PostingsOut, RepeatsOut, SketchDev, SketchOut, TableDev, TableOut

STATUS = {0: "OK", 1: "CUDA", 2: "ARG", 3: "STATE", 4: "TABLE_FULL", 5: "NCCL", 6: "OOM"}


class MdbgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"mdbg status {STATUS.get(status, status)}: {message}")
        self.status = status


@dataclass
class Sketch:
    """CSR of one sketched batch (host copies)."""
    min_offsets: np.ndarray   # uint64 [n_reads+1]
    minimizers: np.ndarray    # uint32
    positions: np.ndarray     # uint32
    directions: np.ndarray    # uint8

    @property
    def n_reads(self) -> int:
        return len(self.min_offsets) - 1

    def read(self, r: int):
        lo, hi = int(self.min_offsets[r]), int(self.min_offsets[r + 1])
        return self.minimizers[lo:hi], self.positions[lo:hi], self.directions[lo:hi]


@dataclass
class CountTable:
    """Finalised k-min-mer abundance table (unordered, like kminmerData_abundance.txt)."""
    k: int
    hashes: np.ndarray        # uint64 [n, 2]: (low64 = Murmur h2, high64 = Murmur h1), the on-disk u128 bytes
    abundances: np.ndarray    # uint32 [n]
    kminmers: np.ndarray      # uint32 [n, k]
    n_instances: int
    n_distinct: int
    checksum: int
    n_rescued: int = 0

    def as_dict(self) -> dict:
        """{(h1, h2): abundance} with h1 = high 64 bits, h2 = low 64 bits."""
        return {(int(h[1]), int(h[0])): int(a) for h, a in zip(self.hashes, self.abundances)}


class Engine:
    """One context = one GPU.  Not thread-safe (serialise calls, as the C ABI requires)."""

    def __init__(self, minimizer_size: int = 15, density: float = 0.005, use_hpc: bool = True,
                 blacklist: np.ndarray | None = None, device: int = 0):
        self._lib = _capi.load()
        self._ctx = C.c_void_p()
        bl = None
        if blacklist is not None and len(blacklist):
            bl = np.ascontiguousarray(blacklist, dtype=np.uint32)
        p = MdbgParams(minimizer_size, float(np.float32(density)), 1 if use_hpc else 0,
                       bl.ctypes.data_as(_capi.u32p) if bl is not None else None, 0 if bl is None else len(bl))
        st = self._lib.mdbg_ctx_create(device, C.byref(p), C.byref(self._ctx))
        if st != 0:
            raise MdbgError(st, self._lib.mdbg_last_error(None).decode())
        self.minimizer_size, self.density, self.use_hpc, self.device = minimizer_size, density, use_hpc, device

    # -- plumbing -------------------------------------------------------------
    def _ck(self, st: int):
        if st != 0:
            raise MdbgError(st, self._lib.mdbg_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.mdbg_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        self._ck(self._lib.mdbg_ctx_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._ck(self._lib.mdbg_ctx_synchronize(self._ctx))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.mdbg_ctx_kernel_launches(self._ctx))

    def bytes_moved(self) -> tuple[int, int]:
        """(H2D, D2H) bytes enqueued by the host-buffer entry points so far."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._lib.mdbg_ctx_bytes_moved(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def allocations(self) -> tuple[int, int]:
        """(device buffers (re)allocated so far, times the two table buffers traded places instead of a reallocation)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._lib.mdbg_ctx_allocations(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def phase_profile(self, on: bool = True):
        """Exclusive per-phase times of the table / collective paths (diagnostic: adds stream synchronisations)."""
        self._ck(self._lib.mdbg_ctx_phase_profile(self._ctx, int(on)))

    def phase_times(self) -> dict:
        ms = (C.c_double * 16)()
        names = (C.c_char_p * 16)()
        n = self._lib.mdbg_ctx_phase_times(self._ctx, ms, names, 16)
        return {names[i].decode(): round(float(ms[i]), 3) for i in range(n) if ms[i] > 0}

    def enable_timing(self, on: bool = True):
        self._ck(self._lib.mdbg_ctx_enable_timing(self._ctx, int(on)))

    def kernel_time_ms(self, which: int) -> float:
        """Device time of the last sketch (0) / insert (1) kernel launch."""
        ms = C.c_float(0)
        self._ck(self._lib.mdbg_ctx_kernel_time_ms(self._ctx, which, C.byref(ms)))
        return float(ms.value)

    # -- sketch ---------------------------------------------------------------
    @staticmethod
    def _copy_sketch(out: SketchOut) -> Sketch:
        n, t = out.n_reads, out.n_minimizers
        mo = np.ctypeslib.as_array(out.min_offsets, shape=(n + 1,)).copy()
        if t:
            m = np.ctypeslib.as_array(out.minimizers, shape=(t,)).copy()
            p = np.ctypeslib.as_array(out.positions, shape=(t,)).copy()
            d = np.ctypeslib.as_array(out.directions, shape=(t,)).copy()
        else:
            m = np.zeros(0, np.uint32); p = np.zeros(0, np.uint32); d = np.zeros(0, np.uint8)
        return Sketch(mo, m, p, d)

    def sketch_batch(self, bases: np.ndarray, offsets: np.ndarray, append_to_store: bool = False,
                     fetch: bool = True) -> Sketch | None:
        """Host reads (ASCII uint8 + uint64 offsets[n+1]) -> minimizer CSR."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = SketchOut()
        self._ck(self._lib.mdbg_sketch_batch(self._ctx, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                             int(append_to_store), C.byref(out) if fetch else None))
        return self._copy_sketch(out) if fetch else None

    def sketch_fastx(self, text: bytes | np.ndarray, is_final: bool = True, append_to_store: bool = False):
        """Raw FASTQ / FASTA text -> (Sketch of its complete records, info dict); record split on the device."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        out, info = SketchOut(), FastxInfo()
        self._ck(self._lib.mdbg_sketch_fastx(self._ctx, buf.ctypes.data if len(buf) else None, len(buf), int(is_final),
                                             int(append_to_store), C.byref(out), C.byref(info)))
        return self._copy_sketch(out), dict(n_records=int(info.n_records), consumed_bytes=int(info.consumed_bytes),
                                            n_bases=int(info.n_bases), format={1: "fastq", 2: "fasta"}.get(info.format))

    def host_pack_reads(self, bases: np.ndarray, offsets: np.ndarray):
        """What a packing reader does, read by read (mdbg_host_pack_read): -> (words u32, read_src u64, ascii spill u8).
        Reads start on 16-byte boundaries of the word array; unpackable reads go to the spill."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        lens = np.diff(offsets.astype(np.int64))
        woff = np.zeros(n + 1, np.int64)
        woff[1:] = np.cumsum(((lens + 63) >> 6) << 2)
        words = np.zeros(int(woff[-1]) + 4, np.uint32)
        src = np.zeros(n, np.uint64)
        spill, n_spill = [], 0
        for r in range(n):
            lo, ln = int(offsets[r]), int(lens[r])
            ok = self._lib.mdbg_host_pack_read(C.c_void_p(bases.ctypes.data + lo), ln, C.c_void_p(words.ctypes.data + 4 * int(woff[r])))
            if ok:
                src[r] = woff[r]
            else:
                src[r] = (1 << 63) | n_spill
                spill.append(bases[lo:lo + ln])
                n_spill += (ln + 15) & ~15
                spill.append(np.zeros(((ln + 15) & ~15) - ln, np.uint8))
        asc = np.concatenate(spill).astype(np.uint8) if spill else np.zeros(0, np.uint8)
        return words, src, asc

    def sketch_batch_packed(self, words: np.ndarray, read_src: np.ndarray, ascii_spill: np.ndarray, offsets: np.ndarray,
                            append_to_store: bool = False, fetch: bool = True):
        """mdbg_sketch_batch_packed: a host batch that was 2-bit packed by the caller."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        read_src = np.ascontiguousarray(read_src, dtype=np.uint64)
        ascii_spill = np.ascontiguousarray(ascii_spill, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = SketchOut()
        self._ck(self._lib.mdbg_sketch_batch_packed(self._ctx, words.ctypes.data, len(words), read_src.ctypes.data,
                                                    ascii_spill.ctypes.data if len(ascii_spill) else None, len(ascii_spill),
                                                    offsets.ctypes.data, len(offsets) - 1, int(append_to_store),
                                                    C.byref(out) if fetch else None))
        return self._copy_sketch(out) if fetch else None

    def sketch_batch_packed_ptr(self, words_ptr: int, n_words: int, read_src: np.ndarray, offsets: np.ndarray,
                                append_to_store: bool = True) -> int:
        """Same from raw (pinned) host pointers, clean reads only; returns the number of minimizers (bench e2e leg)."""
        read_src = np.ascontiguousarray(read_src, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = SketchOut()
        self._ck(self._lib.mdbg_sketch_batch_packed(self._ctx, C.c_void_p(words_ptr), n_words, read_src.ctypes.data, None, 0,
                                                    offsets.ctypes.data, len(offsets) - 1, int(append_to_store), C.byref(out)))
        return int(out.n_minimizers)

    def set_host_packing(self, on: bool | int | None = True):
        """2-bit pack host batches before H2D: True / False / None = automatic (default) / 2 = hybrid (experimental)."""
        self._ck(self._lib.mdbg_ctx_set_host_packing(self._ctx, -1 if on is None else int(on)))

    def last_batch_info(self) -> dict:
        """How the last host batch travelled (pieces, piece-wise tails, packer throughput); see mdbg_batch_info."""
        b = BatchInfo()
        self._ck(self._lib.mdbg_ctx_last_batch_info(self._ctx, C.byref(b)))
        return {"n_pieces": int(b.n_pieces), "n_pieces_pipelined": int(b.n_pieces_pipelined),
                "n_buffer_growths": int(b.n_buffer_growths), "n_direct_pieces": int(b.n_direct_pieces),
                "overflow_fallback": bool(b.overflow_fallback), "packed": bool(b.packed),
                "pack_gb_per_s": float(b.pack_gb_per_s), "pack_isa": (b.pack_isa or b"").decode(),
                "host_threads": int(b.host_threads)}

    def set_read_filters(self, filter_low_complexity: bool = True):
        self._ck(self._lib.mdbg_ctx_set_read_filters(self._ctx, int(filter_low_complexity)))

    def sketch_batch_q(self, bases: np.ndarray, quals: np.ndarray | None, offsets: np.ndarray,
                       append_to_store: bool = False):
        """Sketch + the side outputs of ReadSelectionFunctor (mean quality, complexity, per-minimizer quality).
        Returns (Sketch, dict)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if quals is not None:
            quals = np.ascontiguousarray(quals, dtype=np.uint8)
        out, aux = SketchOut(), AuxOut()
        self._ck(self._lib.mdbg_sketch_batch_q(self._ctx, bases.ctypes.data, quals.ctypes.data if quals is not None else None,
                                               offsets.ctypes.data, len(offsets) - 1, int(append_to_store),
                                               C.byref(out), C.byref(aux)))
        sk = self._copy_sketch(out)
        n, t = sk.n_reads, len(sk.minimizers)
        take = lambda p, m, dt: np.ctypeslib.as_array(p, shape=(m,)).copy() if m else np.zeros(0, dt)
        return sk, dict(mean_quality=take(aux.mean_quality, n, np.float32), complexity=take(aux.complexity, n, np.float64),
                        low_complexity=take(aux.low_complexity, n, np.uint8), qualities=take(aux.qualities, t, np.uint8))

    def sketch_batch_ptr(self, bases_ptr: int, offsets: np.ndarray, append_to_store: bool = True,
                         fetch: bool = True) -> int:
        """Host reads given as a raw pointer (e.g. a pinned buffer) + uint64 offsets; the CSR lands in the
        library's pinned result buffers (no numpy copy).  Returns the number of minimizers."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = SketchOut()
        self._ck(self._lib.mdbg_sketch_batch(self._ctx, C.c_void_p(bases_ptr), offsets.ctypes.data, len(offsets) - 1,
                                             int(append_to_store), C.byref(out) if fetch else None))
        return int(out.n_minimizers) if fetch else 0

    def sketch_batch_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int,
                            append_to_store: bool = False) -> SketchDev:
        out = SketchDev()
        self._ck(self._lib.mdbg_sketch_batch_device(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr),
                                                    n_reads, n_bases, int(append_to_store), C.byref(out)))
        return out

    def set_sketch_variant(self, variant: int):
        """Arithmetic variant of the sketch kernel's unrolled l = 15 block (identical results, sketch.cu)."""
        self._ck(self._lib.mdbg_ctx_set_sketch_variant(self._ctx, int(variant)))

    @property
    def sketch_variant(self) -> int:
        v = C.c_int(0)
        self._ck(self._lib.mdbg_ctx_get_sketch_variant(self._ctx, C.byref(v)))
        return v.value

    def autotune_sketch(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int) -> dict:
        """Run every sketch-kernel variant on this device batch, compare the results byte for byte on the device
        and keep the fastest identical one (variant 0 is the fallback)."""
        out = AutotuneOut()
        self._ck(self._lib.mdbg_ctx_autotune_sketch(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr),
                                                    n_reads, n_bases, C.byref(out)))
        n = out.n_variants
        return {"chosen": int(out.chosen), "identical": [bool(out.identical[i]) for i in range(n)],
                "ms": [float(out.ms[i]) for i in range(n)], "n_reads": int(out.n_reads),
                "n_minimizers": int(out.n_minimizers)}

    def sketch_batch_device_packed(self, d_packed_ptr: int, d_word_offsets_ptr: int, d_offsets_ptr: int, n_reads: int,
                                   n_bases: int, append_to_store: bool = False) -> SketchDev:
        """Reads resident in HBM as 2-bit codes (16 bases per u32, see include/mdbg_b200.h)."""
        out = SketchDev()
        self._ck(self._lib.mdbg_sketch_batch_device_packed(self._ctx, C.c_void_p(d_packed_ptr),
                                                           C.c_void_p(d_word_offsets_ptr), C.c_void_p(d_offsets_ptr),
                                                           n_reads, n_bases, int(append_to_store), C.byref(out)))
        return out

    def pack_device_words(self, n_bases: int, n_reads: int) -> int:
        """u32 words the packed copy of an ASCII device batch needs (closed form, 16-byte aligned read starts)."""
        return int(self._lib.mdbg_pack_device_words(n_bases, n_reads))

    def pack_device(self, d_bases_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int, d_packed_ptr: int,
                    d_read_src_ptr: int):
        """ASCII reads in HBM -> the 2-bit device layout (one streaming kernel); reads with a byte outside ACGT are
        flagged in d_read_src (bit 63 | byte offset) and stay ASCII."""
        self._ck(self._lib.mdbg_pack_device(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads, n_bases,
                                            C.c_void_p(d_packed_ptr), C.c_void_p(d_read_src_ptr)))

    def sketch_batch_device_packed2(self, d_packed_ptr: int, d_read_src_ptr: int, d_bases_ptr: int, d_offsets_ptr: int,
                                    n_reads: int, n_bases: int, append_to_store: bool = False) -> SketchDev:
        """Sketch of a batch resident in HBM in the packed layout; flagged reads are taken from the ASCII buffer."""
        out = SketchDev()
        self._ck(self._lib.mdbg_sketch_batch_device_packed2(self._ctx, C.c_void_p(d_packed_ptr), C.c_void_p(d_read_src_ptr),
                                                            C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr), n_reads,
                                                            n_bases, int(append_to_store), C.byref(out)))
        return out

    def sketch_fetch(self) -> Sketch:
        out = SketchOut()
        self._ck(self._lib.mdbg_sketch_fetch(self._ctx, C.byref(out)))
        return self._copy_sketch(out)

    # -- minimizer-space read store ----------------------------------------------
    def store_clear(self):
        self._ck(self._lib.mdbg_store_clear(self._ctx))

    def store_append(self, minimizers: np.ndarray, min_offsets: np.ndarray):
        minimizers = np.ascontiguousarray(minimizers, dtype=np.uint32)
        min_offsets = np.ascontiguousarray(min_offsets, dtype=np.uint64)
        self._ck(self._lib.mdbg_store_append(self._ctx, minimizers.ctypes.data, min_offsets.ctypes.data,
                                             len(min_offsets) - 1))

    def store_size(self) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._lib.mdbg_store_size(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def store_fetch(self) -> tuple[np.ndarray, np.ndarray]:
        nr, nm = self.store_size()
        offs = np.zeros(nr + 1, dtype=np.uint64)
        mins = np.zeros(max(nm, 1), dtype=np.uint32)
        self._ck(self._lib.mdbg_store_fetch(self._ctx, offs.ctypes.data, mins.ctypes.data))
        return offs, mins[:nm]

    def store_apply_density(self, density: float) -> int:
        """Utils::applyDensityThreshold on every stored read; returns the number of reads that lost minimizers."""
        n = C.c_uint64(0)
        self._ck(self._lib.mdbg_store_apply_density(self._ctx, float(np.float32(density)), C.byref(n)))
        return int(n.value)

    def repetitive_minimizers(self, fraction: float = 0.00001) -> dict:
        """determineRepetitiveMinimizers on the stored reads -> dict(minimizers, counts, n_distinct, ...)."""
        out = RepeatsOut()
        self._ck(self._lib.mdbg_store_repetitive_minimizers(self._ctx, C.c_float(fraction), C.byref(out)))
        n = int(out.n_selected)
        return dict(minimizers=np.ctypeslib.as_array(out.minimizers, shape=(n,)).copy() if n else np.zeros(0, np.uint32),
                    counts=np.ctypeslib.as_array(out.counts, shape=(n,)).copy() if n else np.zeros(0, np.uint32),
                    n_distinct=int(out.n_distinct), min_count_selected=int(out.min_count_selected),
                    n_with_min_count=int(out.n_with_min_count), n_with_min_count_selected=int(out.n_with_min_count_selected))

    def set_blacklist(self, values: np.ndarray | None):
        v = np.ascontiguousarray(values if values is not None else np.zeros(0, np.uint32), dtype=np.uint32)
        self._ck(self._lib.mdbg_ctx_set_blacklist(self._ctx, v.ctypes.data if len(v) else None, len(v)))

    def purge_palindromes(self, first_k: int, last_k: int) -> int:
        n = C.c_uint64(0)
        self._ck(self._lib.mdbg_purge_palindromes(self._ctx, first_k, last_k, C.byref(n)))
        return int(n.value)

    # -- count table ---------------------------------------------------------------
    def count_begin(self, k: int, expected_distinct: int = 0):
        self._ck(self._lib.mdbg_count_begin(self._ctx, k, expected_distinct))

    def count_add_store(self, read_lo: int = 0, read_hi: int = 2 ** 64 - 1):
        self._ck(self._lib.mdbg_count_add_store(self._ctx, read_lo, read_hi))

    def count_add(self, minimizers: np.ndarray, min_offsets: np.ndarray):
        minimizers = np.ascontiguousarray(minimizers, dtype=np.uint32)
        min_offsets = np.ascontiguousarray(min_offsets, dtype=np.uint64)
        self._ck(self._lib.mdbg_count_add(self._ctx, minimizers.ctypes.data, min_offsets.ctypes.data,
                                          len(min_offsets) - 1))

    def count_stats(self, min_abundance: int = 2) -> dict:
        v = [C.c_uint64(0) for _ in range(4)]
        self._ck(self._lib.mdbg_count_stats(self._ctx, min_abundance, *[C.byref(x) for x in v]))
        return dict(n_entries=int(v[0].value), n_distinct=int(v[1].value), n_instances=int(v[2].value),
                    checksum=int(v[3].value))

    def count_finalize_device(self, min_abundance: int = 2) -> dict:
        """Emit the table into device arrays (hashes, abundances, normalized vectors); only the statistics reach the host."""
        out = TableDev()
        self._ck(self._lib.mdbg_count_finalize_device(self._ctx, min_abundance, C.byref(out)))
        return dict(k=int(out.k), n_entries=int(out.n_entries), d_hashes=out.d_hashes, d_abundances=out.d_abundances,
                    d_kminmers=out.d_kminmers, n_instances=int(out.n_instances), n_distinct=int(out.n_distinct),
                    checksum=int(out.checksum), n_rescued=int(out.n_rescued))

    def count_finalize(self, min_abundance: int = 2) -> CountTable:
        out = TableOut()
        self._ck(self._lib.mdbg_count_finalize(self._ctx, min_abundance, C.byref(out)))
        n, k = int(out.n_entries), int(out.k)
        if n:
            h = np.ctypeslib.as_array(out.hashes, shape=(2 * n,)).copy().reshape(n, 2)
            a = np.ctypeslib.as_array(out.abundances, shape=(n,)).copy()
            v = (np.ctypeslib.as_array(out.kminmers, shape=(n * k,)).copy().reshape(n, k) if out.kminmers
                 else np.zeros((0, k), np.uint32))          # keys-only merge: no vectors on this rank
        else:
            h = np.zeros((0, 2), np.uint64); a = np.zeros(0, np.uint32); v = np.zeros((0, k), np.uint32)
        return CountTable(k, h, a, v, int(out.n_instances), int(out.n_distinct), int(out.checksum), int(out.n_rescued))

    def multi_k_run(self, first_k: int = 4, last_k: int = 21, min_abundance: int = 2, rescue: bool = False,
                    merge_mode: int = 0) -> list[dict]:
        """mdbg_multi_k_run: the whole multi-k loop inside the library (merge_mode 0 none / 1 vectors / 2 keys only)."""
        n = last_k - first_k + 1
        st = (KStats * n)()
        self._ck(self._lib.mdbg_multi_k_run(self._ctx, first_k, last_k, min_abundance, int(rescue), merge_mode, st))
        return [dict(k=int(x.k), n_entries=int(x.n_entries), n_distinct=int(x.n_distinct), n_instances=int(x.n_instances),
                     checksum=int(x.checksum), n_reads_rescued=int(x.n_reads_rescued)) for x in st]

    def count_postings(self, min_abundance: int = 2) -> dict:
        """k-min-mer -> (read, window) postings of the current count table (CSR over the emitted keys)."""
        out = PostingsOut()
        self._ck(self._lib.mdbg_count_postings(self._ctx, min_abundance, C.byref(out)))
        n, m = int(out.n_keys), int(out.n_postings)
        return dict(k=int(out.k),
                    hashes=np.ctypeslib.as_array(out.hashes, shape=(2 * n,)).copy().reshape(n, 2) if n else np.zeros((0, 2), np.uint64),
                    offsets=np.ctypeslib.as_array(out.offsets, shape=(n + 1,)).copy(),
                    reads=np.ctypeslib.as_array(out.reads, shape=(m,)).copy() if m else np.zeros(0, np.uint32),
                    windows=np.ctypeslib.as_array(out.windows, shape=(m,)).copy() if m else np.zeros(0, np.uint32))

    def count_rescue(self) -> int:
        """rescueKminmers (CreateMdbg.hpp:4517-4640); returns the number of reads that rescued k-min-mers."""
        n = C.c_uint64(0)
        self._ck(self._lib.mdbg_count_rescue(self._ctx, C.byref(n)))
        return int(n.value)

    # -- multi-k ---------------------------------------------------------------------
    def prev_from_current(self, min_abundance: int = 2):
        self._ck(self._lib.mdbg_prev_from_current(self._ctx, min_abundance))

    def prev_load(self, hashes: np.ndarray, abundances: np.ndarray, clear: bool = True):
        """hashes: uint64 [n, 2] as in CountTable.hashes (low64, high64)."""
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        abundances = np.ascontiguousarray(abundances, dtype=np.uint32)
        self._ck(self._lib.mdbg_prev_load(self._ctx, hashes.ctypes.data, abundances.ctypes.data, len(abundances),
                                          int(clear)))

    def count_add_store_next_k(self, read_lo: int = 0, read_hi: int = 2 ** 64 - 1):
        self._ck(self._lib.mdbg_count_add_store_next_k(self._ctx, read_lo, read_hi))

    # -- edge keys (CreateMdbg::EdgeIndexer, src/graph/CreateMdbg.hpp:4010-4232) ------
    def edges_index(self, min_abundance: int = 2, decode: bool = True) -> dict:
        """Distinct hash128 of the normalized (k-1)-prefix / suffix of every node of the current table.
        -> dict(hashes uint64 [n,2] (low64, high64), n_edges, n_nodes, checksum)."""
        out = EdgesOut()
        self._ck(self._lib.mdbg_edges_index(self._ctx, min_abundance, C.byref(out)))
        n = int(out.n_edges)
        h = np.ctypeslib.as_array(out.hashes, shape=(2 * n,)).reshape(n, 2).copy() if n else np.zeros((0, 2), np.uint64)
        vals = None
        if out.values and not decode:       # raw class words (bit 63 valid, bit 34 multi, bit 33 isPrefix, bit 32 isReversed)
            vals = np.ctypeslib.as_array(out.values, shape=(2 * n,)).reshape(n, 2) if n else np.zeros((0, 2), np.uint64)
            return dict(hashes=h, n_edges=n, n_nodes=int(out.n_nodes), checksum=int(out.checksum), values=None, raw_values=vals)
        if out.values:                      # single context: order-free edge values, two orientation classes per key
            w = np.ctypeslib.as_array(out.values, shape=(2 * n,)).reshape(n, 2).copy() if n else np.zeros((0, 2), np.uint64)
            valid, multi = (w >> np.uint64(63)) & np.uint64(1), (w >> np.uint64(34)) & np.uint64(1)
            single = (valid == 1) & (multi == 0)
            vals = np.zeros((n, 2, 4), np.uint32)          # (count 0/1/2+, minimizer, isReversed, isPrefix)
            vals[..., 0] = np.where(valid == 1, np.where(multi == 1, 2, 1), 0)
            vals[..., 1] = np.where(single, w & np.uint64(0xFFFFFFFF), 0)
            vals[..., 2] = np.where(single, (w >> np.uint64(32)) & np.uint64(1), 0)
            vals[..., 3] = np.where(single, (w >> np.uint64(33)) & np.uint64(1), 0)
        return dict(hashes=h, n_edges=n, n_nodes=int(out.n_nodes), checksum=int(out.checksum), values=vals)

    # -- unitig nodes (CreateMdbg::computeUnitigNodes + computeDeterministicUnitigs, CreateMdbg.cpp:1521-1598, 1001-1043) --
    def unitigs_build(self, min_abundance: int = 2, copy: bool = True) -> dict:
        """Maximal non-branching paths of the current table's node set as normalized minimizer sequences.
        -> dict(offsets [n+1], minimizers, hashes [n,2] (low64, high64), circular [n], order [n]: order[i] = the unitig
        the reference writes as record i of unitigGraph.nodes.bin (unitigIndex 2 * i), n_nodes, n_circular, n_cycle_nodes)."""
        out = _capi.UnitigsOut()
        self._ck(self._lib.mdbg_unitigs_build(self._ctx, min_abundance, C.byref(out)))
        n, t = int(out.n_unitigs), int(out.n_minimizers)
        arr = lambda p, shape, dt: (np.ctypeslib.as_array(p, shape=shape).copy() if copy else np.ctypeslib.as_array(p, shape=shape)) \
            if shape[0] else np.zeros(shape, dt)
        return dict(k=int(out.k), n_nodes=int(out.n_nodes), n_unitigs=n, n_circular=int(out.n_circular),
                    n_cycle_nodes=int(out.n_cycle_nodes),
                    offsets=np.ctypeslib.as_array(out.offsets, shape=(n + 1,)).copy(),
                    minimizers=arr(out.minimizers, (t,), np.uint32),
                    hashes=arr(out.hashes, (2 * n,), np.uint64).reshape(n, 2),
                    circular=arr(out.circular, (n,), np.uint8), order=arr(out.order, (n,), np.uint32),
                    node_abundances=arr(out.node_abundances, (t - n * (int(out.k) - 1),), np.uint32),
                    checksum_nodes=int(out.checksum_nodes), checksum_abundances=int(out.checksum_abundances),
                    n_unitig_edges=int(out.n_unitig_edges), checksum_edges=int(out.checksum_edges),
                    edge_offsets=(np.ctypeslib.as_array(out.edge_offsets, shape=(2 * n + 1,)).copy() if out.edge_offsets else None),
                    edge_targets=(arr(out.edge_targets, (int(out.n_unitig_edges),), np.uint32) if out.edge_offsets else None))

    def unitig_records(self, min_abundance: int = 2) -> dict:
        """The same unitigs in the reference's file order: dict(offsets [n+1], minimizers) = the records of
        unitigGraph.nodes.bin (record i: u32 size, size x u32 minimizers, u32 unitigIndex = 2 * i)."""
        u = self.unitigs_build(min_abundance)
        lens = np.diff(u["offsets"]).astype(np.int64)[u["order"]]
        offs = np.zeros(len(lens) + 1, np.uint64)
        offs[1:] = np.cumsum(lens)
        mins = np.concatenate([u["minimizers"][int(u["offsets"][i]):int(u["offsets"][i + 1])] for i in u["order"]]) \
            if len(lens) else np.zeros(0, np.uint32)
        km1 = u["k"] - 1
        ab = [u["node_abundances"][int(u["offsets"][i]) - int(i) * km1:int(u["offsets"][i + 1]) - (int(i) + 1) * km1] for i in u["order"]]
        return dict(offsets=offs, minimizers=mins.astype(np.uint32), n_circular=u["n_circular"], n_nodes=u["n_nodes"],
                    abundances=np.concatenate(ab).astype(np.uint32) if ab else np.zeros(0, np.uint32),
                    checksum_nodes=u["checksum_nodes"], checksum_abundances=u["checksum_abundances"],
                    edge_offsets=u["edge_offsets"], edge_targets=u["edge_targets"], n_unitig_edges=u["n_unitig_edges"],
                    checksum_edges=u["checksum_edges"])

    # -- multi-GPU -----------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = _capi.load()
        buf = (C.c_uint8 * 128)()
        st = lib.mdbg_nccl_unique_id(buf)
        if st != 0:
            raise MdbgError(st, lib.mdbg_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.mdbg_comm_init(self._ctx, rank, n_ranks, buf))

    def count_merge(self):
        self._ck(self._lib.mdbg_count_merge(self._ctx))

    # -- synthetic reads on the device ------------------------------------------------
    def count_merge_hashes(self):
        """Owner merge of (hash, abundance) only: no k-min-mer vectors on the owner afterwards (multi-k loops)."""
        self._ck(self._lib.mdbg_count_merge_hashes(self._ctx))

    def synth_fill_reads(self, d_bases_ptr: int, d_offsets_ptr: int, d_vstart_ptr: int, d_strand_ptr: int,
                         n_reads: int, read_index_base: int, seed: int, err_q24: int):
        self._ck(self._lib.mdbg_synth_fill_reads(self._ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_offsets_ptr),
                                                 C.c_void_p(d_vstart_ptr), C.c_void_p(d_strand_ptr), n_reads,
                                                 read_index_base, seed, err_q24))


class MinimizerParser:
    """``MinimizerParser(minimizerSize, density, repetitive)`` + ``parse`` of the
    reference (Kmer.hpp:1351-1456), batched: one call sketches many reads."""

    def __init__(self, minimizer_size: int, density: float, repetitive_minimizers: np.ndarray | None = None,
                 use_homopolymer_compression: bool = True, device: int = 0):
        self.engine = Engine(minimizer_size, density, use_homopolymer_compression, repetitive_minimizers, device)

    def parse(self, seq: bytes):
        """Single read -> (minimizers, positions, directions), the three vectors of the reference call."""
        bases = np.frombuffer(seq, dtype=np.uint8)
        sk = self.engine.sketch_batch(bases, np.array([0, len(seq)], dtype=np.uint64))
        return sk.minimizers, sk.positions, sk.directions

    def parse_batch(self, bases: np.ndarray, offsets: np.ndarray) -> Sketch:
        return self.engine.sketch_batch(bases, offsets)


class KminmerCounter:
    """``CreateMdbg::KminmerCounter`` (CreateMdbg.hpp:3591-3883): feed minimizer-space
    reads, get the solid k-min-mer table."""

    def __init__(self, engine: Engine, k: int, expected_distinct: int = 0):
        self.engine, self.k = engine, k
        engine.count_begin(k, expected_distinct)

    def add_reads(self, minimizers: np.ndarray, min_offsets: np.ndarray):
        self.engine.count_add(minimizers, min_offsets)

    def add_store(self):
        self.engine.count_add_store()

    def execute(self, min_abundance: int = 2) -> CountTable:
        return self.engine.count_finalize(min_abundance)
