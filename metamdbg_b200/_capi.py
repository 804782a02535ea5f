"""ctypes binding of libmdbg_b200.so (include/mdbg_b200.h).

The library is the product; this module only declares its C ABI.  It fails
loudly when the shared object is missing or when no CUDA device is usable --
there is no CPU fallback and nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmdbg_b200.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class MdbgParams(C.Structure):
    _fields_ = [("minimizer_size", C.c_uint32), ("density", C.c_float), ("use_hpc", C.c_uint32),
                ("blacklist", u32p), ("n_blacklist", C.c_uint64)]


class SketchOut(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_minimizers", C.c_uint64), ("min_offsets", u64p),
                ("minimizers", u32p), ("positions", u32p), ("directions", u8p)]


class SketchDev(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_minimizers", C.c_uint64), ("d_min_offsets", C.c_void_p),
                ("d_minimizers", C.c_void_p), ("d_positions", C.c_void_p), ("d_directions", C.c_void_p)]


class TableOut(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_entries", C.c_uint64), ("hashes", u64p), ("abundances", u32p),
                ("kminmers", u32p), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64),
                ("checksum", C.c_uint64), ("n_rescued", C.c_uint64)]


class TableDev(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_entries", C.c_uint64), ("d_hashes", C.c_void_p), ("d_abundances", C.c_void_p),
                ("d_kminmers", C.c_void_p), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64),
                ("checksum", C.c_uint64), ("n_rescued", C.c_uint64)]


class FastxInfo(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("consumed_bytes", C.c_uint64), ("n_bases", C.c_uint64), ("format", C.c_int32)]


class KStats(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_entries", C.c_uint64), ("n_distinct", C.c_uint64), ("n_instances", C.c_uint64),
                ("checksum", C.c_uint64), ("n_reads_rescued", C.c_uint64)]


class PostingsOut(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_keys", C.c_uint64), ("n_postings", C.c_uint64), ("hashes", u64p), ("offsets", u64p),
                ("reads", u32p), ("windows", u32p), ("d_hashes", C.c_void_p), ("d_offsets", C.c_void_p),
                ("d_reads", C.c_void_p), ("d_windows", C.c_void_p)]


class RepeatsOut(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("n_selected", C.c_uint64), ("minimizers", u32p), ("counts", u32p),
                ("min_count_selected", C.c_uint32), ("n_with_min_count", C.c_uint64), ("n_with_min_count_selected", C.c_uint64)]


class AuxOut(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("mean_quality", C.POINTER(C.c_float)), ("complexity", C.POINTER(C.c_double)),
                ("low_complexity", u8p), ("qualities", u8p)]


class BatchInfo(C.Structure):
    _fields_ = [("n_pieces", C.c_uint64), ("n_pieces_pipelined", C.c_uint64), ("n_buffer_growths", C.c_uint64),
                ("n_direct_pieces", C.c_uint64), ("overflow_fallback", C.c_int32), ("packed", C.c_int32),
                ("pack_gb_per_s", C.c_double), ("pack_isa", C.c_char_p), ("host_threads", C.c_int32)]


class EdgesOut(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_nodes", C.c_uint64), ("n_edges", C.c_uint64), ("hashes", u64p),
                ("checksum", C.c_uint64), ("values", u64p)]


class UnitigsOut(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_nodes", C.c_uint64), ("n_unitigs", C.c_uint64), ("n_minimizers", C.c_uint64),
                ("n_circular", C.c_uint64), ("n_cycle_nodes", C.c_uint64), ("offsets", u64p), ("minimizers", C.POINTER(C.c_uint32)),
                ("hashes", u64p), ("circular", C.POINTER(C.c_uint8)), ("order", C.POINTER(C.c_uint32)), ("node_abundances", C.POINTER(C.c_uint32)),
                ("n_unitig_edges", C.c_uint64), ("checksum_edges", C.c_uint64), ("edge_offsets", u64p),
                ("edge_targets", C.POINTER(C.c_uint32)), ("checksum_nodes", C.c_uint64), ("checksum_abundances", C.c_uint64), ("d_offsets", C.c_void_p),
                ("d_minimizers", C.c_void_p)]


class AutotuneOut(C.Structure):
    _fields_ = [("n_variants", C.c_int32), ("chosen", C.c_int32), ("identical", C.c_int32 * 4), ("ms", C.c_float * 4),
                ("n_reads", C.c_uint32), ("n_minimizers", C.c_uint64)]


# every symbol include/mdbg_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mdbg_ctx_create": (C.c_int, [C.c_int, C.POINTER(MdbgParams), C.POINTER(C.c_void_p)]),
    "mdbg_ctx_destroy": (None, [C.c_void_p]),
    "mdbg_last_error": (C.c_char_p, [C.c_void_p]),
    "mdbg_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdbg_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "mdbg_ctx_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "mdbg_ctx_bytes_moved": (C.c_int, [C.c_void_p, u64p, u64p]),
    "mdbg_ctx_allocations": (C.c_int, [C.c_void_p, u64p, u64p]),
    "mdbg_ctx_phase_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "mdbg_ctx_phase_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_char_p), C.c_int]),
    "mdbg_ctx_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "mdbg_ctx_kernel_time_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "mdbg_sketch_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(SketchOut)]),
    "mdbg_sketch_fastx": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(SketchOut),
                                    C.POINTER(FastxInfo)]),
    "mdbg_host_pack_read": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "mdbg_sketch_batch_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.c_uint32, C.c_int, C.POINTER(SketchOut)]),
    "mdbg_ctx_set_host_packing": (C.c_int, [C.c_void_p, C.c_int]),
    "mdbg_ctx_last_batch_info": (C.c_int, [C.c_void_p, C.POINTER(BatchInfo)]),
    "mdbg_sketch_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int,
                                           C.POINTER(SketchDev)]),
    "mdbg_ctx_set_sketch_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "mdbg_ctx_get_sketch_variant": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mdbg_ctx_autotune_sketch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                           C.POINTER(AutotuneOut)]),
    "mdbg_sketch_batch_device_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                  C.c_uint64, C.c_int, C.POINTER(SketchDev)]),
    "mdbg_pack_device_words": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "mdbg_pack_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mdbg_sketch_batch_device_packed2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                   C.c_uint64, C.c_int, C.POINTER(SketchDev)]),
    "mdbg_sketch_fetch": (C.c_int, [C.c_void_p, C.POINTER(SketchOut)]),
    "mdbg_ctx_set_read_filters": (C.c_int, [C.c_void_p, C.c_int]),
    "mdbg_sketch_batch_q": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                      C.POINTER(SketchOut), C.POINTER(AuxOut)]),
    "mdbg_store_clear": (C.c_int, [C.c_void_p]),
    "mdbg_store_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mdbg_store_size": (C.c_int, [C.c_void_p, u64p, u64p]),
    "mdbg_store_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdbg_store_apply_density": (C.c_int, [C.c_void_p, C.c_float, u64p]),
    "mdbg_purge_palindromes": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, u64p]),
    "mdbg_multi_k_run": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(KStats)]),
    "mdbg_count_postings": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PostingsOut)]),
    "mdbg_store_repetitive_minimizers": (C.c_int, [C.c_void_p, C.c_float, C.POINTER(RepeatsOut)]),
    "mdbg_ctx_set_blacklist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "mdbg_count_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64]),
    "mdbg_count_add_store": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "mdbg_count_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mdbg_count_finalize": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TableOut)]),
    "mdbg_count_finalize_device": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TableDev)]),
    "mdbg_count_stats": (C.c_int, [C.c_void_p, C.c_uint32, u64p, u64p, u64p, u64p]),
    "mdbg_count_rescue": (C.c_int, [C.c_void_p, u64p]),
    "mdbg_prev_from_current": (C.c_int, [C.c_void_p, C.c_uint32]),
    "mdbg_prev_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "mdbg_count_add_store_next_k": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "mdbg_edges_index": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(EdgesOut)]),
    "mdbg_unitigs_build": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(UnitigsOut)]),
    "mdbg_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mdbg_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mdbg_count_merge": (C.c_int, [C.c_void_p]),
    "mdbg_count_merge_hashes": (C.c_int, [C.c_void_p]),
    "mdbg_synth_fill_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_uint64, C.c_uint64, C.c_uint32]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the engine; raises if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with metamdbg_b200/csrc/Makefile "
                           "(__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(lib, name)          # AttributeError = header/library mismatch
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib
