"""ctypes binding of libccsdt_b200.so (include/ccsdt_b200.h).  No fallback: a missing library is an
ImportError that says how to build it."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CCSDT_B200_LIB") or os.path.join(HERE, "libccsdt_b200.so")  # override: A/B builds in tools/

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


class Options(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("sub", C.c_int32 * 3), ("stages", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("overlap", C.c_int32), ("stagger", C.c_int32), ("verbose", C.c_int32), ("symmetry", C.c_int32),
                ("exec_tilesize", C.c_int32), ("prefetch_tasks", C.c_int32), ("block_budget_bytes", C.c_int64),
                ("watchdog_ms", C.c_int32), ("check_symmetry", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("tasks_run", C.c_int64), ("kernel_launches", C.c_int64), ("counted_flops", C.c_double),
                ("seconds_total", C.c_double), ("seconds_kernel", C.c_double),
                ("seconds_staging", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("blocks_fetched", C.c_int64), ("evaluated_flops", C.c_double), ("blocks_evicted", C.c_int64),
                ("seconds_fetch", C.c_double), ("seconds_host_wait", C.c_double), ("executed_flops", C.c_double),
                ("blocks_from_peers", C.c_int64), ("peer_bytes", C.c_int64)]


class MemoryEstimate(C.Structure):
    _fields_ = [("exec_max_hole_tile", C.c_int64), ("exec_max_particle_tile", C.c_int64), ("panel_bytes", C.c_int64),
                ("s1_bytes", C.c_int64), ("task_block_bytes", C.c_int64), ("tensor_bytes", C.c_int64 * 5),
                ("minimum_bytes", C.c_int64)]


FETCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _u32p, _dp, C.c_size_t)

SIGNATURES = {
    "ccsdt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "ccsdt_destroy": (C.c_int, [C.c_void_p]),
    "ccsdt_last_error": (C.c_char_p, [C.c_void_p]),
    "ccsdt_default_options": (C.c_int, [C.POINTER(Options)]),
    "ccsdt_set_options": (C.c_int, [C.c_void_p, C.POINTER(Options)]),
    "ccsdt_tiles": (C.c_int, [C.c_int64] * 5 + [_i64p, _i32p, _i32p, C.c_int]),
    "ccsdt_enumerate": (C.c_int64, [C.c_int, C.c_int, _i32p, C.c_int, _i64p, _dp, C.c_int64, _i64p]),
    "ccsdt_task_terms": (C.c_int, [C.c_int, C.c_int, _i32p, _i64p, C.c_int, _i64p, _u8p, _u8p, _u8p]),
    "ccsdt_count_ops": (C.c_int, [C.c_int, C.c_int, _i32p, _i64p, C.c_int, C.POINTER(C.c_longdouble)]),
    "ccsdt_partition": (C.c_int, [C.c_int, C.c_int, _i32p, _i64p, C.c_int, C.c_int, _i32p, C.c_int64]),
    "ccsdt_check_memory": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ccsdt_estimate_memory": (C.c_int, [C.c_int] * 4 + [_i64p, _i32p, C.c_int, C.POINTER(MemoryEstimate)]),
    "ccsdt_box_weight": (C.c_int, [C.c_int, _i32p]),
    "ccsdt_set_space": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [_i64p, _i32p, _dp, C.c_int]),
    "ccsdt_put_dense": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "ccsdt_put_dense_async": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "ccsdt_put_block": (C.c_int, [C.c_void_p, C.c_int, _u32p, _dp]),
    "ccsdt_set_fetch": (C.c_int, [C.c_void_p, FETCH_FN, C.c_void_p]),
    "ccsdt_set_synthetic": (C.c_int, [C.c_void_p, C.c_uint64]),
    "ccsdt_clear_blocks": (C.c_int, [C.c_void_p]),
    "ccsdt_release_cached": (C.c_int, []),
    "ccsdt_exec_tiles": (C.c_int, [C.c_void_p, _i64p, _i32p, _i32p, C.c_int]),
    "ccsdt_num_tasks": (C.c_int64, [C.c_void_p]),
    "ccsdt_make_exec_tiles": (C.c_int, [C.c_int] * 4 + [_i64p, _i32p, C.c_int, _i64p, _i32p, _i32p, C.c_int]),
    "ccsdt_split_request": (C.c_int64, [C.c_int] * 4 + [_i64p, _i32p, _i64p, _i32p, C.c_int, _u32p, _i64p, C.c_int64]),
    "ccsdt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ccsdt_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "ccsdt_comm_allreduce": (C.c_int, [C.c_void_p, _dp]),
    "ccsdt_comm_destroy": (C.c_int, [C.c_void_p]),
    "ccsdt_share_attach": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "ccsdt_share_detach": (C.c_int, [C.c_void_p]),
    "ccsdt_task_counter_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_i64p)]),
    "ccsdt_task_counter_close": (C.c_int, [_i64p, C.c_char_p, C.c_int]),
    "ccsdt_put_cholesky": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "ccsdt_set_task_counter": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ccsdt_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _dp, _dp, C.POINTER(Stats)]),
    "ccsdt_run_tasks": (C.c_int, [C.c_void_p, _i64p, C.c_int64, _dp, _dp, C.POINTER(Stats)]),
    "ccsdt_probe_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "ccsdt_probe_mainloop": (C.c_int, [C.c_int] * 6 + [_dp]),
    "ccsdt_probe_dmma_layout": (C.c_int, [C.c_int, _dp, _dp, _dp]),
    "ccsdt_probe_tma_swizzle": (C.c_int, [C.c_int, _dp, C.c_int]),
    "ccsdt_synth_block": (C.c_int, [C.c_int, C.c_uint64, C.c_int] + [C.c_int] * 4 + [_i64p, _i64p, _dp]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `make -C exachem_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
