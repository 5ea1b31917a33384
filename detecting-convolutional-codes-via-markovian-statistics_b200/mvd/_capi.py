"""ctypes binding of libmvd.so (include/mvd.h).  No compute happens in Python.

The library is built in-tree (``build.py`` / ``__graft_entry__.build()``) next to this package as
``libmvd.so``.  If it is missing, or no CUDA device is usable, everything here raises -- there is
no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_N = 4
MAX_M = 6
MAX_K = 3
ABI_VERSION = 3          # must equal MVD_ABI_VERSION of include/mvd.h: the struct layouts below mirror that header

SRC_PHILOX, SRC_BITSTREAM = 0, 1
ENGINE_AUTO, ENGINE_ACS, ENGINE_FSM = 0, 1, 2
ENGINES = {"auto": ENGINE_AUTO, "acs": ENGINE_ACS, "fsm": ENGINE_FSM}

E_UNKNOWN_STATE = -6
OPT_FORCE_GENERIC = 1
OPT_NO_PAIR = 2
OPT_LEARN_WARM = 3
OPT_NO_FSM1 = 4
OPT_SPLIT = 5
OPT_NO_ANTIPODAL = 6
OPT_ASYNC_DETECT = 7
OPT_SPLIT_SEQUENTIAL = 8
OPT_SPLIT_CHUNK = 9

LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libmvd.so")

EXPORTS = (
    "mvd_abi_version", "mvd_create", "mvd_destroy", "mvd_last_error", "mvd_set_stream", "mvd_synchronize",
    "mvd_set_code", "mvd_set_states", "mvd_enumerate_states", "mvd_get_states", "mvd_set_loglik",
    "mvd_learn_counts", "mvd_detect", "mvd_trace", "mvd_acs_hash", "mvd_last_kernel_ms", "mvd_launch_count",
    "mvd_int_peak", "mvd_device_info", "mvd_set_option", "mvd_last_kernel_kind", "mvd_learn_stats",
    "mvd_enumerate_states_gpu", "mvd_bfs_levels", "mvd_chernoff_rho", "mvd_chernoff_rho_dense", "mvd_parity_detect", "mvd_host_log_table", "mvd_acs_final",
    "mvd_copy_stats", "mvd_async_stats", "mvd_host_p1_edge_tables", "mvd_split_stats", "mvd_set_code_tables", "mvd_set_encoders",
)


class MvdError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmvd error {code}: {message}")
        self.code = code


class UnknownStateError(MvdError, KeyError):
    """The device met a metric vector that is not in the state table (reference: KeyError at
    Pd_plotter.py:112)."""


class Src(C.Structure):
    _fields_ = [("mode", C.c_int32), ("bits_on_device", C.c_int32), ("seed", C.c_uint64),
                ("bits", C.c_void_p), ("bits_words", C.c_uint64)]


class Segment(C.Structure):
    _fields_ = [("N", C.c_uint32), ("threshold", C.c_uint32), ("stream", C.c_uint32), ("table", C.c_uint32),
                ("enc_taps", C.c_uint32 * MAX_N), ("decide", C.c_uint32), ("random_input", C.c_uint32),
                ("trial_begin", C.c_uint64), ("trial_end", C.c_uint64), ("bits_offset", C.c_uint64)]


class ParitySegment(C.Structure):
    _fields_ = [("N", C.c_uint32), ("m", C.c_uint32), ("n", C.c_uint32), ("threshold", C.c_uint32), ("stream", C.c_uint32),
                ("decide", C.c_uint32), ("enc_taps", C.c_uint32 * MAX_N), ("tmpl", C.c_uint32 * MAX_N), ("gamma", C.c_double),
                ("trial_begin", C.c_uint64), ("trial_end", C.c_uint64), ("bits_offset", C.c_uint64)]


class BfsStats(C.Structure):
    _fields_ = [("S", C.c_uint32), ("frontier", C.c_uint32), ("iterations", C.c_uint32), ("launches", C.c_uint32),
                ("candidates", C.c_uint64), ("closed", C.c_int32), ("max_metric", C.c_int32), ("ms", C.c_float),
                ("reserved", C.c_uint32)]


BFS_INSTALL, BFS_COUNT_ONLY = 1, 2

_lib = None


def load():
    """Load libmvd.so and declare prototypes (idempotent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build the CUDA library first "
                          f"(python __graft_entry__.py build).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    P = C.POINTER
    lib.mvd_abi_version.restype = i32
    if lib.mvd_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {lib.mvd_abi_version()}, this binding expects {ABI_VERSION}: "
                          f"rebuild it (python __graft_entry__.py)")
    lib.mvd_create.argtypes = [P(vp), i32]
    lib.mvd_destroy.argtypes = [vp]
    lib.mvd_last_error.argtypes = [vp]
    lib.mvd_last_error.restype = C.c_char_p
    lib.mvd_set_stream.argtypes = [vp, vp]
    lib.mvd_synchronize.argtypes = [vp]
    lib.mvd_set_code.argtypes = [vp, i32, i32, i32, P(u32)]
    lib.mvd_set_states.argtypes = [vp, u32, vp, vp]
    lib.mvd_enumerate_states.argtypes = [vp, u32, P(u32)]
    lib.mvd_enumerate_states_gpu.argtypes = [vp, u32, u32, u32, P(BfsStats)]
    lib.mvd_bfs_levels.argtypes = [vp, vp, u32, P(u32)]
    lib.mvd_host_log_table.argtypes = [vp, vp, u64]
    lib.mvd_parity_detect.argtypes = [vp, P(Src), P(ParitySegment), u32, vp, vp]
    lib.mvd_chernoff_rho_dense.argtypes = [vp, u32, u32, vp, vp, vp, u32, C.c_double, u32, vp, vp]
    lib.mvd_chernoff_rho.argtypes = [vp, u32, u32, vp, vp, vp, vp, vp, vp, u32, C.c_double, u32, vp, vp]
    lib.mvd_get_states.argtypes = [vp, vp, vp]
    lib.mvd_set_loglik.argtypes = [vp, u32, vp, vp]
    lib.mvd_learn_counts.argtypes = [vp, P(Src), P(Segment), u32, u32, i32, vp]
    lib.mvd_detect.argtypes = [vp, P(Src), P(Segment), u32, i32, vp, vp, vp]
    lib.mvd_trace.argtypes = [vp, P(Src), P(Segment), i32, vp, vp]
    lib.mvd_acs_hash.argtypes = [vp, P(Src), P(Segment), vp, vp]
    lib.mvd_acs_final.argtypes = [vp, P(Src), P(Segment), vp]
    lib.mvd_last_kernel_ms.argtypes = [vp, P(C.c_float)]
    lib.mvd_launch_count.argtypes = [vp, P(u64)]
    lib.mvd_copy_stats.argtypes = [vp, P(u64), P(u64)]
    lib.mvd_async_stats.argtypes = [vp, P(C.c_double), P(u64)]
    lib.mvd_host_p1_edge_tables.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, vp]
    lib.mvd_int_peak.argtypes = [vp, P(C.c_double), P(C.c_double)]
    lib.mvd_set_option.argtypes = [vp, i32, C.c_int64]
    lib.mvd_last_kernel_kind.argtypes = [vp, P(i32)]
    lib.mvd_learn_stats.argtypes = [vp, P(u32)]
    lib.mvd_split_stats.argtypes = [vp, P(u64), P(u64)]
    lib.mvd_set_code_tables.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.mvd_set_encoders.argtypes = [vp, C.c_uint32, vp, vp]
    lib.mvd_device_info.argtypes = [vp, P(i32), P(i32), P(u64), C.c_char_p, i32]
    for name in EXPORTS:
        getattr(lib, name)            # AttributeError here = header / library mismatch
        if name not in ("mvd_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(lib, ctx, rc: int):
    if rc == 0:
        return
    msg = lib.mvd_last_error(ctx)
    text = msg.decode(errors="replace") if msg else ""
    if rc == E_UNKNOWN_STATE:
        raise UnknownStateError(rc, text)
    raise MvdError(rc, text)
