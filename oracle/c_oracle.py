"""ctypes view of oracle/_build/libmvd_oracle.so (oracle/mvd_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product never touches it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmvd_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "mvd_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.mvdo_philox.argtypes = [vp, vp, vp]
        L.mvdo_trial_words.argtypes = [u64, u32, u64, u32, i32, u32, vp, vp]
        L.mvdo_encoder_branch.argtypes = [i32, i32, vp, i32, i32, vp, vp]
        L.mvdo_build_trellis.argtypes = [vp, i32, i32, vp, vp]
        L.mvdo_metric_step.argtypes = [vp, vp, i32, vp, i32, vp]
        L.mvdo_table_create.argtypes = [vp, u32, i32]
        L.mvdo_table_create.restype = vp
        L.mvdo_table_destroy.argtypes = [vp]
        L.mvdo_table_lookup.argtypes = [vp, vp]
        L.mvdo_enumerate.argtypes = [vp, i32, i32, u32, vp, vp, vp]
        L.mvdo_simulate.argtypes = [vp, vp, i32, i32, u32, vp, vp, vp, vp, vp, vp]
        L.mvdo_count_transitions.argtypes = [vp, vp, u32, u32, i32, u32, vp, vp]
        L.mvdo_log_prob.argtypes = [vp, vp, u32, i32, vp]
        L.mvdo_log_prob.restype = C.c_double
        L.mvdo_run_trials.argtypes = [vp, vp, i32, i32, u32, u32, u64, u32, u64, u64, vp, vp, vp, i32, vp]
        L.mvdo_run_trials.restype = C.c_int64
        L.mvdo_learn_chain.argtypes = [vp, vp, i32, i32, u32, u32, u32, u64, u32, u64, vp, vp, vp]
        L.mvdo_trial_words_k.argtypes = [u64, u32, u64, u32, i32, i32, u32, vp, vp]
        L.mvdo_metric_step_tab.argtypes = [vp, vp, i32, i32, vp, i32, vp]
        L.mvdo_simulate_tab.argtypes = [vp, vp, vp, vp, i32, i32, i32, u32, vp, vp, vp, vp, vp, vp]
        L.mvdo_run_trials_tab.argtypes = [vp, vp, vp, vp, i32, i32, i32, u32, u32, u64, u32, u64, u64, vp, vp, vp, i32, vp]
        L.mvdo_run_trials_tab.restype = C.c_int64
        L.mvdo_learn_chain_tab.argtypes = [vp, vp, vp, vp, i32, i32, i32, u32, u32, u32, u64, u32, u64, vp, vp]
        _lib = L
    return _lib


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def philox(ctr, key):
    c, k, o = _u32(ctr), _u32(key), np.zeros(4, dtype=np.uint32)
    lib().mvdo_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return [int(v) for v in o]


def trial_words(seed, stream, trial, N, n, T):
    nblk = (N + 31) // 32
    U = np.zeros(nblk, dtype=np.uint32)
    E = np.zeros((n, nblk), dtype=np.uint32)
    lib().mvdo_trial_words(seed, stream, trial, N, n, T, U.ctypes.data, E.ctypes.data)
    return U, E


def build_trellis(taps, n, m):
    t = _u32(taps)
    prev = np.zeros((1 << m, 2), dtype=np.uint8)
    blab = np.zeros((1 << m, 2), dtype=np.uint8)
    lib().mvdo_build_trellis(t.ctypes.data, n, m, prev.ctypes.data, blab.ctypes.data)
    return prev, blab


def metric_step(prev, blab, m, D, r):
    d = np.ascontiguousarray(D, dtype=np.int32)
    o = np.zeros_like(d)
    lib().mvdo_metric_step(prev.ctypes.data, blab.ctypes.data, m, d.ctypes.data, int(r), o.ctypes.data)
    return o


def enumerate_states(taps, n, m, max_states=1 << 22):
    t = _u32(taps)
    met = np.zeros((max_states, 1 << m), dtype=np.uint8)
    nxt = np.zeros((max_states, 1 << n), dtype=np.uint32)
    S = C.c_uint32()
    rc = lib().mvdo_enumerate(t.ctypes.data, n, m, max_states, met.ctypes.data, nxt.ctypes.data, C.byref(S))
    if rc != 0:
        raise MemoryError("oracle enumeration exceeded max_states")
    return met[:S.value].copy(), nxt[:S.value].copy()


class Table:
    def __init__(self, metrics, m):
        self.metrics = np.ascontiguousarray(metrics, dtype=np.uint8)
        self.m = m
        self.S = self.metrics.shape[0]
        self.h = lib().mvdo_table_create(self.metrics.ctypes.data, self.S, m)

    def __del__(self):
        if getattr(self, "h", None):
            lib().mvdo_table_destroy(self.h)
            self.h = None


def simulate(dec_taps, enc_taps, n, m, N, U, E, table, want_metrics=False):
    dt, et = _u32(dec_taps), _u32(enc_taps)
    U, E = _u32(U), _u32(E)
    idx = np.zeros(N + 1, dtype=np.int32)
    rseq = np.zeros(N + 1, dtype=np.uint8)
    met = np.zeros((N + 1, 1 << m), dtype=np.uint8) if want_metrics else None
    rc = lib().mvdo_simulate(dt.ctypes.data, et.ctypes.data, n, m, N, U.ctypes.data, E.ctypes.data, table.h,
                             idx.ctypes.data, rseq.ctypes.data, met.ctypes.data if want_metrics else None)
    if rc != 0:
        raise KeyError(f"metric vector at step {-rc - 1} not in the state table")
    return idx, rseq[:N], met


def count_transitions(idx, rseq, N, burn, n, S, dense=False):
    edge = np.zeros((S, 1 << n), dtype=np.uint64)
    dn = np.zeros((S, S), dtype=np.float64) if dense else None
    r = np.ascontiguousarray(np.concatenate([rseq, [0]]), dtype=np.uint8)
    i = np.ascontiguousarray(idx, dtype=np.int32)
    lib().mvdo_count_transitions(i.ctypes.data, r.ctypes.data, N, burn, n, S, edge.ctypes.data,
                                 dn.ctypes.data if dense else None)
    return edge, dn


def log_prob(idx, rseq, N, n, Tedge):
    i = np.ascontiguousarray(idx, dtype=np.int32)
    r = np.ascontiguousarray(np.concatenate([rseq, [0]]), dtype=np.uint8)
    t = np.ascontiguousarray(Tedge, dtype=np.float64)
    return float(lib().mvdo_log_prob(i.ctypes.data, r.ctypes.data, N, n, t.ctypes.data))


def run_trials(dec_taps, enc_taps, n, m, N, T, seed, stream, trial_begin, trial_end, table, P1edge, Tref_edge,
               decide, want_logp=False):
    dt, et = _u32(dec_taps), _u32(enc_taps)
    p1 = np.ascontiguousarray(P1edge, dtype=np.float64)
    tr = np.ascontiguousarray(Tref_edge, dtype=np.float64)
    lp = np.zeros((trial_end - trial_begin, 2), dtype=np.float64) if want_logp else None
    s = lib().mvdo_run_trials(dt.ctypes.data, et.ctypes.data, n, m, N, T, seed, stream, trial_begin, trial_end,
                              table.h, p1.ctypes.data, tr.ctypes.data, decide, lp.ctypes.data if want_logp else None)
    if s < 0:
        raise KeyError("a metric vector was not in the state table")
    return (int(s), lp) if want_logp else int(s)


def learn_chain(dec_taps, enc_taps, n, m, length, burn, T, seed, stream, trial, table, dense=False):
    dt, et = _u32(dec_taps), _u32(enc_taps)
    edge = np.zeros((table.S, 1 << n), dtype=np.uint64)
    dn = np.zeros((table.S, table.S), dtype=np.float64) if dense else None
    rc = lib().mvdo_learn_chain(dt.ctypes.data, et.ctypes.data, n, m, length, burn, T, seed, stream, trial,
                                table.h, edge.ctypes.data, dn.ctypes.data if dense else None)
    if rc != 0:
        raise KeyError("a metric vector was not in the state table")
    return edge, dn


def acs_hash(dec_taps, enc_taps, n, m, N, U, E):
    dt, et = _u32(dec_taps), _u32(enc_taps)
    U, E = _u32(U), _u32(E)
    h = C.c_uint64()
    fin = np.zeros(1 << m, dtype=np.uint8)
    L = lib()
    L.mvdo_acs_hash.argtypes = [C.c_void_p] * 2 + [C.c_int, C.c_int, C.c_uint32] + [C.c_void_p] * 4
    rc = L.mvdo_acs_hash(dt.ctypes.data, et.ctypes.data, n, m, N, U.ctypes.data, E.ctypes.data, C.byref(h), fin.ctypes.data)
    if rc != 0:
        raise OverflowError("relative metric exceeded 15")
    return int(h.value), fin


# ---- codes given as tables (k <= 3): prev / blab / enc_next / enc_out uint8 [2^m, 2^k] (mvd.codes.trellis_arrays / encoder_tables)
def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def trial_words_k(seed, stream, trial, N, k, n, T):
    nblk = (N + 31) // 32
    U = np.zeros((k, nblk), dtype=np.uint32)
    E = np.zeros((n, nblk), dtype=np.uint32)
    lib().mvdo_trial_words_k(seed, stream, trial, N, k, n, T, U.ctypes.data, E.ctypes.data)
    return U, E


def metric_step_tab(prev, blab, m, k, D, r):
    p, b = _u8(prev), _u8(blab)
    d = np.ascontiguousarray(D, dtype=np.int32)
    o = np.zeros(1 << m, dtype=np.int32)
    lib().mvdo_metric_step_tab(p.ctypes.data, b.ctypes.data, m, k, d.ctypes.data, int(r), o.ctypes.data)
    return tuple(int(v) for v in o)


def simulate_tab(prev, blab, enc_next, enc_out, k, n, m, N, U, E, table, want_metrics=False):
    p, b, en, eo = _u8(prev), _u8(blab), _u8(enc_next), _u8(enc_out)
    U, E = _u32(U), _u32(E)
    idx = np.zeros(N + 1, dtype=np.int32)
    rseq = np.zeros(N + 1, dtype=np.uint8)
    met = np.zeros((N + 1, 1 << m), dtype=np.uint8) if want_metrics else None
    rc = lib().mvdo_simulate_tab(p.ctypes.data, b.ctypes.data, en.ctypes.data, eo.ctypes.data, k, n, m, N, U.ctypes.data,
                                 E.ctypes.data, table.h, idx.ctypes.data, rseq.ctypes.data,
                                 met.ctypes.data if want_metrics else None)
    if rc != 0:
        raise KeyError(f"metric vector at step {-rc - 1} not in the state table")
    return idx, rseq[:N], met


def run_trials_tab(prev, blab, enc_next, enc_out, k, n, m, N, T, seed, stream, trial_begin, trial_end, table, P1edge, Tref_edge,
                   decide, want_logp=False):
    p, b, en, eo = _u8(prev), _u8(blab), _u8(enc_next), _u8(enc_out)
    p1 = np.ascontiguousarray(P1edge, dtype=np.float64)
    tr = np.ascontiguousarray(Tref_edge, dtype=np.float64)
    lp = np.zeros((trial_end - trial_begin, 2), dtype=np.float64) if want_logp else None
    s = lib().mvdo_run_trials_tab(p.ctypes.data, b.ctypes.data, en.ctypes.data, eo.ctypes.data, k, n, m, N, T, seed, stream,
                                  trial_begin, trial_end, table.h, p1.ctypes.data, tr.ctypes.data, decide,
                                  lp.ctypes.data if want_logp else None)
    if s < 0:
        raise KeyError("a metric vector was not in the state table")
    return (int(s), lp) if want_logp else int(s)


def learn_chain_tab(prev, blab, enc_next, enc_out, k, n, m, length, burn, T, seed, stream, trial, table):
    p, b, en, eo = _u8(prev), _u8(blab), _u8(enc_next), _u8(enc_out)
    edge = np.zeros((table.S, 1 << n), dtype=np.uint64)
    rc = lib().mvdo_learn_chain_tab(p.ctypes.data, b.ctypes.data, en.ctypes.data, eo.ctypes.data, k, n, m, length, burn, T, seed,
                                    stream, trial, table.h, edge.ctypes.data)
    if rc != 0:
        raise KeyError("a metric vector was not in the state table")
    return edge
