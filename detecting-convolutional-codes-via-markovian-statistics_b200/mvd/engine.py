"""High-level host driver over libmvd.so: one :class:`Detector` per (decoder code, GPU).

Call order mirrors the reference's experiment (Pd_plotter.py:176-235):
``Detector(gen1, ...)``  -> trellis + Markov state table (host) -> device tables;
``learn_counts``         -> Pd_plotter.py:149-163 on the GPU;
``set_models``           -> Pd_plotter.py:166-167 (host float64) + log tables to the GPU;
``detect``               -> Pd_plotter.py:210-223 on the GPU, one launch for a whole sweep.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _capi, bitsource, codes


@dataclass
class Seg:
    """One hypothesis of one sweep point (or one learning chain); see ``mvd_segment``."""
    N: int
    threshold: int = 0
    stream: int = 0
    table: int = 0
    enc_taps: Sequence[int] = ()
    decide: int = 0
    random_input: bool = True
    trial_begin: int = 0
    trial_end: int = 1
    bits_offset: int = 0

    @property
    def ntrials(self) -> int:
        return self.trial_end - self.trial_begin


# numpy view of ``mvd_segment`` (include/mvd.h; layout checked against the header in tests/test_capi_exports.py):
# sweeps of hundreds of segments are filled column-wise instead of one ctypes attribute at a time
SEG_DTYPE = np.dtype({"names": ["N", "threshold", "stream", "table", "enc_taps", "decide", "random_input", "trial_begin",
                                "trial_end", "bits_offset"],
                      "formats": ["<u4", "<u4", "<u4", "<u4", ("<u4", (_capi.MAX_N,)), "<u4", "<u4", "<u8", "<u8", "<u8"],
                      "offsets": [0, 4, 8, 12, 16, 32, 36, 40, 48, 56], "itemsize": 64})
assert SEG_DTYPE.itemsize == C.sizeof(_capi.Segment)


def fresh_seed() -> int:
    """seed=None: a fresh 64-bit Philox key from numpy's global generator (the reference idiom: no reseed,
    alpha_exponent.py:105-106) -- never a silent fixed stream."""
    return int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))


def _log_table(values: np.ndarray) -> np.ndarray:
    """Elementwise ``math.log(max(v, 1e-300))`` (Pd_plotter.py:114-115) with libm's log -- the
    same function the reference calls per step, so device sums add bit-identical terms."""
    flat = np.ascontiguousarray(values, dtype=np.float64).ravel()
    out = np.empty_like(flat)               # the same libm call, made from C (mvd_host_log_table; bit-equal, tests)
    _capi.check(_capi.load(), None, _capi.load().mvd_host_log_table(flat.ctypes.data, out.ctypes.data, flat.size))
    return out.reshape(np.shape(values))


class Detector:
    def __init__(self, gen1, k: int, n: int, m: int, device: int = 0, table: Optional[codes.StateTable] = None,
                 enumerate_with: str = "python", max_states: int = 1 << 22):
        self.lib = _capi.load()
        self.gen1 = codes.freeze_generator(gen1)
        self.k, self.n, self.m = int(k), int(n), int(m)
        self.R = 1 << self.n
        self.ctx = C.c_void_p()
        rc = self.lib.mvd_create(C.byref(self.ctx), int(device))
        if rc != 0:
            _capi.check(self.lib, None, rc)
        self.device = int(device)
        self.table_code = self.k != 1
        if self.table_code:
            # k > 1 inputs per step: the reference's trellis / branch functions are generic in k (viterbi_markov.py:82-132);
            # their results go to the device as tables (mvd_set_code_tables), hypotheses are encoder-table indices
            if self.k > _capi.MAX_K:
                self.close()
                raise _capi.MvdError(-3, f"the device path supports k <= {_capi.MAX_K} inputs per step")
            prev, blab = codes.trellis_arrays(self.gen1, self.m, self.k, self.n)
            self._ck(self.lib.mvd_set_code_tables(self.ctx, self.k, self.n, self.m, np.ascontiguousarray(prev, dtype=np.uint8).ctypes.data,
                                                  np.ascontiguousarray(blab, dtype=np.uint8).ctypes.data))
            self._enc_index, self._enc_next, self._enc_out = {}, [], []
            self.dec_taps = self.taps_of(self.gen1)               # encoder 0 = the decoder's own code
            if table is None and enumerate_with != "python":
                enumerate_with = "python"                         # the C++ / GPU enumerations need the tap-mask form
        else:
            self.dec_taps = codes.tap_masks(self.gen1, self.m, self.k)
            taps = (C.c_uint32 * len(self.dec_taps))(*self.dec_taps)
            self._ck(self.lib.mvd_set_code(self.ctx, self.k, self.n, self.m, taps))
        if table is not None:
            self.table = table
            self._upload_states()
        elif enumerate_with == "gpu":
            self.bfs_stats = self._enumerate_gpu(int(max_states), install=True)
            self.table = self._fetch_states(self.bfs_stats["S"])
        elif enumerate_with == "lib":
            S = C.c_uint32()
            self._ck(self.lib.mvd_enumerate_states(self.ctx, int(max_states), C.byref(S)))
            met = np.empty((S.value, 1 << self.m), dtype=np.uint8)
            nxt = np.empty((S.value, self.R), dtype=np.uint32)
            self._ck(self.lib.mvd_get_states(self.ctx, met.ctypes.data, nxt.ctypes.data))
            mult = (nxt[:, :, None] == nxt[:, None, :]).sum(axis=2).astype(np.uint8)
            self.table = codes.StateTable(self.k, self.n, self.m, met, nxt, mult)
        else:
            self.table = codes.enumerate_states(self.gen1, self.m, self.k, self.n, max_states=max_states)
            self._upload_states()
        self.ntables = 0

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc: int):
        _capi.check(self.lib, self.ctx, rc)

    def _upload_states(self):
        met = np.ascontiguousarray(self.table.metrics, dtype=np.uint8)
        nxt = np.ascontiguousarray(self.table.nxt, dtype=np.uint32)
        self._ck(self.lib.mvd_set_states(self.ctx, self.table.S, met.ctypes.data, nxt.ctypes.data))

    def _fetch_states(self, S: int) -> codes.StateTable:
        met = np.empty((S, 1 << self.m), dtype=np.uint8)
        nxt = np.empty((S, self.R), dtype=np.uint32)
        self._ck(self.lib.mvd_get_states(self.ctx, met.ctypes.data, nxt.ctypes.data))
        if S * self.R * self.R <= 1 << 26:
            mult = (nxt[:, :, None] == nxt[:, None, :]).sum(axis=2).astype(np.uint8)
        else:
            mult = np.zeros_like(nxt, dtype=np.uint8)
            for r in range(self.R):
                mult += (nxt == nxt[:, r:r + 1]).astype(np.uint8)
        return codes.StateTable(self.k, self.n, self.m, met, nxt, mult)

    def _enumerate_gpu(self, max_states: int, install: bool = True, count_only: bool = False, chunk_parents: int = 0,
                       allow_partial: bool = False) -> dict:
        """enumerate_markov_states_allzero (viterbi_markov.py:166-195) on the GPU; see mvd_enumerate_states_gpu."""
        st = _capi.BfsStats()
        flags = (_capi.BFS_INSTALL if install else 0) | (_capi.BFS_COUNT_ONLY if count_only else 0)
        rc = self.lib.mvd_enumerate_states_gpu(self.ctx, int(max_states), flags, int(chunk_parents), C.byref(st))
        if rc != 0 and not (allow_partial and rc == -5 and st.S > 0):
            self._ck(rc)
        nl = C.c_uint32()
        self.lib.mvd_bfs_levels(self.ctx, None, 0, C.byref(nl))
        lev = np.zeros(nl.value, dtype=np.uint32)
        self.lib.mvd_bfs_levels(self.ctx, lev.ctypes.data, nl.value, C.byref(nl))
        return dict(S=int(st.S), frontier=int(st.frontier), iterations=int(st.iterations), launches=int(st.launches),
                    candidates=int(st.candidates), closed=bool(st.closed), max_metric=int(st.max_metric), ms=float(st.ms),
                    levels=lev.tolist())

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.mvd_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def S(self) -> int:
        return self.table.S

    def taps_of(self, gen) -> List[int]:
        """What goes into ``Seg.enc_taps`` for the encoder ``gen``: its tap masks (k = 1), or -- for a code given as tables --
        ``[index]`` of its (enc_next, enc_out) tables, registered with ``mvd_set_encoders`` on first use."""
        gen = codes.freeze_generator(gen)
        if not self.table_code:
            return codes.tap_masks(gen, self.m, self.k)
        if gen not in self._enc_index:
            enc_out, enc_next = codes.encoder_tables(gen, self.m, self.k)
            self._enc_index[gen] = len(self._enc_next)
            self._enc_next.append(np.asarray(enc_next, dtype=np.uint8))
            self._enc_out.append(np.asarray(enc_out, dtype=np.uint8))
            nx, ou = np.ascontiguousarray(np.stack(self._enc_next)), np.ascontiguousarray(np.stack(self._enc_out))
            self._ck(self.lib.mvd_set_encoders(self.ctx, len(self._enc_next), nx.ctypes.data, ou.ctypes.data))
        return [self._enc_index[gen]]

    def segment_array(self, N, threshold, stream, table, enc_taps, decide, trial_begin, trial_end, random_input=1,
                      bits_offset=0) -> np.ndarray:
        """``mvd_segment`` records as one numpy structured array; every argument broadcasts over the segments
        (``enc_taps``: [nsegs, n] or [n])."""
        n = len(np.atleast_1d(np.asarray(N)))
        a = np.zeros(n, dtype=SEG_DTYPE)
        a["N"], a["threshold"], a["table"], a["decide"] = N, threshold, table, decide
        a["stream"] = np.asarray(stream, dtype=np.uint64) & 0xFFFFFFFF
        taps = np.atleast_2d(np.asarray(enc_taps, dtype=np.uint32))
        a["enc_taps"][:, :taps.shape[1]] = taps
        a["random_input"], a["trial_begin"], a["trial_end"], a["bits_offset"] = random_input, trial_begin, trial_end, bits_offset
        return a

    def _segments(self, segs):
        if isinstance(segs, np.ndarray):                         # prebuilt records (segment_array)
            if segs.dtype != SEG_DTYPE or not segs.flags.c_contiguous:
                raise TypeError("segment arrays must be contiguous with dtype engine.SEG_DTYPE")
            return C.cast(segs.ctypes.data, C.POINTER(_capi.Segment))
        arr = (_capi.Segment * len(segs))()
        for a, s in zip(arr, segs):
            a.N, a.threshold, a.stream, a.table = int(s.N), int(s.threshold), int(s.stream) & 0xFFFFFFFF, int(s.table)
            taps = list(s.enc_taps) if len(s.enc_taps) else list(self.dec_taps)
            for j in range(_capi.MAX_N):
                a.enc_taps[j] = int(taps[j]) if j < len(taps) else 0
            a.decide = int(s.decide)
            a.random_input = 1 if s.random_input else 0
            a.trial_begin, a.trial_end, a.bits_offset = int(s.trial_begin), int(s.trial_end), int(s.bits_offset)
        return arr

    @staticmethod
    def _src(seed: Optional[int], bits, bits_device_ptr: Optional[int] = None, bits_words: int = 0):
        src = _capi.Src()
        if bits is None and bits_device_ptr is None:
            src.mode = _capi.SRC_PHILOX
            src.seed = (fresh_seed() if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF
            return src, None
        src.mode = _capi.SRC_BITSTREAM
        src.seed = 0
        if bits_device_ptr is not None:
            src.bits_on_device = 1
            src.bits = int(bits_device_ptr)
            src.bits_words = int(bits_words)
            return src, None
        keep = np.ascontiguousarray(bits, dtype=np.uint32)
        src.bits_on_device = 0
        src.bits = keep.ctypes.data
        src.bits_words = keep.size // 4
        return src, keep

    # ------------------------------------------------------------------ hot path
    def learn_counts(self, segs: Sequence[Seg], burn: int, seed: Optional[int] = None, bits=None,
                     engine: str = "auto", bits_device_ptr=None, bits_words=0) -> np.ndarray:
        """Edge counts uint64 [nsegs, S, R] of the steps t >= burn (Pd_plotter.py:158-163)."""
        src, keep = self._src(seed, bits, bits_device_ptr, bits_words)
        out = np.zeros((len(segs), self.S, self.R), dtype=np.uint64)
        self._ck(self.lib.mvd_learn_counts(self.ctx, C.byref(src), self._segments(segs), len(segs), int(burn),
                                           _capi.ENGINES[engine], out.ctypes.data))
        del keep
        return out

    def set_models(self, p1_edge_tables: Iterable[np.ndarray], tref_edge: Optional[np.ndarray] = None):
        """Upload ``log P1`` (one table per distinct p) and ``log T(1/2)``, edge-indexed [S, R]."""
        if isinstance(p1_edge_tables, np.ndarray) and p1_edge_tables.ndim == 3:
            tabs = np.ascontiguousarray(p1_edge_tables, dtype=np.float64).reshape(-1, self.S, self.R)    # one threaded log pass
        else:
            tabs = np.stack([np.asarray(t, dtype=np.float64).reshape(self.S, self.R) for t in p1_edge_tables])
        if tref_edge is None:
            tref_edge = codes.tref_half_table(self.table)
        self.logP1 = _log_table(tabs)
        self.logTref = np.ascontiguousarray(_log_table(np.asarray(tref_edge, dtype=np.float64).reshape(self.S, self.R)))
        self._ck(self.lib.mvd_set_loglik(self.ctx, len(tabs), self.logP1.ctypes.data, self.logTref.ctypes.data))
        self.ntables = len(tabs)

    def set_loglik(self, logP1: np.ndarray, logTref: np.ndarray):
        """Upload the log tables themselves (``mvd_set_loglik``): ``logP1`` [ntables, S, R], ``logTref`` [S, R] -- what
        :meth:`set_models` derives from probabilities with the reference's ``math.log(max(P, 1e-300))`` (Pd_plotter.py:114)."""
        self.logP1 = np.ascontiguousarray(logP1, dtype=np.float64).reshape(-1, self.S, self.R)
        self.logTref = np.ascontiguousarray(logTref, dtype=np.float64).reshape(self.S, self.R)
        self._ck(self.lib.mvd_set_loglik(self.ctx, len(self.logP1), self.logP1.ctypes.data, self.logTref.ctypes.data))
        self.ntables = len(self.logP1)

    def detect(self, segs, seed: Optional[int] = None, bits=None, engine: str = "auto",
               want_logp: bool = False, d_tallies_ptr: Optional[int] = None, bits_device_ptr=None, bits_words=0,
               host_tallies: bool = True):
        """Successes per segment (uint64 [nsegs]) and optionally (logp1, logp_ref) per trial.  ``segs``: a list of
        :class:`Seg` or a :meth:`segment_array`.  ``d_tallies_ptr``: device uint64[nsegs] that receives this call's
        tallies (for a device-side allreduce); with ``host_tallies=False`` nothing but the error flag comes back to
        the host and ``None`` is returned in place of the tallies."""
        src, keep = self._src(seed, bits, bits_device_ptr, bits_words)
        if not host_tallies and not d_tallies_ptr:
            raise ValueError("host_tallies=False needs d_tallies_ptr")
        tallies = np.zeros(len(segs), dtype=np.uint64) if host_tallies else None
        logp = None
        if want_logp:
            ntr = int((segs["trial_end"] - segs["trial_begin"]).sum()) if isinstance(segs, np.ndarray) else sum(s.ntrials for s in segs)
            logp = np.zeros((ntr, 2), dtype=np.float64)
        self._ck(self.lib.mvd_detect(self.ctx, C.byref(src), self._segments(segs), len(segs), _capi.ENGINES[engine],
                                     tallies.ctypes.data if host_tallies else None, logp.ctypes.data if want_logp else None,
                                     C.c_void_p(d_tallies_ptr) if d_tallies_ptr else None))
        del keep
        return (tallies, logp) if want_logp else tallies

    def trace(self, seg: Seg, seed: Optional[int] = None, bits=None, engine: str = "acs", want_metrics: bool = True):
        """State indices uint32 [ntrials, N+1] and metric vectors uint8 [ntrials, N+1, 2^m]."""
        src, keep = self._src(seed, bits)
        idx = np.zeros((seg.ntrials, seg.N + 1), dtype=np.uint32)
        met = np.zeros((seg.ntrials, seg.N + 1, 1 << self.m), dtype=np.uint8) if want_metrics else None
        self._ck(self.lib.mvd_trace(self.ctx, C.byref(src), self._segments([seg]), _capi.ENGINES[engine],
                                    idx.ctypes.data, met.ctypes.data if want_metrics else None))
        del keep
        return idx, met

    def acs_hash(self, seg: Seg, seed: Optional[int] = None, bits=None, want_final: bool = True):
        src, keep = self._src(seed, bits)
        h = np.zeros(seg.ntrials, dtype=np.uint64)
        fin = np.zeros((seg.ntrials, 1 << self.m), dtype=np.uint8) if want_final else None
        self._ck(self.lib.mvd_acs_hash(self.ctx, C.byref(src), self._segments([seg]), h.ctypes.data,
                                       fin.ctypes.data if want_final else None))
        del keep
        return h, fin

    def acs_final(self, seg: Seg, seed: Optional[int] = None) -> np.ndarray:
        """Final metric vectors uint8 [ntrials, 2^m] of the Eq. 4-5 recursion, throughput form (``mvd_acs_final``)."""
        src, _ = self._src(seed, None)
        fin = np.zeros((seg.ntrials, 1 << self.m), dtype=np.uint8)
        self._ck(self.lib.mvd_acs_final(self.ctx, C.byref(src), self._segments([seg]), fin.ctypes.data))
        return fin

    # ------------------------------------------------------------------ error exponent (alpha_exponent.py)
    def chernoff_rho_edges(self, nxt, lp1, lp2, lb1, lb2, u_vals, tol: float = 1e-14, max_iter: int = 100000):
        """rho(M(u)) for every u (Eq. 7) from edge-form log tensors; see ``mvd_chernoff_rho``."""
        nxt = np.ascontiguousarray(nxt, dtype=np.uint32)
        K, R = nxt.shape
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (lp1, lp2, lb1, lb2, u_vals)]
        assert arrs[0].shape == (K, R) and arrs[1].shape == (K, R) and arrs[2].shape == (K,) and arrs[3].shape == (K,)
        nu = arrs[4].size
        rho = np.zeros(nu, dtype=np.float64)
        iters = np.zeros(nu, dtype=np.uint32)
        self._ck(self.lib.mvd_chernoff_rho(self.ctx, K, R, nxt.ctypes.data, *(a.ctypes.data for a in arrs), nu, float(tol),
                                           int(max_iter), rho.ctypes.data, iters.ctypes.data))
        return rho, iters

    def chernoff_rho_dense(self, logP1, logP2, u_vals, tol: float = 1e-14, max_iter: int = 100000):
        """rho(M(u)) for dense K x K x R log tensors; see ``mvd_chernoff_rho_dense``."""
        a = np.ascontiguousarray(logP1, dtype=np.float64)
        b = np.ascontiguousarray(logP2, dtype=np.float64)
        K, K2, R = a.shape
        assert K == K2 and b.shape == a.shape
        u = np.ascontiguousarray(u_vals, dtype=np.float64)
        rho = np.zeros(u.size, dtype=np.float64)
        iters = np.zeros(u.size, dtype=np.uint32)
        self._ck(self.lib.mvd_chernoff_rho_dense(self.ctx, K, R, a.ctypes.data, b.ctypes.data, u.ctypes.data, u.size,
                                                 float(tol), int(max_iter), rho.ctypes.data, iters.ctypes.data))
        return rho, iters

    # ------------------------------------------------------------------ introspection
    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.mvd_last_kernel_ms(self.ctx, C.byref(ms)))
        return float(ms.value)

    def copy_stats(self):
        """(host -> device bytes, device -> host bytes) this context has copied so far, counted at the copy call sites."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.mvd_copy_stats(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def launch_count(self) -> int:
        v = C.c_uint64()
        self._ck(self.lib.mvd_launch_count(self.ctx, C.byref(v)))
        return int(v.value)

    def split_sequential(self, on: bool = True):
        """Split path: add every log-likelihood term one by one in step order (``MVD_OPT_SPLIT_SEQUENTIAL``) instead of
        re-associating the float64 additions inside a binade -- identical results, the check of the re-association.  ``on=2``:
        re-association without the class counting of the second sum (both sums as recurrences)."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_SPLIT_SEQUENTIAL, int(on)))

    def split_chunk(self, steps: int = 0):
        """Steps per chunk of the split path (``MVD_OPT_SPLIT_CHUNK``): 0 = chosen per call, else 256, 512 or 1024."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_SPLIT_CHUNK, int(steps)))

    def split_stats(self):
        """(sub-chunks of the last split launch, sub-chunks whose terms were added one by one)."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.mvd_split_stats(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def async_detect(self, on: bool = True):
        """``detect(..., d_tallies_ptr=..., host_tallies=False)`` calls only queue their work (``MVD_OPT_ASYNC_DETECT``);
        :meth:`synchronize` waits for them.  Switching it off drains what is in flight."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_ASYNC_DETECT, 1 if on else 0))

    def synchronize(self):
        """Wait for everything queued on the context's stream; raises the KeyError analogue of any asynchronous launch."""
        self._ck(self.lib.mvd_synchronize(self.ctx))

    def async_stats(self):
        """(summed kernel ms, launches) of the asynchronous detection launches drained since the last call."""
        ms, n = C.c_double(), C.c_uint64()
        self._ck(self.lib.mvd_async_stats(self.ctx, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def set_stream(self, cuda_stream_ptr: int):
        """Run this context's work on the caller's CUDA stream (``mvd_set_stream``), e.g. a torch stream's ``cuda_stream``."""
        self._ck(self.lib.mvd_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    def force_generic(self, on: bool = True):
        """Route detection through the generic (checked) kernels instead of the fast ones."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_FORCE_GENERIC, 1 if on else 0))

    def no_pair(self, on=True):
        """Fast kernels with one trial per thread (True / 1), two per thread even for few trials (2),
        or the automatic choice (False / 0).  Only m = 2 has the choice."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_NO_PAIR, int(on)))

    def no_fsm1(self, on: bool = True):
        """NEXT-table walk with separate log-likelihood and NEXT tables instead of the one-load entry."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_NO_FSM1, 1 if on else 0))

    def no_antipodal(self, on: bool = True):
        """Two-trials-per-thread m = 2 kernel with the general branch-metric table even for codes whose
        generators all have their first and last tap set (the complement-label short cut is the default there)."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_NO_ANTIPODAL, 1 if on else 0))

    def split_trials(self, mode: int = 0):
        """Long trials split along the time axis (NEXT-table engine): 0 = automatic, 1 = whenever possible, 2 = never."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_SPLIT, int(mode)))

    def learn_warm(self, steps: int = 128):
        """Warm-up steps of the chunk-parallel learning chains (0 = speculate cold: every chunk is
        repaired by the fix-up pass; results are identical)."""
        self._ck(self.lib.mvd_set_option(self.ctx, _capi.OPT_LEARN_WARM, int(steps)))

    def learn_dirty_chunks(self) -> int:
        v = C.c_uint32()
        self._ck(self.lib.mvd_learn_stats(self.ctx, C.byref(v)))
        return int(v.value)

    def last_kernel_kind(self) -> int:
        """0 = generic kernel; else 1 + lookup kind (0 direct, 1 hash, 2 NEXT walk) + 16 * log2(row stride)
        (+ 256 for the two-trials-per-thread kernel)."""
        v = C.c_int()
        self._ck(self.lib.mvd_last_kernel_kind(self.ctx, C.byref(v)))
        return int(v.value)

    def int_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.mvd_int_peak(self.ctx, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def device_info(self) -> dict:
        sm, khz, smem = C.c_int(), C.c_int(), C.c_uint64()
        name = C.create_string_buffer(128)
        self._ck(self.lib.mvd_device_info(self.ctx, C.byref(sm), C.byref(khz), C.byref(smem), name, 128))
        return dict(sm_count=sm.value, clock_khz=khz.value, smem_per_block_optin=int(smem.value), name=name.value.decode())


class HashOnlyDetector(Detector):
    """Decoder without a Markov state table (m = 5, 6): only :meth:`acs_hash` is usable."""

    def __init__(self, gen1, k, n, m, device=0):
        self.lib = _capi.load()
        self.gen1 = codes.freeze_generator(gen1)
        self.k, self.n, self.m = int(k), int(n), int(m)
        self.R = 1 << self.n
        self.ctx = C.c_void_p()
        rc = self.lib.mvd_create(C.byref(self.ctx), int(device))
        if rc != 0:
            _capi.check(self.lib, None, rc)
        self.dec_taps = codes.tap_masks(self.gen1, self.m, self.k)
        taps = (C.c_uint32 * len(self.dec_taps))(*self.dec_taps)
        self._ck(self.lib.mvd_set_code(self.ctx, self.k, self.n, self.m, taps))
        self.table = None
        self.ntables = 0

    @property
    def S(self):
        return 0


class BareContext(Detector):
    """A device context without a code: entry points that need no trellis (Chernoff spectral radius)."""

    def __init__(self, device: int = 0):
        self.lib = _capi.load()
        self.ctx = C.c_void_p()
        rc = self.lib.mvd_create(C.byref(self.ctx), int(device))
        if rc != 0:
            _capi.check(self.lib, None, rc)
        self.device = int(device)
        self.table = None
        self.ntables = 0


def philox_threshold(p: float) -> int:
    return bitsource.bsc_threshold(p)
