"""comp_parity -- drop-in for the reference's parity-template baseline detector (paper section IV), B200 edition.

Reference names kept: ``encode_convolutional`` (comp_parity.py:65-86), ``parity_satisfaction_fraction`` (:93-116),
``parity_detector`` (:123-132) -- host scalars for single sequences -- and a ``__main__`` that prints the
parity equation and the H1 accuracy like the reference's (:139-181).

New: the Monte-Carlo loop itself (:165-176) runs on the GPU (``mvd_parity_detect``: one thread per trial,
encoder / BSC / template / popcount all bit-parallel), for both hypotheses and a whole (N, p) sweep in one
launch, and ``__main__`` also writes ``results_parity/Pd_parity_results.csv`` (columns ``N,p,Pd,Pc``) -- the
file README.md:190-192 and ``plots_compare.py --baseline`` expect but the reference never writes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np

from mvd import _capi, bitsource, dist
from parity_eqn_check import build_parity_system, nullspace_mod2, parity_vector_to_equation, parse_poly_token, split_parity_vector

DEFAULTS = {"N": 200, "p": 0.1, "gamma": 0.6, "trials": 1000, "seed": 12345,            # reference :158-162
            "p_vec": [0.001, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5], "save_dir": "results_parity"}
PARITY_STREAM_BASE = 0x40000000       # stream tags of the parity trials: base + 2 * point + hypothesis


def encode_convolutional(u_bits: List[int], generators: List[List[List[int]]], m: int):
    """n output streams of ``len(u_bits) + m`` bits: feed-forward convolution with a zero tail (reference :65-86)."""
    K = len(u_bits)
    T = K + m
    streams = []
    for row in generators:
        taps = [s for s, bit in enumerate(row[0]) if bit]
        streams.append([sum(u_bits[t - s] for s in taps if 0 <= t - s < K) & 1 for t in range(T)])
    return streams


def parity_satisfaction_fraction(y: List[List[int]], template: List[Tuple[int, int]]):
    """Fraction of positions t >= max delay with XOR_{(j,s)} y[j][t-s] == 0 (reference :93-116)."""
    T = len(y[0])
    reach = max(s for _, s in template)
    total = max(0, T - reach)
    if total == 0:
        return 0.0
    ok = sum(1 for t in range(reach, T) if sum(y[j][t - s] for j, s in template) % 2 == 0)
    return ok / total


def parity_detector(y: List[List[int]], template: List[Tuple[int, int]], gamma: float):
    """``(decide_H1, P_hat)`` with decide_H1 = P_hat >= gamma (reference :123-132)."""
    frac = parity_satisfaction_fraction(y, template)
    return frac >= gamma, frac


# ----------------------------------------------------------------------------- template construction
def template_from_generators(generators, m: int, deg_h: int | None = None, which: int = 0):
    """The reference's choice (:139-157): first nullspace vector of the parity system with deg_h = m + 3
    -> ``(template [(j, s), ...], h_vec)``."""
    deg_h = m + 3 if deg_h is None else int(deg_h)
    basis = nullspace_mod2(build_parity_system(generators, deg_h))
    if len(basis) <= which:
        raise ValueError("no parity-check vector of this degree")
    h_vec = split_parity_vector(basis[which], len(generators), deg_h)
    return [(j, s) for j, poly in enumerate(h_vec) for s, bit in enumerate(poly) if bit], h_vec


def _tap_masks(generators) -> List[int]:
    return [sum((int(b) & 1) << s for s, b in enumerate(row[0])) for row in generators]


def _template_masks(template: Sequence[Tuple[int, int]], n: int) -> List[int]:
    masks = [0] * n
    for j, s in template:
        if not 0 <= s <= 31:
            raise ValueError("template delays above 31 are not supported on the device")
        masks[j] ^= 1 << s                      # a repeated term cancels, as it does in the XOR of :108-109
    return masks


# ----------------------------------------------------------------------------- GPU Monte-Carlo
class ParityContext:
    """Device context for ``mvd_parity_detect`` (no trellis needed)."""

    def __init__(self, device: int = 0):
        self.lib = _capi.load()
        self.ctx = C.c_void_p()
        rc = self.lib.mvd_create(C.byref(self.ctx), int(device))
        if rc != 0:
            _capi.check(self.lib, None, rc)

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.mvd_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, segs: Sequence[dict], seed=None, bits=None, want_satisfied=False):
        """``segs``: dicts with N, m, taps (masks), tmpl (masks), gamma, decide, threshold, stream, trial_begin,
        trial_end, bits_offset.  Returns tallies uint64 [nsegs] (and satisfied counts per trial)."""
        arr = (_capi.ParitySegment * len(segs))()
        total = 0
        for a, s in zip(arr, segs):
            a.N, a.m, a.n = int(s["N"]), int(s["m"]), len(s["taps"])
            a.threshold, a.stream, a.decide = int(s.get("threshold", 0)), int(s.get("stream", 0)) & 0xFFFFFFFF, int(s.get("decide", 0))
            for j in range(_capi.MAX_N):
                a.enc_taps[j] = int(s["taps"][j]) if j < len(s["taps"]) else 0
                a.tmpl[j] = int(s["tmpl"][j]) if j < len(s["tmpl"]) else 0
            a.gamma = float(s["gamma"])
            a.trial_begin, a.trial_end = int(s.get("trial_begin", 0)), int(s.get("trial_end", 1))
            a.bits_offset = int(s.get("bits_offset", 0))
            total += a.trial_end - a.trial_begin
        src = _capi.Src()
        keep = None
        if bits is None:
            src.mode, src.seed = _capi.SRC_PHILOX, int(seed or 0) & 0xFFFFFFFFFFFFFFFF
        else:
            keep = np.ascontiguousarray(bits, dtype=np.uint32)
            src.mode, src.bits_on_device, src.bits, src.bits_words = _capi.SRC_BITSTREAM, 0, keep.ctypes.data, keep.size // 4
        tallies = np.zeros(len(segs), dtype=np.uint64)
        sat = np.zeros(total, dtype=np.uint32) if want_satisfied else None
        _capi.check(self.lib, self.ctx, self.lib.mvd_parity_detect(self.ctx, C.byref(src), arr, len(segs), tallies.ctypes.data,
                                                                   sat.ctypes.data if want_satisfied else None))
        del keep
        return (tallies, sat) if want_satisfied else tallies

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _capi.check(self.lib, self.ctx, self.lib.mvd_last_kernel_ms(self.ctx, C.byref(ms)))
        return float(ms.value)


_CTX = {}


def _context(device=0) -> ParityContext:
    if device not in _CTX:
        _CTX[device] = ParityContext(device)
    return _CTX[device]


def run_parity_experiment(generators_h1, generators_h2, m, N_list, p_vec, gamma, trials, seed, *, template=None,
                          deg_h=None, device=None, trial_offset=0, details=None):
    """Parity-template detector over all (N, p) points -> DataFrame[N, p, Pd, Pc] (N-major, p-minor rows, the
    layout of Pd_plotter's CSV): ``Pd`` = P(decide H1 | H1 sent), ``Pc`` = mean of that and P(decide H2 | H2
    sent).  The template is built from ``generators_h1`` as the reference does (:139-157) unless given.
    Trials shard across ranks under torchrun; tallies are combined with one allreduce."""
    import pandas as pd

    if device is None:
        device = int(os.environ.get("LOCAL_RANK", 0))
    if template is None:
        template, _ = template_from_generators(generators_h1, m, deg_h)
    n = len(generators_h1)
    tmpl = _template_masks(template, n)
    taps = (_tap_masks(generators_h1), _tap_masks(generators_h2))
    rank, ws = dist.world()
    begin, end = dist.shard_range(int(trials), rank, ws, offset=int(trial_offset))
    segs, points = [], []
    for N in N_list:
        for p in p_vec:
            q = len(points)
            for h in (0, 1):
                segs.append(dict(N=int(N), m=int(m), taps=taps[h], tmpl=tmpl, gamma=float(gamma), decide=h,
                                 threshold=bitsource.bsc_threshold(float(p)), stream=PARITY_STREAM_BASE + 2 * q + h,
                                 trial_begin=begin, trial_end=end))
            points.append((N, p))
    ctx = _context(device)
    tallies = ctx.run(segs, seed=int(seed))
    kernel_ms = ctx.last_kernel_ms()
    tallies = dist.allreduce_sum(tallies.astype(np.int64))
    rows = []
    for q, (N, p) in enumerate(points):
        s1, s2 = int(tallies[2 * q]), int(tallies[2 * q + 1])
        rows.append({"N": N, "p": p, "Pd": s1 / trials, "Pc": (s1 + s2) / (2 * trials)})
    if details is not None:
        details.update(tallies=tallies, kernel_ms=kernel_ms, template=template,
                       steps=2 * sum(int(N) + int(m) for N, _ in points) * int(trials))
    return pd.DataFrame(rows, columns=["N", "p", "Pd", "Pc"])


if __name__ == "__main__":
    g1 = [parse_poly_token("7")]                                   # reference :137-139
    g2 = [parse_poly_token("5")]
    generators = [g1, g2]
    m = 2
    template, h_vec = template_from_generators(generators, m)
    print("Using parity equation:")
    print(parity_vector_to_equation(h_vec))

    # the reference's own experiment (:158-181): H1 only, one point
    alt = [[parse_poly_token("6")], [parse_poly_token("5")]]      # the second hypothesis of the hybrid experiment
    df = run_parity_experiment(generators, alt, m, [DEFAULTS["N"]], [DEFAULTS["p"]], DEFAULTS["gamma"], DEFAULTS["trials"],
                               DEFAULTS["seed"], template=template)
    print(f"Baseline parity detector accuracy: {df['Pd'][0]:.3f}")

    # the sweep plots_compare.py reads (README.md:190-192)
    from Pd_plotter import N_SPECTRUM_BY_M
    sweep = run_parity_experiment(generators, alt, m, [DEFAULTS["N"]] + N_SPECTRUM_BY_M[m], DEFAULTS["p_vec"], DEFAULTS["gamma"],
                                  10 * DEFAULTS["trials"], DEFAULTS["seed"], template=template)
    os.makedirs(DEFAULTS["save_dir"], exist_ok=True)
    out_csv = os.path.join(DEFAULTS["save_dir"], "Pd_parity_results.csv")
    sweep.to_csv(out_csv, index=False)
    print("Saved results to", out_csv)
