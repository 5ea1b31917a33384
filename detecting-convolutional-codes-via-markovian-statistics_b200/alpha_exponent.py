"""alpha_exponent -- drop-in for the reference module of the same name (Section III-C, Eq. 7), B200 edition.

Public names and signatures follow the reference's ``alpha_exponent.py``:

    spectral_radius, learn_transition_tensor, compute_error_exponent, fit_error_exponent, _encoder_step

As shipped the reference module cannot be imported or run: it imports ``octal_to_taps`` and
``simulate_markov_sequence`` from ``viterbi_markov`` (alpha_exponent.py:58,62 -- neither exists) and calls
``enumerate_markov_states_allzero(decoder_taps, m)`` / ``build_trellis(decoder_taps, m)`` with two arguments
(:109,:116) against the four/three-argument definitions (viterbi_markov.py:166,118).  This module keeps the
call contract and runs the work on the GPU:

* ``learn_transition_tensor``: the 305 000-step chain (:119-141) is one chunk-parallel learning launch
  (``mvd_learn_counts``; the same recursion, the r-indexed histogram *is* the edge-count table because
  ``j = NEXT[i][r]``), Laplace + normalisation (:144-145) in float64 on the host;
* ``compute_error_exponent``: the 401 eigenvalue problems (:170-176) become one launch of a power-iteration
  kernel, one thread block per ``u`` (``mvd_chernoff_rho_dense``); :func:`error_exponent_from_edges` does the
  same from the edge tables in O(K R) per product (``mvd_chernoff_rho``), which also works at m = 4 where
  the dense K x K x R tensor (1.7 TB at K = 232 567) cannot exist.

``taps`` arguments are rate-1/n tap lists ``[g_0, g_1, ...]`` with ``g_j[t]`` the tap of output j on the
register position t (``g_j[0]`` = current input), i.e. ``generator_matrix[j][0]`` of ``viterbi_markov``.

Encoder bit order: the reference's ``_encoder_step`` (:220-234) shifts the register the *other* way round from
``viterbi_markov.branch_output_and_next_state`` -- new input into the top bit, taps 1..m paired with the oldest
... newest past input.  Tap lists whose positions 1..m read the same backwards ([1,1,1]) are unaffected; for any other ([1,0,1], [1,1,0]) the chain is
that of the tap list with positions 1..m reversed.  ``learn_transition_tensor`` reproduces this literally
(:func:`effective_encoder_taps`); pass ``reference_bit_order=False`` for the convention of the rest of the
repository.
"""
from __future__ import annotations

import numpy as np

import viterbi_markov as vm
from mvd import bitsource, codes
from mvd.engine import BareContext, Seg

DENSE_TENSOR_LIMIT = 2048       # largest K for which the dense K x K x R tensor is materialised


def _generator_matrix(taps):
    return [[list(int(b) for b in g)] for g in taps]


def _encoder_step(state, u, taps, m):
    """One encoder update exactly as the reference writes it (alpha_exponent.py:220-234): outputs from
    ``[u, s_0, ..., s_{m-1}]`` (``s_i`` = bit i of ``state``), next state ``(u << (m-1)) | (state >> 1)``."""
    x = [int(u)] + [(state >> i) & 1 for i in range(m)]
    y = tuple(sum(int(g[i]) & x[i] for i in range(len(g))) & 1 for g in taps)
    return y, (((int(u) << (m - 1)) | (state >> 1)) if m > 0 else 0)


def effective_encoder_taps(taps, m, reference_bit_order=True):
    """Tap masks (bit d = tap on the input d steps ago) of the chain ``_encoder_step`` generates: with the
    reference's register, bit i of the state holds u_{t-(m-i)}, so list position 1 + i acts on delay m - i."""
    masks = []
    for g in taps:
        g = [int(b) & 1 for b in g]
        if len(g) > m + 1:
            raise ValueError("tap list longer than m + 1")
        v = g[0] if g else 0
        for pos in range(1, len(g)):
            delay = (m + 1 - pos) if reference_bit_order else pos
            v |= g[pos] << delay
        masks.append(v)
    return masks


def spectral_radius(A):
    """Largest eigenvalue magnitude of a (small, dense) matrix (alpha_exponent.py:69-76)."""
    return float(np.abs(np.linalg.eigvals(np.asarray(A, dtype=float))).max())


# ----------------------------------------------------------------------------- learning
def learn_transition_edges(encoder_taps, decoder_taps, m, p, length=300_000, burn_in=5_000, seed=None, *,
                           reference_bit_order=True, device=0, stream=bitsource.ALPHA_STREAM, trial=0):
    """GPU chain of ``burn_in + length`` steps -> ``(edge_counts uint64 [K, R], state table)``;
    ``edge_counts[i, r]`` = counted steps that left state i on received word r (alpha_exponent.py:133-141)."""
    n = len(decoder_taps)
    det = vm._detector(codes.freeze_generator(_generator_matrix(decoder_taps)), 1, n, int(m), device)
    if seed is None:                                    # no reseed: draw the key from numpy's global stream (:105-106)
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    seg = Seg(N=int(burn_in) + int(length), threshold=bitsource.bsc_threshold(float(p)), stream=stream,
              enc_taps=effective_encoder_taps(encoder_taps, int(m), reference_bit_order),
              trial_begin=int(trial), trial_end=int(trial) + 1)
    counts = det.learn_counts([seg], burn=int(burn_in), seed=int(seed))[0]
    return counts, det.table


def edges_to_tensor(table, edge_counts, laplace=1.0):
    """Dense smoothed joint tensor ``C[i, j, r]`` with the reference's two statements (alpha_exponent.py:144-145)."""
    K, R = table.S, table.R
    if K > DENSE_TENSOR_LIMIT:
        raise MemoryError(f"dense {K} x {K} x {R} tensor refused; use error_exponent_from_edges")
    Cijr = np.zeros((K, K, R), dtype=float)
    rows = np.repeat(np.arange(K), R)
    cols = np.tile(np.arange(R), K)
    Cijr[rows, table.nxt.reshape(-1), cols] = np.asarray(edge_counts, dtype=float).reshape(-1)
    Cijr += laplace
    Cijr /= np.maximum(Cijr.sum(axis=(1, 2), keepdims=True), 1.0)
    return Cijr


def learn_transition_tensor(encoder_taps, decoder_taps, m, p, length=300_000, burn_in=5_000, laplace=1.0, seed=None, *,
                            reference_bit_order=True, device=0):
    """``(C, states, sidx, all_r)`` with ``C[i, j, r]`` ~ P(D_t = j, Y_t = r | D_{t-1} = i), Laplace-smoothed
    (reference alpha_exponent.py:83-149).  The chain runs on the GPU; see the module docstring."""
    counts, table = learn_transition_edges(encoder_taps, decoder_taps, m, p, length, burn_in, seed,
                                           reference_bit_order=reference_bit_order, device=device)
    states = table.state_tuples()
    sidx = {s: i for i, s in enumerate(states)}
    all_r = [codes.bits_of_label(r, table.n) for r in range(table.R)]
    return edges_to_tensor(table, counts, laplace), states, sidx, all_r


# ----------------------------------------------------------------------------- Eq. 7
_CTX = {}


def _context(device=0):
    if device not in _CTX:
        _CTX[device] = BareContext(device)
    return _CTX[device]


def _pick_minimum(u_vals, rho):
    rho = np.maximum(np.asarray(rho, dtype=float), 1e-300)          # alpha_exponent.py:174
    q = int(np.argmin(rho))                                         # first minimum, like the strict '<' of :176
    return float(-np.log(rho[q])), float(u_vals[q])


def compute_error_exponent(P1_ijr, P2_ijr, u_grid=401, *, device=0, details=None):
    """``(I_err, best_u)``: Eq. 7 on a grid of ``u_grid`` points in [0, 1] (reference alpha_exponent.py:155-184).
    The spectral radii are computed on the GPU (one block per u, power iteration on the positive matrix M(u))."""
    P1 = np.clip(np.asarray(P1_ijr, dtype=float), 1e-300, 1.0)
    P2 = np.clip(np.asarray(P2_ijr, dtype=float), 1e-300, 1.0)
    u_vals = np.linspace(0.0, 1.0, int(u_grid))
    rho, iters = _context(device).chernoff_rho_dense(np.log(P1), np.log(P2), u_vals)
    if details is not None:
        details.update(u=u_vals, rho=rho, iters=iters, kernel_ms=_context(device).last_kernel_ms())
    return _pick_minimum(u_vals, rho)


def edge_log_tensors(table, edge_counts, laplace=1.0):
    """Edge form of ``log clip(C, 1e-300, 1)``: ``(lp [K, R], lb [K])`` -- log probability of the edge entries
    ``(i, NEXT[i][r], r)`` and of row i's background entries (every other ``(j, r)``)."""
    K, R = table.S, table.R
    c = np.asarray(edge_counts, dtype=float).reshape(K, R)
    denom = np.maximum(c.sum(axis=1) + laplace * K * R, 1.0)
    lp = np.log(np.clip((c + laplace) / denom[:, None], 1e-300, 1.0))
    lb = np.log(np.clip(np.full(K, float(laplace)) / denom, 1e-300, 1.0))
    return lp, lb


def error_exponent_from_edges(table, counts1, counts2, laplace=1.0, u_grid=401, *, device=0, tol=1e-14,
                              max_iter=100000, details=None):
    """Eq. 7 straight from the two edge-count tables (no K x K x R tensor): ``(I_err, best_u)``."""
    lp1, lb1 = edge_log_tensors(table, counts1, laplace)
    lp2, lb2 = edge_log_tensors(table, counts2, laplace)
    u_vals = np.linspace(0.0, 1.0, int(u_grid))
    ctx = _context(device)
    rho, iters = ctx.chernoff_rho_edges(table.nxt, lp1, lp2, lb1, lb2, u_vals, tol=tol, max_iter=max_iter)
    if details is not None:
        details.update(u=u_vals, rho=rho, iters=iters, kernel_ms=ctx.last_kernel_ms())
    return _pick_minimum(u_vals, rho)


# ----------------------------------------------------------------------------- tail fit
def fit_error_exponent(N_vals, P_e_vals, tail_cap=0.2):
    """Least-squares fit of ``log P_e = log A - I N`` on the tail ``0 < P_e <= tail_cap``
    (reference alpha_exponent.py:191-213): ``(I_emp, A)``, or ``(0.0, nan)`` with fewer than 3 tail points."""
    N = np.asarray(N_vals, dtype=float)
    Pe = np.asarray(P_e_vals, dtype=float)
    tail = (Pe > 0) & (Pe <= tail_cap)
    if int(tail.sum()) < 3:
        return 0.0, float("nan")
    design = np.column_stack([np.ones(int(tail.sum())), -N[tail]])
    coef, *_ = np.linalg.lstsq(design, np.log(Pe[tail]), rcond=None)
    return float(coef[1]), float(np.exp(coef[0]))
