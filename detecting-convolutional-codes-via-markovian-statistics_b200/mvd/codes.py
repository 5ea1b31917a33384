"""Host-side code / trellis / Markov-state tables for the hybrid detector hot path.

Everything here is *host* work that the north star keeps on the CPU: it turns a
generator matrix into the flat integer tables the CUDA kernels consume.

Semantics follow the reference (citations into /root/reference):

* encoder branch         viterbi_markov.py:82-106  (``branch_output_and_next_state``)
* incoming-branch lists  viterbi_markov.py:118-132 (``build_trellis``)
* Eq. 4-5 metric update  viterbi_markov.py:139-159 (``viterbi_metric_step``)
* BFS state enumeration  viterbi_markov.py:166-195 (``enumerate_markov_states_allzero``)

Conventions used by every table in this package
-----------------------------------------------
* an n-bit branch label / received word ``(o_0, ..., o_{n-1})`` is the integer
  ``sum(o_j << (n-1-j))`` -- i.e. the position of the tuple in
  ``itertools.product([0,1], repeat=n)`` (viterbi_markov.py:175), first output = MSB;
* encoder state bit ``i`` is the input from ``i+1`` steps ago (bit 0 = newest),
  exactly the LSB-first register of viterbi_markov.py:60-75,102-104;
* Markov-state index = BFS discovery order (viterbi_markov.py:189-192).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAX_N_OUT = 4          # device limit on n (received-word alphabet 2^n <= 16)
MAX_M_DEVICE = 6       # device limit on memory m (64 trellis states)


# --------------------------------------------------------------------------- generator handling
def freeze_generator(generator_matrix) -> Tuple[Tuple[Tuple[int, ...], ...], ...]:
    """Nested lists -> hashable nested tuples ``gen[j][i][tap]`` (n x k x taps)."""
    return tuple(tuple(tuple(int(t) for t in taps) for taps in row) for row in generator_matrix)


def label_of(bits: Sequence[int]) -> int:
    """Tuple of output bits -> integer label (first element is the MSB)."""
    v = 0
    for b in bits:
        v = (v << 1) | (int(b) & 1)
    return v


def bits_of_label(label: int, n: int) -> Tuple[int, ...]:
    return tuple((label >> (n - 1 - j)) & 1 for j in range(n))


def encoder_branch(state: int, input_bits: Sequence[int], gen, m: int, k: int) -> Tuple[Tuple[int, ...], int]:
    """One encoder branch: (output tuple, next state).

    Restates viterbi_markov.py:82-106 with integer arithmetic.  The register seen by
    *every* input i is ``[u_i, s_0, ..., s_{m-1}]`` (for k>1 this is the reference's
    own quirk, kept on purpose), truncated to the tap-list length
    (``min(len(taps), len(x))``, viterbi_markov.py:97).
    """
    outs = []
    for row in gen:                      # one row per output j
        acc = 0
        for i in range(k):
            taps = row[i]
            reg = (int(input_bits[i]) & 1) | ((state & ((1 << m) - 1)) << 1)   # bit t <-> x[t]
            for t in range(min(len(taps), m + 1)):
                acc ^= (int(taps[t]) & 1) & (reg >> t)
        outs.append(acc & 1)
    if m > 0:
        keep = max(0, m - k)
        regs = [int(b) & 1 for b in input_bits] + [(state >> i) & 1 for i in range(keep)]
        regs = regs[:m]
        nxt = 0
        for pos, b in enumerate(regs):
            nxt |= b << pos
    else:
        nxt = 0
    return tuple(outs), nxt


def tap_masks(gen, m: int, k: int) -> List[int]:
    """k = 1 only: per-output mask whose bit t is the tap on u_{now - t}."""
    if k != 1:
        raise ValueError("tap masks are defined for k = 1 codes only")
    masks = []
    for row in gen:
        taps = row[0]
        v = 0
        for t in range(min(len(taps), m + 1)):
            v |= (int(taps[t]) & 1) << t
        masks.append(v)
    return masks


def encoder_tables(gen, m: int, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """``enc_out[s, u]`` (label) and ``enc_next[s, u]`` for all states / inputs.

    ``u`` is the index of the input tuple in ``itertools.product([0,1], repeat=k)``.
    """
    ns = 1 << m
    nu = 1 << k
    enc_out = np.zeros((ns, nu), dtype=np.uint8)
    enc_next = np.zeros((ns, nu), dtype=np.uint8)
    for s in range(ns):
        for ui, u in enumerate(itertools.product((0, 1), repeat=k)):
            out, nxt = encoder_branch(s, u, gen, m, k)
            enc_out[s, ui] = label_of(out)
            enc_next[s, ui] = nxt
    return enc_out, enc_next


# --------------------------------------------------------------------------- trellis
def incoming_branches(gen, m: int, k: int) -> Dict[int, List[Tuple[int, Tuple[int, ...], Tuple[int, ...]]]]:
    """``incoming[ns] = [(ps, u, out), ...]`` in the reference's insertion order
    (ps ascending, then u in product order; viterbi_markov.py:124-130)."""
    inc: Dict[int, list] = {s: [] for s in range(1 << m)}
    for ps in range(1 << m):
        for u in itertools.product((0, 1), repeat=k):
            out, ns = encoder_branch(ps, u, gen, m, k)
            inc[ns].append((ps, u, out))
    return inc


def trellis_arrays(gen, m: int, k: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Dense decoder trellis: ``prev[ns, b]`` and ``blab[ns, b]`` (branch label).

    Requires the regular shape every shift-register code has (2^k incoming branches per
    state); raises otherwise.
    """
    inc = incoming_branches(gen, m, k)
    nb = 1 << k
    prev = np.zeros((1 << m, nb), dtype=np.uint8)
    blab = np.zeros((1 << m, nb), dtype=np.uint8)
    for ns, lst in inc.items():
        if len(lst) != nb:
            raise ValueError(f"irregular trellis: state {ns} has {len(lst)} incoming branches")
        for b, (ps, _u, out) in enumerate(lst):
            if len(out) != n:
                raise ValueError("generator matrix does not have n output rows")
            prev[ns, b] = ps
            blab[ns, b] = label_of(out)
    return prev, blab


def metric_step(d_prev: Sequence[int], incoming, y: Sequence[int]) -> Tuple[int, ...]:
    """Eq. 4 (add-compare-select) then Eq. 5 (subtract the minimum).

    Same result as viterbi_markov.py:139-159 for integer metrics.
    """
    nxt = []
    for ns in range(len(d_prev)):
        best = None
        for ps, _u, out in incoming[ns]:
            cand = d_prev[ps] + sum(1 for a, b in zip(out, y) if a != b)
            if best is None or cand < best:
                best = cand
        nxt.append(best)
    lo = min(nxt)
    return tuple(int(v - lo) for v in nxt)


# --------------------------------------------------------------------------- Markov state table
@dataclass
class StateTable:
    """Flat form of ``enumerate_markov_states_allzero`` (viterbi_markov.py:166-195).

    ``metrics[i]`` is the relative-metric vector of Markov state ``i`` (BFS order),
    ``nxt[i, r]`` the state reached on received word ``r``; ``mult[i, r]`` is the number
    of received words that lead from ``i`` to ``nxt[i, r]`` (= ``len(transitions[i][j])``),
    so that ``T(1/2)[i, nxt[i, r]] = mult[i, r] / 2^n``.
    """
    k: int
    n: int
    m: int
    metrics: np.ndarray            # uint8 [S, 2^m]
    nxt: np.ndarray                # uint32 [S, 2^n]
    mult: np.ndarray = field(default=None)   # uint8 [S, 2^n]

    @property
    def S(self) -> int:
        return int(self.metrics.shape[0])

    @property
    def R(self) -> int:
        return 1 << self.n

    @property
    def max_metric(self) -> int:
        return int(self.metrics.max())

    def state_tuples(self) -> List[Tuple[int, ...]]:
        return [tuple(int(v) for v in row) for row in self.metrics]


def _vector_step(front: np.ndarray, prev: np.ndarray, blab: np.ndarray, n: int) -> np.ndarray:
    """All successors of all frontier vectors: returns int16 [F, 2^n, 2^m]."""
    R = 1 << n
    labels = np.arange(R, dtype=np.int64)
    # hamming distance between every branch label and every received word
    x = blab.astype(np.int64)[None, :, :] ^ labels[:, None, None]          # [R, ns, nb]
    dist = np.zeros_like(x)
    for j in range(n):
        dist += (x >> j) & 1
    cand = front.astype(np.int64)[:, prev.astype(np.int64)]                 # [F, ns, nb]
    tot = cand[:, None, :, :] + dist[None, :, :, :]                         # [F, R, ns, nb]
    acs = tot.min(axis=3)
    acs -= acs.min(axis=2, keepdims=True)
    return acs.astype(np.int16)


def enumerate_states(gen, m: int, k: int, n: int, max_states: int | None = None) -> StateTable:
    """Breadth-first closure of the all-zero metric vector under all 2^n received words.

    Level-synchronous, vectorised per level, but the *index assignment* walks candidates
    in (parent index, received word) order -- the order in which the reference's deque
    BFS (viterbi_markov.py:183-193) discovers them -- so indices are identical.
    """
    prev, blab = trellis_arrays(gen, m, k, n)
    nstates = 1 << m
    R = 1 << n
    start = bytes(nstates)
    index: Dict[bytes, int] = {start: 0}
    rows: List[np.ndarray] = [np.zeros(nstates, dtype=np.uint8)]
    nxt_rows: List[np.ndarray] = []
    level = np.zeros((1, nstates), dtype=np.uint8)
    while level.shape[0]:
        succ = _vector_step(level, prev, blab, n)                            # [F, R, ns]
        if succ.max() > 255:
            raise OverflowError("relative metric exceeds 8 bits")
        succ8 = succ.astype(np.uint8)
        F = level.shape[0]
        flat = succ8.reshape(F * R, nstates)
        keys = [row.tobytes() for row in flat]
        out = np.empty(F * R, dtype=np.uint32)
        fresh: List[int] = []
        for pos, key in enumerate(keys):
            j = index.get(key)
            if j is None:
                j = len(rows)
                index[key] = j
                rows.append(flat[pos])
                fresh.append(pos)
                if max_states is not None and j >= max_states:
                    raise MemoryError(f"more than {max_states} Markov states")
            out[pos] = j
        nxt_rows.append(out.reshape(F, R))
        level = flat[fresh] if fresh else np.zeros((0, nstates), dtype=np.uint8)
    metrics = np.stack(rows).astype(np.uint8)
    nxt = np.concatenate(nxt_rows, axis=0).astype(np.uint32)
    assert nxt.shape[0] == metrics.shape[0]
    # multiplicity of the (i -> j) edge = number of r with nxt[i, r] == j
    mult = (nxt[:, :, None] == nxt[:, None, :]).sum(axis=2).astype(np.uint8)
    return StateTable(k=k, n=n, m=m, metrics=metrics, nxt=nxt, mult=mult)


def transitions_from_table(tab: StateTable):
    """Rebuild the reference's ``transitions[i][j] = [r, ...]`` mapping and ``all_r``."""
    from collections import defaultdict

    all_r = list(itertools.product((0, 1), repeat=tab.n))
    trans = defaultdict(lambda: defaultdict(list))
    nxt = tab.nxt
    for i in range(tab.S):
        row = trans[i]
        for r in range(tab.R):
            row[int(nxt[i, r])].append(all_r[r])
    return trans, all_r


def table_from_transitions(states, transitions, n: int, k: int = 1) -> StateTable:
    """Inverse of :func:`transitions_from_table` (accepts the reference's own structures)."""
    S = len(states)
    nstates = len(states[0])
    m = nstates.bit_length() - 1
    metrics = np.array(states, dtype=np.uint8).reshape(S, nstates)
    nxt = np.zeros((S, 1 << n), dtype=np.uint32)
    seen = np.zeros((S, 1 << n), dtype=bool)
    for i in range(S):
        for j, rlist in transitions[i].items():
            for r in rlist:
                ri = label_of(r)
                nxt[i, ri] = j
                seen[i, ri] = True
    if not seen.all():
        raise ValueError("transition structure is not closed under all received words")
    mult = (nxt[:, :, None] == nxt[:, None, :]).sum(axis=2).astype(np.uint8)
    return StateTable(k=k, n=n, m=m, metrics=metrics, nxt=nxt, mult=mult)


# --------------------------------------------------------------------------- log-likelihood tables
def tref_half_table(tab: StateTable) -> np.ndarray:
    """Edge form of ``T(p = 1/2)``: ``tref[i, r] = mult[i, r] / 2^n`` (float64).

    Bit-identical to ``evaluate_symbolic_T(T, p, 0.5)[i, nxt[i, r]]`` (Pd_plotter.py:89-99):
    all values are dyadic and every row sums to exactly 1.
    """
    return tab.mult.astype(np.float64) / float(1 << tab.n)


def t_edge_table(tab: StateTable, p: float) -> np.ndarray:
    """Edge form of the theoretical transition matrix ``T(p)`` (Eq. 6) for any crossover ``p``,
    without sympy: ``T[i, nxt[i, r]] = sum over the received words r' that lead i to the same
    state of p^w(r') (1-p)^(n-w(r'))``, rows renormalised as Pd_plotter.py:97-99 does.

    Same numbers as ``evaluate_symbolic_T(*build_symbolic_T(...)[::-1], p)`` (viterbi_markov.py:
    202-230 + Pd_plotter.py:89-99) on the edges, to float64 rounding; O(S 2^n) instead of the
    S x S sympy matrix, so it also works at S = 10^5 (m = 4) where the reference cannot run.
    """
    n, R = tab.n, tab.R
    w = np.array([bin(r).count("1") for r in range(R)])
    weight = np.array([float(p) ** int(wr) * (1.0 - float(p)) ** int(n - wr) for wr in w])
    same = tab.nxt[:, :, None] == tab.nxt[:, None, :]                    # [S, r, r']
    edge = (same * weight[None, None, :]).sum(axis=2)
    # row sum over *distinct* successors: each successor counted once
    first = np.ones((tab.S, R), dtype=bool)
    for r in range(1, R):
        first[:, r] = ~(tab.nxt[:, :r] == tab.nxt[:, r:r + 1]).any(axis=1)
    rows = (edge * first).sum(axis=1, keepdims=True)
    rows[rows == 0] = 1.0
    return edge / rows


DENSE_LIMIT = 2048      # largest S for which the dense S x S replay of the reference is built


def dense_counts_from_edges(tab: StateTable, edge_counts: np.ndarray) -> np.ndarray:
    """The reference's ``counts`` matrix (Pd_plotter.py:158-163) from edge counts (small S)."""
    S = tab.S
    ec = np.asarray(edge_counts, dtype=np.float64).reshape(S, tab.R)
    flat = (np.arange(S, dtype=np.int64)[:, None] * S + tab.nxt).reshape(-1)     # (i, j) -> i * S + j; duplicates add up
    return np.bincount(flat, weights=ec.reshape(-1), minlength=S * S).reshape(S, S)


def p1_dense(tab: StateTable, edge_counts: np.ndarray, laplace: float) -> np.ndarray:
    """Smoothed, row-normalised estimate with the reference's own two numpy statements
    (Pd_plotter.py:166-167) on the dense matrix -- bit-identical by construction."""
    P = dense_counts_from_edges(tab, edge_counts) + laplace
    P /= P.sum(axis=1, keepdims=True)
    return P


def p1_from_edge_counts(tab: StateTable, edge_counts: np.ndarray, laplace: float) -> np.ndarray:
    """Edge form ``P1[i, r] = P[i, nxt[i, r]]`` of Pd_plotter.py:166-167.

    ``edge_counts[i, r]`` = counted steps that left state ``i`` on received word ``r``.  Up to
    ``DENSE_LIMIT`` states the dense matrix is replayed exactly as the reference builds it and
    gathered; above it (where the reference itself cannot allocate S x S) the closed form
    ``(c_ij + laplace) / (row_i + laplace * S)`` is used, with ``c_ij`` summed over every r that
    reaches j -- identical for dyadic ``laplace`` (1.0 is the only value the reference uses),
    within 2 ulp otherwise.
    """
    S, R = tab.S, tab.R
    ec = np.asarray(edge_counts, dtype=np.float64).reshape(S, R)
    if S <= DENSE_LIMIT:
        P = p1_dense(tab, ec, laplace)
        return np.ascontiguousarray(P[np.arange(S)[:, None], tab.nxt])
    return p1_tables_from_edge_counts(tab, np.asarray(edge_counts).reshape(1, S, R), laplace)[0]


def p1_closed_form_numpy(tab: StateTable, edge_counts: np.ndarray, laplace: float) -> np.ndarray:
    """The closed form in numpy (one table) -- what ``mvd_host_p1_edge_tables`` computes on threads; kept as its check."""
    S, R = tab.S, tab.R
    ec = np.asarray(edge_counts, dtype=np.float64).reshape(S, R)
    cij = np.zeros((S, R))
    for r2, mask in enumerate(_same_successor_masks(tab)):       # c_ij = sum of the counts of every r' that reaches j
        cij += mask * ec[:, r2:r2 + 1]
    denom = ec.sum(axis=1) + laplace * S
    return (cij + laplace) / denom[:, None]


def _exact_row_sums(counts: np.ndarray, laplace: float, S: int) -> bool:
    """True when every partial sum of a dense row ``counts + laplace`` is exact in float64 whatever the order of the
    additions (integer counts, ``laplace`` a multiple of 2^-10, everything below 2^40): the reference's
    ``P.sum(axis=1)`` (Pd_plotter.py:167) is then exactly ``row + laplace * S`` and the closed form divides the same two
    numbers -- bit-identical to replaying the dense matrix, without building it (0.1 ms per sweep at S = 31)."""
    lam = float(laplace)
    if not (0.0 <= lam < 2.0 ** 20) or lam * 1024.0 != np.floor(lam * 1024.0) or S >= 1 << 20:
        return False
    return counts.dtype.kind in "ui" and (counts.size == 0 or int(counts.max()) < 1 << 36)


def p1_tables_from_edge_counts(tab: StateTable, edge_counts: np.ndarray, laplace: float) -> np.ndarray:
    """``p1_from_edge_counts`` for a stack of count tables [T, S, R] -> float64 [T, S, R].  Above ``DENSE_LIMIT`` states
    the closed form runs in C on the host's threads (``mvd_host_p1_edge_tables``; bit-equal to the numpy statements,
    tests): seven tables of S = 150 743 take 3 ms instead of 66."""
    S, R = tab.S, tab.R
    counts = np.asarray(edge_counts)
    T = counts.shape[0]
    if S <= DENSE_LIMIT and not _exact_row_sums(counts, laplace, S):
        return np.stack([p1_from_edge_counts(tab, counts[t], laplace) for t in range(T)])
    from . import _capi
    lib = _capi.load()
    c64 = np.ascontiguousarray(counts, dtype=np.uint64).reshape(T, S, R)
    nxt = np.ascontiguousarray(tab.nxt, dtype=np.uint32)
    out = np.empty((T, S, R), dtype=np.float64)
    _capi.check(lib, None, lib.mvd_host_p1_edge_tables(c64.ctypes.data, nxt.ctypes.data, S, R, T, float(laplace), out.ctypes.data))
    return out


def _same_successor_masks(tab: StateTable):
    """``masks[r2][i, r] = 1.0`` iff ``nxt[i, r2] == nxt[i, r]`` -- cached on the table (built once per code)."""
    masks = getattr(tab, "_same_masks", None)
    if masks is None:
        masks = [(tab.nxt[:, r2:r2 + 1] == tab.nxt).astype(np.float64) for r2 in range(tab.R)]
        tab._same_masks = masks
    return masks


def dense_from_edges(tab: StateTable, edge_values: np.ndarray, fill: float = 0.0) -> np.ndarray:
    """Scatter an edge table [S, R] into a dense S x S matrix (small S only)."""
    S = tab.S
    dense = np.full((S, S), fill, dtype=np.float64)
    for r in range(tab.R):
        dense[np.arange(S), tab.nxt[:, r]] = edge_values[:, r]
    return dense
