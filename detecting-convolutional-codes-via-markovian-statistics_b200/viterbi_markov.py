"""viterbi_markov -- drop-in for the reference module of the same name, B200 edition.

Same public names and argument orders as the reference's ``viterbi_markov.py`` so that
``import viterbi_markov as vm`` (Pd_plotter.py:61) keeps working:

    state_bits_from_int, bits_to_int, branch_output_and_next_state, hamming_distance,
    build_trellis, viterbi_metric_step, enumerate_markov_states_allzero, build_symbolic_T,
    simulate_markov_sequence, octal_to_taps (the last two are imported by the reference but never defined)

Host-side by design (north star): trellis construction, Markov-state enumeration and the symbolic
T(p) stay on the CPU; they are thin views over :mod:`mvd.codes`.  ``simulate_markov_sequence`` --
which the reference calls (Pd_plotter.py:149,212,219) but never defines -- runs on the GPU
(trace mode of libmvd.so) and fails loudly without it.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from mvd import bitsource, codes


def octal_to_taps(octal, m=None):
    """Octal generator -> tap list, coefficient of D^0 first.  ``alpha_exponent.py:58`` imports this name from
    ``viterbi_markov`` but the reference never defines it; the convention here is the one of the only octal
    parser the reference ships, ``parity_eqn_check.parse_poly_token`` (:80-83, LSB-first), padded to ``m + 1``
    taps when ``m`` is given.  (Pd_plotter.py:247-248 writes its tap lists MSB-first -- "6" there is [1,1,0] --
    so always check which convention a tap list came from; the tap lists themselves are authoritative.)"""
    val = int(str(octal), 8)
    taps = [(val >> i) & 1 for i in range(max(val.bit_length(), 1))]
    if m is not None:
        if len(taps) > m + 1:
            raise ValueError(f"octal {octal} needs more than m + 1 = {m + 1} taps")
        taps += [0] * (m + 1 - len(taps))
    return taps


def state_bits_from_int(state_int, m):
    """Register contents, LSB-first (reference viterbi_markov.py:60-66)."""
    return [(state_int >> pos) & 1 for pos in range(m)]


def bits_to_int(bit_list):
    """LSB-first bit list -> integer (reference viterbi_markov.py:70-75)."""
    return sum((int(b) & 1) << pos for pos, b in enumerate(bit_list))


def branch_output_and_next_state(state_int, input_bits, generator_matrix, m, k):
    """One encoder branch (reference viterbi_markov.py:82-106)."""
    return codes.encoder_branch(int(state_int), tuple(input_bits), generator_matrix, m, k)


def hamming_distance(a, b):
    """d_H over the common prefix of a and b (reference viterbi_markov.py:109-111)."""
    return sum(1 for x, y in zip(a, b) if x != y)


def build_trellis(generator_matrix, m, k):
    """``incoming[ns] = [(ps, u, out), ...]`` (reference viterbi_markov.py:118-132)."""
    return codes.incoming_branches(generator_matrix, m, k)


def viterbi_metric_step(D_prev, trellis, y_t):
    """Eq. 4-5 on the host, for single steps (reference viterbi_markov.py:139-159).
    The GPU kernels run the same recursion in registers; this is the API-compatible scalar."""
    return codes.metric_step(D_prev, trellis, y_t)


@lru_cache(maxsize=32)
def _state_table(gen_frozen, m, k, n):
    return codes.enumerate_states(gen_frozen, m, k, n)


def state_table(generator_matrix, m, k, n) -> codes.StateTable:
    """Flat (numpy) form of the enumeration, cached per code; what the GPU consumes."""
    return _state_table(codes.freeze_generator(generator_matrix), m, k, n)


def enumerate_markov_states_allzero(generator_matrix, m, k, n):
    """``(states, transitions, all_r)`` with the reference's structures and BFS index order
    (reference viterbi_markov.py:166-195)."""
    tab = state_table(generator_matrix, m, k, n)
    transitions, all_r = codes.transitions_from_table(tab)
    return tab.state_tuples(), transitions, all_r


def build_symbolic_T(states, transitions, all_r, normalize=True):
    """Symbolic T(p), Eq. 6 (reference viterbi_markov.py:202-230): returns ``(p, T)`` with ``T`` an
    S x S sympy matrix.  Only the non-zero entries are ever touched, so this takes O(edges)
    instead of O(S^2) simplifications."""
    import sympy as sp

    p = sp.symbols("p")
    S = len(states)
    n = len(all_r[0])
    weight = {tuple(r): p ** sum(r) * (1 - p) ** (n - sum(r)) for r in all_r}
    T = sp.zeros(S, S)
    for i in range(S):
        row = {j: sum(weight[tuple(r)] for r in rl) for j, rl in transitions[i].items()}
        total = sum(row.values())
        for j, val in row.items():
            entry = val / total if (normalize and total != 0) else val
            T[i, j] = sp.simplify(entry)
    return p, T


def numeric_T(states, transitions, all_r, p_val):
    """Numeric T(p_val) as a dense S x S float64 matrix, straight from ``transitions`` -- what
    ``evaluate_symbolic_T(T, p, p_val)`` (reference Pd_plotter.py:89-99) returns for the matrix of
    ``build_symbolic_T`` (reference viterbi_markov.py:202-230), without the sympy round trip
    (44 s at S = 435, infeasible at S >= 10^4).  The GPU consumes the edge form,
    :func:`mvd.codes.t_edge_table`."""
    tab = codes.table_from_transitions(states, transitions, len(all_r[0]))
    return codes.dense_from_edges(tab, codes.t_edge_table(tab, float(p_val)))


def simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input=True, seed=None, *,
                             decoder_matrix=None, u_bits=None, e_bits=None, stream=0, trial=0,
                             engine="acs", device=0):
    """The simulator the reference calls but does not ship (Pd_plotter.py:149-155,212,219).

    Encodes ``length`` info bits with ``generator_matrix``, passes them through BSC(``p_val``) and
    runs the relative-metric recursion of ``decoder_matrix`` (default: the same code) on the GPU.
    Returns ``{"metrics": [D_0, ..., D_length], "states": indices}``; D_0 is all-zero.

    Bits: ``u_bits`` [length] / ``e_bits`` [length][n] if given (verification mode), otherwise the
    on-device MVD-PHILOX-2 stream keyed by ``(seed, stream, trial)``; ``seed=None`` draws a fresh
    64-bit key from numpy's global generator (the reference idiom: no reseed,
    alpha_exponent.py:105-106).
    """
    from mvd.engine import Detector, Seg

    dec = generator_matrix if decoder_matrix is None else decoder_matrix
    det = _detector(codes.freeze_generator(dec), k, n, m, device)
    if k != 1:
        engine = "fsm"      # codes with k > 1 inputs go to the device as tables and run the Markov-state walk (mvd_set_code_tables)
    seg = Seg(N=int(length), threshold=bitsource.bsc_threshold(float(p_val)), stream=stream,
              enc_taps=det.taps_of(generator_matrix), random_input=bool(random_input),
              trial_begin=int(trial), trial_end=int(trial) + 1)
    if u_bits is not None or e_bits is not None:
        if k == 1:
            u = np.zeros((1, length), dtype=np.uint8) if u_bits is None else np.asarray(u_bits, dtype=np.uint8).reshape(1, length)
        else:                                                     # input tuples per step -> [1, k, length]
            u = np.zeros((1, k, length), dtype=np.uint8) if u_bits is None else \
                np.asarray(u_bits, dtype=np.uint8).reshape(length, k).T.reshape(1, k, length)
        e = np.zeros((1, n, length), dtype=np.uint8) if e_bits is None else \
            np.asarray(e_bits, dtype=np.uint8).reshape(length, n).T.reshape(1, n, length)
        idx, met = det.trace(seg, bits=bitsource.pack_bitstreams(u, e), engine=engine)
    else:
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
        idx, met = det.trace(seg, seed=int(seed), engine=engine)
    return {"metrics": [tuple(int(v) for v in row) for row in met[0]], "states": idx[0]}


_DETECTORS = {}


def _detector(gen_frozen, k, n, m, device=0):
    """One GPU context per (decoder code, device), reused across calls."""
    from mvd.engine import Detector

    key = (gen_frozen, k, n, m, device)
    det = _DETECTORS.get(key)
    if det is None:
        if m >= 4:          # S = 2.5e4 .. 2.3e5: enumerate on the GPU (reference index order, mvd_enumerate_states_gpu)
            det = Detector(gen_frozen, k, n, m, device=device, enumerate_with="gpu", max_states=1 << 20)
        else:
            det = Detector(gen_frozen, k, n, m, device=device, table=state_table(gen_frozen, m, k, n))
        _DETECTORS[key] = det
    return det
