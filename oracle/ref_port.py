"""ref_port.py -- pure-Python restatement of the reference's hybrid-detector path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs, never by the product.

It keeps the reference's *data structures and cost profile* (tuples of Python ints, a dict from
metric tuple to state index, a dense S x S numpy matrix, ``math.log`` per step) so that timing it
on host cores is an honest stand-in for the reference, which cannot travel to the GPU box.
Each function cites the lines of /root/reference it follows.

Parity status: pinned against the reference's own functions executed in the build container
(oracle/make_golden.py -> tests/golden/*.json).  Unpinned: the random-bit source -- the
reference calls ``vm.simulate_markov_sequence`` (Pd_plotter.py:149,212,219) but does not ship it;
``simulate_markov_sequence`` below restates the call contract with the MVD-PHILOX-2 source.
"""
from __future__ import annotations

import itertools
import math
from collections import defaultdict, deque

import numpy as np

M32 = 0xFFFFFFFF
LEARN_STREAM = 0xFFFFFFFF

# Pd_plotter.py:78-83
N_SPECTRUM_BY_M = {1: [5, 10, 20, 50, 100, 200], 2: [500], 3: [500], 4: [50, 100, 200, 300, 500]}


# ----------------------------------------------------------------------------- L1: trellis core
def branch(state, u_bits, G, m, k):
    """viterbi_markov.py:82-106 -- GF(2) convolution of [u_i, s_0..s_{m-1}] with every tap list,
    then the LSB-first shift-register update."""
    sbits = [(state >> i) & 1 for i in range(m)]                    # :60-66
    outs = []
    for g in G:                                                     # :89
        bit = 0
        for i in range(k):                                          # :91
            reg = [u_bits[i]] + sbits                               # :92
            for t in range(min(len(g[i]), len(reg))):               # :94 (truncating)
                bit ^= g[i][t] & reg[t]
        outs.append(bit)
    regs = (list(u_bits) + sbits[:max(0, m - k)])[:m] if m > 0 else []   # :102-103
    nxt = sum((b & 1) << i for i, b in enumerate(regs))             # :70-75
    return tuple(outs), nxt


def trellis_of(G, m, k):
    """viterbi_markov.py:118-132 -- incoming[ns] = [(ps, u, out), ...]."""
    inc = {s: [] for s in range(1 << m)}
    for s in range(1 << m):
        for u in itertools.product([0, 1], repeat=k):
            out, ns = branch(s, u, G, m, k)
            inc[ns].append((s, u, out))
    return inc


def metric_step(D, trellis, y):
    """viterbi_markov.py:139-159 -- Eq. 4 then Eq. 5."""
    new = []
    for ns in range(len(D)):
        best = math.inf
        for ps, _, out in trellis[ns]:
            v = D[ps] + sum(a != b for a, b in zip(out, y))         # :109-111
            if v < best:
                best = v
        new.append(best)
    lo = min(new)
    return tuple(int(v - lo) for v in new)


def enumerate_states(G, m, k, n):
    """viterbi_markov.py:166-195 -- BFS closure, index = discovery order."""
    trellis = trellis_of(G, m, k)
    all_r = list(itertools.product([0, 1], repeat=n))               # :175
    start = (0,) * (1 << m)                                         # :177
    index = {start: 0}
    states = [start]
    trans = defaultdict(lambda: defaultdict(list))
    todo = deque([start])
    while todo:
        cur = todo.popleft()
        ci = index[cur]
        for r in all_r:
            nx = metric_step(cur, trellis, r)
            if nx not in index:
                index[nx] = len(states)
                states.append(nx)
                todo.append(nx)
            trans[ci][index[nx]].append(r)                          # :193
    return states, trans, all_r


def numeric_T(states, trans, all_r, p):
    """Eq. 6 evaluated numerically: what evaluate_symbolic_T(build_symbolic_T(...), p)
    returns (viterbi_markov.py:217-228 + Pd_plotter.py:89-99), without sympy."""
    S = len(states)
    n = len(all_r[0])
    T = np.zeros((S, S))
    for i in range(S):
        for j, rl in trans[i].items():
            T[i, j] = sum(p ** sum(r) * (1 - p) ** (n - sum(r)) for r in rl)
    rs = T.sum(axis=1, keepdims=True)
    rs[rs == 0] = 1.0
    return T / rs


# ----------------------------------------------------------------------------- bit source
def _philox(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        a = 0xD2511F53 * c0
        b = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((b >> 32) ^ c1 ^ k0) & M32, b & M32, ((a >> 32) ^ c3 ^ k1) & M32, a & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def threshold_of(p):
    return min(int(math.floor(p * 4294967296.0 + 0.5)), M32)


def philox_bits(seed, stream, trial, N, n, T, k=1):
    """MVD-PHILOX-2 (mvd/bitsource.py docstring): info bits u[N] (k = 1) or input tuples u[N][k] (k > 1: input i from
    slot 32 + i) and flips e[N][n] of one trial.  Every Philox call is addressed by position: c0 = (block << 6) | slot."""
    key = (seed & M32, (seed >> 32) & M32)

    def call(block, slot):
        return _philox((((block << 6) | slot) & M32, trial & M32, (trial >> 32) & M32, stream & M32), key)

    def lazy(block, j, vmask):
        und, e = vmask, 0
        if T == 0:
            return 0
        dmin = (T & -T).bit_length() - 1
        d, k = 31, 0
        while d >= dmin and und:
            for w in call(block, 8 * j + k):
                if d < dmin:
                    break
                if (T >> d) & 1:
                    e |= und & ~w & M32
                    und &= w
                else:
                    und &= ~w & M32
                d -= 1
            k += 1
        return e

    u = [0] * N if k == 1 else [[0] * k for _ in range(N)]
    e = [[0] * n for _ in range(N)]
    for sb in range((N + 127) // 128):
        uw = call(4 * sb, 32)
        uk = [uw] + [call(4 * sb, 32 + i) for i in range(1, k)]
        for w in range(4):
            t0 = 128 * sb + 32 * w
            if t0 >= N:
                break
            valid = min(32, N - t0)
            vmask = M32 if valid == 32 else (1 << valid) - 1
            ew = [lazy(4 * sb + w, j, vmask) for j in range(n)]
            for b in range(valid):
                if k == 1:
                    u[t0 + b] = (uw[w] >> b) & 1
                else:
                    u[t0 + b] = [(uk[i][w] >> b) & 1 for i in range(k)]
                for j in range(n):
                    e[t0 + b][j] = (ew[j] >> b) & 1
    return u, e


# ----------------------------------------------------------------------------- the missing simulator
def simulate_markov_sequence(generator_matrix, m, k, n, length, p_val, random_input=True, seed=None,
                             *, decoder_matrix=None, u_bits=None, e_bits=None, stream=0, trial=0,
                             step=metric_step, branch_fn=branch, trellis_fn=trellis_of):
    """Contract of the absent ``vm.simulate_markov_sequence`` (call sites Pd_plotter.py:149-155,
    212,219): returns ``{"metrics": [D_0 .. D_length]}``; D_0 all-zero (viterbi_markov.py:177),
    encoder state 0 (alpha_exponent.py:123), k = 1 info bits uniform, flips iid Bernoulli(p),
    recursion on the *decoder* trellis (H1).  Bits come from ``u_bits``/``e_bits`` when
    given, else from MVD-PHILOX-2 keyed by (seed, stream, trial).

    ``step``/``branch_fn``/``trellis_fn`` let make_golden.py run this driver on the reference's
    own functions.
    """
    dec = generator_matrix if decoder_matrix is None else decoder_matrix
    trellis = trellis_fn(dec, m, k)
    if u_bits is None or e_bits is None:
        gu, ge = philox_bits(0 if seed is None else int(seed), stream, trial, length, n, threshold_of(p_val), k)
        u_bits = gu if u_bits is None else u_bits
        e_bits = ge if e_bits is None else e_bits
    D = tuple([0] * (1 << m))
    metrics = [D]
    received = []
    enc = 0
    for t in range(length):
        u = (int(u_bits[t]) if random_input else 0,) if k == 1 else tuple(int(b) if random_input else 0 for b in u_bits[t])
        out, enc = branch_fn(enc, u, generator_matrix, m, k)
        r = tuple(int(o) ^ int(f) for o, f in zip(out, e_bits[t]))
        D = step(list(D), trellis, r)
        metrics.append(D)
        received.append(r)
    return {"metrics": metrics, "received": received}


# ----------------------------------------------------------------------------- L2: detector statistics
def log_prob_sequence(metrics, state_index, T):
    """Pd_plotter.py:106-116."""
    lp = 0.0
    for t in range(len(metrics) - 1):
        pij = max(T[state_index[metrics[t]], state_index[metrics[t + 1]]], 1e-300)
        lp += math.log(pij)
    return lp


def learn_P1(G1, k, n, m, p, learn_len, learn_burn, laplace, seed):
    """Pd_plotter.py:123-169 (without the lru_cache): returns (states, index, P, counts)."""
    states, _, _ = enumerate_states(G1, m, k, n)
    index = {s: i for i, s in enumerate(states)}
    S = len(states)
    L = max(5000, 200 * S) if learn_len is None else learn_len      # :143-146
    sim = simulate_markov_sequence(G1, m, k, n, L, p, True, seed, stream=LEARN_STREAM, trial=0)
    mets = sim["metrics"]
    counts = np.zeros((S, S))
    for t in range(learn_burn, len(mets) - 1):                      # :158-163
        counts[index[mets[t]], index[mets[t + 1]]] += 1.0
    P = counts + laplace                                            # :166-167
    P /= P.sum(axis=1, keepdims=True)
    return states, index, P, counts


def run_experiment(k, n, m, gen1, gen2, num_iter, p_vec, learn_len, learn_burn, laplace, seed,
                   N_spectrum=None, trial_offset=0, collect=None):
    """Pd_plotter.py:176-235 with the stream convention of the product's run_experiment:
    point id = position in the (N major, p minor) sweep; trial streams 2*point + {0: H1, 1: H2};
    trial ids trial_offset .. trial_offset + num_iter - 1.  Returns a list of row dicts that also
    carry the raw tallies (s1, s2)."""
    states, trans, all_r = enumerate_states(gen1, m, k, n)
    T_ref = numeric_T(states, trans, all_r, 0.5)                    # :193-194
    rows = []
    spectrum = N_SPECTRUM_BY_M.get(m, [50, 100, 200]) if N_spectrum is None else list(N_spectrum)
    learned = {}
    point = 0
    for N in spectrum:
        for p in p_vec:
            if p not in learned:                                    # lru_cache, :123
                learned[p] = learn_P1(gen1, k, n, m, p, learn_len, learn_burn, laplace, seed)
            _, index, P1, _ = learned[p]
            s1 = s2 = 0
            for it in range(num_iter):
                tr = trial_offset + it
                a = simulate_markov_sequence(gen1, m, k, n, N, p, True, seed, decoder_matrix=gen1,
                                             stream=2 * point, trial=tr)["metrics"]
                l1, l1r = log_prob_sequence(a, index, P1), log_prob_sequence(a, index, T_ref)
                s1 += l1 > l1r                                      # :215
                b = simulate_markov_sequence(gen2, m, k, n, N, p, True, seed, decoder_matrix=gen1,
                                             stream=2 * point + 1, trial=tr)["metrics"]
                l2, l2r = log_prob_sequence(b, index, P1), log_prob_sequence(b, index, T_ref)
                s2 += l2 <= l2r                                     # :222
                if collect is not None:
                    collect.append((point, tr, l1, l1r, l2, l2r))
            rows.append({"N": N, "p": p, "Pd": s1 / num_iter, "Pc": (s1 + s2) / (2 * num_iter),
                         "s1": int(s1), "s2": int(s2)})
            point += 1
    return rows


# ----------------------------------------------------------------------------- alpha_exponent.py (Eq. 7)
ALPHA_STREAM = 0xFFFFFFFE


def alpha_encoder_step(state, u, taps, m):
    """alpha_exponent.py:220-234, literally: x = [u, s_0..s_{m-1}], next = (u << (m-1)) | (state >> 1)."""
    x = [u] + [(state >> i) & 1 for i in range(m)]
    y = tuple(sum(g[i] & x[i] for i in range(len(g))) % 2 for g in taps)
    return y, (((u << (m - 1)) | (state >> 1)) if m > 0 else 0)


def learn_transition_tensor(encoder_taps, decoder_taps, m, p, length=300_000, burn_in=5_000, laplace=1.0,
                            seed=0, stream=ALPHA_STREAM, trial=0):
    """alpha_exponent.py:83-149 with its two stale calls adapted to the shipped signatures
    (enumerate_markov_states_allzero(decoder_taps, m) -> (G, m, 1, n); build_trellis(decoder_taps, m) ->
    (G, m, 1)) and np.random replaced by the MVD-PHILOX-2 bits of (seed, stream, trial): per step one info
    bit (:130,:138) then one flip per output in order (:132,:140).  Returns (C, states, sidx, all_r, counts)."""
    n = len(decoder_taps)
    G = [[list(g)] for g in decoder_taps]
    states, _, all_r = enumerate_states(G, m, 1, n)                 # :109
    sidx = {s: i for i, s in enumerate(states)}
    K, R = len(states), len(all_r)
    r_index = {r: i for i, r in enumerate(all_r)}
    trellis = trellis_of(G, m, 1)                                   # :116
    u_bits, e_bits = philox_bits(seed, stream, trial, burn_in + length, n, threshold_of(p))
    cur = tuple([0] * (1 << m))                                     # :119
    cur_i = sidx[cur]
    enc = 0                                                         # :123
    C = np.zeros((K, K, R))
    for t in range(burn_in):                                        # :129-135
        y, enc = alpha_encoder_step(enc, u_bits[t], encoder_taps, m)
        r = tuple(b ^ f for b, f in zip(y, e_bits[t]))
        cur = metric_step(list(cur), trellis, r)
        cur_i = sidx.get(cur, cur_i)
    for t in range(burn_in, burn_in + length):                      # :137-149
        y, enc = alpha_encoder_step(enc, u_bits[t], encoder_taps, m)
        r = tuple(b ^ f for b, f in zip(y, e_bits[t]))
        nxt = metric_step(list(cur), trellis, r)
        nxt_i = sidx.get(nxt, None)
        if nxt_i is not None:
            C[cur_i, nxt_i, r_index[r]] += 1.0
            cur, cur_i = nxt, nxt_i
    counts = C.copy()
    C += laplace                                                    # :144-145
    C /= np.maximum(C.sum(axis=(1, 2), keepdims=True), 1.0)
    return C, states, sidx, all_r, counts


def chernoff_matrix(P1_ijr, P2_ijr, u):
    """M(u) of Eq. 7 as alpha_exponent.py:167-171 forms it."""
    P1 = np.clip(P1_ijr, 1e-300, 1.0)
    P2 = np.clip(P2_ijr, 1e-300, 1.0)
    return np.sum((P1 ** u) * (P2 ** (1.0 - u)), axis=2)


def compute_error_exponent(P1_ijr, P2_ijr, u_grid=401):
    """alpha_exponent.py:155-184: grid search of -log rho(M(u)) with LAPACK eigenvalues; also returns rho(u)."""
    best, best_u, rhos = None, None, []
    for u in np.linspace(0.0, 1.0, u_grid):
        rho = max(float(np.max(np.abs(np.linalg.eigvals(chernoff_matrix(P1_ijr, P2_ijr, u))))), 1e-300)
        rhos.append(rho)
        if best is None or rho < best:
            best, best_u = rho, u
    return float(-np.log(best)), float(best_u), rhos


# ----------------------------------------------------------------------------- comp_parity.py (section IV baseline)
PARITY_STREAM_BASE = 0x40000000


def encode_convolutional(u_bits, generators, m):
    """comp_parity.py:65-86 -- time-domain feed-forward encoder, n streams of len(u) + m bits."""
    n, T = len(generators), len(u_bits) + m
    out = [[0] * T for _ in range(n)]
    for t in range(T):
        for j in range(n):
            acc = 0
            for shift, bit in enumerate(generators[j][0]):          # :80-82
                if bit and 0 <= t - shift < len(u_bits):
                    acc ^= u_bits[t - shift]
            out[j][t] = acc
    return out


def parity_satisfaction(y, template):
    """comp_parity.py:93-116 -- (satisfied, total) over t in [max_delay, T)."""
    T = len(y[0])
    max_delay = max(s for _, s in template)                         # :101
    satisfied = total = 0
    for t in range(max_delay, T):                                   # :107-113
        x = 0
        for j, s in template:
            x ^= y[j][t - s]
        total += 1
        satisfied += x == 0
    return satisfied, total


def parity_trial(generators, m, template, gamma, N, p, seed, stream, trial):
    """One iteration of comp_parity.py:165-176 on MVD-PHILOX-2 bits: (decide_H1, satisfied, total)."""
    n = len(generators)
    u, e = philox_bits(seed, stream, trial, N + m, n, threshold_of(p))
    v = encode_convolutional(u[:N], generators, m)                  # :167
    y = [[v[j][t] ^ e[t][j] for t in range(N + m)] for j in range(n)]   # :171
    sat, total = parity_satisfaction(y, template)
    frac = sat / total if total > 0 else 0.0                        # :116
    return frac >= gamma, sat, total                                # :131-132


def timed_steps(gen1, gen2, m, k, n, N, p, num_iter, seed, trial_offset=0):
    """bench helper: run ``num_iter`` iterations of the trial loop (Pd_plotter.py:210-223) of one
    point with a fixed small learned P1; returns (steps done, tallies)."""
    rows = run_experiment(k, n, m, gen1, gen2, num_iter, [p], None, 200, 1.0, seed,
                          N_spectrum=[N], trial_offset=trial_offset)
    return 2 * N * num_iter, (rows[0]["s1"], rows[0]["s2"])
