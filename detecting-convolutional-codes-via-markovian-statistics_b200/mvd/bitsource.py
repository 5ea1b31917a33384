"""Bit sources for the detector: the MVD-PHILOX-2 stream spec and the bitstream layout.

The reference never defines where its random bits come from: the simulator it calls
(``vm.simulate_markov_sequence``, Pd_plotter.py:149,212,219) is absent, only the idiom is
visible elsewhere (uniform info bits, iid Bernoulli(p) flips, alpha_exponent.py:130-132).
This module *defines* the two sources the CUDA kernels consume; the CPU oracle implements the
same spec independently, so that even throughput-mode (on-device RNG) tallies are bit-exact.

MVD-PHILOX-2
------------
* generator: Philox4x32-10 (Salmon et al., SC'11), key = (seed & 2^32-1, seed >> 32),
  counter = (c0, trial & 2^32-1, trial >> 32, stream).  Every call is addressed by *position*:
  ``c0 = (b << 6) | slot`` with ``b`` the index of a block of 32 trellis steps (b < 2^26), so any
  block of any trial can be generated independently (chunk-parallel learning chains, several
  threads per long trial) and skipping a call never moves the rest of the stream.
* info bits: the four info words of the superblock of blocks 4s..4s+3 come from ONE call with
  ``b = 4s, slot = 32``; bit t of word w is the info bit of step 128 s + 32 w + t.  Codes with k > 1 inputs per step
  (viterbi_markov.py:89-104 is generic in k) take the words of input i from ``slot = 32 + i`` of the same block
  (i < k <= 3; input 0 is the k = 1 definition), and the step's input tuple is (u_0, ..., u_{k-1}).
* BSC flips: for block b and output j (0 <= j < n <= 4) a *lazy Bernoulli word* E_j (bit t = flip of
  output j at step 32 b + t), whose k-th call (k = 0..7) has ``slot = 8 j + k``.
* lazy Bernoulli word for threshold T (P(flip) = T / 2^32), restricted to the valid lanes
  ``vmask`` of the block: ``und = vmask, e = 0``; walk the threshold bits d = 31 .. ctz(T); call k
  supplies the 4 words of levels d = 31 - 4k .. 28 - 4k and is made only while ``und != 0``.
  At level d with random word w: if bit d of T is set, ``e |= und & ~w; und &= w`` else
  ``und &= ~w``.  Lanes still undecided at the end are 0.  This is exactly
  ``(32-bit uniform) < T`` evaluated MSB-first with early termination (expected 6.3 levels for 32
  lanes instead of 32 uniform words).

(MVD-PHILOX-2, the first draft, numbered calls sequentially per trial; the data-dependent number
of lazy calls made every block's position depend on all earlier blocks.)

Bitstream layout (verification mode, also the HBM-bound path)
-------------------------------------------------------------
``bits`` is an array of 128-bit words (4 x uint32, little-endian lanes x,y,z,w), indexed
``[(sb * (k + n) + c) * ntrials + trial]`` where c = i < k is the info stream of input i (k = 1: c = 0) and
c = k + j the flip stream of output j: consecutive trials are adjacent, so a warp reads 512 contiguous bytes.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF

LEARN_STREAM = 0xFFFFFFFF      # stream tag of the learning chain (trial id 0)
ALPHA_STREAM = 0xFFFFFFFE      # stream tag of the joint-tensor chains of alpha_exponent.py (trial id 0)


def philox4x32_10(ctr: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    """Scalar Philox4x32-10 on Python ints."""
    c0, c1, c2, c3 = (int(c) & MASK32 for c in ctr)
    k0, k1 = (int(k) & MASK32 for k in key)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0 = (k0 + PHILOX_W0) & MASK32
        k1 = (k1 + PHILOX_W1) & MASK32
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, seed: int) -> np.ndarray:
    """Vectorised Philox4x32-10: arrays of counters -> uint32 [..., 4]."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.broadcast_to(np.asarray(c1, dtype=np.uint64), c0.shape).copy()
    c2 = np.broadcast_to(np.asarray(c2, dtype=np.uint64), c0.shape).copy()
    c3 = np.broadcast_to(np.asarray(c3, dtype=np.uint64), c0.shape).copy()
    c0 = c0.copy()
    k0 = np.uint64(seed & MASK32)
    k1 = np.uint64((seed >> 32) & MASK32)
    m32 = np.uint64(MASK32)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c0
        p1 = np.uint64(PHILOX_M1) * c2
        n0 = ((p1 >> s32) ^ c1 ^ k0) & m32
        n2 = ((p0 >> s32) ^ c3 ^ k1) & m32
        c1 = p1 & m32
        c3 = p0 & m32
        c0, c2 = n0, n2
        k0 = (k0 + np.uint64(PHILOX_W0)) & m32
        k1 = (k1 + np.uint64(PHILOX_W1)) & m32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def bsc_threshold(p: float) -> int:
    """32-bit flip threshold: P(flip) = T / 2^32, |T/2^32 - p| <= 2^-33 (clipped at 2^32 - 1)."""
    if not (0.0 <= p <= 1.0):
        raise ValueError("p must be in [0, 1]")
    t = int(math.floor(p * 4294967296.0 + 0.5))
    return min(t, MASK32)


def stream_call(seed: int, stream: int, trial: int, block: int, slot: int) -> Tuple[int, int, int, int]:
    """The Philox call at position (block, slot) of the stream (seed, stream, trial)."""
    key = (seed & MASK32, (seed >> 32) & MASK32)
    c0 = ((block << 6) | slot) & MASK32
    return philox4x32_10((c0, trial & MASK32, (trial >> 32) & MASK32, stream & MASK32), key)


INFO_SLOT = 32


def lazy_bernoulli_word(seed: int, stream: int, trial: int, block: int, j: int, T: int, vmask: int) -> int:
    und = vmask & MASK32
    e = 0
    if T == 0:
        return 0
    dmin = (T & -T).bit_length() - 1
    d = 31
    k = 0
    while d >= dmin and und:
        words = stream_call(seed, stream, trial, block, 8 * j + k)
        k += 1
        for w in words:
            if d < dmin:
                break
            if (T >> d) & 1:
                e |= und & ~w & MASK32
                und &= w
            else:
                und &= ~w & MASK32
            d -= 1
    return e


def trial_words(seed: int, stream: int, trial: int, N: int, n: int, T: int) -> Tuple[np.ndarray, np.ndarray]:
    """All info / flip words of one trial under MVD-PHILOX-2.

    Returns ``U`` uint32 [nblk] and ``E`` uint32 [n, nblk] with nblk = ceil(N / 32).
    """
    nblk = (N + 31) // 32
    U = np.zeros(nblk, dtype=np.uint32)
    E = np.zeros((n, nblk), dtype=np.uint32)
    for sb in range((N + 127) // 128):
        uw = stream_call(seed, stream, trial, 4 * sb, INFO_SLOT)
        for w in range(4):
            blk = 4 * sb + w
            t0 = 32 * blk
            if t0 >= N:
                break
            U[blk] = uw[w]
            valid = min(32, N - t0)
            vmask = MASK32 if valid == 32 else (1 << valid) - 1
            for j in range(n):
                E[j, blk] = lazy_bernoulli_word(seed, stream, trial, blk, j, T, vmask)
    return U, E


def trial_words_k(seed: int, stream: int, trial: int, N: int, k: int, n: int, T: int) -> Tuple[np.ndarray, np.ndarray]:
    """:func:`trial_words` for k inputs per step: ``U`` uint32 [k, nblk] (input i from slot 32 + i), ``E`` uint32 [n, nblk]."""
    nblk = (N + 31) // 32
    U0, E = trial_words(seed, stream, trial, N, n, T)
    U = np.zeros((k, nblk), dtype=np.uint32)
    U[0] = U0
    for i in range(1, k):
        for sb in range((N + 127) // 128):
            uw = stream_call(seed, stream, trial, 4 * sb, INFO_SLOT + i)
            for w in range(4):
                if 32 * (4 * sb + w) < N:
                    U[i, 4 * sb + w] = uw[w]
    return U, E


def words_to_bits(words: np.ndarray, N: int) -> np.ndarray:
    """uint32 [..., nblk] -> uint8 bits [..., N], LSB-first inside each word."""
    w = np.ascontiguousarray(words, dtype="<u4")
    bits = np.unpackbits(w.view(np.uint8), axis=-1, bitorder="little")
    return bits[..., :N]


def bits_to_words(bits: np.ndarray) -> np.ndarray:
    """uint8 bits [..., N] -> uint32 words [..., ceil(N/128)*4], LSB-first, zero padded."""
    bits = np.asarray(bits, dtype=np.uint8)
    N = bits.shape[-1]
    pad = (-N) % 128
    if pad:
        bits = np.concatenate([bits, np.zeros(bits.shape[:-1] + (pad,), dtype=np.uint8)], axis=-1)
    by = np.packbits(bits, axis=-1, bitorder="little")
    return np.ascontiguousarray(by).view("<u4")


def pack_bitstreams(u_bits: np.ndarray, e_bits: np.ndarray) -> np.ndarray:
    """Host bit arrays -> device layout.

    ``u_bits`` uint8 [ntrials, N] (k = 1) or [ntrials, k, N]; ``e_bits`` uint8 [ntrials, n, N].
    Returns uint32 [nsb, k + n, ntrials, 4] (C-contiguous) == 128-bit words indexed
    ``(sb * (k + n) + c) * ntrials + trial``.
    """
    u_bits = np.asarray(u_bits, dtype=np.uint8)
    e_bits = np.asarray(e_bits, dtype=np.uint8)
    if u_bits.ndim == 2:
        u_bits = u_bits[:, None, :]
    ntr, k, N = u_bits.shape
    n = e_bits.shape[1]
    assert e_bits.shape == (ntr, n, N)
    uw = bits_to_words(u_bits)                      # [ntr, k, nsb*4]
    ew = bits_to_words(e_bits)                      # [ntr, n, nsb*4]
    nsb = uw.shape[-1] // 4
    out = np.empty((nsb, k + n, ntr, 4), dtype=np.uint32)
    out[:, :k] = uw.reshape(ntr, k, nsb, 4).transpose(2, 1, 0, 3)
    out[:, k:] = ew.reshape(ntr, n, nsb, 4).transpose(2, 1, 0, 3)
    return np.ascontiguousarray(out)


def philox_bitstreams(seed: int, stream: int, trial_begin: int, ntrials: int, N: int, n: int, T: int):
    """Materialise MVD-PHILOX-2 as host bit arrays (small cases; pure Python per trial)."""
    u = np.zeros((ntrials, N), dtype=np.uint8)
    e = np.zeros((ntrials, n, N), dtype=np.uint8)
    for i in range(ntrials):
        U, E = trial_words(seed, stream, trial_begin + i, N, n, T)
        u[i] = words_to_bits(U, N)
        e[i] = words_to_bits(E, N)
    return u, e
