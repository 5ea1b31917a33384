"""parity_eqn_check -- drop-in for the reference module of the same name (paper section IV): parity-check
polynomial vectors h(D) with sum_j h_j(D) g_j(D) = 0 over GF(2), by Gaussian elimination.

Same public names and results as the reference's ``parity_eqn_check.py``:
``parse_poly_token`` (:60-84), ``nullspace_mod2`` (:91-139), ``build_parity_system`` (:146-181),
``parity_vector_to_equation`` (:188-200).  Host-side algebra (a 2(deg_h+1)-column system); the Monte-Carlo
that consumes the template runs on the GPU (:mod:`comp_parity`)."""
from __future__ import annotations

import re
from typing import List

import numpy as np


def parse_poly_token(token: str) -> List[int]:
    """Generator polynomial token -> coefficient list, D^0 first.  Accepts a comma-separated tap list
    (returned as given), a string of binary digits written MSB-first, or an octal number -- tried in this
    order, so "101" is binary and "7" octal (reference :60-84)."""
    token = token.strip()
    if "," in token:
        return [int(part) for part in token.split(",")]
    if re.fullmatch(r"[01]+", token):
        return [int(ch) for ch in reversed(token)]
    if re.fullmatch(r"[0-7]+", token):
        value = int(token, 8)
        return [(value >> pos) & 1 for pos in range(value.bit_length())]
    raise ValueError(f"Cannot parse polynomial token: {token}")


def nullspace_mod2(A: np.ndarray) -> np.ndarray:
    """Basis of {x : A x = 0 mod 2}, one vector per free column (ascending), read off the reduced row echelon
    form (unique, so the basis equals the reference's, :91-139).  Rows are eliminated as Python integers."""
    A = np.asarray(A, dtype=np.uint8) & 1
    nrows, ncols = A.shape
    rows = [int("".join(str(int(b)) for b in A[r, ::-1]), 2) if ncols else 0 for r in range(nrows)]   # bit c = column c
    pivots: List[int] = []
    rank = 0
    for col in range(ncols):
        if rank >= nrows:
            break
        hit = next((r for r in range(rank, nrows) if (rows[r] >> col) & 1), None)
        if hit is None:
            continue
        rows[rank], rows[hit] = rows[hit], rows[rank]
        for r in range(nrows):
            if r != rank and (rows[r] >> col) & 1:
                rows[r] ^= rows[rank]
        pivots.append(col)
        rank += 1
    free = [c for c in range(ncols) if c not in pivots]
    if not free:
        return np.zeros((0, ncols), dtype=np.uint8)
    basis = np.zeros((len(free), ncols), dtype=np.uint8)
    for b, f in enumerate(free):
        basis[b, f] = 1
        for r, pc in enumerate(pivots):
            if (rows[r] >> f) & 1:
                basis[b, pc] = 1
    return basis


def build_parity_system(generators: List[List[List[int]]], deg_h: int) -> np.ndarray:
    """Coefficient matrix of sum_j h_j(D) g_{j,i}(D) = 0 for every input i: one row per (i, power of D), one
    column per unknown h_{j,s} at index j (deg_h + 1) + s (reference :146-181)."""
    n, k = len(generators), len(generators[0])
    deg_g = max(len(g) - 1 for out in generators for g in out)
    top = deg_g + deg_h
    A = np.zeros((k * (top + 1), n * (deg_h + 1)), dtype=np.uint8)
    for i in range(k):
        for j in range(n):
            for u, tap in enumerate(generators[j][i]):
                if not tap:
                    continue
                for s in range(deg_h + 1):                 # h_{j,s} D^s * D^u lands on power s + u
                    A[i * (top + 1) + s + u, j * (deg_h + 1) + s] ^= 1
    return A


def parity_vector_to_equation(h_vec: List[List[int]]) -> str:
    """h(D) -> "v0[t-0] ⊕ v1[t-2] ⊕ ... = 0" (reference :188-200)."""
    terms = [f"v{j}[t-{s}]" for j, poly in enumerate(h_vec) for s, bit in enumerate(poly) if bit]
    return " ⊕ ".join(terms) + " = 0"


def split_parity_vector(row, n: int, deg_h: int) -> List[List[int]]:
    """Nullspace row -> [h_0, ..., h_{n-1}], each of deg_h + 1 coefficients (the slicing of reference :216-219)."""
    return [[int(v) for v in row[j * (deg_h + 1):(j + 1) * (deg_h + 1)]] for j in range(n)]


if __name__ == "__main__":
    gens = [[parse_poly_token("7")], [parse_poly_token("5")]]      # reference :209-212
    deg_h = 5
    basis = nullspace_mod2(build_parity_system(gens, deg_h))
    print(f"Found {len(basis)} parity-check vectors")
    for row in basis:
        print(parity_vector_to_equation(split_parity_vector(row, len(gens), deg_h)))
