"""Seeded synthetic parity-check matrices for the BASELINE configs the reference's own
constructor cannot express (it only builds 802.16e codes with N = 24 z, MyLdpc.cpp:55).

Everything is returned as CSR (row_ptr, col_idx) with ascending columns inside a row -- the
order that fixes the reference's fp32 summation order (MyLdpc.cpp:188-191).  Pure numpy,
host-side setup only; no decode arithmetic lives here.
"""
from __future__ import annotations

import numpy as np


def _csr_from_pairs(M: int, rows: np.ndarray, cols: np.ndarray):
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    row_ptr = np.zeros(M + 1, dtype=np.int32)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    return row_ptr, cols.astype(np.int32)


def regular_code(N: int = 8192, dv: int = 3, dc: int = 6, seed: int = 36):
    """Regular (dv, dc) code by the socket/permutation construction: N*dv variable sockets are
    randomly matched to M*dc check sockets; repeated (row, col) pairs are repaired by swaps.
    Returns (M, N, K, row_ptr, col_idx) with K = N - M (design rate)."""
    assert (N * dv) % dc == 0
    M = N * dv // dc
    rng = np.random.default_rng(seed)
    var_sock = np.repeat(np.arange(N, dtype=np.int64), dv)
    chk_sock = np.repeat(np.arange(M, dtype=np.int64), dc)
    perm = rng.permutation(N * dv)
    cols = var_sock[perm]
    rows = chk_sock.copy()
    for _ in range(1000):
        key = rows * N + cols
        _, first, counts = np.unique(key, return_index=True, return_counts=True)
        if np.all(counts == 1):
            break
        dup = np.setdiff1d(np.arange(key.size), first)
        # swap each duplicate's column with a random other socket
        other = rng.integers(0, key.size, size=dup.size)
        cols[dup], cols[other] = cols[other].copy(), cols[dup].copy()
    else:  # pragma: no cover
        raise RuntimeError("could not remove parallel edges")
    row_ptr, col_idx = _csr_from_pairs(M, rows, cols)
    return M, N, N - M, row_ptr, col_idx


def ira_code(N: int = 64800, K: int = 32400, seed: int = 5, deg_hi: int = 8, n_hi: int = 12960, deg_lo: int = 3):
    """Irregular repeat-accumulate code of DVB-S2 rate-1/2 size: K information columns
    (n_hi of degree deg_hi, the rest deg_lo) spread evenly over the M = N - K checks, plus a
    dual-diagonal (accumulator) parity part: check r touches parity r and r-1.
    Returns (M, N, K, row_ptr, col_idx).  Systematic encoder: ira_encode()."""
    M = N - K
    rng = np.random.default_rng(seed)
    deg = np.full(K, deg_lo, dtype=np.int64)
    deg[:n_hi] = deg_hi
    sockets = np.repeat(np.arange(K, dtype=np.int64), deg)
    E = sockets.size
    # near-uniform check degrees: deal the permuted sockets round-robin over the checks
    for _ in range(1000):
        perm = rng.permutation(E)
        cols = sockets[perm]
        rows = np.arange(E, dtype=np.int64) % M
        key = rows * K + cols
        if np.unique(key).size == E:
            break
        # repair duplicates locally
        for _ in range(100):
            key = rows * K + cols
            _, first, counts = np.unique(key, return_index=True, return_counts=True)
            if np.all(counts == 1):
                break
            dup = np.setdiff1d(np.arange(E), first)
            other = rng.integers(0, E, size=dup.size)
            cols[dup], cols[other] = cols[other].copy(), cols[dup].copy()
        if np.unique(rows * K + cols).size == E:
            break
    else:  # pragma: no cover
        raise RuntimeError("could not build IRA info part")
    prow = np.concatenate([np.arange(M), np.arange(1, M)]).astype(np.int64)
    pcol = np.concatenate([K + np.arange(M), K + np.arange(0, M - 1)]).astype(np.int64)
    rows = np.concatenate([rows, prow])
    cols = np.concatenate([cols, pcol])
    row_ptr, col_idx = _csr_from_pairs(M, rows, cols)
    return M, N, K, row_ptr, col_idx


def ira_encode(M: int, N: int, K: int, row_ptr, col_idx, info_bits: np.ndarray) -> np.ndarray:
    """Systematic encoding for ira_code(): p_0 = s_0, p_r = p_{r-1} + s_r where s = H_info u.
    info_bits: [ncw, K] 0/1.  Returns codeword bits [ncw, N] uint8."""
    u = np.asarray(info_bits, dtype=np.uint8).reshape(-1, K)
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    info_mask = col_idx < K
    s = np.zeros((u.shape[0], M), dtype=np.uint8)
    np.add.at(s.T, rows[info_mask], u.T[col_idx[info_mask]])
    s &= 1
    p = np.bitwise_xor.accumulate(s, axis=1)
    return np.concatenate([u, p], axis=1).astype(np.uint8)


def syndrome(M: int, row_ptr, col_idx, bits: np.ndarray) -> np.ndarray:
    """H x (mod 2) for bits [ncw, N]; returns [ncw, M] uint8."""
    b = np.asarray(bits, dtype=np.uint8)
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    s = np.zeros((b.shape[0], M), dtype=np.uint32)
    np.add.at(s.T, rows, b.T[col_idx])
    return (s & 1).astype(np.uint8)


def gf2_systematic_encoder(M: int, N: int, K: int, row_ptr, col_idx):
    """Generic systematic encoder for H = [A | B] with B (the last M columns) invertible:
    returns the K x M matrix G_p (uint8) with parity = info @ G_p mod 2 (so H [u | p]^T = 0).
    Dense GF(2) elimination -- fine for the 802.16e sizes.  This is what the reference's
    forEncoder/encode compute through Eigen (MyLdpc.cpp:137-165, 633-682): the parity bits
    are uniquely determined by H, so any correct solver yields the reference's codewords."""
    assert N - K == M
    H = np.zeros((M, N), dtype=np.uint8)
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    H[rows, col_idx] = 1
    A, B = H[:, :K].copy(), H[:, K:].copy()
    # solve B X = A  (X is M x K), Gauss-Jordan on [B | A]
    aug = np.concatenate([B, A], axis=1)
    for c in range(M):
        piv = np.nonzero(aug[c:, c])[0]
        if piv.size == 0:
            raise ValueError("parity part of H is singular")
        pr = c + piv[0]
        if pr != c:
            aug[[c, pr]] = aug[[pr, c]]
        others = np.nonzero(aug[:, c])[0]
        others = others[others != c]
        aug[others] ^= aug[c]
    X = aug[:, M:]  # p = X u
    return X.T.copy()  # K x M


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """[ncw, n] 0/1 -> [ncw, ceil(n/8)] bytes, LSB first (the reference's bit order)."""
    return np.packbits(np.asarray(bits, dtype=np.uint8), axis=-1, bitorder="little")


def unpack_bits(bytes_: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.asarray(bytes_, dtype=np.uint8), axis=-1, bitorder="little")[..., :n]
