"""Oracle unit tests on hand-checkable parity-check matrices (no reference needed)."""
import numpy as np
import pytest

import oracle


def _csr(H):
    H = np.asarray(H)
    rp = np.concatenate([[0], np.cumsum(H.sum(1))]).astype(np.int32)
    ci = np.concatenate([np.nonzero(r)[0] for r in H]).astype(np.int32)
    return H.shape[0], H.shape[1], rp, ci


H_HAMMING = [[1, 1, 0, 1, 1, 0, 0],
             [1, 0, 1, 1, 0, 1, 0],
             [0, 1, 1, 1, 0, 0, 1]]


def test_noiseless_word_stops_after_one_iteration():
    M, N, rp, ci = _csr(H_HAMMING)
    o = oracle.Oracle(M, N, 4, rp, ci, times=40)
    cw = np.array([1, 0, 1, 1, 0, 1, 0])  # H cw = 0
    assert (np.array(H_HAMMING) @ cw % 2).sum() == 0
    y = (1.0 - 2.0 * cw).astype(np.float32)[None, :]
    info, iters, hard, post = o.decode(y)
    assert iters[0] == 1 and np.array_equal(hard[0], cw)
    assert info[0, 0] == (1 | 4 | 8)  # bits 0,2,3 set, LSB first


def test_first_iteration_by_hand():
    """One iteration on the (7,4) Hamming matrix, worked by hand from MyLdpc.cpp:705-735."""
    M, N, rp, ci = _csr(H_HAMMING)
    y = np.array([[0.9, -0.3, 1.2, 0.5, -0.1, 2.0, 0.7]], dtype=np.float32)
    o = oracle.Oracle(M, N, 4, rp, ci, times=1)
    _, iters, hard, post = o.decode(y)
    f = np.float32
    # row 0 = {0,1,3,4}: R to col 0 = sign(-0.3*0.5*-0.1=+) * min(0.3,0.5,0.1) = +0.1; ...
    R = {
        (0, 0): f(0.1), (0, 1): f(-0.1), (0, 3): f(0.1), (0, 4): f(-0.3),
        (1, 0): f(0.5), (1, 2): f(0.5), (1, 3): f(0.9), (1, 5): f(0.5),
        (2, 1): f(0.5), (2, 2): f(-0.3), (2, 3): f(-0.3), (2, 6): f(-0.3),
    }
    want = y[0].copy()
    for (r, c), v in sorted(R.items()):  # ascending row order per column
        want[c] = f(want[c] + v)
    assert np.array_equal(post[0], want)
    assert np.array_equal(hard[0], (~(want > 0)).astype(np.uint8))
    assert iters[0] == 1


def test_zero_posterior_decides_one_and_degree_one_check_gives_1000():
    # a single check touching only column 0 (degree 1): R = +1000 (b starts at 1000, MyLdpc.cpp:708)
    H = [[1, 0, 0], [0, 1, 1]]
    M, N, rp, ci = _csr(H)
    o = oracle.Oracle(M, N, 3, rp, ci, times=1)
    y = np.array([[-2.0, 0.5, -0.5]], dtype=np.float32)
    _, _, hard, post = o.decode(y)
    assert post[0, 0] == np.float32(998.0)
    # col 1: 0.5 + (-0.5) = 0 -> bit 1 (tmp > 0 is false, MyLdpc.cpp:729-733); col 2: -0.5 + 0.5 = 0 -> 1
    assert post[0, 1] == 0 and post[0, 2] == 0 and hard[0, 1] == 1 and hard[0, 2] == 1


def test_cap_is_inclusive_and_literal_equals_fast():
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    rng = np.random.default_rng(0)
    y = (1 + 0.9 * rng.standard_normal((64, 576))).astype(np.float32)
    for cap in (1, 3, 40):
        o = oracle.Oracle(M, 576, 432, rp, ci, times=cap)
        a = o.decode(y, literal=True)
        b = o.decode(y, literal=False, threads=3)
        assert all(np.array_equal(x, z) for x, z in zip(a, b))
        assert a[1].max() == cap and a[1].min() >= 1


def test_size_helpers_and_stream_packing():
    L = oracle.lib()
    assert L.oracle_getCodeSize(432, 1000) == 19           # ceil(1000 / 54), MyLdpc.cpp:628-631
    assert L.oracle_getPostCodeLength(432, 576, 1000) == 19 * 576
    assert L.oracle_getPriorCodeLength(432, 576, 1000) == 19 * 72
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    o = oracle.Oracle(M, 576, 432, rp, ci)
    y = -np.ones((2, 576), dtype=np.float32)  # all-one hard decision never satisfies -> cap; bits all 1
    stream, iters, hard, _ = o.decode_stream(y, 100)
    assert stream.shape == (100,) and np.all(stream == 0xFF) and np.all(hard == 1)


def test_sum_product_restatement_basics():
    """oracle_sp_expf is a ~2 ulp exp; the sum-product restatement decodes clean words in one iteration
    and corrects more noise than min-sum on the same inputs (it is the better decoder)."""
    import math
    L = oracle.lib()
    for x in (-80.0, -8.0, -0.3, 0.0, 0.7, 8.0, 20.0, 80.0):
        assert abs(L.oracle_sp_expf(x) / math.exp(x) - 1.0) < 5e-7
    assert L.oracle_sp_expf(100.0) == float("inf") and L.oracle_sp_expf(-200.0) == 0.0
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    o = oracle.Oracle(M, 576, 432, rp, ci)
    rng = np.random.default_rng(3)
    clean = np.ones((4, 576), dtype=np.float32)
    info, iters, hard, p0, p1 = oracle.decode_sp(o, clean)
    assert np.all(iters == 1) and not hard.any()
    y = (1 + 0.58 * rng.standard_normal((128, 576))).astype(np.float32)
    sp = oracle.decode_sp(o, y)
    ms = o.decode(y, literal=False)
    assert sp[2].sum() <= ms[2].sum()


def test_layered_restatement_basics():
    """The layered (TDMP) restatement: clean words stop at iteration 1, it needs fewer iterations than the
    flooding schedule on the same noisy words, rejects layers whose rows share a column, and an all-zero
    input keeps every bit at its initial 0 (the P == 0 rule of hardDecisionTDMP)."""
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    o = oracle.Oracle(M, 576, 432, rp, ci)
    info, iters, hard, post = oracle.decode_tdmp(o, np.ones((3, 576), dtype=np.float32), 24)
    assert np.all(iters == 1) and not hard.any()
    rng = np.random.default_rng(5)
    y = (1 + 0.55 * rng.standard_normal((256, 576))).astype(np.float32)
    td = oracle.decode_tdmp(o, y, 24)
    ms = o.decode(y, literal=False)
    assert td[1].mean() < 0.8 * ms[1].mean()
    assert td[2].sum() <= ms[2].sum()
    z0 = oracle.decode_tdmp(o, np.zeros((1, 576), dtype=np.float32), 24)
    assert z0[1][0] == 1 and not z0[2].any()
    assert oracle.lib().oracle_tdmp_layering_ok(o._t, 24) == 1
    assert oracle.lib().oracle_tdmp_layering_ok(o._t, 48) == 0
    with pytest.raises(ValueError):
        oracle.decode_tdmp(o, y[:1], 48)


def _gf2_codewords(rp, ci, M, N, K, count, seed):
    """Random codewords of H = [A | B] (B = the last M columns): parity p solves B p = A u over GF(2)."""
    H = np.zeros((M, N), dtype=np.uint8)
    for r in range(M):
        H[r, ci[rp[r]:rp[r + 1]]] = 1
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 2, (count, K)).astype(np.uint8)
    aug = np.concatenate([H[:, K:], (H[:, :K] @ u.T) % 2], axis=1).astype(np.uint8)   # [B | A u]
    for col in range(M):                                                                # Gauss-Jordan mod 2
        piv = col + int(np.nonzero(aug[col:, col])[0][0])
        if piv != col:
            aug[[col, piv]] = aug[[piv, col]]
        rows = np.nonzero(aug[:, col])[0]
        rows = rows[rows != col]
        aug[rows] ^= aug[col]
    cw = np.concatenate([u, aug[:, M:].T], axis=1).astype(np.uint8)
    assert not ((H.astype(np.int64) @ cw.T.astype(np.int64)) % 2).any()
    return cw


@pytest.mark.parametrize("name,num,den", [("3/4B", 3, 4), ("1/2", 1, 2), ("5/6", 5, 6)])
def test_channel_symmetry_of_min_sum_and_layered(name, num, den):
    """A size-independent property of the algorithm (Coder::decodeCPU, MyLdpc.cpp:684-784, and the layered schedule): sending
    codeword c over the same noise instead of the all-zero word flips exactly the signs of c's positions -- every message,
    posterior and hard bit -- so hard bits are XORed with c, posteriors negated there (bit for bit: fp32 negation is exact),
    and the iteration counts are equal.  (Exact zeros, where `P > 0 ? 0 : 1` is not symmetric, do not occur in random noise.)"""
    N = 576
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    cw = _gf2_codewords(rp, ci, M, N, K, 48, seed=5)
    rng = np.random.default_rng(6)
    sig = {"3/4B": 0.52, "1/2": 0.75, "5/6": 0.45}[name]
    y0 = (1.0 + sig * rng.standard_normal((48, N))).astype(np.float32)
    y0[40:] = (1.0 + 1.2 * rng.standard_normal((8, N))).astype(np.float32)      # words that run into the cap
    s = (1.0 - 2.0 * cw).astype(np.float32)
    y1 = y0 * s
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    for what, dec in (("min-sum", lambda y: o.decode(y)), ("layered", lambda y: oracle.decode_tdmp(o, y, N // 24))):
        i0, t0, h0, p0 = dec(y0)
        i1, t1, h1, p1 = dec(y1)
        assert not (p0 == 0).any()
        assert np.array_equal(t0, t1), what + ": iteration counts"
        assert np.array_equal(h1, h0 ^ cw), what + ": hard bits"
        assert np.array_equal(p1, p0 * s), what + ": posteriors"
        assert np.array_equal(np.unpackbits(i1, axis=1, bitorder="little")[:, :K], h1[:, :K]), what + ": info bytes"
        assert (t0[:40] < 40).sum() >= 30 and (t0[40:] == 40).all()   # both regimes are in the sample
