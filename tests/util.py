"""Shared helpers for the parity tests (seeded channel, oracle comparison)."""
import numpy as np


def awgn_llr(ncw, N, sigma, seed, bits=None):
    """Reference Coder::test semantics (MyLdpc.cpp:1061-1078): 0 -> +1.0, 1 -> -1.0, plus noise."""
    rng = np.random.default_rng(seed)
    s = np.ones((ncw, N), dtype=np.float32) if bits is None else (1.0 - 2.0 * bits.astype(np.float32))
    return (s + np.float32(sigma) * rng.standard_normal((ncw, N), dtype=np.float32)).astype(np.float32)


def sigma_from_ebn0(ebn0_db, rate):
    return float(np.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0))))


def assert_parity(res, ref, N, tol=1e-4, what=""):
    """res: dict from Decoder (numpy), ref: (info, iters, hard_bytes, post) from the oracle."""
    info, iters, hard, post = ref
    assert np.array_equal(res["iters"], iters), what + " iteration counts differ at " + str(np.nonzero(res["iters"] != iters)[0][:8])
    assert np.array_equal(res["info"], info), what + " info bytes differ"
    if "hard" in res and hard is not None:
        packed = np.packbits(hard, axis=1, bitorder="little")
        assert np.array_equal(res["hard"], packed), what + " hard bits differ"
    if "post" in res and post is not None:
        a, b = res["post"], post
        assert np.all(np.isfinite(a) == np.isfinite(b))
        # north_star tolerance: <= 1e-4 relative.  (The kernels reproduce the fp32 op order, so the
        # values are in fact identical; the exact check below is the stronger statement.)
        np.testing.assert_allclose(a, b, rtol=tol, atol=0.0, err_msg=what + " posterior outside tolerance")
        assert np.array_equal(a, b), what + " posterior not bit-identical (max rel %g)" % float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))
