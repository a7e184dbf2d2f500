"""The SHIPPED drop-in (C++ `class Coder`, libmyldpc_b200.so, reached through include/MyLdpc_c.h) against the
REFERENCE'S OWN `Coder` (oracle/_ref: MyLdpc.cpp + decodeCL.c compiled unmodified, OpenCL variants executed on the
CPU): identical postCode floats into both `decode(postCode, srcCode, srcLength, deType)` calls, srcCode bytes and
per-word iteration counts compared, for all six decodeTypes, ragged srcLength included."""
import numpy as np
import pytest

import oracle
from oracle import ref
from tests.util import awgn_llr, sigma_from_ebn0

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available("cl"), reason="oracle/_ref not built (needs /root/reference at build time)")]

RATES = [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (4, "3/4B", 3, 4), (5, "5/6", 5, 6)]
CPU, MS, SP, TDMP, TDMPCL, MSCL = range(6)  # enum decodeType, MyLdpc.h:37-39


def _inputs(N, rate, ncw, seed, ebn0s=(1.5, 2.5, 3.5)):
    _, name, num, den = RATES[rate]
    K = N * num // den
    per = (ncw + len(ebn0s) - 1) // len(ebn0s)
    y = np.concatenate([awgn_llr(per, N, sigma_from_ebn0(e, num / den), seed=seed + i) for i, e in enumerate(ebn0s)])[:ncw]
    return K, np.ascontiguousarray(y)


def _b200_decode(K, N, rate, y, src_len, de_type, batch):
    import myldpccppapi_b200 as m
    c = m.Coder(K, N, rate)
    c.forDecoder(batch)
    c.addDecodeType(de_type)
    out = np.zeros(src_len + 1, dtype=np.uint8)
    assert c.decode(y.reshape(-1), out, src_len, de_type) == 0
    return out[:src_len], c.lastIterations.copy(), c


@pytest.mark.parametrize("rate", range(6))
@pytest.mark.parametrize("de_type", [CPU, MS])
def test_min_sum_bytes_and_iterations_equal_the_reference(rate, de_type):
    N, ncw = 576, 45
    K, y = _inputs(N, rate, ncw, seed=700 + rate)
    kb = K // 8
    src_len = ncw * kb - 11  # the stream ends inside the last codeword (MyLdpc.cpp:620-631, 765-774)
    got, iters, _ = _b200_decode(K, N, rate, y, src_len, de_type, batch=16)
    r = ref.RefCoder(K, N, rate, opt="cl")
    if de_type == CPU:
        want = r.decode_cpu(y, src_len)                       # the reference's Coder::decodeCPU
        _, times = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, src_len, MS, batch=1)
    else:
        want, times = r.decode_cl(y, src_len, MS, batch=1)    # the reference's OpenCL min-sum, one word per chunk
    assert np.array_equal(got, want)
    assert np.array_equal(iters, times)


@pytest.mark.parametrize("rate", range(6))
def test_sum_product_bytes_and_iterations_equal_the_reference(rate):
    N, ncw = 576, 30
    K, y = _inputs(N, rate, ncw, seed=800 + rate, ebn0s=(1.0, 2.0, 3.0, 4.5))
    kb = K // 8
    src_len = ncw * kb - 5
    got, iters, c = _b200_decode(K, N, rate, y, src_len, SP, batch=8)
    assert c.lastAlgorithm == 1
    want, times = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, src_len, SP, batch=1)
    assert np.array_equal(got, want)
    assert np.array_equal(iters, times)


@pytest.mark.parametrize("rate,N", [(1, 576), (5, 576), (1, 1152), (5, 960)])
def test_layered_bytes_and_iterations_equal_the_reference_host_loop(rate, N):
    """DecodeTDMP against the reference's own decodeOnceTDMP loop, on the codes where that loop is sound (one row
    weight: rates 2/3A and 5/6; see tests/test_oracle_vs_refcl.py for the others)."""
    ncw = 24
    K, y = _inputs(N, rate, ncw, seed=900 + rate)
    kb = K // 8
    src_len = ncw * kb - 3
    got, iters, c = _b200_decode(K, N, rate, y, src_len, TDMP, batch=8)
    assert c.lastAlgorithm == 2
    want, times = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, src_len, TDMP, batch=1)
    assert np.array_equal(got, want)
    assert np.array_equal(iters, times)


@pytest.mark.parametrize("rate", [0, 2, 3, 4, 5])
def test_fused_layered_kernel_bytes_equal_the_reference(rate):
    """DecodeTDMPCL against the reference's fused kernel decodeOnceTDMP (decodeCL.c:307-426) run work-group by
    work-group.  That kernel takes message signs from a float product and decides bit = (P < 0): it differs from the
    layered schedule only when a message or posterior is exactly zero, which happens in words that never converge;
    every word that converges must carry the reference kernel's bytes, and so must most of the rest."""
    N, ncw = 576, 36
    K, y = _inputs(N, rate, ncw, seed=1000 + rate, ebn0s=(2.0, 3.0, 4.0))
    kb = K // 8
    got, iters, _ = _b200_decode(K, N, rate, y, ncw * kb, TDMPCL, batch=12)
    want, _ = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, ncw * kb, TDMPCL, batch=12)
    same = (got.reshape(ncw, kb) == want.reshape(ncw, kb)).all(axis=1)
    assert same[iters < 40].all()
    assert (iters < 40).sum() >= ncw // 3


@pytest.mark.parametrize("rate", [0, 2, 3, 4, 5])
def test_fused_min_sum_kernel_bytes_equal_the_reference(rate):
    """DecodeMSCL against the reference's fused kernel decodeOnceMS (decodeCL.c:432-567): cap 120, not 40."""
    N, ncw = 576, 36
    K, y = _inputs(N, rate, ncw, seed=1100 + rate, ebn0s=(2.0, 2.8, 3.6))
    kb = K // 8
    got, iters, _ = _b200_decode(K, N, rate, y, ncw * kb, MSCL, batch=12)
    want, _ = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, ncw * kb, MSCL, batch=12)
    same = (got.reshape(ncw, kb) == want.reshape(ncw, kb)).all(axis=1)
    assert same[iters < 120].all()
    assert iters.max() <= 120
    # the same stream with DecodeMS stops at 40
    _, it40, _ = _b200_decode(K, N, rate, y, ncw * kb, MS, batch=12)
    assert np.array_equal(np.minimum(iters, 40), it40)


@pytest.mark.parametrize("rate,N", [(0, 576), (2, 576), (3, 576), (4, 576), (5, 576), (4, 1152), (1, 2304)])
@pytest.mark.parametrize("de_type", [TDMPCL, MSCL])
def test_fused_kernel_arithmetic_is_byte_identical_on_request(rate, N, de_type):
    """setFusedKernelArithmetic(true): every word -- zero messages, zero posteriors and erased inputs included -- carries
    the bytes of the reference's fused OpenCL kernel executed work-group by work-group.  (Rate 2/3A only at z = 96: for
    other z the reference's fused kernels expand a different matrix than Coder::initCheckMatrix.)"""
    import myldpccppapi_b200 as m
    ncw = 12 if N > 1152 else 30
    K, y = _inputs(N, rate, ncw, seed=1200 + rate, ebn0s=(1.2, 2.4, 3.6))
    y[0, ::11] = 0.0
    y[1, 5::13] = -0.0
    kb = K // 8
    c = m.Coder(K, N, rate)
    c.setFusedKernelArithmetic(True)
    c.forDecoder(ncw)
    c.addDecodeType(de_type)
    out = np.zeros(ncw * kb + 1, dtype=np.uint8)
    assert c.decode(y.reshape(-1), out, ncw * kb, de_type) == 0
    assert c.lastAlgorithm == (4 if de_type == TDMPCL else 3)
    want, _ = ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, ncw * kb, de_type, batch=6)
    assert np.array_equal(out[:-1], want)
    _, name, num, den = RATES[rate]
    rp, ci, M = oracle.wimax_H(N, name)
    fused = oracle.decode_fused(oracle.Oracle(M, N, K, rp, ci), y, N // 24, layered=(de_type == TDMPCL))
    assert np.array_equal(c.lastIterations, fused[1])


def test_strict_decode_type_and_fallback_query():
    """DecodeSP on a code that does not fit the on-chip layout runs the any-size sum-product kernel (no fallback).
    DecodeTDMP on a code without a layer structure cannot be layered: by default it is decoded with flooding min-sum and
    lastAlgorithm says so; strict mode returns LDPC_B200_ERR_UNSUPPORTED instead."""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    y = awgn_llr(4, N, 0.7, seed=3)
    c = m.Coder.from_csr(M, N, K, rp, ci)
    c.forDecoder(4)
    c.addDecodeType(SP)
    out = np.zeros(4 * K // 8 + 1, dtype=np.uint8)
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    assert c.decode(y.reshape(-1), out, 4 * K // 8, SP) == 0
    assert c.lastAlgorithm == 1
    sp = oracle.decode_sp(o, y)
    assert np.array_equal(out[:-1], sp[0].reshape(-1)) and np.array_equal(c.lastIterations, sp[1])
    assert c.decode(y.reshape(-1), out, 4 * K // 8, TDMP) == 0
    assert c.lastAlgorithm == 0 and "min-sum" in c.lastError()
    want = o.decode(y, literal=False)[0]
    assert np.array_equal(out[:-1], want.reshape(-1))
    c.setStrictDecodeType(True)
    with pytest.raises(m.LdpcError) as e:
        c.decode(y.reshape(-1), out, 4 * K // 8, TDMP)
    assert e.value.code == -3
    assert c.decode(y.reshape(-1), out, 4 * K // 8, MS) == 0 and c.lastAlgorithm == 0


def test_step_times_are_reported():
    import myldpccppapi_b200 as m
    N, rate = 576, 4
    K, y = _inputs(N, rate, 4000, seed=5)
    c = m.Coder(K, N, rate)
    c.forDecoder(4000)
    c.addDecodeType(MS)
    out = np.zeros(4000 * K // 8 + 1, dtype=np.uint8)
    c.decode(y.reshape(-1), out, 4000 * K // 8, MS)
    t = c.lastStepTimes()
    assert t["wall_s"] > 0 and t["kernel_s"] > 0 and t["h2d_s"] > 0 and t["d2h_s"] > 0
    assert t["kernel_s"] <= t["wall_s"] * 1.05


def test_set_devices_gives_identical_bytes():
    """Coder::setDevices shards contiguous codeword ranges over the GPUs of the box with no collective: bytes and
    iteration counts must not depend on the device count (needs >= 2 GPUs; skipped on a one-GPU box)."""
    import torch
    import myldpccppapi_b200 as m
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("one GPU")
    N, rate = 576, 4
    K, y = _inputs(N, rate, 5003, seed=77)
    kb = K // 8
    src_len = 5003 * kb - 9
    base = None
    for g in [1, 2, 4, 8]:
        if g > ndev:
            break
        c = m.Coder(K, N, rate)
        c.setDevices(list(range(g)))
        c.forDecoder(5003)
        c.addDecodeType(MS)
        out = np.zeros(src_len + 1, dtype=np.uint8)
        c.decode(y.reshape(-1), out, src_len, MS)
        cur = (out[:src_len].copy(), c.lastIterations.copy())
        if base is None:
            base = cur
            want = oracle.Oracle(N - K, N, K, *oracle.wimax_H(N, "3/4B")[:2], times=40).decode_stream(y.reshape(-1), src_len)
            assert np.array_equal(cur[0], want[0]) and np.array_equal(cur[1], want[1])
        assert np.array_equal(cur[0], base[0]) and np.array_equal(cur[1], base[1]), "differs with %d devices" % g
