"""Pins the sum-product, layered and GPU-min-sum restatements to the REFERENCE'S OWN OpenCL CODE.

oracle/_ref/libmyldpc_refcl.so is the reference's MyLdpc.cpp AND decodeCL.c, both compiled unmodified
(oracle/Makefile); oracle/shim/cl_exec.h makes setArg / enqueueNDRangeKernel / enqueueRead/WriteBuffer
execute the kernels on the CPU, so the reference's own Coder::decode(..., DecodeMS | DecodeSP | DecodeTDMP |
DecodeTDMPCL | DecodeMSCL) host loops (MyLdpc.cpp:786-1059) run.  With forDecoder(1) the "Time=" line the
reference prints per chunk is its per-codeword iteration count.  Skipped where _ref was never built."""
import numpy as np
import pytest

import oracle
from oracle import ref
from tests.util import awgn_llr, sigma_from_ebn0

pytestmark = pytest.mark.skipif(not ref.available("cl"), reason="oracle/_ref/libmyldpc_refcl.so not built (needs /root/reference at build time)")

RATES = [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (4, "3/4B", 3, 4), (5, "5/6", 5, 6)]
MS, SP, TDMP, TDMPCL, MSCL = 1, 2, 3, 4, 5  # enum decodeType, MyLdpc.h:37-39


def _code(N, rate):
    _, name, num, den = RATES[rate]
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    return K, M, rp, ci, num / den


def _ref_decode(K, N, rate, y, src_len, de_type, batch):
    return ref.RefCoder(K, N, rate, opt="cl").decode_cl(y, src_len, de_type, batch=batch)


@pytest.mark.parametrize("rate", range(6))
@pytest.mark.parametrize("N", [576, 960])
def test_reference_opencl_min_sum_is_the_cpu_decode(rate, N):
    """DecodeMS (decodeInitMS / refreshRMS / refreshPostPMS / checkResult / refreshQMS / toChar under
    decodeOnceMS): same bytes as the reference's own decodeCPU and as the oracle, and -- one word per
    chunk -- the same stopping iteration as the oracle reports."""
    K, M, rp, ci, R = _code(N, rate)
    kb = K // 8
    y = np.concatenate([awgn_llr(6, N, sigma_from_ebn0(e, R), seed=N + 10 * rate + i) for i, e in enumerate((1.0, 2.2, 3.5))])
    ncw = y.shape[0]
    info, iters, _, _ = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y)
    got, times = _ref_decode(K, N, rate, y, ncw * kb, MS, batch=1)
    assert np.array_equal(got, info.reshape(-1))
    assert np.array_equal(times, iters)
    cpu = ref.RefCoder(K, N, rate, opt="cl").decode_cpu(y, ncw * kb)  # Coder::decodeCPU of the same build
    assert np.array_equal(got, cpu)
    # chunks of 5 words and a stream that ends inside a codeword (MyLdpc.cpp:577-616): the count the reference
    # prints per chunk is the chunk's maximum
    src_len = ncw * kb - 7
    got2, times2 = _ref_decode(K, N, rate, y, src_len, MS, batch=5)
    assert np.array_equal(got2, info.reshape(-1)[:src_len])
    assert times2.tolist() == [int(iters[i:i + 5].max()) for i in range(0, ncw, 5)]


@pytest.mark.parametrize("rate", range(6))
def test_reference_opencl_sum_product_equals_oracle(rate):
    """DecodeSP (decodeInit / refreshR / hardDecision / checkResult / refreshQ under decodeOnceSP): bytes and
    per-word iteration counts of oracle_decode_sp_batch.  exp() is the one operation that is not the reference's
    own (an OpenCL built-in without a canonical value): both sides use oracle_sp_expf."""
    N = 576
    K, M, rp, ci, R = _code(N, rate)
    kb = K // 8
    y = np.concatenate([awgn_llr(5, N, sigma_from_ebn0(e, R), seed=77 + 10 * rate + i) for i, e in enumerate((1.0, 2.0, 3.0, 4.5))])
    ncw = y.shape[0]
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    info, iters, hard, p0, p1 = oracle.decode_sp(o, y)
    got, times = _ref_decode(K, N, rate, y, ncw * kb, SP, batch=1)
    assert np.array_equal(got, info.reshape(-1))
    assert np.array_equal(times, iters)
    assert iters.min() < 40 and (rate == 0 or iters.max() == 40)  # both regimes are in the sample
    got2, times2 = _ref_decode(K, N, rate, y, ncw * kb - 3, SP, batch=8)
    assert np.array_equal(got2, info.reshape(-1)[:ncw * kb - 3])
    assert times2.tolist() == [int(iters[i:i + 8].max()) for i in range(0, ncw, 8)]


def test_reference_opencl_sum_product_special_values():
    """Saturating inputs: exp(8y) overflows for y > 11.09 (q0 = inf/inf = NaN), underflows for y < -12.9."""
    N, rate = 576, 4
    K, M, rp, ci, R = _code(N, rate)
    y = awgn_llr(6, N, 0.55, seed=5)
    y[0, ::7] = 12.0
    y[1, ::5] = -13.5
    y[2, :40] = 0.0
    y[3, 100:140] = -0.0
    y[4, ::9] *= 30.0
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    info, iters, _, _, _ = oracle.decode_sp(o, y)
    got, times = _ref_decode(K, N, rate, y, 6 * K // 8, SP, batch=1)
    assert np.array_equal(got, info.reshape(-1)) and np.array_equal(times, iters)


def _layer_sizes(rp, z):
    L = (len(rp) - 1) // z
    shipped = [int(rp[b + z] - rp[b]) for b in range(L)]      # MyLdpc.cpp:907,958: hRowRange[blockRow + z] - hRowRange[blockRow]
    true = [int(rp[(b + 1) * z] - rp[b * z]) for b in range(L)]
    return shipped, true


@pytest.mark.parametrize("rate,N", [(1, 576), (5, 576), (1, 1152), (5, 960)])
def test_reference_host_looped_tdmp_equals_layered_oracle_on_uniform_codes(rate, N):
    """DecodeTDMP as shipped (decodeOnceTDMP host loop, MyLdpc.cpp:889-976, kernels decodeCL.c:203-300).  The loop
    sizes layer b as hRowRange[b + z] - hRowRange[b], which is the layer's edge count exactly when all rows have the
    same weight -- rates 2/3A (10) and 5/6 (20).  There the reference's own run gives the bytes AND the per-word
    iteration counts of oracle_decode_tdmp_batch."""
    K, M, rp, ci, R = _code(N, rate)
    z = N // 24
    shipped, true = _layer_sizes(rp, z)
    assert shipped == true
    kb = K // 8
    y = np.concatenate([awgn_llr(5, N, sigma_from_ebn0(e, R), seed=31 + rate + i) for i, e in enumerate((1.5, 2.5, 3.5))])
    ncw = y.shape[0]
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    info, iters, _, _ = oracle.decode_tdmp(o, y, z)
    got, times = _ref_decode(K, N, rate, y, ncw * kb, TDMP, batch=1)
    assert np.array_equal(got, info.reshape(-1))
    assert np.array_equal(times, iters)
    assert iters.min() < 40


@pytest.mark.parametrize("rate", [0, 2, 3, 4])
def test_reference_host_looped_tdmp_breaks_on_mixed_row_weights(rate):
    """Where row weights are mixed (rates 1/2, 2/3B, 3/4A, 3/4B -- Test.cpp's code among them) the shipped loop
    launches the wrong number of work-items from the second layer on: rows of a layer are left out or rows of the
    next layer are processed early with stale lQ (and lQ is indexed below its start; the executor's guard bands
    absorb that).  Documented here: layer sizes differ from the first layer whose neighbours have another weight,
    and the reference's own output is no longer the layered schedule's -- words that the layered oracle decodes in a
    few iterations run into the cap."""
    N = 576
    K, M, rp, ci, R = _code(N, rate)
    z = N // 24
    shipped, true = _layer_sizes(rp, z)
    assert shipped != true and shipped[0] == true[0] and sum(true) == rp[-1]
    kb = K // 8
    y = awgn_llr(16, N, sigma_from_ebn0(3.0, R), seed=400 + rate)
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    info, iters, _, _ = oracle.decode_tdmp(o, y, z)
    got, times = _ref_decode(K, N, rate, y, 16 * kb, TDMP, batch=1)
    assert not np.array_equal(times, iters)
    assert (times > iters).sum() > (times < iters).sum()  # the broken schedule converges later, if at all


@pytest.mark.parametrize("rate,N", [(0, 576), (2, 576), (3, 576), (4, 576), (5, 576), (4, 1152), (1, 2304)])
@pytest.mark.parametrize("layered", [False, True])
def test_fused_kernels_equal_their_restatement(rate, N, layered):
    """DecodeMSCL (decodeOnceMS, decodeCL.c:432-567, cap 120) and DecodeTDMPCL (decodeOnceTDMP, :307-426, cap 40), run
    work-group by work-group with fibers at barrier(): bytes equal oracle_decode_fused_batch on every word, zero
    messages / posteriors and erased inputs included.  (Rate 2/3A only at z = 96: the kernels expand the seed with
    p*z/96 where Coder::initCheckMatrix uses p % z, MyLdpc.cpp:90-94 -- for other z they decode another code.)
    On words without zero events the bytes are those of flooding min-sum at cap 120 / the layered oracle."""
    K, M, rp, ci, R = _code(N, rate)
    z = N // 24
    kb = K // 8
    n_each = 2 if N > 1152 else 4
    y = np.concatenate([awgn_llr(n_each, N, sigma_from_ebn0(e, R), seed=9 + rate + i) for i, e in enumerate((1.2, 2.4, 3.6))])
    y[0, ::11] = 0.0          # erasures: exact-zero Q in the first iteration
    y[1, 5::13] = -0.0
    ncw = y.shape[0]
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    fused = oracle.decode_fused(o, y, z, layered)
    got, _ = _ref_decode(K, N, rate, y, ncw * kb, TDMPCL if layered else MSCL, batch=3)
    assert np.array_equal(got, fused[0].reshape(-1))
    if layered:
        clean = oracle.decode_tdmp(o, y, z)[0]
    else:
        clean = oracle.Oracle(M, N, K, rp, ci, times=120).decode(y)[0]
    same = (clean == fused[0]).all(axis=1)
    assert same[2:].sum() >= (ncw - 2) // 2, "fused-kernel arithmetic should differ from the clean schedule only on zero events"


def test_fused_min_sum_runs_120_iterations():
    """decodeOnceMS's cap is the literal 120 (decodeCL.c:479), not Coder::times = 40: a word that needs 41..120
    iterations decodes under DecodeMSCL and fails under DecodeMS / DecodeCPU."""
    N, rate = 576, 4
    K, M, rp, ci, R = _code(N, rate)
    y = awgn_llr(400, N, 0.6, seed=120)
    it120 = oracle.Oracle(M, N, K, rp, ci, times=120).decode(y, literal=False, want_post=False)[1]
    late = np.nonzero((it120 > 40) & (it120 < 120))[0]
    assert late.size > 0
    w = y[late[:3]]
    o40 = oracle.Oracle(M, N, K, rp, ci, times=40).decode(w)[0]
    o120 = oracle.Oracle(M, N, K, rp, ci, times=120).decode(w)[0]
    got, _ = _ref_decode(K, N, rate, w, w.shape[0] * K // 8, MSCL, batch=2)
    assert np.array_equal(got, o120.reshape(-1)) and not np.array_equal(got, o40.reshape(-1))
