"""Guard-band test of every kernel family (stands in for the memcheck run compute-sanitizer cannot do on this pool,
profiles/r02_sanitizer_closed.txt): each output buffer of ldpc_b200_decode_device sits between two 64 KB bands of a
sentinel byte, the channel values between two bands of NaN.  A store outside an output buffer changes a sentinel; a
load outside the input that reaches the arithmetic turns the word's result into something the oracle does not give.
Ragged batch sizes (1, 37, 333 words) so that the last group / warp / CTA of every layout is partly empty."""
import numpy as np
import pytest

import oracle
from tests.util import awgn_llr, sigma_from_ebn0

pytestmark = pytest.mark.gpu

GUARD = 65536          # bytes on either side of every buffer
SENTINEL = 0xA5


def _guarded(torch, nbytes, skew=0):
    """A uint8 CUDA buffer of nbytes between two guard bands; the buffer starts `skew` bytes past a 16-byte boundary."""
    whole = torch.full((GUARD + skew + nbytes + GUARD,), SENTINEL, dtype=torch.uint8, device="cuda")
    return whole, whole[GUARD + skew:GUARD + skew + nbytes]


def _bands_intact(whole, nbytes, skew=0):
    lo = bool((whole[:GUARD + skew] == SENTINEL).all().item())
    hi = bool((whole[GUARD + skew + nbytes:] == SENTINEL).all().item())
    return lo, hi


def _decode_guarded(dec, y, want_post=True, skewed=False):
    """skewed: every buffer at the weakest alignment its element type allows (floats and counts 4 bytes past a 16-byte
    boundary, packed bits 1 byte past) -- what a caller slicing into a larger array hands over."""
    import torch
    ncw, N = y.shape
    KB, NB = dec.KB, dec.NB
    # channel values between NaN bands
    gf = GUARD // 4 + (1 if skewed else 0)
    yin = torch.full((gf + ncw * N + gf,), float("nan"), dtype=torch.float32, device="cuda")
    yin[gf:gf + ncw * N] = torch.from_numpy(y).cuda().reshape(-1)
    llr = yin[gf:gf + ncw * N].view(ncw, N)
    skew = {"info": 1, "hard": 1, "iters": 4, "post": 4} if skewed else {"info": 0, "hard": 0, "iters": 0, "post": 0}
    sizes = {"info": ncw * KB, "hard": ncw * NB, "iters": ncw * 4}
    if want_post:
        sizes["post"] = ncw * N * 4
    whole, out = {}, {}
    for k, nb in sizes.items():
        whole[k], v = _guarded(torch, nb, skew[k])
        out[k] = {"info": lambda: v.view(ncw, KB), "hard": lambda: v.view(ncw, NB), "iters": lambda: v.view(torch.int32),
                  "post": lambda: v.view(torch.float32).view(ncw, N)}[k]()
    dec.decode_device(llr, want_hard=True, want_post=want_post, out=out)
    torch.cuda.synchronize()
    for k, nb in sizes.items():
        lo, hi = _bands_intact(whole[k], nb, skew[k])
        assert lo and hi, "%s: a store landed %s the buffer (%d words)" % (k, "below" if not lo else "above", ncw)
    assert bool(torch.isnan(yin[:gf]).all().item()) and bool(torch.isnan(yin[gf + ncw * N:]).all().item()), "input guard bands written"
    return {k: v.cpu().numpy() for k, v in out.items()}


def _same(res, ref, what, post=True):
    info, iters, hard, pst = ref
    assert np.array_equal(res["iters"], iters), what + ": iteration counts"
    assert np.array_equal(res["info"], info), what + ": info bytes"
    assert np.array_equal(res["hard"], np.packbits(hard, axis=1, bitorder="little")), what + ": hard bits"
    if post and pst is not None:
        assert np.array_equal(res["post"], pst), what + ": posteriors"


def _inputs(N, rate, seed):
    y = np.concatenate([awgn_llr(150, N, sigma_from_ebn0(2.0, rate), seed=seed), awgn_llr(150, N, sigma_from_ebn0(4.0, rate), seed=seed + 1),
                        awgn_llr(33, N, 1.2, seed=seed + 2)])
    return y   # 333 words


SIZES = (1, 37, 333)


@pytest.mark.parametrize("path,name", [(0, "lane_smem"), (1, "lane_global"), (3, "lane16"), (4, "group"), (5, "cluster"), (7, "qc"), (8, "warp")])
def test_guard_bands_default_code_every_path(default_code, path, name):
    import myldpccppapi_b200 as m
    c = default_code
    y = _inputs(c["N"], 0.75, 100 + path)
    ref = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40).decode(y)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    dec.set_path(path)
    assert dec.info()["path_name"] == name
    for n in SIZES:
        _same(_decode_guarded(dec, y[:n]), tuple(r[:n] for r in ref), "%s, %d words" % (name, n))
    _same(_decode_guarded(dec, y[:37], skewed=True), tuple(r[:37] for r in ref), "%s, buffers at odd offsets" % name)


@pytest.mark.parametrize("N", [576, 768])
def test_guard_bands_warp_per_codeword_kernel(N):
    """ldpc_ms_qcw_kernel (prefetches the next word's channel values into registers: the last word has no next)."""
    import myldpccppapi_b200 as m
    K = N * 3 // 4
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    y = _inputs(N, 0.75, 200 + N)
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
    dec = m.Decoder.wimax(K, N, 4)
    dec.reserve(4096)
    dec.set_option("qc_et", 1)
    for n in SIZES:
        _same(_decode_guarded(dec, y[:n]), tuple(r[:n] for r in ref), "qcw N=%d, %d words" % (N, n))
    _same(_decode_guarded(dec, y[:37], skewed=True), tuple(r[:37] for r in ref), "qcw N=%d, buffers at odd offsets" % N)
    assert dec.info()["kernel_variant"] == 1


@pytest.mark.parametrize("z,rate,name,num,den", [(36, 0, "1/2", 1, 2), (60, 4, "3/4B", 3, 4), (92, 5, "5/6", 5, 6), (44, 1, "2/3A", 2, 3)])
def test_guard_bands_group_of_warps_kernel(z, rate, name, num, den):
    """ldpc_ms_qcm_kernel, one and several codewords per group (z = 36 / 44 run three / two side by side by default)."""
    import myldpccppapi_b200 as m
    N = 24 * z
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    y = _inputs(N, num / den, 300 + z)[:120]
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
    dec = m.Decoder.wimax(K, N, rate)
    for n in (1, 37, 120):
        _same(_decode_guarded(dec, y[:n]), tuple(r[:n] for r in ref), "qcm z=%d, %d words" % (z, n))
    _same(_decode_guarded(dec, y[:37], skewed=True), tuple(r[:37] for r in ref), "qcm z=%d, buffers at odd offsets" % z)


def test_guard_bands_long_codes():
    """Group kernel (one codeword per CTA) and the streamed kernel on the regular (3,6) N = 8192 code."""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    y = awgn_llr(11, N, 0.84, seed=5)
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
    for path, name in [(4, "group"), (6, "stream"), (1, "lane_global")]:
        dec = m.Decoder(M, N, K, rp, ci)
        dec.set_path(path)
        assert dec.info()["path_name"] == name
        for n in (1, 11):
            _same(_decode_guarded(dec, y[:n]), tuple(r[:n] for r in ref), "reg36 %s, %d words" % (name, n))
        _same(_decode_guarded(dec, y[:3], skewed=True), tuple(r[:3] for r in ref), "reg36 %s, buffers at odd offsets" % name)


@pytest.mark.parametrize("N,big", [(576, False), (576, True), (1440, False), (1440, True)])
def test_guard_bands_sum_product_and_layered(N, big, monkeypatch):
    """DecodeSP (on-chip group kernel, quasi-cyclic kernel, any-size kernel) and DecodeTDMP (on-chip, any-size)."""
    import myldpccppapi_b200 as m
    K = N * 3 // 4
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    y = _inputs(N, 0.75, 400 + N)[:85]
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    if big:
        monkeypatch.setenv("LDPC_B200_SP_BIG", "1")
        monkeypatch.setenv("LDPC_B200_TDMP_G", "32")
    dec = m.Decoder.wimax(K, N, 4)
    dec.set_algorithm(1)
    sp = oracle.decode_sp(o, y)
    for n in (1, 37, 85):
        _same(_decode_guarded(dec, y[:n], want_post=False), (sp[0][:n], sp[1][:n], sp[2][:n], None), "sum-product N=%d big=%d, %d words" % (N, big, n), post=False)
    dec.set_algorithm(2)
    td = oracle.decode_tdmp(o, y, N // 24)
    for n in (1, 37, 85):
        _same(_decode_guarded(dec, y[:n]), tuple(r[:n] for r in td), "layered N=%d big=%d, %d words" % (N, big, n))
    _same(_decode_guarded(dec, y[:37], skewed=True), tuple(r[:37] for r in td), "layered N=%d big=%d, buffers at odd offsets" % (N, big))
    dec.set_algorithm(1)
    _same(_decode_guarded(dec, y[:37], want_post=False, skewed=True), (sp[0][:37], sp[1][:37], sp[2][:37], None),
          "sum-product N=%d big=%d, buffers at odd offsets" % (N, big), post=False)


# ---- host-buffer calls --------------------------------------------------------------------------------------------------
def _host_guarded(nbytes, skew, pinned):
    """(whole uint8 array, view of nbytes) in host memory; pinned = page-locked through torch."""
    total = GUARD + skew + nbytes + GUARD
    if pinned:
        import torch
        whole = torch.empty(total, dtype=torch.uint8).pin_memory().numpy()
    else:
        whole = np.empty(total, dtype=np.uint8)
    whole[:] = SENTINEL
    return whole, whole[GUARD + skew:GUARD + skew + nbytes]


def _host_intact(whole, nbytes, skew):
    return bool((whole[:GUARD + skew] == SENTINEL).all()) and bool((whole[GUARD + skew + nbytes:] == SENTINEL).all())


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("skewed", [False, True])
def test_guard_bands_host_buffers(default_code, pinned, skewed):
    """ldpc_b200_decode_host (persistent launch fed by the copy stream when pinned, staged chunks when pageable; tiny
    chunks; the warp-per-codeword kernel) and ldpc_b200_decode_host_packed (float16 / int8): the caller's host buffers sit
    between sentinel bands, at 16-byte alignment or at the weakest alignment of their element type."""
    import myldpccppapi_b200 as m
    c = default_code
    N, K = c["N"], c["K"]
    y = np.concatenate([_inputs(N, 0.75, 500), _inputs(N, 0.75, 510), _inputs(N, 0.75, 520)])   # 999 words
    ref = oracle.Oracle(c["M"], N, K, c["row_ptr"], c["col_idx"], times=40).decode(y, literal=False)
    dec = m.Decoder.wimax(K, N, c["rate"])
    KB, NB = dec.KB, dec.NB

    def run(n, fmt=np.float32, scale=1.0, what=""):
        esz = np.dtype(fmt).itemsize
        sk = {"llr": esz if skewed else 0, "info": 1 if skewed else 0, "hard": 1 if skewed else 0, "iters": 4 if skewed else 0,
              "post": 4 if skewed else 0}
        sizes = {"llr": n * N * esz, "info": n * KB, "hard": n * NB, "iters": n * 4, "post": n * N * 4}
        whole, view = {}, {}
        for k in sizes:
            whole[k], view[k] = _host_guarded(sizes[k], sk[k], pinned)
        src = view["llr"].view(fmt).reshape(n, N)
        src[:] = y[:n] if fmt == np.float32 else (np.clip(np.round(y[:n] / scale), -127, 127) if fmt == np.int8 else y[:n]).astype(fmt)
        out = {"info": view["info"].reshape(n, KB), "hard": view["hard"].reshape(n, NB), "iters": view["iters"].view(np.int32),
               "post": view["post"].view(np.float32).reshape(n, N)}
        before = whole["llr"].copy()
        if fmt == np.float32:
            dec.decode_host(src, want_hard=True, want_post=True, out=out)
            want = tuple(r[:n] for r in ref)
        else:
            dec.decode_host_packed(src, scale=scale, want_hard=True, want_post=True, out=out)
            wide = (src.astype(np.float32) * np.float32(scale)).astype(np.float32)   # the contract: the fp32 call on the widened floats
            want = oracle.Oracle(c["M"], N, K, c["row_ptr"], c["col_idx"], times=40).decode(wide, literal=False)
        for k in ("info", "hard", "iters", "post"):
            assert _host_intact(whole[k], sizes[k], sk[k]), "%s %s: a byte outside the buffer changed (%d words)" % (what, k, n)
        assert np.array_equal(whole["llr"], before), what + ": the caller's channel values were written"
        _same(out, want, "%s, %d words" % (what, n))

    for n in (1, 37, 999):
        run(n, what="host fp32")
    dec.set_option("stream_chunk", 8)
    run(333, what="host fp32, 8-word chunks")
    dec.set_option("stream_chunk", 0)
    dec.reserve(4096)
    dec.set_option("qc_et", 1)
    run(999, what="host fp32, warp-per-codeword kernel")
    run(37, fmt=np.float16, what="host fp16")
    run(333, fmt=np.int8, scale=1.0 / 32, what="host int8")
    dec.set_option("qc_et", -1)
    run(999, fmt=np.float16, what="host fp16, handle's own kernel choice")


@pytest.mark.parametrize("de_type", [1, 2, 3, 4, 5])   # DecodeMS, DecodeSP, DecodeTDMP, DecodeTDMPCL, DecodeMSCL (MyLdpc.h:37-39)
def test_guard_bands_coder_decode_ragged_src_length(de_type):
    """The drop-in Coder::decode with a stream that ends inside the last codeword: the reference's guard is
    `charOffset <= srcLength` and may touch srcCode[srcLength] (MyLdpc.cpp:765-774); the drop-in writes exactly srcLength
    bytes.  postCode is read only (getPostCodeLength floats, nothing behind them that matters)."""
    import myldpccppapi_b200 as m
    N, K, rate = 576, 432, 4
    ncw = 45
    y = np.ascontiguousarray(_inputs(N, 0.75, 600 + de_type)[:ncw])
    src_len = ncw * (K // 8) - 11
    whole, dst = _host_guarded(src_len, 1, False)
    gf = GUARD // 4
    yin = np.full(gf + ncw * N + gf, np.nan, dtype=np.float32)
    yin[gf:gf + ncw * N] = y.reshape(-1)
    coder = m.Coder(K, N, rate)
    coder.forDecoder(16)
    coder.addDecodeType(de_type)
    assert coder.decode(yin[gf:gf + ncw * N], dst, src_len, de_type) == 0
    assert _host_intact(whole, src_len, 1), "Coder::decode wrote outside srcCode[0, srcLength)"
    plain = np.zeros(src_len + 1, dtype=np.uint8)
    c2 = m.Coder(K, N, rate)
    c2.forDecoder(64)
    c2.addDecodeType(de_type)
    assert c2.decode(y.reshape(-1), plain, src_len, de_type) == 0
    assert np.array_equal(dst, plain[:src_len]) and plain[src_len] == 0
    assert np.array_equal(coder.lastIterations, c2.lastIterations)
