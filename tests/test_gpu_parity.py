"""GPU parity tests: the CUDA decoder, called through the C-ABI, against the CPU oracle on
the same seeded inputs.  Bar: hard decisions and iteration counts bit-exact; posterior within
1e-4 relative (and in fact identical)."""
import numpy as np
import pytest

import oracle
from tests.util import awgn_llr, assert_parity, sigma_from_ebn0

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _run_device(dec, llr_np, **kw):
    torch = _torch()
    d = torch.from_numpy(llr_np).cuda()
    out = dec.decode_device(d, want_hard=True, want_post=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("sigma", [0.4, 0.55, 0.62, 0.7, 1.0])
def test_default_code_parity(default_code, sigma):
    """BASELINE config 1: Test.cpp's code, 40-iteration cap, sigma from cliff to no-convergence."""
    import myldpccppapi_b200 as m
    c = default_code
    ncw = 1024 + 7  # ragged: not a multiple of 32
    llr = awgn_llr(ncw, c["N"], sigma, seed=int(sigma * 1000))
    ref = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    res = _run_device(dec, llr)
    assert_parity(res, ref, c["N"], what="sigma=%g" % sigma)
    host = dec.decode_host(llr, want_hard=True, want_post=True)
    assert_parity(host, ref, c["N"], what="host sigma=%g" % sigma)


@pytest.mark.parametrize("path,name", [(0, "lane_smem"), (1, "lane_global"), (3, "lane16"), (4, "group"), (5, "cluster"), (7, "qc"), (8, "warp")])
def test_default_code_every_kernel_path(default_code, path, name):
    """Every kernel family (shared-memory lane, global-workspace lane, lane16, group, cluster, quasi-cyclic) gives the oracle's bits."""
    import myldpccppapi_b200 as m
    c = default_code
    llr = np.concatenate([awgn_llr(150, c["N"], 0.62, seed=11), awgn_llr(150, c["N"], 0.52, seed=12),
                          awgn_llr(21, c["N"], 1.0, seed=13)])
    ref = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    dec.set_path(path)
    assert dec.info()["path_name"] == name
    assert_parity(_run_device(dec, llr), ref, c["N"], what=name)


def test_regular_3_6_every_kernel_path():
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    llr = awgn_llr(40, N, 0.84, seed=2)
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(llr, literal=False)
    for path, name in [(1, "lane_global"), (4, "group"), (5, "cluster"), (6, "stream"), (8, "warp")]:
        dec = m.Decoder(M, N, K, rp, ci)
        dec.set_path(path)
        assert dec.info()["path_name"] == name
        assert_parity(_run_device(dec, llr), ref, N, what="reg36 " + name)


@pytest.mark.parametrize("rate,name,num,den", [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3),
                                               (3, "3/4A", 3, 4), (5, "5/6", 5, 6)])
def test_other_wimax_rates(rate, name, num, den):
    import myldpccppapi_b200 as m
    N = 1152
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    sig = sigma_from_ebn0(2.5, num / den)
    llr = awgn_llr(200, N, sig, seed=rate + 100)
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(llr)
    dec = m.Decoder.wimax(K, N, rate)
    assert_parity(_run_device(dec, llr), ref, N, what="rate " + name)


@pytest.mark.parametrize("N,rate,name,num,den", [(2304, 0, "1/2", 1, 2), (2304, 4, "3/4B", 3, 4), (2304, 5, "5/6", 5, 6),
                                                 (576, 0, "1/2", 1, 2), (576, 2, "2/3B", 2, 3), (576, 5, "5/6", 5, 6)])
def test_largest_and_smallest_wimax_codes(N, rate, name, num, den):
    """The ends of the reference's code family (z = 96 and z = 24, MyLdpc.cpp:55): whichever kernel path the plan
    picks for them must give the oracle's bits, counts and posteriors."""
    import myldpccppapi_b200 as m
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    llr = np.concatenate([awgn_llr(48, N, sigma_from_ebn0(2.0, num / den), seed=7 * rate + N),
                          awgn_llr(48, N, sigma_from_ebn0(3.5, num / den), seed=7 * rate + N + 1)])
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(llr, literal=False)
    dec = m.Decoder.wimax(K, N, rate)
    assert_parity(_run_device(dec, llr), ref, N, what="N=%d rate %s (%s)" % (N, name, dec.info()["path_name"]))
    assert_parity(dec.decode_host(llr, want_hard=True, want_post=True), ref, N, what="host N=%d rate %s" % (N, name))


def test_ebn0_sweep_iteration_counts(default_code):
    """BASELINE config 4: Eb/N0 0..4 dB with syndrome early termination, iteration-count parity."""
    import myldpccppapi_b200 as m
    c = default_code
    o = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    rng = np.random.default_rng(4)
    Gp = m.codes.gf2_systematic_encoder(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"])
    for i, ebn0 in enumerate(np.arange(0.0, 4.01, 0.5)):
        u = rng.integers(0, 2, (256, c["K"])).astype(np.uint8)
        cwb = np.concatenate([u, (u.astype(np.int64) @ Gp.astype(np.int64) % 2).astype(np.uint8)], axis=1)
        assert m.codes.syndrome(c["M"], c["row_ptr"], c["col_idx"], cwb).sum() == 0
        llr = awgn_llr(256, c["N"], sigma_from_ebn0(ebn0, 0.75), seed=40 + i, bits=cwb)
        ref = o.decode(llr)
        assert_parity(_run_device(dec, llr), ref, c["N"], what="EbN0=%g" % ebn0)


def test_iteration_cap_and_no_early_termination(default_code):
    import myldpccppapi_b200 as m
    c = default_code
    llr = awgn_llr(128, c["N"], 0.6, seed=77)
    for cap in (1, 2, 7, 50):
        ref = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=cap).decode(llr)
        dec = m.Decoder.wimax(c["K"], c["N"], c["rate"], max_iter=cap)
        assert_parity(_run_device(dec, llr), ref, c["N"], what="cap=%d" % cap)
    # early termination off: every word runs to the cap; a noiseless word keeps its bits
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"], max_iter=5, early_termination=False)
    res = _run_device(dec, np.ones((40, c["N"]), dtype=np.float32))
    assert np.all(res["iters"] == 5) and not res["info"].any()


def test_edge_inputs(default_code):
    """Zeros, signed zeros, infinities, huge values (the 1000 clamp), and an empty batch."""
    import myldpccppapi_b200 as m
    torch = _torch()
    c = default_code
    N = c["N"]
    rng = np.random.default_rng(5)
    llr = awgn_llr(96, N, 0.7, seed=5)
    llr[0, :] = 0.0
    llr[1, :] = -0.0
    llr[2, ::3] = 0.0
    llr[3, ::5] = -0.0
    llr[4, :] = 5000.0 * np.sign(llr[4, :])          # above the 1000 clamp
    llr[5, ::7] = np.inf
    llr[6, ::11] = -np.inf
    llr[7, :] = np.float32(1e-40)                      # denormals
    llr[8, :] = rng.choice(np.array([-0.5, 0.5, 1.0, -1.0, 0.25], dtype=np.float32), N)  # exact ties / cancellations
    llr[9, :] = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), N)
    ref = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr)
    dec = m.Decoder.wimax(c["K"], N, c["rate"])
    res = _run_device(dec, llr)
    # compare posteriors with == semantics (signed zeros compare equal; inf-inf gives NaN on both sides)
    info, iters, hard, post = ref
    assert np.array_equal(res["iters"], iters)
    assert np.array_equal(res["info"], info)
    assert np.array_equal(res["hard"], np.packbits(hard, axis=1, bitorder="little"))
    both_nan = np.isnan(res["post"]) & np.isnan(post)
    assert np.array_equal(res["post"][~both_nan], post[~both_nan])
    # empty batch is a no-op
    out = dec.decode_device(torch.empty((0, N), dtype=torch.float32, device="cuda"))
    assert out["info"].shape[0] == 0


def test_regular_3_6_parity():
    """BASELINE config 3 code (regular (3,6), N=8192) on a seeded sample, incl. non-convergent words."""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    dec = m.Decoder(M, N, K, rp, ci)
    for sigma, seed in [(0.7, 1), (0.84, 2), (1.0, 3)]:
        llr = awgn_llr(48, N, sigma, seed=seed)
        ref = o.decode(llr, literal=False)
        assert_parity(_run_device(dec, llr), ref, N, what="reg36 sigma=%g" % sigma)


def test_ira_64800_parity():
    """BASELINE config 5 code (irregular, N=64800, 50-iteration cap) on a small seeded sample."""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.ira_code()
    o = oracle.Oracle(M, N, K, rp, ci, times=50)
    dec = m.Decoder(M, N, K, rp, ci, max_iter=50)
    rng = np.random.default_rng(9)
    u = rng.integers(0, 2, (34, K)).astype(np.uint8)
    cwb = m.codes.ira_encode(M, N, K, rp, ci, u)
    assert dec.info()["path_name"] == "stream"
    lg = m.Decoder(M, N, K, rp, ci, max_iter=50)
    lg.set_path(1)     # compressed check state in the global workspace (any degree)
    assert lg.info()["path_name"] == "lane_global"
    clus = m.Decoder(M, N, K, rp, ci, max_iter=50)
    clus.set_path(5)   # one codeword per 8-CTA cluster, state in distributed shared memory
    assert clus.info()["path_name"] == "cluster"
    for sigma, seed in [(0.8, 1), (0.97, 2)]:
        llr = awgn_llr(34, N, sigma, seed=seed, bits=cwb)
        ref = o.decode(llr, literal=False)
        assert_parity(_run_device(dec, llr), ref, N, what="ira stream sigma=%g" % sigma)
        assert_parity(_run_device(lg, llr), ref, N, what="ira lane_global sigma=%g" % sigma)
        assert_parity(_run_device(clus, llr), ref, N, what="ira cluster sigma=%g" % sigma)


def test_roundtrip_full_batch_properties(default_code):
    """BASELINE config 2 size (65,536 words): size-independent properties at full scale --
    encode -> BPSK+AWGN -> decode recovers the payload at high SNR, results do not depend on the
    position of a word in the batch, and a seeded 2,048-word slice equals the oracle."""
    import myldpccppapi_b200 as m
    torch = _torch()
    c = default_code
    N, K = c["N"], c["K"]
    ncw = 65536
    rng = np.random.default_rng(65536)
    Gp = m.codes.gf2_systematic_encoder(c["M"], N, K, c["row_ptr"], c["col_idx"])
    u = rng.integers(0, 2, (ncw, K)).astype(np.uint8)
    cwb = np.concatenate([u, ((u.astype(np.float32) @ Gp.astype(np.float32)) % 2).astype(np.uint8)], axis=1)
    bits = torch.from_numpy(m.codes.pack_bits(cwb)).cuda()
    llr = m.synth_llr(ncw, N, 0.45, seed=2, bits=bits)
    dec = m.Decoder.wimax(K, N, c["rate"])
    out = dec.decode_device(llr)
    info = out["info"].cpu().numpy()
    assert np.array_equal(info, m.codes.pack_bits(u)), "payload not recovered at sigma=0.45"
    assert int(out["iters"].max()) < 40
    # permutation invariance: decode the batch in reversed order
    rev = torch.flip(llr, dims=[0]).contiguous()
    out2 = dec.decode_device(rev)
    assert torch.equal(torch.flip(out2["info"], dims=[0]), out["info"])
    assert torch.equal(torch.flip(out2["iters"], dims=[0]), out["iters"])
    # oracle on a slice of the very same floats, at a noisier point
    llr2 = m.synth_llr(2048, N, 0.62, seed=3, bits=bits[:2048].contiguous())
    ref = oracle.Oracle(c["M"], N, K, c["row_ptr"], c["col_idx"], times=40).decode(llr2.cpu().numpy())
    o2 = dec.decode_device(llr2, want_hard=True, want_post=True)
    torch.cuda.synchronize()
    assert_parity({k: v.cpu().numpy() for k, v in o2.items()}, ref, N, what="slice")


def test_qc_runtime_profile_kernel(default_code, monkeypatch):
    """The quasi-cyclic kernel with a run-time profile (ldpc_qcg.cuh) on Test.cpp's code (forced: the compiled profile
    would take it) and on codes with padded slots; bits, counts and posteriors of the oracle, device and host paths."""
    import myldpccppapi_b200 as m
    c = default_code
    N = c["N"]
    monkeypatch.setenv("LDPC_B200_QC_GENERIC", "1")
    llr = np.concatenate([awgn_llr(300, N, 0.62, seed=51), awgn_llr(300, N, 0.5, seed=52), awgn_llr(37, N, 1.0, seed=53)])
    ref = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr, literal=False)
    dec = m.Decoder.wimax(c["K"], N, c["rate"])
    inf = dec.info()
    assert inf["path_name"] == "qc" and inf["codewords_per_cta"] == 8
    assert_parity(_run_device(dec, llr), ref, N, what="run-time profile")
    assert_parity(dec.decode_host(llr, want_hard=True, want_post=True), ref, N, what="run-time profile, host")
    # other sizes and rates: first with the run-time profile forced (padded slots, G = 4 and 2), then as planned
    # (compiled profiles for z = 24, 32, 40, 48, 64, 80, 96; run-time profile or the group kernel for the other z)
    cases = [(960, 3, "3/4A", 3, 4, 4), (1152, 2, "2/3B", 2, 3, 4), (2304, 0, "1/2", 1, 2, 2), (768, 5, "5/6", 5, 6, 4),
             (1536, 4, "3/4B", 3, 4, 2), (1920, 1, "2/3A", 2, 3, 2), (2304, 5, "5/6", 5, 6, 2),
             # block sizes whose groups do not divide evenly over the warps (empty slots): z = 72, 56, 36, 92
             (1728, 4, "3/4B", 3, 4, None), (1344, 0, "1/2", 1, 2, None), (864, 3, "3/4A", 3, 4, None), (2208, 2, "2/3B", 2, 3, None)]
    for forced in (True, False):
        if not forced:
            monkeypatch.delenv("LDPC_B200_QC_GENERIC")
        for NN, rate, name, num, den, g in cases[:4] if forced else cases:
            K = NN * num // den
            rp, ci, M = oracle.wimax_H(NN, name)
            y = awgn_llr(96, NN, sigma_from_ebn0(2.5, num / den), seed=NN + rate)
            r2 = oracle.Oracle(M, NN, K, rp, ci, times=40).decode(y, literal=False)
            d2 = m.Decoder.wimax(K, NN, rate)
            i2 = d2.info()
            if g is not None and not (forced and rate == 5):  # (check degree 20 is beyond the run-time profile: other kernels)
                assert i2["path_name"] == "qc" and i2["codewords_per_cta"] == g, i2
            assert_parity(_run_device(d2, y), r2, NN, what="N=%d rate %s forced=%s (%s)" % (NN, name, forced, i2["path_name"]))


@pytest.mark.parametrize("env", [{"LDPC_B200_QC_RING": "1"}, {"LDPC_B200_QC_PREFER_G": "4"}, {"LDPC_B200_QC_RING": "1", "LDPC_B200_QC_PREFER_G": "4"}])
def test_qc_alternative_kernels(default_code, monkeypatch, env):
    """The opt-in variants of the quasi-cyclic path (profiles/r02_ring_kernel.md): the ring kernel (codewords staged by
    bulk asynchronous copies behind mbarriers, refill off the loop top) and the half-size-CTA profile (24, 4, 6).  Bits,
    counts and posteriors of the oracle on the device path, the streamed host path and tiny batches."""
    import myldpccppapi_b200 as m
    torch = _torch()
    c = default_code
    N = c["N"]
    for k, v in env.items():
        monkeypatch.setenv(k, v)   # plan-time switches: read once, at create
    llr = np.concatenate([awgn_llr(900, N, 0.62, seed=61), awgn_llr(1200, N, 0.5, seed=62), awgn_llr(301, N, 1.0, seed=63)])
    ref = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr, literal=False)
    dec = m.Decoder.wimax(c["K"], N, c["rate"])
    inf = dec.info()
    assert inf["path_name"] == "qc" and inf["codewords_per_cta"] == (4 if "LDPC_B200_QC_PREFER_G" in env else 8)
    assert_parity(_run_device(dec, llr), ref, N, what="device")
    pinned = torch.from_numpy(llr).pin_memory().numpy()
    assert_parity(dec.decode_host(pinned, want_hard=True, want_post=True), ref, N, what="streamed, pinned")
    dec.set_option("stream_chunk", 8)
    assert_parity(dec.decode_host(pinned[:700], want_hard=True, want_post=True), tuple(r[:700] for r in ref), N, what="streamed, tiny chunks")
    dec.set_option("stream_chunk", 0)
    for n in (1, 5, 9, 2401):
        assert_parity(_run_device(dec, llr[:n]), tuple(r[:n] for r in ref), N, what="%d words" % n)
    dec.set_early_termination(False)
    dec.set_max_iter(7)
    ref7 = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=7).decode(llr[:500], literal=False)
    out = _run_device(dec, llr[:500])
    assert np.array_equal(out["iters"], np.full(500, 7))
    # without early termination every word runs the cap: bits are those of the capped oracle for the words it did not stop early
    late = ref7[1] == 7
    assert np.array_equal(out["info"][late], ref7[0][late])


@pytest.mark.parametrize("rate,name,num,den", [(4, "3/4B", 3, 4), (0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (5, "5/6", 5, 6)])
def test_qc_group_of_warps_kernel(rate, name, num, den, monkeypatch):
    """ldpc_ms_qcm_kernel (a codeword per group of ceil(z / 32) warps, run-time tables): the twelve block sizes of the
    reference's family without a compiled lockstep profile (z = 28 ... 92) as planned, and -- forced -- one, two and
    three warps per codeword at the profiled sizes z = 24, 48, 96.  Bits, counts and posteriors of the oracle on the
    device path, the streamed host path, tiny batches, a short cap; special values."""
    import myldpccppapi_b200 as m
    torch = _torch()
    for z, forced in [(28, False), (36, False), (44, False), (52, False), (56, False), (60, False), (68, False), (72, False),
                      (76, False), (84, False), (88, False), (92, False), (24, True), (48, True), (96, True)]:
        N = 24 * z
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        if forced:
            monkeypatch.setenv("LDPC_B200_QCM_ALWAYS", "1")
        else:
            monkeypatch.delenv("LDPC_B200_QCM_ALWAYS", raising=False)
        nw = 70 if z in (36, 60, 96, 24) else 24
        y = np.concatenate([awgn_llr(nw, N, sigma_from_ebn0(2.2, num / den), seed=N + rate), awgn_llr(nw, N, sigma_from_ebn0(4.0, num / den), seed=N + rate + 1),
                            awgn_llr(5, N, 1.3, seed=N + rate + 2)])
        y[3] = 0.0
        y[4, ::3] = 0.0
        y[5] = np.where(np.arange(N) % 2 == 0, -0.0, 0.0)
        y[6, :7] = [np.inf, -np.inf, 1e30, -1e30, 1e-40, -1e-40, 1000.0]
        ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
        dec = m.Decoder.wimax(K, N, rate)
        inf = dec.info()
        assert inf["path_name"] == "qc" and inf["threads_per_cta"] == 32 * inf["codewords_per_cta"] * ((z + 31) // 32), inf
        what = "z=%d rate %s" % (z, name)
        assert_parity(_run_device(dec, y), ref, N, what=what)
        if z in (36, 60, 96, 24):
            assert_parity(dec.decode_host(y, want_hard=True, want_post=True), ref, N, what=what + " host")
            pinned = torch.from_numpy(y).pin_memory().numpy()
            dec.set_option("stream_chunk", 8)
            assert_parity(dec.decode_host(pinned, want_hard=False, want_post=False), ref, N, what=what + " streamed, tiny chunks")
            dec.set_option("stream_chunk", 0)
            for n in (1, 3, 17):
                assert_parity(_run_device(dec, y[:n]), tuple(r[:n] for r in ref), N, what=what + " %d words" % n)
            dec.set_max_iter(3)
            rc = oracle.Oracle(M, N, K, rp, ci, times=3).decode(y, literal=False)
            assert_parity(_run_device(dec, y), rc, N, what=what + " cap 3")


@pytest.mark.parametrize("pack", [2, 3])
@pytest.mark.parametrize("rate,name,num,den", [(4, "3/4B", 3, 4), (0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (5, "5/6", 5, 6)])
def test_qc_group_of_warps_kernel_several_codewords_per_group(rate, name, num, den, pack, monkeypatch):
    """ldpc_ms_qcm_multi_kernel: two / three codewords side by side on one group of warps (forced for every launch).
    Words of one group stop at different iterations (mixed Eb/N0), the last group of a batch is short, special values:
    bits, counts and posteriors of the oracle on the device path and the streamed host path."""
    import myldpccppapi_b200 as m
    torch = _torch()
    monkeypatch.setenv("LDPC_B200_QCM_PACK", str(pack))
    for z in (36, 44, 68, 76, 92) if pack == 2 else (36, 28):
        N = 24 * z
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        rng = np.random.default_rng(N + rate)
        y = np.concatenate([awgn_llr(40, N, sigma_from_ebn0(2.4, num / den), seed=N + rate), awgn_llr(40, N, sigma_from_ebn0(4.0, num / den), seed=N + rate + 1),
                            awgn_llr(7, N, 1.3, seed=N + rate + 2)])
        y = y[rng.permutation(len(y))]   # fast and slow words inside one group
        y[3] = 0.0
        y[4, ::3] = 0.0
        y[5] = np.where(np.arange(N) % 2 == 0, -0.0, 0.0)
        y[6, :7] = [np.inf, -np.inf, 1e30, -1e30, 1e-40, -1e-40, 1000.0]
        ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
        dec = m.Decoder.wimax(K, N, rate)
        dec.set_option("qcm_multi_pct", 0)
        what = "z=%d rate %s pack %d" % (z, name, pack)
        assert_parity(_run_device(dec, y), ref, N, what=what)
        assert dec.info()["kernel_variant"] == 4, dec.info()
        for n in (1, 2, 3, 4, 17):
            assert_parity(_run_device(dec, y[:n]), tuple(r[:n] for r in ref), N, what=what + " %d words" % n)
        if z == 36:
            assert_parity(dec.decode_host(y, want_hard=True, want_post=True), ref, N, what=what + " host")
            pinned = torch.from_numpy(y).pin_memory().numpy()
            dec.set_option("stream_chunk", 8)
            assert_parity(dec.decode_host(pinned, want_hard=False, want_post=False), ref, N, what=what + " streamed, tiny chunks")
            dec.set_option("stream_chunk", 0)
            assert dec.info()["kernel_variant"] == 4
            dec.set_max_iter(3)
            rc = oracle.Oracle(M, N, K, rp, ci, times=3).decode(y, literal=False)
            assert_parity(_run_device(dec, y), rc, N, what=what + " cap 3")
            dec.set_max_iter(5)
            dec.set_early_termination(False)   # every word runs to the cap; a noiseless word keeps its bits
            res = _run_device(dec, np.ones((11, N), dtype=np.float32))
            assert np.all(res["iters"] == 5) and not res["info"].any()


@pytest.mark.parametrize("N", [576, 768])
@pytest.mark.parametrize("rate,name,num,den", [(4, "3/4B", 3, 4), (0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (5, "5/6", 5, 6)])
def test_qc_early_termination_kernel(rate, name, num, den, N):
    """ldpc_ms_qcw_kernel (a warp per codeword, syndrome from packed hard bits, no message clearing): bits, iteration
    counts and posteriors of the oracle in every regime -- forced on (option qc_et = 1) over inputs from clean to
    hopeless, on the device path, the streamed host path, tiny batches, caps 1 / 2 / 7 and the special values; and chosen
    by the handle itself (qc_et = -1) after a launch whose words stopped early, dropped again after one whose words did not."""
    import myldpccppapi_b200 as m
    torch = _torch()
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    s0 = {0: 0.85, 1: 0.68, 2: 0.68, 3: 0.6, 4: 0.6, 5: 0.52}[rate]   # around the waterfall of each rate
    llr = np.concatenate([awgn_llr(900, N, s0, seed=71 + rate), awgn_llr(1500, N, s0 * 0.85, seed=72 + rate),
                          awgn_llr(301, N, 1.2, seed=73 + rate), awgn_llr(200, N, s0 * 0.6, seed=74 + rate)])
    llr[5] = 0.0                                  # erased word: every posterior stays 0, bits all 1
    llr[6, ::3] = 0.0
    llr[7] = np.where(np.arange(N) % 2 == 0, -0.0, 0.0)
    llr[8] = np.abs(llr[8]) + 0.1                 # all-zero codeword received cleanly: stops at iteration 1
    llr[9, :7] = [np.inf, -np.inf, 1e30, -1e30, 1e-40, -1e-40, 1000.0]
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(llr, literal=False)
    dec = m.Decoder.wimax(K, N, rate)
    assert dec.info()["path_name"] == "qc"
    dec.reserve(4096)
    assert dec.info()["et_available"] == 1
    dec.set_option("qc_et", 1)
    assert_parity(_run_device(dec, llr), ref, N, what="rate %s device" % name)
    assert dec.info()["kernel_variant"] == 1
    pinned = torch.from_numpy(llr).pin_memory().numpy()
    assert_parity(dec.decode_host(pinned, want_hard=True, want_post=True), ref, N, what="streamed, pinned")
    assert_parity(dec.decode_host(llr, want_hard=False, want_post=False), ref, N, what="host, pageable, info only")
    dec.set_option("stream_chunk", 8)
    assert_parity(dec.decode_host(pinned[:700], want_hard=True, want_post=True), tuple(r[:700] for r in ref), N, what="streamed, tiny chunks")
    dec.set_option("stream_chunk", 0)
    for n in (1, 3, 5, 9, 2401):
        assert_parity(_run_device(dec, llr[:n]), tuple(r[:n] for r in ref), N, what="%d words" % n)
    for cap in (1, 2, 7):
        dec.set_max_iter(cap)
        rc = oracle.Oracle(M, N, K, rp, ci, times=cap).decode(llr[:1200], literal=False)
        assert_parity(_run_device(dec, llr[:1200]), rc, N, what="cap %d" % cap)
    dec.set_max_iter(40)
    # the handle's own choice: the lockstep kernel while the regime is unknown, afterwards the warp-per-codeword kernel
    # whenever the previous launch's words stopped early on average (z = 24: mean <= 97 % of the cap; z = 32: always,
    # it is the faster kernel in every regime there)
    dec.set_option("qc_et", -1)
    dec.set_option("qc_et_every", 1)   # (default: the counts are sampled after every fourth launch)
    easy, hard = llr[900:2400], llr[2400:2701]
    seen = []
    for part, lo, hi in ((hard, 2400, 2701), (easy, 900, 2400), (easy, 900, 2400), (hard, 2400, 2701), (hard, 2400, 2701)):
        out = _run_device(dec, part)
        assert_parity(out, tuple(r[lo:hi] for r in ref), N, what="auto")
        seen.append(dec.info()["kernel_variant"])
    assert seen == ([0, 0, 1, 1, 0] if N == 576 else [1, 1, 1, 1, 1]), seen


def test_qc_lockstep_hands_over_to_the_group_of_warps_kernel():
    """A block size with a compiled lockstep profile but no warp-per-codeword kernel (z = 48): after a launch whose words
    stopped early the handle decodes with ldpc_ms_qcm_kernel (kernel_variant 3), and goes back to the lockstep kernel
    after one whose words ran long; the oracle's bits, counts and posteriors either way."""
    import myldpccppapi_b200 as m
    N, K, rate = 1152, 864, 4
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    easy = awgn_llr(600, N, sigma_from_ebn0(4.2, 0.75), seed=91)
    hard = awgn_llr(200, N, 1.1, seed=92)
    orc = oracle.Oracle(M, N, K, rp, ci, times=40)
    ref_e, ref_h = orc.decode(easy, literal=False), orc.decode(hard, literal=False)
    dec = m.Decoder.wimax(K, N, rate)
    dec.reserve(1024)
    assert dec.info()["et_available"] == 2
    dec.set_option("qc_et_every", 1)
    seen = []
    for part, ref in ((hard, ref_h), (easy, ref_e), (easy, ref_e), (hard, ref_h), (hard, ref_h)):
        assert_parity(_run_device(dec, part), ref, N, what="auto")
        seen.append(dec.info()["kernel_variant"])
    assert seen == [0, 0, 3, 3, 0], seen
    dec.set_option("qc_et", 1)
    assert_parity(dec.decode_host(easy, want_hard=True, want_post=True), ref_e, N, what="forced, host")
    assert dec.info()["kernel_variant"] == 3


@pytest.mark.parametrize("N,rate,name,num,den", [(1152, 4, "3/4B", 3, 4), (1632, 0, "1/2", 1, 2), (2304, 5, "5/6", 5, 6), (1824, 1, "2/3A", 2, 3)])
def test_any_size_sum_product_and_layered_kernels(N, rate, name, num, den, monkeypatch):
    """The reference's kernels have no size limit (decodeCL.c:25-62, 203-292).  Codes beyond the on-chip layouts --
    most of the reference's own family -- run the quasi-cyclic sum-product kernel (ldpc_spq.cuh: a group of warps per
    codeword) and, forced, ldpc_sp_big_kernel (messages in a global workspace); layered: ldpc_tdmp_big_kernel where the
    on-chip layout does not fit.  Bits and iteration counts of the sum-product oracle; bits, counts and posteriors of the
    layered oracle."""
    import myldpccppapi_b200 as m
    torch = _torch()
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    y = np.concatenate([awgn_llr(20, N, sigma_from_ebn0(e, num / den), seed=N + i) for i, e in enumerate((1.5, 2.5, 4.0))])
    y = y[:53]  # a ragged last group of 32
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    dec = m.Decoder.wimax(K, N, rate)
    dec.set_algorithm(1)
    out = dec.decode_device(torch.from_numpy(y).cuda(), want_hard=True)
    torch.cuda.synchronize()
    sp = oracle.decode_sp(o, y)
    assert np.array_equal(out["iters"].cpu().numpy(), sp[1]) and np.array_equal(out["info"].cpu().numpy(), sp[0])
    assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(sp[2], axis=1, bitorder="little"))
    assert dec.info()["kernel_variant"] == 5   # the quasi-cyclic sum-product kernel (ldpc_spq.cuh)
    host = dec.decode_host(y)
    assert np.array_equal(host["iters"], sp[1]) and np.array_equal(host["info"], sp[0])
    for n in (1, 2, 17):
        part = dec.decode_device(torch.from_numpy(y[:n]).cuda(), want_hard=True)
        assert np.array_equal(part["iters"].cpu().numpy(), sp[1][:n]) and np.array_equal(part["info"].cpu().numpy(), sp[0][:n])
    dec.set_max_iter(3)
    sp3 = oracle.decode_sp(oracle.Oracle(M, N, K, rp, ci, times=3), y)
    out3 = dec.decode_device(torch.from_numpy(y).cuda(), want_hard=True)
    assert np.array_equal(out3["iters"].cpu().numpy(), sp3[1]) and np.array_equal(out3["info"].cpu().numpy(), sp3[0])
    dec.set_max_iter(40)
    monkeypatch.setenv("LDPC_B200_SP_BIG", "1")   # the any-size kernel (ldpc_big.cuh): same bytes
    big = m.Decoder.wimax(K, N, rate)
    monkeypatch.delenv("LDPC_B200_SP_BIG")
    big.set_algorithm(1)
    outb = big.decode_device(torch.from_numpy(y).cuda(), want_hard=True)
    assert big.info()["kernel_variant"] != 5
    assert np.array_equal(outb["iters"].cpu().numpy(), sp[1]) and np.array_equal(outb["info"].cpu().numpy(), sp[0])
    dec.set_algorithm(2)
    td = oracle.decode_tdmp(o, y, N // 24)
    assert_parity(_run_device(dec, y), td, N, what="layered, N=%d" % N)
    dec.set_max_iter(3)
    td3 = oracle.decode_tdmp(oracle.Oracle(M, N, K, rp, ci, times=3), y, N // 24)
    assert_parity(_run_device(dec, y), td3, N, what="layered cap 3, N=%d" % N)


def test_any_size_kernels_equal_the_on_chip_ones(default_code, monkeypatch):
    """Test.cpp's code through the any-size kernels (forced) gives what the on-chip kernels give: one contract, two layouts."""
    import myldpccppapi_b200 as m
    c = default_code
    N = c["N"]
    y = np.concatenate([awgn_llr(100, N, 0.62, seed=71), awgn_llr(100, N, 0.5, seed=72), awgn_llr(41, N, 0.9, seed=73)])
    o = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=40)
    monkeypatch.setenv("LDPC_B200_SP_BIG", "1")
    monkeypatch.setenv("LDPC_B200_TDMP_G", "32")
    dec = m.Decoder.wimax(c["K"], N, c["rate"])
    dec.set_algorithm(1)
    sp = oracle.decode_sp(o, y)
    out = _run_device_nopost(dec, y)
    assert np.array_equal(out["iters"], sp[1]) and np.array_equal(out["info"], sp[0])
    dec.set_algorithm(2)
    assert_parity(_run_device(dec, y), oracle.decode_tdmp(o, y, 24), N, what="layered, any-size kernel")
    # long synthetic codes: the regular (3,6) code has no layer structure, sum-product runs
    M2, N2, K2, rp2, ci2 = m.codes.regular_code()
    y2 = awgn_llr(5, N2, 0.75, seed=9)
    d2 = m.Decoder(M2, N2, K2, rp2, ci2)
    d2.set_algorithm(1)
    sp2 = oracle.decode_sp(oracle.Oracle(M2, N2, K2, rp2, ci2, times=40), y2)
    o2 = _run_device_nopost(d2, y2)
    assert np.array_equal(o2["iters"], sp2[1]) and np.array_equal(o2["info"], sp2[0])


def _run_device_nopost(dec, llr_np):
    torch = _torch()
    out = dec.decode_device(torch.from_numpy(llr_np).cuda(), want_hard=True)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def test_streamed_host_pipeline(default_code):
    """Host buffers on the quasi-cyclic path: one persistent launch fed by a copy stream.  Ragged sizes, pageable and
    pinned inputs, optional outputs asked for only on a later call, several launches per call (small batch cap) and
    the 3-stream pipeline (option no_streamed) must all give the oracle's bytes, counts and posteriors."""
    import myldpccppapi_b200 as m
    torch = _torch()
    c = default_code
    N = c["N"]
    ncw = 3001
    llr = np.concatenate([awgn_llr(1500, N, 0.6, seed=41), awgn_llr(1501, N, 0.5, seed=42)])
    ref = oracle.Oracle(c["M"], N, c["K"], c["row_ptr"], c["col_idx"], times=40).decode(llr, literal=False)
    dec = m.Decoder.wimax(c["K"], N, c["rate"])
    assert dec.info()["path_name"] == "qc"
    first = dec.decode_host(llr)                                    # pageable input: the chunked 3-stream pipeline
    assert np.array_equal(first["info"], ref[0]) and np.array_equal(first["iters"], ref[1])
    dec.set_option("staged_min_kb", 0)              # ... staged by host threads through pinned buffers
    assert_parity(dec.decode_host(llr, want_hard=True, want_post=True), ref, N, what="staged pageable input")
    dec.set_option("stream_chunk", 8)               # 376 chunks through the ring of 8 staging buffers
    assert_parity(dec.decode_host(llr, want_hard=True, want_post=True), ref, N, what="staged, tiny chunks")
    dec.set_option("stream_batch_kb", 1024)         # several launches
    assert_parity(dec.decode_host(llr[:1000], want_hard=True, want_post=True), tuple(r[:1000] for r in ref), N, what="staged, batches")
    dec.set_option("stream_chunk", 0); dec.set_option("stream_batch_kb", 0); dec.set_option("staged_min_kb", 8192)
    dec.set_option("streamed_pageable", 1)          # ... and the persistent launch fed from pageable memory
    first = dec.decode_host(llr)                                    # info + iters only
    assert np.array_equal(first["info"], ref[0]) and np.array_equal(first["iters"], ref[1])
    assert_parity(dec.decode_host(llr, want_hard=True, want_post=True), ref, N, what="streamed, all outputs")
    dec.set_option("streamed_pageable", 0)
    pinned = torch.from_numpy(llr).pin_memory().numpy()
    assert_parity(dec.decode_host(pinned, want_hard=True, want_post=True), ref, N, what="streamed, pinned")
    for n in (1, 9):                                                # fewer words than one CTA holds
        assert_parity(dec.decode_host(pinned[:n], want_hard=True, want_post=True), tuple(r[:n] for r in ref), N, what="streamed, %d words" % n)
    dec.set_option("stream_batch_kb", 1024)         # 455 words per launch: 7 launches
    assert_parity(dec.decode_host(pinned, want_hard=True, want_post=True), ref, N, what="streamed, 7 launches")
    dec.set_option("stream_chunk", 8)               # one word group per copy
    assert_parity(dec.decode_host(pinned[:500], want_hard=True, want_post=True), tuple(r[:500] for r in ref), N, what="streamed, tiny chunks")
    dec.set_option("stream_batch_kb", 0); dec.set_option("stream_chunk", 0)
    dec.set_option("no_streamed", 1)
    assert_parity(dec.decode_host(pinned, want_hard=True, want_post=True), ref, N, what="3-stream pipeline")
    assert ncw == llr.shape[0]


def test_coder_api_roundtrip(default_code):
    """Test.cpp-shaped run through the Coder mirror: payload -> encode -> test() channel -> decode."""
    import myldpccppapi_b200 as m
    c = default_code
    K, N = c["K"], c["N"]
    srcLength = 1000  # not a multiple of K/8 = 54: the last codeword is partly padding
    src = np.array([ord("a") + i % 26 for i in range(srcLength)], dtype=np.uint8)  # Test.cpp:43-45
    coder = m.Coder(K, N, m.rate_3_4_b)
    coder.forDecoder(8)
    coder.addDecodeType(m.DecodeMS)
    ncw = coder.getCodeSize(srcLength)
    assert ncw == 19 and coder.getPostCodeLength(srcLength) == ncw * N
    padded = np.zeros(ncw * K // 8, dtype=np.uint8)
    padded[:srcLength] = src
    u = m.codes.unpack_bits(padded.reshape(ncw, K // 8), K)
    Gp = m.codes.gf2_systematic_encoder(c["M"], N, K, c["row_ptr"], c["col_idx"])
    cwb = np.concatenate([u, (u.astype(np.int64) @ Gp.astype(np.int64) % 2).astype(np.uint8)], axis=1)
    post = awgn_llr(ncw, N, 10 ** (-6.0 / 20), seed=6, bits=cwb).reshape(-1)   # snr = 6 dB, Test.cpp:56-59
    new_src = np.zeros(srcLength + 1, dtype=np.uint8)
    assert coder.decode(post, new_src, srcLength, m.DecodeMS) == 0
    assert np.array_equal(new_src[:srcLength], src)                           # ErrNum == 0, Test.cpp:105-110
    ref_bytes, ref_iters, _, _ = oracle.Oracle(c["M"], N, K, c["row_ptr"], c["col_idx"]).decode_stream(post, srcLength)
    assert np.array_equal(new_src[:srcLength], ref_bytes)
    assert np.array_equal(coder.lastIterations, ref_iters)


def test_coder_decode_sp_uses_the_sum_product_kernel(default_code):
    """Coder::decode(..., DecodeSP) (MyLdpc.cpp:571-618 -> decodeOnceSP) runs the sum-product kernel; the same
    stream decoded with DecodeCPU runs min-sum.  At a noise level where the two disagree on iteration counts the
    Coder must reproduce each oracle."""
    import myldpccppapi_b200 as m
    c = default_code
    K, N = c["K"], c["N"]
    ncw = 64
    srcLength = ncw * K // 8
    post = awgn_llr(ncw, N, 0.62, seed=21).reshape(-1)
    o = oracle.Oracle(c["M"], N, K, c["row_ptr"], c["col_idx"])
    sp_info, sp_iters, _, _, _ = oracle.decode_sp(o, post.reshape(ncw, N))
    ms_info, ms_iters, _, _ = o.decode(post.reshape(ncw, N))
    assert not np.array_equal(sp_iters, ms_iters)
    coder = m.Coder(K, N, m.rate_3_4_b)
    coder.forDecoder(ncw)
    coder.addDecodeType(m.DecodeSP)
    out = np.zeros(srcLength + 1, dtype=np.uint8)
    assert coder.decode(post, out, srcLength, m.DecodeSP) == 0
    assert np.array_equal(coder.lastIterations, sp_iters)
    assert np.array_equal(out[:srcLength], sp_info.reshape(-1))
    assert coder.decode(post, out, srcLength, m.DecodeCPU) == 0
    assert np.array_equal(coder.lastIterations, ms_iters)
    assert np.array_equal(out[:srcLength], ms_info.reshape(-1))
    # DecodeTDMP / DecodeTDMPCL -> the layered kernel
    td_info, td_iters, _, _ = oracle.decode_tdmp(o, post.reshape(ncw, N), N // 24)
    assert not np.array_equal(td_iters, ms_iters)
    for t in (m.DecodeTDMP, m.DecodeTDMPCL):
        assert coder.decode(post, out, srcLength, t) == 0
        assert np.array_equal(coder.lastIterations, td_iters)
        assert np.array_equal(out[:srcLength], td_info.reshape(-1))


@pytest.mark.parametrize("N,rate,name,num,den", [(576, 4, "3/4B", 3, 4), (1440, 0, "1/2", 1, 2)])
def test_packed_channel_values(N, rate, name, num, den):
    """ldpc_b200_decode_host_packed: float16 / int8 channel values widened on the device give exactly what the fp32 call
    gives on the widened floats ((float)x * scale, one fp32 multiplication) -- bits, counts and posteriors of the oracle;
    pinned and pageable buffers, many chunks, a ragged tail."""
    import myldpccppapi_b200 as m
    torch = _torch()
    K = N * num // den
    rp, ci, M = oracle.wimax_H(N, name)
    ncw = 3001
    y = np.concatenate([awgn_llr(2000, N, sigma_from_ebn0(3.0, num / den), seed=N), awgn_llr(1001, N, sigma_from_ebn0(1.0, num / den), seed=N + 1)])
    orc = oracle.Oracle(M, N, K, rp, ci, times=40)
    dec = m.Decoder.wimax(K, N, rate)
    h = y.astype(np.float16)
    for scale in (1.0, 0.37):
        wide = (h.astype(np.float32) * np.float32(scale)).astype(np.float32)
        ref = orc.decode(wide, literal=False)
        assert_parity(dec.decode_host_packed(h, scale=scale, want_hard=True, want_post=True), ref, N, what="float16 pageable scale %g" % scale)
    ref = orc.decode(h.astype(np.float32), literal=False)
    pinned = torch.from_numpy(h).pin_memory()
    assert_parity(dec.decode_host_packed(pinned, want_hard=True, want_post=True), ref, N, what="float16 pinned")
    assert_parity(dec.decode_host(h.astype(np.float32), want_hard=True, want_post=True), ref, N, what="fp32 call on the widened values")
    q = np.clip(np.rint(y * 8.0), -127, 127).astype(np.int8)         # 4-bit fraction, as a receiver front end would deliver
    wide = (q.astype(np.float32) * np.float32(0.125)).astype(np.float32)
    ref = orc.decode(wide, literal=False)
    assert_parity(dec.decode_host_packed(q, scale=0.125, want_hard=True, want_post=True), ref, N, what="int8")
    assert_parity(dec.decode_host_packed(torch.from_numpy(q).pin_memory(), scale=0.125), ref, N, what="int8 pinned, info only")
    with pytest.raises(ValueError):
        dec.decode_host_packed(y)                                     # fp32 belongs to decode_host


def _run_cli(name, *args):
    import pathlib
    import subprocess
    exe = pathlib.Path(__file__).resolve().parents[1] / "myldpccppapi_b200" / "bin" / name
    if not exe.exists():
        pytest.skip(name + " not built")
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return dict(line.split("=", 1) for line in r.stdout.splitlines() if "=" in line), r.stdout


def test_cpp_coder_cli_roundtrip():
    """C++ drop-in `Coder` (include/MyLdpc.h, libmyldpc_b200.so) driven by our Test.cpp-shaped CLI."""
    kv, out = _run_cli("mytest", 54000, 256, 6.5, "MS", 7)
    assert kv["ErrNum"] == "0", out
    assert 1.0 <= float(kv["MeanIterations"]) < 10.0
    kv, out = _run_cli("mytest", 1000, 8, 0, "CPU", 7)   # sigma = 1: nothing converges, cap reached
    assert float(kv["MeanIterations"]) == 40.0 and int(kv["ErrNum"]) > 0
    # DecodeSP reaches the sum-product kernel through the C++ Coder: same seed, different iteration profile
    kv_ms, _ = _run_cli("mytest", 54000, 256, 5.0, "MS", 11)
    kv_sp, _ = _run_cli("mytest", 54000, 256, 5.0, "SP", 11)
    assert float(kv_sp["MeanIterations"]) != float(kv_ms["MeanIterations"])
    assert float(kv_sp["MeanIterations"]) < float(kv_ms["MeanIterations"])
    kv_td, _ = _run_cli("mytest", 54000, 256, 5.0, "TDMP", 11)
    assert float(kv_td["MeanIterations"]) < float(kv_ms["MeanIterations"])
    assert int(kv_td["ErrNum"]) <= int(kv_ms["ErrNum"])


def test_first_decode_of_a_fresh_process_is_not_slow():
    """CUDA loads a kernel at its first launch; a kernel first launched behind the spinning persistent launch of the
    host-buffer path made the first Coder::decode of a process sit out the kernel's 4 s input wait.  addDecodeType (and,
    for handles nobody reserved on, the first decode) launches every kernel once: in a FRESH process the first decode of
    32,768 words from malloc'd memory -- and the first C-ABI call on a bare handle -- stay far below a second."""
    import pathlib
    import subprocess
    import sys
    root = pathlib.Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "tools" / "setdevices_first_call.py"), "32768", "coder", "cabi"], capture_output=True, text=True,
                       timeout=300, cwd=str(root), env=dict(__import__("os").environ, PYTHONPATH=str(root)))
    assert r.returncode == 0, r.stdout + r.stderr
    firsts = [float(line.split("first call")[1].split("ms")[0]) for line in r.stdout.splitlines() if "first call" in line and "phases" not in line]
    assert len(firsts) == 2, r.stdout
    assert max(firsts) < 500.0, r.stdout


def test_reference_test_cpp_runs_against_the_drop_in():
    """The reference's OWN Test.cpp, compiled unmodified against include/MyLdpc.h at build time
    (myldpccppapi_b200/_build.py: build_harness), decodes its payload without byte errors."""
    for alg in ("MS", "SP", "CPU", "TDMP"):
        kv, out = _run_cli("MyTest_reference_harness", 5400, 64, 7, alg)
        assert kv["ErrNum"] == "0", out


@pytest.mark.parametrize("sigma", [0.45, 0.55, 0.6, 0.66, 0.9])
def test_sum_product_matches_restated_oracle(default_code, sigma):
    """DecodeSP: the probability-domain sum-product kernel against the oracle's restatement of the
    reference's OpenCL kernels (decodeCL.c:3-108 / decodeOnceSP).  Both sides take every product in the
    reference's list order with IEEE fp32 operations and share one exp routine, so hard decisions and
    iteration counts must be identical.  (No CPU sum-product exists in the reference: this parity is
    GPU == restated oracle.)"""
    import myldpccppapi_b200 as m
    c = default_code
    llr = awgn_llr(300 + 5, c["N"], sigma, seed=int(sigma * 1000) + 1)
    llr[0, :] = 0.0
    llr[1, ::4] = -0.0
    o = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40)
    info, iters, hard, p0, p1 = oracle.decode_sp(o, llr)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    dec.set_algorithm(1)
    torch = _torch()
    out = dec.decode_device(torch.from_numpy(llr).cuda(), want_hard=True)
    torch.cuda.synchronize()
    assert np.array_equal(out["iters"].cpu().numpy(), iters)
    assert np.array_equal(out["info"].cpu().numpy(), info)
    assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(hard, axis=1, bitorder="little"))
    # switching back to min-sum on the same handle still matches the min-sum oracle
    dec.set_algorithm(0)
    ref = o.decode(llr)
    assert_parity(_run_device(dec, llr), ref, c["N"], what="min-sum after sum-product")


@pytest.mark.parametrize("rate,name,num,den", [(4, "3/4B", 3, 4), (0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (5, "5/6", 5, 6)])
def test_quasi_cyclic_sum_product_kernel(rate, name, num, den, monkeypatch):
    """ldpc_sp_qcm_kernel (ldpc_spq.cuh) on every rate of the reference's family: forced on the short code the on-chip
    group kernel also holds (z = 24), and as planned where one, two and three warps serve a codeword (z = 28, 36, 68; idle
    lanes in the last warp).  Bits and iteration counts of the sum-product oracle (pinned to the reference's own kernels,
    tests/test_oracle_vs_refcl.py); special values; the streamed host path."""
    import myldpccppapi_b200 as m
    torch = _torch()
    monkeypatch.setenv("LDPC_B200_SP_QC", "1")
    for z in (24, 28, 36, 68):
        N = 24 * z
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        y = np.concatenate([awgn_llr(14, N, sigma_from_ebn0(2.0, num / den), seed=N + rate), awgn_llr(14, N, sigma_from_ebn0(4.0, num / den), seed=N + rate + 1),
                            awgn_llr(5, N, 1.3, seed=N + rate + 2)])
        y[3] = 0.0
        y[4, ::3] = 0.0
        y[5] = np.where(np.arange(N) % 2 == 0, -0.0, 0.0)
        y[6, :7] = [20.0, -20.0, 1e30, -1e30, 1e-40, -1e-40, 11.1]   # exp(8y) overflows / underflows
        o = oracle.Oracle(M, N, K, rp, ci, times=40)
        sp = oracle.decode_sp(o, y)
        dec = m.Decoder.wimax(K, N, rate)
        dec.set_algorithm(1)
        out = dec.decode_device(torch.from_numpy(y).cuda(), want_hard=True)
        torch.cuda.synchronize()
        what = "z=%d rate %s" % (z, name)
        assert dec.info()["kernel_variant"] == 5, what
        assert np.array_equal(out["iters"].cpu().numpy(), sp[1]), what
        assert np.array_equal(out["info"].cpu().numpy(), sp[0]), what
        assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(sp[2], axis=1, bitorder="little")), what
        if z in (24, 36):
            host = dec.decode_host(torch.from_numpy(y).pin_memory().numpy())
            assert np.array_equal(host["iters"], sp[1]) and np.array_equal(host["info"], sp[0]), what + " host"


def test_sum_product_other_rates_and_caps():
    import myldpccppapi_b200 as m
    torch = _torch()
    for rate, name, num, den in [(0, "1/2", 1, 2), (5, "5/6", 5, 6)]:
        N = 576
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        llr = awgn_llr(130, N, sigma_from_ebn0(2.0, num / den), seed=rate + 7)
        for cap in (1, 3, 40):
            o = oracle.Oracle(M, N, K, rp, ci, times=cap)
            info, iters, hard, _, _ = oracle.decode_sp(o, llr)
            dec = m.Decoder.wimax(K, N, rate, max_iter=cap)
            dec.set_algorithm(1)
            out = dec.decode_device(torch.from_numpy(llr).cuda(), want_hard=True)
            torch.cuda.synchronize()
            assert np.array_equal(out["iters"].cpu().numpy(), iters), (name, cap)
            assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(hard, axis=1, bitorder="little")), (name, cap)


def test_sum_product_has_no_posterior_output():
    """(the reference's sum-product keeps probabilities, not a posterior LLR: the one output DecodeSP cannot give)"""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    dec = m.Decoder(M, N, K, rp, ci)
    dec.set_algorithm(1)
    with pytest.raises(m.LdpcError) as e:
        dec.decode_device(_torch().zeros((4, N), dtype=_torch().float32, device="cuda"), want_post=True)
    assert e.value.code == -3


@pytest.mark.parametrize("sigma", [0.45, 0.55, 0.6, 0.66, 0.9])
def test_layered_min_sum_matches_restated_oracle(default_code, sigma):
    """DecodeTDMP: the layered min-sum kernel against the oracle's restatement of the schedule the reference's
    TDMP kernels intend (decodeCL.c:203-292 / decodeOnceTDMP, layer bookkeeping repaired).  Same IEEE fp32
    operations in the same order on both sides: bits, iteration counts and posteriors must be identical.
    (The reference has no runnable TDMP: this parity is GPU == restated oracle.)"""
    import myldpccppapi_b200 as m
    c = default_code
    llr = awgn_llr(300 + 5, c["N"], sigma, seed=int(sigma * 1000) + 2)
    llr[0, :] = 0.0           # P == 0 everywhere: every bit keeps its initial 0, the word is "clean" at iteration 1
    llr[1, ::4] = -0.0
    llr[2, :] = -1.0
    llr[2, ::7] = 0.0
    o = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40)
    info, iters, hard, post = oracle.decode_tdmp(o, llr, c["N"] // 24)
    dec = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    dec.set_algorithm(2)
    torch = _torch()
    out = dec.decode_device(torch.from_numpy(llr).cuda(), want_hard=True, want_post=True)
    torch.cuda.synchronize()
    assert np.array_equal(out["iters"].cpu().numpy(), iters)
    assert np.array_equal(out["info"].cpu().numpy(), info)
    assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(hard, axis=1, bitorder="little"))
    assert np.array_equal(out["post"].cpu().numpy(), post)
    dec.set_algorithm(0)
    assert_parity(_run_device(dec, llr), o.decode(llr), c["N"], what="min-sum after layered")


def test_layered_other_rates_sizes_and_caps():
    import myldpccppapi_b200 as m
    torch = _torch()
    for rate, name, num, den, N in [(0, "1/2", 1, 2, 576), (1, "2/3A", 2, 3, 672), (5, "5/6", 5, 6, 576), (3, "3/4A", 3, 4, 1152)]:
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        llr = awgn_llr(70, N, sigma_from_ebn0(2.2, num / den), seed=rate + 17)
        for cap in (1, 2, 40):
            o = oracle.Oracle(M, N, K, rp, ci, times=cap)
            info, iters, hard, post = oracle.decode_tdmp(o, llr, N // 24)
            dec = m.Decoder.wimax(K, N, rate, max_iter=cap)
            dec.set_algorithm(2)
            out = dec.decode_device(torch.from_numpy(llr).cuda(), want_hard=True, want_post=True)
            torch.cuda.synchronize()
            assert np.array_equal(out["iters"].cpu().numpy(), iters), (name, N, cap)
            assert np.array_equal(out["hard"].cpu().numpy(), np.packbits(hard, axis=1, bitorder="little")), (name, N, cap)
            assert np.array_equal(out["post"].cpu().numpy(), post), (name, N, cap)
    # early termination off: every word runs to the cap
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    dec = m.Decoder.wimax(432, 576, 4, max_iter=7)
    dec.set_algorithm(2)
    dec.set_early_termination(False)
    out = dec.decode_device(torch.from_numpy(awgn_llr(40, 576, 0.4, seed=3)).cuda())
    assert (out["iters"].cpu().numpy() == 7).all()


def test_layered_needs_a_layer_height():
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    dec = m.Decoder(M, N, K, rp, ci)
    with pytest.raises(m.LdpcError) as e:
        dec.set_algorithm(2)
    assert e.value.code == -3
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    dec = m.Decoder(M, 576, 432, rp, ci)          # CSR constructor: z unknown until told
    with pytest.raises(m.LdpcError):
        dec.set_algorithm(2)
    dec.set_layer_height(24)
    dec.set_algorithm(2)
    with pytest.raises(m.LdpcError):              # 16 rows per layer: rows of a circulant collide with the next block row
        dec.set_layer_height(16)


@pytest.mark.parametrize("rate,name,num,den", [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4),
                                               (4, "3/4B", 3, 4), (5, "5/6", 5, 6)])
def test_device_encoder(rate, name, num, den):
    """Coder::forEncoder / Coder::encode on the device: parity bits of H c = 0.  Codewords equal the host GF(2)
    solver's, have zero syndrome, and -- for the five rates the reference encodes correctly -- equal the reference's
    own Coder::encode output (oracle/_ref, when it was built)."""
    import myldpccppapi_b200 as m
    torch = _torch()
    for N in (576, 960):
        K = N * num // den
        rp, ci, M = oracle.wimax_H(N, name)
        rng = np.random.default_rng(N + rate)
        ncw = 257
        u = rng.integers(0, 2, (ncw, K)).astype(np.uint8)
        info = np.packbits(u, axis=1, bitorder="little")
        dec = m.Decoder.wimax(K, N, rate)
        cw_dev = dec.encode_device(torch.from_numpy(info).cuda()).cpu().numpy()
        cw_host = dec.encode_host(info)
        assert np.array_equal(cw_dev, cw_host)
        bits = np.unpackbits(cw_dev, axis=1, bitorder="little")[:, :N]
        assert np.array_equal(bits[:, :K], u)
        assert not m.codes.syndrome(M, rp, ci, bits).any()
        Gp = m.codes.gf2_systematic_encoder(M, N, K, rp, ci)
        assert np.array_equal(bits[:, K:], (u.astype(np.int64) @ Gp.astype(np.int64) % 2).astype(np.uint8))
        from oracle import ref
        if ref.available() and name != "3/4B" and N == 576:
            rc = ref.RefCoder(K, N, rate)
            src = rng.integers(1, 256, (8, K // 8)).astype(np.uint8)  # no NUL: encodeOnce copies with strncpy (MyLdpc.cpp:661)
            assert np.array_equal(rc.encode(src.reshape(-1)).reshape(8, N // 8), dec.encode_host(src))


def test_device_roundtrip_encode_channel_decode(default_code):
    """Test.cpp's loop entirely on the device: random payload -> encode -> BPSK + AWGN (6 dB) -> decode, all three
    algorithms recover every info byte; the channel generator reproduces its values for a given seed."""
    import myldpccppapi_b200 as m
    torch = _torch()
    c = default_code
    N, K = c["N"], c["K"]
    ncw = 20000
    g = torch.Generator(device="cuda").manual_seed(5)
    info = torch.randint(0, 256, (ncw, K // 8), dtype=torch.uint8, device="cuda", generator=g)
    dec = m.Decoder.wimax(K, N, c["rate"])
    cw = dec.encode_device(info)
    llr = m.synth_llr(ncw, N, 10 ** (-6.0 / 20), seed=99, bits=cw)
    assert torch.equal(llr, m.synth_llr(ncw, N, 10 ** (-6.0 / 20), seed=99, bits=cw))
    for alg in (0, 1, 2):
        dec.set_algorithm(alg)
        out = dec.decode_device(llr)
        assert torch.equal(out["info"], info), alg
        assert int(out["iters"].max()) < 40


def test_coder_shards_over_every_visible_gpu(default_code):
    """Coder::setDevices ([B200] addition): the codewords of one decode() call are split into contiguous shards,
    one per GPU, with no collective; results equal the single-GPU ones.  With one visible GPU the same device is
    listed twice, which exercises the same sharding code (two handles, two host threads)."""
    import myldpccppapi_b200 as m
    torch = _torch()
    ngpu = torch.cuda.device_count()
    shards = max(ngpu, 2)
    kv1, _ = _run_cli("mytest", 54000, 256, 5.5, "MS", 23, 1)
    kvn, out = _run_cli("mytest", 54000, 256, 5.5, "MS", 23, shards, ngpu)
    assert kvn["ErrNum"] == kv1["ErrNum"] and kvn["MeanIterations"] == kv1["MeanIterations"], out


def test_chunked_pipeline_stages_pageable_input(default_code):
    """ldpc_b200_decode_host on the chunked pipeline (codes that are not quasi-cyclic; sum-product and layered decoding) with
    a pageable caller buffer: host threads stage pieces of it through the pinned ring (decode_host_chunked_staged) instead
    of the driver.  Same bytes, counts and posteriors as the oracle -- chunks smaller than a ring slot, chunks of several
    pieces, a ragged last chunk, one staging thread and many, plain memcpy staging, and the driver-staged loop (option off)."""
    import myldpccppapi_b200 as m
    M, N, K, rp, ci = m.codes.regular_code()
    y = awgn_llr(150, N, 0.84, seed=21)          # 150 words x 32 KB, pageable
    ref = oracle.Oracle(M, N, K, rp, ci, times=40).decode(y, literal=False)
    dec = m.Decoder(M, N, K, rp, ci)
    assert dec.info()["path_name"] != "qc"
    dec.set_option("staged_min_kb", 0)           # (default: calls of 8 MB or more)
    for batch, threads, nt in ((16, 6, 1), (37, 3, 0), (64, 1, 1), (200, 6, 1)):   # 200 words = 6.5 MB: two pieces per chunk
        dec.reserve(batch)
        dec.set_option("stage_threads", threads)
        dec.set_option("stage_nt", nt)
        assert_parity(dec.decode_host(y, want_hard=True, want_post=True), ref, N, what="staged chunks of %d words, %d threads" % (batch, threads))
        assert_parity(dec.decode_host(y[:1], want_hard=True, want_post=True), tuple(r[:1] for r in ref), N, what="one word")
    dec.set_option("chunk_stage", 0)
    assert_parity(dec.decode_host(y, want_hard=True, want_post=True), ref, N, what="driver-staged")
    dec.set_option("chunk_stage", 1)
    t = dec.timing(reset=True)
    dec.decode_host(y)
    t = dec.timing()
    assert t["calls"] == 1 and t["codewords"] == 150 and t["h2d_s"] > 0 and t["kernel_s"] > 0
    # sum-product and layered decoding of Test.cpp's code run the chunked pipeline too
    c = default_code
    yd = np.concatenate([awgn_llr(400, c["N"], 0.62, seed=31), awgn_llr(301, c["N"], 0.5, seed=32)])
    o = oracle.Oracle(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"], times=40)
    d2 = m.Decoder.wimax(c["K"], c["N"], c["rate"])
    d2.set_option("staged_min_kb", 0)
    d2.set_algorithm(1)
    sp = oracle.decode_sp(o, yd)
    out = d2.decode_host(yd, want_hard=True)
    assert np.array_equal(out["iters"], sp[1]) and np.array_equal(out["info"], sp[0])
    d2.set_algorithm(2)
    assert_parity(d2.decode_host(yd, want_hard=True, want_post=True), oracle.decode_tdmp(o, yd, 24), c["N"], what="layered, staged")
