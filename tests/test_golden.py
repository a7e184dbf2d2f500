"""Golden fixtures generated FROM THE REFERENCE ITSELF (tests/golden/make_golden.py: the reference's
MyLdpc.cpp compiled unmodified into oracle/_ref).  CPU part: the oracle restatement must reproduce the
reference's decodeCPU bytes at the reference's cap and at every cap 1..40 (per-iteration hard
decisions -> stopping iteration).  GPU part: the CUDA decoder must reproduce the same fixtures."""
import pathlib

import numpy as np
import pytest

import oracle

G = np.load(pathlib.Path(__file__).parent / "golden" / "default_code_golden.npz")
N, K = 576, 432
NSETS = len(G["sigmas"])


def _oracle(times=40):
    return oracle.Oracle(len(G["row_ptr"]) - 1, N, K, G["row_ptr"], G["col_idx"], times=times)


def test_golden_H_is_the_reference_H():
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    assert np.array_equal(rp, G["row_ptr"]) and np.array_equal(ci, G["col_idx"])


@pytest.mark.parametrize("s", range(NSETS))
def test_oracle_reproduces_reference_bytes(s):
    llr = G["llr_%d" % s]
    info, iters, hard, post = _oracle().decode(llr)
    assert np.array_equal(info.reshape(-1), G["ref_bytes_%d" % s])
    # stream API (Coder::decode(..., DecodeCPU) restated) gives the same bytes
    stream, it2, _, _ = _oracle().decode_stream(llr, llr.shape[0] * K // 8)
    assert np.array_equal(stream, G["ref_bytes_%d" % s]) and np.array_equal(it2, iters)
    assert np.array_equal(iters, G["oracle_iters_%d" % s])
    assert np.array_equal(post, G["oracle_post_%d" % s])
    assert np.array_equal(np.packbits(hard, axis=1, bitorder="little"), G["oracle_hard_%d" % s])


@pytest.mark.parametrize("s", range(NSETS))
def test_oracle_matches_reference_at_every_cap(s):
    """The reference was run with times = 1..40; the oracle with the same caps must give the same
    bytes, and its iteration count must be the first cap at which the reference's output stops being
    'still iterating' (cap-inclusive stop, no syndrome check before iteration 1)."""
    llr = G["llr_%d" % s]
    caps = G["ref_bytes_cap_%d" % s]
    ncw = llr.shape[0]
    iters40 = G["oracle_iters_%d" % s]
    for t in range(40):
        info, iters, _, _ = _oracle(times=t + 1).decode(llr, want_post=False)
        assert np.array_equal(info.reshape(-1), caps[t]), "cap %d" % (t + 1)
        assert np.array_equal(iters, np.minimum(iters40, t + 1))
    # a word that stopped at iteration i keeps the same bytes for every larger cap
    per_word = caps.reshape(40, ncw, K // 8)
    for b in range(ncw):
        i = int(iters40[b])
        assert all(np.array_equal(per_word[t, b], per_word[i - 1, b]) for t in range(i - 1, 40))


@pytest.mark.gpu
@pytest.mark.parametrize("s", range(NSETS))
def test_gpu_reproduces_reference_bytes(s):
    import torch
    import myldpccppapi_b200 as m
    llr = G["llr_%d" % s]
    dec = m.Decoder.wimax(K, N, m.rate_3_4_b)
    out = dec.decode_device(torch.from_numpy(llr).cuda(), want_hard=True, want_post=True)
    torch.cuda.synchronize()
    assert np.array_equal(out["info"].cpu().numpy().reshape(-1), G["ref_bytes_%d" % s])
    assert np.array_equal(out["iters"].cpu().numpy(), G["oracle_iters_%d" % s])
    assert np.array_equal(out["hard"].cpu().numpy(), G["oracle_hard_%d" % s])
    assert np.array_equal(out["post"].cpu().numpy(), G["oracle_post_%d" % s])
    # and at a few other caps, against the reference's per-cap bytes
    for cap in (1, 2, 5, 17, 39):
        dec.set_max_iter(cap)
        o2 = dec.decode_device(torch.from_numpy(llr).cuda())
        assert np.array_equal(o2["info"].cpu().numpy().reshape(-1), G["ref_bytes_cap_%d" % s][cap - 1])
