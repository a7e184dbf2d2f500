#!/usr/bin/env python3
"""BASELINE config 4: Eb/N0 sweep on Test.cpp's code with syndrome early termination.

For every Eb/N0 point: random valid codewords -> BPSK + AWGN on the GPU (ldpc_b200_synth_llr), decode
with the CUDA decoder, report FER/BER on the info bits, mean iterations, the iteration histogram and the
decode throughput; the first `--check` words of every point are also decoded by the CPU oracle and must
match bit for bit (bits and iteration counts).  Writes a markdown table (stdout) -- see profiles/.
"""
import argparse
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncw", type=int, default=65536)
    ap.add_argument("--check", type=int, default=512)
    ap.add_argument("--points", default="0,0.5,1,1.5,2,2.5,3,3.5,4")
    args = ap.parse_args()
    import torch
    import myldpccppapi_b200 as m
    import oracle

    N, K, rate = 576, 432, m.rate_3_4_b
    rp, ci, M = m.wimax_csr(K, N, rate)
    Gp = m.codes.gf2_systematic_encoder(M, N, K, rp, ci).astype(np.float32)
    dec = m.Decoder(M, N, K, rp, ci, max_iter=40)
    orc = oracle.Oracle(M, N, K, rp, ci, times=40)
    rng = np.random.default_rng(2024)
    u = rng.integers(0, 2, (args.ncw, K)).astype(np.uint8)
    cw = np.concatenate([u, ((u.astype(np.float32) @ Gp) % 2).astype(np.uint8)], axis=1)
    bits = torch.from_numpy(m.codes.pack_bits(cw)).cuda()
    want = torch.from_numpy(m.codes.pack_bits(u)).cuda()
    print("| Eb/N0 dB | sigma | mean iters | cap hits | FER | info BER | ms / %d words | info Gbit/s | oracle match (%d words) | iteration histogram (1,2,3,4,5-8,9-16,17-39,40) |" % (args.ncw, args.check))
    print("|---|---|---|---|---|---|---|---|---|---|")
    for i, e in enumerate(float(x) for x in args.points.split(",")):
        sigma = float(np.sqrt(1.0 / (2.0 * 0.75 * 10.0 ** (e / 10.0))))
        llr = m.synth_llr(args.ncw, N, sigma, seed=100 + i, bits=bits)
        out = dec.decode_device(llr)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            dec.decode_device(llr, out=out)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        it = out["iters"].cpu().numpy()
        diff = (out["info"] ^ want)
        wrong_words = int((diff.reshape(args.ncw, -1).max(dim=1).values != 0).sum().item())
        wrong_bits = int(np.unpackbits(diff.cpu().numpy()).sum())
        nchk = min(args.check, args.ncw)
        ref = orc.decode(llr[:nchk].cpu().numpy(), want_post=False, want_hard=False)
        ok = bool(np.array_equal(ref[0], out["info"][:nchk].cpu().numpy()) and np.array_equal(ref[1], it[:nchk]))
        h = [int((it == k).sum()) for k in (1, 2, 3, 4)] + [int(((it >= a) & (it <= b)).sum()) for a, b in ((5, 8), (9, 16), (17, 39))] + [int((it == 40).sum())]
        print("| %.1f | %.4f | %.2f | %.2f%% | %.2e | %.2e | %.3f | %.2f | %s | %s |" % (
            e, sigma, it.mean(), 100.0 * (it == 40).mean(), wrong_words / args.ncw, wrong_bits / (args.ncw * K), ms,
            args.ncw * K / (ms * 1e-3) / 1e9, "yes" if ok else "NO", " ".join(map(str, h))))
        assert ok, "GPU and oracle differ at Eb/N0 = %g" % e


if __name__ == "__main__":
    main()
