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
    ap.add_argument("--algorithm", default="ms", choices=["ms", "sp", "tdmp"],
                    help="ms = flooding min-sum (DecodeMS), sp = sum-product (DecodeSP), tdmp = layered min-sum (DecodeTDMP)")
    args = ap.parse_args()
    import torch
    import myldpccppapi_b200 as m
    import oracle

    N, K, rate = 576, 432, m.rate_3_4_b
    rp, ci, M = m.wimax_csr(K, N, rate)
    dec = m.Decoder.wimax(K, N, rate, max_iter=40)
    dec.set_algorithm({"ms": 0, "sp": 1, "tdmp": 2}[args.algorithm])
    orc = oracle.Oracle(M, N, K, rp, ci, times=40)

    def oracle_decode(y):
        if args.algorithm == "sp":
            r = oracle.decode_sp(orc, y)
        elif args.algorithm == "tdmp":
            r = oracle.decode_tdmp(orc, y, N // 24)
        else:
            r = orc.decode(y, want_post=False, want_hard=False)
        return r[0], r[1]

    rng = np.random.default_rng(2024)
    u = rng.integers(0, 2, (args.ncw, K)).astype(np.uint8)
    want = torch.from_numpy(m.codes.pack_bits(u)).cuda()
    bits = dec.encode_device(want)  # systematic encoder on the device
    print("algorithm: %s\n" % args.algorithm)
    print("| Eb/N0 dB | sigma | mean iters | cap hits | FER | info BER | ms / %d words | info Gbit/s | oracle match (%d words) | iteration histogram (1,2,3,4,5-8,9-16,17-39,40) |" % (args.ncw, args.check))
    print("|---|---|---|---|---|---|---|---|---|---|")
    for i, e in enumerate(float(x) for x in args.points.split(",")):
        sigma = float(np.sqrt(1.0 / (2.0 * 0.75 * 10.0 ** (e / 10.0))))
        llr = m.synth_llr(args.ncw, N, sigma, seed=100 + i, bits=bits)
        out = dec.decode_device(llr)
        for _ in range(4):   # (the handle samples the iteration counts after every fourth launch and picks its kernel from them)
            dec.decode_device(llr, out=out)
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
        ref = oracle_decode(llr[:nchk].cpu().numpy())
        ok = bool(np.array_equal(ref[0], out["info"][:nchk].cpu().numpy()) and np.array_equal(ref[1], it[:nchk]))
        h = [int((it == k).sum()) for k in (1, 2, 3, 4)] + [int(((it >= a) & (it <= b)).sum()) for a, b in ((5, 8), (9, 16), (17, 39))] + [int((it == 40).sum())]
        print("| %.1f | %.4f | %.2f | %.2f%% | %.2e | %.2e | %.3f | %.2f | %s | %s |" % (
            e, sigma, it.mean(), 100.0 * (it == 40).mean(), wrong_words / args.ncw, wrong_bits / (args.ncw * K), ms,
            args.ncw * K / (ms * 1e-3) / 1e9, "yes" if ok else "NO", " ".join(map(str, h))))
        assert ok, "GPU and oracle differ at Eb/N0 = %g" % e


if __name__ == "__main__":
    main()
