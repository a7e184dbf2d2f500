"""Lockstep quasi-cyclic kernel (compiled profile) against the group-of-warps kernel (ldpc_qcm.cuh) over the Eb/N0 range,
at a block size that has both.  usage: PYTHONPATH=. python tools/qcm_vs_lockstep.py --N 1152 [--rate 4] [--points sigma1,2,3,3.5,4]"""
import argparse
import os

import numpy as np
import torch

import myldpccppapi_b200 as m


def timed(dec, x, out, reps=8):
    for _ in range(3):
        dec.decode_device(x, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dec.decode_device(x, out=out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1152)
    ap.add_argument("--rate", type=int, default=4)
    ap.add_argument("--ncw", type=int, default=32768)
    ap.add_argument("--points", default="sigma1,2,2.5,3,3.5,4,5")
    args = ap.parse_args()
    N = args.N
    num, den = {0: (1, 2), 1: (2, 3), 2: (2, 3), 3: (3, 4), 4: (3, 4), 5: (5, 6)}[args.rate]
    K = N * num // den
    main_dec = m.Decoder.wimax(K, N, args.rate, max_iter=40)
    main_dec.set_option("qc_et", 0)
    os.environ["LDPC_B200_QCM_ALWAYS"] = "1"
    alt = m.Decoder.wimax(K, N, args.rate, max_iter=40)
    del os.environ["LDPC_B200_QCM_ALWAYS"]
    print("main:", main_dec.info()["codewords_per_cta"], "x", main_dec.info()["threads_per_cta"], " qcm:", alt.info()["codewords_per_cta"], "x", alt.info()["threads_per_cta"])
    rng = np.random.default_rng(7)
    bits = main_dec.encode_device(torch.from_numpy(m.codes.pack_bits(rng.integers(0, 2, (args.ncw, K)).astype(np.uint8))).cuda())
    print("| point | mean iters | share of cap | lockstep ms | group-of-warps ms | ratio | identical |")
    print("|---|---|---|---|---|---|---|")
    for i, pt in enumerate(args.points.split(",")):
        sigma = 1.0 if pt == "sigma1" else float(np.sqrt(1.0 / (2.0 * (num / den) * 10.0 ** (float(pt) / 10.0))))
        x = m.synth_llr(args.ncw, N, sigma, seed=300 + i, bits=bits)
        o1 = main_dec.decode_device(x)
        o2 = alt.decode_device(x)
        t1, t2 = timed(main_dec, x, o1), timed(alt, x, o2)
        same = bool(torch.equal(o1["info"], o2["info"]) and torch.equal(o1["iters"], o2["iters"]))
        it = float(o1["iters"].float().mean())
        print("| %s | %.2f | %.0f %% | %.3f | %.3f | %.2fx | %s |" % (pt, it, 100 * it / 40, t1, t2, t1 / t2, same))


if __name__ == "__main__":
    main()
