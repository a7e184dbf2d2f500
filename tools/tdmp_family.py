"""Layered min-sum (DecodeTDMP) over the reference's family: device-resident time at the cap and at 3.5 dB, and which
kernel ran (on-chip ldpc_tdmp_kernel or the any-size ldpc_tdmp_big_kernel).  usage: PYTHONPATH=. python tools/tdmp_family.py"""
import argparse

import numpy as np
import torch

import myldpccppapi_b200 as m

RATES = {0: ("1/2", 1, 2), 1: ("2/3A", 2, 3), 2: ("2/3B", 2, 3), 3: ("3/4A", 3, 4), 4: ("3/4B", 3, 4), 5: ("5/6", 5, 6)}


def timed(dec, x, out, reps=3):
    dec.decode_device(x, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dec.decode_device(x, out=out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ap = argparse.ArgumentParser()
ap.add_argument("--zs", default=",".join(str(z) for z in range(24, 97, 8)))
ap.add_argument("--rates", default="0,4,5")
ap.add_argument("--ncw", type=int, default=16384)
args = ap.parse_args()
print("| z | N | rate | cap: ms (Gbit/s) | 3.5 dB: ms (Gbit/s, mean it) | min-sum flooding at the cap, ms |")
print("|---|---|---|---|---|---|")
for z in (int(v) for v in args.zs.split(",")):
    N = 24 * z
    for rate in (int(v) for v in args.rates.split(",")):
        name, num, den = RATES[rate]
        K = N * num // den
        dec = m.Decoder.wimax(K, N, rate, max_iter=40)
        x = m.synth_llr(args.ncw, N, 1.0, seed=z + rate)
        o = dec.decode_device(x)
        t_ms = timed(dec, x, o)
        dec.set_layer_height(z)
        dec.set_algorithm(2)
        o = dec.decode_device(x)
        t1 = timed(dec, x, o)
        sigma = float(np.sqrt(1.0 / (2.0 * (num / den) * 10.0 ** 0.35)))
        y = m.synth_llr(args.ncw, N, sigma, seed=z + rate + 1)
        o2 = dec.decode_device(y)
        t2 = timed(dec, y, o2)
        print("| %d | %d | %s | %.3f (%.2f) | %.3f (%.2f, %.1f) | %.3f |" % (z, N, name, t1, args.ncw * K / t1 / 1e6, t2, args.ncw * K / t2 / 1e6,
                                                                      float(o2["iters"].float().mean()), t_ms), flush=True)
