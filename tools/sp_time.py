"""Sum-product (DecodeSP) device-resident timing over the reference's family: the quasi-cyclic kernel (ldpc_spq.cuh),
the on-chip group kernel (ldpc_sp.cuh, short codes only) and the any-size kernel (ldpc_big.cuh), checked equal.
usage: PYTHONPATH=. python tools/sp_time.py [--zs 24,32,48,60,96] [--rates 0,4,5] [--ncw 16384]"""
import argparse
import os

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


def make(K, N, rate, env):
    for k, v in env.items():
        os.environ[k] = v
    dec = m.Decoder.wimax(K, N, rate, max_iter=40)
    for k in env:
        del os.environ[k]
    dec.set_algorithm(1)
    return dec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zs", default="24,32,36,48,60,76,96")
    ap.add_argument("--rates", default="0,4,5")
    ap.add_argument("--ncw", type=int, default=16384)
    ap.add_argument("--big-ncw", type=int, default=2048, help="words timed through the any-size kernel (slow)")
    args = ap.parse_args()
    print("| z | N | rate | regime | quasi-cyclic kernel ms (Gbit/s) | on-chip group kernel | any-size kernel (scaled to the batch) | identical |")
    print("|---|---|---|---|---|---|---|---|")
    for z in (int(v) for v in args.zs.split(",")):
        N = 24 * z
        for rate in (int(v) for v in args.rates.split(",")):
            name, num, den = RATES[rate]
            K = N * num // den
            qc = make(K, N, rate, {"LDPC_B200_SP_QC": "1"})
            grp = make(K, N, rate, {"LDPC_B200_SP_QC": "0"})
            big = make(K, N, rate, {"LDPC_B200_SP_BIG": "1"})
            for regime, sigma in (("cap", 1.0), ("3.5 dB", float(np.sqrt(1.0 / (2.0 * (num / den) * 10.0 ** 0.35))))):
                x = m.synth_llr(args.ncw, N, sigma, seed=z + rate)
                o1 = qc.decode_device(x)
                t1 = timed(qc, x, o1)
                v1 = qc.info()["kernel_variant"]
                o2 = grp.decode_device(x)
                t2 = timed(grp, x, o2)
                on_chip = grp.info()["path_name"] == "group" and t2 < 50 * t1
                xb = x[: args.big_ncw]
                o3 = big.decode_device(xb)
                t3 = timed(big, xb, o3, reps=1) * args.ncw / args.big_ncw
                same = all(torch.equal(o1[k], o2[k]) for k in ("info", "iters")) and all(torch.equal(o1[k][: args.big_ncw], o3[k]) for k in ("info", "iters"))
                print("| %d | %d | %s | %s (%.1f it) | %.3f (%.2f)%s | %.3f (%.2f) | %.1f (%.2f) | %s |" % (
                    z, N, name, regime, float(o1["iters"].float().mean()), t1, args.ncw * K / t1 / 1e6, "" if v1 == 5 else " [variant %d]" % v1,
                    t2, args.ncw * K / t2 / 1e6, t3, args.ncw * K / t3 / 1e6, same), flush=True)
            del qc, grp, big


if __name__ == "__main__":
    main()
