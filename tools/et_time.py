"""The early-termination kernel (ldpc_qc_et.cuh) against the main quasi-cyclic kernel on Test.cpp's code, 65,536 words,
cap 40, device-resident, over the Eb/N0 range of BASELINE config 4 and the fixed-40 headline (sigma 1.0).  Random valid
codewords (device encoder) + AWGN; both kernels must give identical bytes and iteration counts.
usage: PYTHONPATH=. python tools/et_time.py [--rate 4] [--points 0,1,2,2.5,3,3.5,4,5]"""
import argparse

import numpy as np
import torch

import myldpccppapi_b200 as m


def timed(dec, x, out, reps):
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
    ap.add_argument("--ncw", type=int, default=65536)
    ap.add_argument("--rate", type=int, default=4)
    ap.add_argument("--N", type=int, default=576)
    ap.add_argument("--points", default="sigma1,0,1,2,2.5,3,3.5,4,5")
    args = ap.parse_args()
    N = args.N
    num, den = {0: (1, 2), 1: (2, 3), 2: (2, 3), 3: (3, 4), 4: (3, 4), 5: (5, 6)}[args.rate]
    K = N * num // den
    dec = m.Decoder.wimax(K, N, args.rate, max_iter=40)
    dec.reserve(args.ncw)
    assert dec.info()["et_available"] == 1
    rng = np.random.default_rng(2024)
    u = rng.integers(0, 2, (args.ncw, K)).astype(np.uint8)
    bits = dec.encode_device(torch.from_numpy(m.codes.pack_bits(u)).cuda())
    print("| point | sigma | mean iters | main kernel ms | ET kernel ms | speed-up | ET info Gbit/s | identical |")
    print("|---|---|---|---|---|---|---|---|")
    for i, pt in enumerate(args.points.split(",")):
        sigma = 1.0 if pt == "sigma1" else float(np.sqrt(1.0 / (2.0 * (num / den) * 10.0 ** (float(pt) / 10.0))))
        x = m.synth_llr(args.ncw, N, sigma, seed=100 + i, bits=bits)
        res = {}
        for mode in (0, 1):
            dec.set_option("qc_et", mode)
            out = dec.decode_device(x)
            ms = timed(dec, x, out, 10)
            assert dec.info()["kernel_variant"] == mode
            res[mode] = (ms, out["info"].clone(), out["iters"].clone())
        same = bool(torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2]))
        it = float(res[0][2].float().mean())
        print("| %s | %.4f | %.2f | %.3f | %.3f | %.2fx | %.1f | %s |" % (pt if pt == "sigma1" else pt + " dB", sigma, it, res[0][0], res[1][0],
                                                                       res[0][0] / res[1][0], args.ncw * K / res[1][0] / 1e6, same))
    dec.set_option("qc_et", -1)


if __name__ == "__main__":
    main()
