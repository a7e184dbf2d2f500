"""Device-resident timing of every decode algorithm on one of the reference's codes (default Test.cpp's: N=576, rate 3/4B).

usage: PYTHONPATH=. python tools/alg_time.py [N rate ncw [sigma ...]]
Prints ms per batch, decoded info Gbit/s and the mean iteration count for min-sum (DecodeMS / CPU), sum-product
(DecodeSP), layered min-sum (DecodeTDMP) and the two fused-kernel arithmetics (DecodeMSCL / DecodeTDMPCL on request);
for Test.cpp's code also sum-product / layered through the any-size kernels."""
import os
import sys

import torch

import myldpccppapi_b200 as m

N = int(sys.argv[1]) if len(sys.argv) > 1 else 576
rate = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ncw = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
sigmas = [float(x) for x in sys.argv[4:]] or [1.0, 0.6, 0.5]
num, den = {0: (1, 2), 1: (2, 3), 2: (2, 3), 3: (3, 4), 4: (3, 4), 5: (5, 6)}[rate]
K = N * num // den
names = {0: "min-sum", 1: "sum-product", 2: "layered", 3: "fused-MS arithmetic (cap 120)", 4: "fused-layered arithmetic"}


def run(dec, tag):
    for sigma in sigmas:
        llr = m.synth_llr(ncw, N, sigma, seed=1)
        for alg in (0, 1, 2, 3, 4):
            dec.set_algorithm(alg)
            dec.set_max_iter(120 if alg == 3 else 40)
            for _ in range(2):
                out = dec.decode_device(llr)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                out = dec.decode_device(llr)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"N {N} rate {rate} {tag:9s} sigma {sigma:4.2f} {names[alg]:30s} {ms:9.3f} ms  {ncw * K / ms / 1e6:7.2f} Gbit/s  "
                  f"mean iterations {out['iters'].float().mean().item():6.2f}", flush=True)


run(m.Decoder.wimax(K, N, rate), "default")
if N == 576:
    os.environ["LDPC_B200_SP_BIG"] = "1"
    os.environ["LDPC_B200_TDMP_G"] = "32"
    run(m.Decoder.wimax(K, N, rate), "any-size")
