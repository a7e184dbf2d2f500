"""Device-resident timing of the three decode algorithms on Test.cpp's code (N=576, K=432, rate 3/4B).

usage: PYTHONPATH=. python tools/alg_time.py [ncw] [sigma ...]
Prints ms per batch, decoded info Gbit/s and the mean iteration count for min-sum (DecodeMS),
sum-product (DecodeSP) and layered min-sum (DecodeTDMP)."""
import sys

import torch

import myldpccppapi_b200 as m

ncw = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sigmas = [float(x) for x in sys.argv[2:]] or [1.0, 0.6, 0.5]
names = {0: "min-sum", 1: "sum-product", 2: "layered"}
dec = m.Decoder.wimax(432, 576, 4)
for sigma in sigmas:
    llr = m.synth_llr(ncw, 576, sigma, seed=1)
    for alg in (0, 1, 2):
        dec.set_algorithm(alg)
        for _ in range(3):
            out = dec.decode_device(llr)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = dec.decode_device(llr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"sigma {sigma:4.2f} {names[alg]:12s} {ms:8.3f} ms  {ncw * 432 / ms / 1e6:7.2f} Gbit/s  "
              f"mean iterations {out['iters'].float().mean().item():6.2f}")
