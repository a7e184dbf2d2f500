"""One device-resident decode of a code of the reference's family (for ncu captures).
usage: PYTHONPATH=. python tools/one_decode.py N rate ncw sigma [algorithm]"""
import sys

import torch

import myldpccppapi_b200 as m

N, rate, ncw, sigma = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
alg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
num, den = {0: (1, 2), 1: (2, 3), 2: (2, 3), 3: (3, 4), 4: (3, 4), 5: (5, 6)}[rate]
dec = m.Decoder.wimax(N * num // den, N, rate, max_iter=40)
dec.set_option("no_warm", 1)
if alg:
    dec.set_algorithm(alg)
x = m.synth_llr(ncw, N, sigma, seed=1)
out = dec.decode_device(x)
torch.cuda.synchronize()
print(dec.info()["path_name"], dec.info()["kernel_variant"], float(out["iters"].float().mean()))
