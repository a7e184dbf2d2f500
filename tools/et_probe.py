"""One decode of 65,536 words of Test.cpp's code at a given Eb/N0 with the quasi-cyclic kernel variant forced (ncu target).
usage: PYTHONPATH=. python tools/et_probe.py <ebn0_dB> <qc_et 0|1> [launches]"""
import sys

import numpy as np
import torch

import myldpccppapi_b200 as m

e, mode = float(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N, K = 576, 432
dec = m.Decoder.wimax(K, N, 4, max_iter=40)
dec.reserve(65536)
dec.set_option("qc_et", mode)
rng = np.random.default_rng(2024)
bits = dec.encode_device(torch.from_numpy(m.codes.pack_bits(rng.integers(0, 2, (65536, K)).astype(np.uint8))).cuda())
sigma = float(np.sqrt(1.0 / (2.0 * 0.75 * 10.0 ** (e / 10.0))))
x = m.synth_llr(65536, N, sigma, seed=107, bits=bits)
for _ in range(reps):
    out = dec.decode_device(x)
torch.cuda.synchronize()
print("mean iterations %.2f kernel variant %d" % (float(out["iters"].float().mean()), dec.info()["kernel_variant"]))
