"""Per-launch times and kernel choice of one handle walked through the Eb/N0 points of tools/ebn0_sweep.py (auto mode).
usage: PYTHONPATH=. python tools/auto_regime_probe.py"""
import numpy as np
import torch

import myldpccppapi_b200 as m

N, K = 576, 432
dec = m.Decoder.wimax(K, N, 4, max_iter=40)
rng = np.random.default_rng(2024)
u = rng.integers(0, 2, (65536, K)).astype(np.uint8)
bits = dec.encode_device(torch.from_numpy(m.codes.pack_bits(u)).cuda())
for i, e in enumerate((0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0)):
    sigma = float(np.sqrt(1.0 / (2.0 * 0.75 * 10.0 ** (e / 10.0))))
    x = m.synth_llr(65536, N, sigma, seed=100 + i, bits=bits)
    out = dec.decode_device(x)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dec.decode_device(x, out=out); b.record(); torch.cuda.synchronize()
        ts.append((round(a.elapsed_time(b), 3), dec.info()["kernel_variant"]))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        dec.decode_device(x, out=out)
    b.record(); torch.cuda.synchronize()
    print(e, round(float(out["iters"].float().mean()), 2), ts, "5 back to back: %.3f ms each" % (a.elapsed_time(b) / 5), flush=True)
