"""One device-resident decode of a WiMAX code for profiling: usage PYTHONPATH=. python tools/family_run.py N rate [ncw]"""
import sys, torch, myldpccppapi_b200 as m
N, rate = int(sys.argv[1]), int(sys.argv[2])
ncw = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
num, den = {0: (1, 2), 1: (2, 3), 2: (2, 3), 3: (3, 4), 4: (3, 4), 5: (5, 6)}[rate]
dec = m.Decoder.wimax(N * num // den, N, rate)
x = m.synth_llr(ncw, N, 1.0, seed=1)
for _ in range(3):
    out = dec.decode_device(x)
torch.cuda.synchronize()
print(dec.info()["path_name"], dec.info()["codewords_per_cta"], float(out["iters"].float().mean()))
