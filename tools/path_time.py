"""Device-resident decode time of every kernel family on cfg2 (Test.cpp's code, 65,536 words, sigma 1.0, cap 40)
and on the regular (3,6) N=8192 code (16,384 words).  usage: PYTHONPATH=. python tools/path_time.py"""
import torch, myldpccppapi_b200 as m


def timed(dec, x, reps=5):
    for _ in range(2): dec.decode_device(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = dec.decode_device(x)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, float(out["iters"].float().mean())


N, K = 576, 432
x = m.synth_llr(65536, N, 1.0, seed=1)
print("cfg2: Test.cpp's code, 65,536 words, sigma 1.0, cap 40")
for path in (7, 4, 3, 0, 8):
    dec = m.Decoder.wimax(K, N, 4)
    dec.set_path(path)
    ms, it = timed(dec, x)
    print("  %-10s %8.3f ms  %6.2f Gbit/s  mean iterations %.2f" % (dec.info()["path_name"], ms, 65536 * K / ms / 1e6, it))
M2, N2, K2, rp, ci = m.codes.regular_code()
x2 = m.synth_llr(16384, N2, 1.0, seed=2)
print("cfg3: regular (3,6) N=8192, 16,384 words, sigma 1.0, cap 40")
for path in (4, 6, 8):
    dec = m.Decoder(M2, N2, K2, rp, ci)
    dec.set_path(path)
    ms, it = timed(dec, x2, reps=3)
    print("  %-10s %8.3f ms  %6.2f Gbit/s  mean iterations %.2f" % (dec.info()["path_name"], ms, 16384 * K2 / ms / 1e6, it))
