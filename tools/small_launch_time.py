"""Device-resident decode time of small launches (the chunks of the host pipeline), per kernel path."""
import sys, torch, myldpccppapi_b200 as m
N, K = 576, 432
llr = m.synth_llr(65536, N, 1.0, seed=1)
for path in (7, 4):
    dec = m.Decoder.wimax(K, N, 4)
    dec.set_path(path)
    for ncw in (296, 1184, 1776, 2368, 4736, 9472, 65536):
        x = llr[:ncw]
        for _ in range(3): dec.decode_device(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): dec.decode_device(x)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(dec.info()["path_name"], ncw, "%.4f ms" % ms, "%.2f Gbit/s" % (ncw * K / ms / 1e6))
