"""Do consecutive decode launches on different streams overlap?  (device-resident, CUDA events)"""
import torch, myldpccppapi_b200 as m
N, K = 576, 432
llr = m.synth_llr(65536, N, 1.0, seed=1)
streams = [torch.cuda.Stream() for _ in range(3)]
for path in (7, 4):
    dec = m.Decoder.wimax(K, N, 4)
    dec.set_path(path)
    for ncw in (1184, 1776, 2368):
        nch = 36
        xs = [llr[i * ncw:(i + 1) * ncw] for i in range(nch)]
        outs = [dec.decode_device(x) for x in xs]
        torch.cuda.synchronize()
        for mode in ("one stream", "three streams"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for s in streams: s.wait_stream(torch.cuda.current_stream())
            for i, x in enumerate(xs):
                st = torch.cuda.current_stream() if mode == "one stream" else streams[i % 3]
                dec.decode_device(x, out=outs[i], stream=st)
            for s in streams: torch.cuda.current_stream().wait_stream(s)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print(dec.info()["path_name"], ncw, mode, "%.3f ms for %d launches = %.1f us each, %.2f Gbit/s" % (ms, nch, ms / nch * 1e3, nch * ncw * K / ms / 1e6))
