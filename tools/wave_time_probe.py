import torch, myldpccppapi_b200 as m
N, K = 576, 432
llr = m.synth_llr(65536*2, N, 1.0, seed=1)
dec = m.Decoder.wimax(K, N, 4)
def timeit(fn, n):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for ncw in (2368, 4736, 9472, 18944, 37888, 65536):
    nch = min(36, 131072 // ncw)
    xs = [llr[i * ncw:(i + 1) * ncw] for i in range(nch)]
    outs = [dec.decode_device(x) for x in xs]
    t1 = timeit(lambda i: dec.decode_device(xs[0]), nch)
    t2 = timeit(lambda i: dec.decode_device(xs[0], out=outs[0]), nch)
    t3 = timeit(lambda i: dec.decode_device(xs[i], out=outs[i]), nch)
    w = ncw / 2368
    print(ncw, "waves %.1f" % w, "same x fresh out %.1f us/wave | same x same out %.1f | distinct %.1f" % (t1 / w * 1e3, t2 / w * 1e3, t3 / w * 1e3), "mean it", float(outs[0]["iters"].float().mean()))
