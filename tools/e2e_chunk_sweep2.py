import time, numpy as np, torch, myldpccppapi_b200 as m
N,K=576,432; ncw=65536
llr_d = m.synth_llr(ncw, N, 1.0, seed=1)
llr_h = torch.empty((ncw,N),dtype=torch.float32).pin_memory(); llr_h.copy_(llr_d); torch.cuda.synchronize()
y = llr_h.numpy()
for chunk in (1184, 1776, 2368):
    dec = m.Decoder.wimax(K,N,4); dec.reserve(chunk)
    out={"info": torch.empty((ncw,K//8),dtype=torch.uint8).pin_memory().numpy(), "iters": torch.empty((ncw,),dtype=torch.int32).pin_memory().numpy()}
    for _ in range(3): dec.decode_host(y, out=out)
    t0=time.perf_counter()
    for _ in range(10): dec.decode_host(y, out=out)
    ms=(time.perf_counter()-t0)/10*1e3
    print(dec.info()["path_name"], chunk, round(ms,3), "ms", round(ncw*K/ms/1e6,3), "Gbit/s")
