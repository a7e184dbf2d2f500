import torch, numpy as np, myldpccppapi_b200 as m
dec = m.Decoder.wimax(432, 576, 4)
for sigma in (1.0, 0.5):
    llr = m.synth_llr(65536, 576, sigma, seed=1) if hasattr(m,'synth_llr') else None
    for alg in (0,1):
        dec.set_algorithm(alg)
        for _ in range(2): out = dec.decode_device(llr)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = dec.decode_device(llr)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/3
        print("sigma",sigma,"alg",alg,"ms",ms,"Gbit/s",65536*432/ms/1e6,"mean iters",out["iters"].float().mean().item())
