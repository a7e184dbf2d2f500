import time, os, numpy as np, torch, myldpccppapi_b200 as m
N,K=576,432; ncw=65536
llr_d = m.synth_llr(ncw, N, 1.0, seed=1)
y = llr_d.cpu().numpy().copy()          # pageable
for mode in ("chunked", "staged"):
    if mode == "staged": os.environ["LDPC_B200_STAGED_MIN_KB"] = "0"
    dec = m.Decoder.wimax(K,N,4)
    out = {"info": np.empty((ncw,K//8),dtype=np.uint8), "iters": np.empty((ncw,),dtype=np.int32)}
    ts=[]
    for i in range(6):
        t0=time.perf_counter(); dec.decode_host(y, out=out); ts.append((time.perf_counter()-t0)*1e3)
    print(mode, " ".join("%.1f"%t for t in ts), "ms; last: %.2f Gbit/s" % (ncw*K/ts[-1]/1e6), "iters", out["iters"].mean())
