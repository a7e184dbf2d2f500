import time, torch, myldpccppapi_b200 as m
M,N,K,rp,ci = m.codes.regular_code()
llr = m.synth_llr(64, N, 0.8, seed=1)
t0=time.perf_counter(); dec = m.Decoder(M,N,K,rp,ci); out = dec.decode_device(llr); torch.cuda.synchronize(); print("create + first decode: %.2f s" % (time.perf_counter()-t0))
