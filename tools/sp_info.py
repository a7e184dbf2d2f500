import os, torch, myldpccppapi_b200 as m
for z in (24, 28, 32, 36):
    for env in ({}, {"LDPC_B200_SP_QC": "0"}):
        os.environ.update(env)
        N = 24*z; K = N*3//4
        d = m.Decoder.wimax(K, N, 4); d.set_algorithm(1)
        for k in env: del os.environ[k]
        x = m.synth_llr(4096, N, 1.0, seed=1)
        o = d.decode_device(x); torch.cuda.synchronize()
        a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); d.decode_device(x, out=o); b.record(); torch.cuda.synchronize()
        i = d.info()
        print(z, env, i["path_name"], i["codewords_per_cta"], i["threads_per_cta"], "variant", i["kernel_variant"], "%.2f ms" % a.elapsed_time(b))
