"""First and steady-state decode of a FRESH process (CUDA loads a kernel's code at its first launch, so only a new
process shows what a user's first call costs).  Scenarios, each in its own subprocess:
  coder      Coder::decode through Coder::setDevices on all visible GPUs (one process, one host thread per GPU, malloc'd buffers)
  cabi       ldpc_b200_decode_host with a pinned buffer on a handle nobody called ldpc_b200_reserve on
  <deType>   coder with DecodeSP / DecodeTDMP / DecodeMSCL
usage: PYTHONPATH=. python tools/setdevices_first_call.py [words_per_gpu] [scenario ...]"""
import subprocess
import sys
import time


def run(scenario, per):
    import numpy as np
    import torch

    import myldpccppapi_b200 as m

    N, K = 576, 432
    rng = np.random.default_rng(3)
    if scenario == "cabi":
        post = torch.from_numpy(1.0 + rng.standard_normal((per, N)).astype(np.float32)).pin_memory()
        t0 = time.perf_counter(); dec = m.Decoder.wimax(K, N, m.rate_3_4_b); setup = time.perf_counter() - t0
        t0 = time.perf_counter(); dec.decode_host(post); first = time.perf_counter() - t0
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); dec.decode_host(post); ts.append(time.perf_counter() - t0)
        print("%-10s gpus 1  setup %.1f ms  first call %.1f ms  steady %.2f ms" % (scenario, setup * 1e3, first * 1e3, min(ts) * 1e3), flush=True)
        return
    de = {"coder": m.DecodeMS, "DecodeSP": m.DecodeSP, "DecodeTDMP": m.DecodeTDMP, "DecodeMSCL": m.DecodeMSCL}[scenario]
    n = torch.cuda.device_count()
    tot = per * n
    post = (1.0 + rng.standard_normal((tot, N)).astype(np.float32)).reshape(-1)
    src_len = tot * (K // 8)
    src = np.zeros(src_len + 1, dtype=np.uint8)
    coder = m.Coder(K, N, m.rate_3_4_b)
    coder.setDevices(list(range(n)))
    t0 = time.perf_counter(); coder.forDecoder(tot); coder.addDecodeType(de); setup = time.perf_counter() - t0
    t0 = time.perf_counter(); coder.decode(post, src, src_len, de); first = time.perf_counter() - t0
    print("           first call phases (s):", {k: round(v, 5) for k, v in coder.lastStepTimes().items()}, flush=True)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); coder.decode(post, src, src_len, de); ts.append(time.perf_counter() - t0)
    print("           steady phases (s):", {k: round(v, 5) for k, v in coder.lastStepTimes().items()}, " calls 2..6 (ms):", [round(t * 1e3, 2) for t in ts], flush=True)
    print("%-10s gpus %d  setup %.1f ms  first call %.1f ms  steady %.2f ms" % (scenario, n, setup * 1e3, first * 1e3, min(ts) * 1e3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        run(sys.argv[2], int(sys.argv[3]))
    else:
        args = sys.argv[1:]
        per = int(args[0]) if args and args[0].isdigit() else 65536
        scen = [a for a in args if not a.isdigit()] or ["coder", "cabi", "DecodeSP", "DecodeTDMP", "DecodeMSCL"]
        for s in scen:
            subprocess.run([sys.executable, __file__, "--one", s, str(per if s in ("coder", "cabi") else min(per, 8192))], check=False)
