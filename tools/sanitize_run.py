"""Small decode of every kernel family, for compute-sanitizer (memcheck / racecheck) runs.

usage: PYTHONPATH=. compute-sanitizer --tool racecheck python tools/sanitize_run.py [which ...]
which: ms (quasi-cyclic path) group host sp tdmp enc reg36 stream lane16 lane_smem (default: all)"""
import sys

import numpy as np
import torch

import myldpccppapi_b200 as m

which = set(sys.argv[1:]) or {"ms", "group", "host", "sp", "tdmp", "enc", "reg36", "stream", "lane16", "lane_smem"}
rng = np.random.default_rng(0)


def llr(ncw, N, sigma):
    return torch.from_numpy((1 + sigma * rng.standard_normal((ncw, N))).astype(np.float32)).cuda()


dec = m.Decoder.wimax(432, 576, 4, max_iter=6)
y = llr(700, 576, 0.58)
for name, alg, path in (("ms", 0, -1), ("group", 0, 4), ("sp", 1, -1), ("tdmp", 2, -1), ("lane16", 0, 3), ("lane_smem", 0, 0)):
    if name in which:
        dec.set_algorithm(alg)
        dec.set_path(path)
        out = dec.decode_device(y, want_hard=True)
        torch.cuda.synchronize()
        print(name, "mean iterations", out["iters"].float().mean().item())
dec.set_algorithm(0)
dec.set_path(-1)
if "host" in which:  # host buffers: one persistent launch fed by the copy stream
    res = dec.decode_host(y.cpu().numpy(), want_hard=True)
    print("host", dec.info()["path_name"], "mean iterations", float(res["iters"].mean()))
if "enc" in which:
    info = torch.randint(0, 256, (300, 54), dtype=torch.uint8, device="cuda")
    cw = dec.encode_device(info)
    torch.cuda.synchronize()
    print("enc", cw.shape)
if "reg36" in which or "stream" in which:
    M, N, K, rp, ci = m.codes.regular_code()
    d2 = m.Decoder(M, N, K, rp, ci, max_iter=4)
    y2 = llr(40, N, 0.8)
    for name, path in (("reg36", -1), ("stream", 6)):
        if name in which:
            d2.set_path(path)
            out = d2.decode_device(y2)
            torch.cuda.synchronize()
            print(name, "mean iterations", out["iters"].float().mean().item())
