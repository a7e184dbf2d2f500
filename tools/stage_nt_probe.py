"""Pageable host buffer -> ldpc_b200_decode_host (staged through the pinned ring by host threads): plain memcpy against
non-temporal stores (option stage_nt), a few thread counts.  Test.cpp's code, 65,536 words, fixed 40 iterations and Eb/N0 4 dB."""
import time

import numpy as np

import myldpccppapi_b200 as m

import sys

N, K, ncw = 576, 432, 65536
THREADS = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else (4, 8)
MODES = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 1)
for sigma, what in ((1.0, "cap 40"), (0.515, "4 dB")):
    y = m.synth_llr(ncw, N, sigma, seed=1).cpu().numpy().copy()   # pageable
    dec = m.Decoder.wimax(K, N, 4)
    dec.reserve(ncw)
    out = {"info": np.empty((ncw, K // 8), dtype=np.uint8), "iters": np.empty((ncw,), dtype=np.int32)}
    ref = None
    for _ in range(3):
        dec.decode_host(y, out=out)
    for threads in THREADS:
        dec.set_option("stage_threads", threads)
        res = {}
        for rnd in range(3):
            for nt in MODES:
                dec.set_option("stage_nt", nt)
                ts = []
                for i in range(8):
                    t0 = time.perf_counter()
                    dec.decode_host(y, out=out)
                    ts.append((time.perf_counter() - t0) * 1e3)
                res.setdefault(nt, []).extend(ts)
                if ref is None:
                    ref = out["info"].copy()
                assert np.array_equal(ref, out["info"])
        for nt in MODES:
            a = np.array(res[nt])
            print("%s  threads %d  stage_nt %d: median %.2f ms  min %.2f  (%.1f GB/s at the median)" % (what, threads, nt, np.median(a), a.min(), ncw * N * 4 / np.median(a) / 1e6))
