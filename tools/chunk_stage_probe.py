"""Pageable caller buffer on the chunked pipeline (regular (3,6) N = 8192 code, 16,384 words = 537 MB): the driver's own
staging of cudaMemcpyAsync (option chunk_stage = 0) against host threads with streaming stores (1)."""
import time

import numpy as np

import myldpccppapi_b200 as m

M, N, K, rp, ci = m.codes.regular_code()
ncw = 16384
y = m.synth_llr(ncw, N, 1.0, seed=1).cpu().numpy().copy()   # pageable
dec = m.Decoder(M, N, K, rp, ci)
out = {"info": np.empty((ncw, K // 8), dtype=np.uint8), "iters": np.empty((ncw,), dtype=np.int32)}
ref = None
for mode in (1, 0, 1, 0):
    dec.set_option("chunk_stage", mode)
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        dec.decode_host(y, out=out)
        ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None:
        ref = out["info"].copy()
    assert np.array_equal(ref, out["info"])
    print("chunk_stage %d: %s ms  (%.1f GB/s at the best; the kernel alone needs %.1f ms)" % (mode, " ".join("%.1f" % t for t in ts), ncw * N * 4 / min(ts) / 1e6, 86.1 * ncw / 131072))
