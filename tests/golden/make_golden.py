#!/usr/bin/env python3
"""Generate tests/golden/default_code_golden.npz from the REFERENCE ITSELF.

Run in the build container only (needs oracle/_ref, i.e. /root/reference mounted at build time):
    python tests/golden/make_golden.py
Inputs are seeded BPSK-AWGN channel values for Test.cpp's code (z=24, N=576, K=432, rate_3_4_b,
Test.cpp:19-26).  Expected outputs come from the reference's own Coder::decode(..., DecodeCPU)
compiled unmodified (oracle/_ref/libmyldpc_ref.so, -O0 like the reference's Makefile):
  * ref_bytes[s]        : srcCode after decodeCPU with the reference's cap (times = 40)
  * ref_bytes_cap[s][t] : srcCode after decodeCPU with times = t+1, t = 0..39 -- the hard decisions
                          after every iteration, from which the stopping iteration follows
The oracle's iteration counts / posteriors are stored next to them (oracle_*) for the GPU tests;
tests/test_golden.py checks that they are consistent with the reference-derived arrays.
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from oracle import ref  # noqa: E402

N, K, RATE = 576, 432, 4
SIGMAS = [0.5, 0.57, 0.62, 1.0]
NCW = 16


def main():
    assert ref.available(), "oracle/_ref is not built (needs /root/reference)"
    rp, ci, M = oracle.wimax_H(N, "3/4B")
    o = oracle.Oracle(M, N, K, rp, ci, times=40)
    from myldpccppapi_b200 import codes
    Gp = codes.gf2_systematic_encoder(M, N, K, rp, ci).astype(np.int64)
    out = {"sigmas": np.array(SIGMAS, dtype=np.float32), "row_ptr": rp, "col_idx": ci}
    src_len = NCW * K // 8
    for s, sigma in enumerate(SIGMAS):
        rng = np.random.default_rng(1000 + s)
        if s == 0:
            bits = np.zeros((NCW, N), dtype=np.float32)          # all-zero codeword
        elif s == len(SIGMAS) - 1:
            bits = rng.integers(0, 2, (NCW, N)).astype(np.float32)  # arbitrary signs, not codewords
        else:
            u = rng.integers(0, 2, (NCW, K)).astype(np.int64)     # random valid codewords
            bits = np.concatenate([u, u @ Gp % 2], axis=1).astype(np.float32)
        llr = ((1.0 - 2.0 * bits) + sigma * rng.standard_normal((NCW, N))).astype(np.float32)
        c = ref.RefCoder(K, N, RATE)
        ref_bytes = c.decode_cpu(llr, src_len)
        caps = np.zeros((40, src_len), dtype=np.uint8)
        for t in range(40):
            c.set_times(t + 1)
            caps[t] = c.decode_cpu(llr, src_len)
        info, iters, hard, post = o.decode(llr)
        out["llr_%d" % s] = llr
        out["ref_bytes_%d" % s] = ref_bytes
        out["ref_bytes_cap_%d" % s] = caps
        out["oracle_iters_%d" % s] = iters
        out["oracle_hard_%d" % s] = np.packbits(hard, axis=1, bitorder="little")
        out["oracle_post_%d" % s] = post
        print("sigma", sigma, "mean iters", iters.mean(), "ref==oracle", np.array_equal(ref_bytes, info.reshape(-1)))
    path = pathlib.Path(__file__).with_name("default_code_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
