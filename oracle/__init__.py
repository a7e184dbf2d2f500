"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/liboracle.so (plain-C restatement of the reference's CPU min-sum
decode, reference MyLdpc.cpp:52-109,167-222,620-631,684-784) and, when it has been built,
oracle/_ref/libmyldpc_ref.so (the reference's own MyLdpc.cpp compiled unmodified).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under myldpccppapi_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None
_REF = None

RATE = {"1/2": 0, "2/3A": 1, "2/3B": 2, "3/4A": 3, "3/4B": 4, "5/6": 5}


def build(verbose: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    r = subprocess.run(["make", "-C", str(_HERE), "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        p32 = C.POINTER(C.c_int32)
        L.oracle_wimax_H.argtypes = [C.c_int, C.c_int, C.POINTER(p32), C.POINTER(p32), C.POINTER(C.c_int)]
        L.oracle_wimax_H.restype = C.c_int
        L.oracle_tables_build.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_tables_build.restype = C.c_void_p
        L.oracle_tables_free.argtypes = [C.c_void_p]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_decodeCPU.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_decodeCPU.restype = C.c_int
        L.oracle_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_decode_batch.restype = C.c_int
        L.oracle_bpsk.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_decode_sp_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_decode_sp_batch.restype = C.c_int
        L.oracle_decode_tdmp_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_decode_tdmp_batch.restype = C.c_int
        L.oracle_decode_fused_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_decode_fused_batch.restype = C.c_int
        L.oracle_tdmp_layering_ok.argtypes = [C.c_void_p, C.c_int]
        L.oracle_tdmp_layering_ok.restype = C.c_int
        L.oracle_sp_expf.argtypes = [C.c_float]
        L.oracle_sp_expf.restype = C.c_float
        for f in ("oracle_getCodeSize",):
            getattr(L, f).argtypes = [C.c_int, C.c_int]
        for f in ("oracle_getPostCodeLength", "oracle_getPriorCodeLength"):
            getattr(L, f).argtypes = [C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def wimax_H(N: int, rate) -> tuple[np.ndarray, np.ndarray, int]:
    """CSR (row_ptr, col_idx) and M of the reference's H for code length N (initCheckMatrix)."""
    L = lib()
    rate = RATE.get(rate, rate)
    rp, ci, M = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int()
    nnz = L.oracle_wimax_H(N, int(rate), C.byref(rp), C.byref(ci), C.byref(M))
    if nnz < 0:
        raise ValueError("bad rate")
    row_ptr = np.ctypeslib.as_array(rp, shape=(M.value + 1,)).copy()
    col_idx = np.ctypeslib.as_array(ci, shape=(nnz,)).copy()
    L.oracle_free(rp)
    L.oracle_free(ci)
    return row_ptr, col_idx, M.value


class Oracle:
    """Reference-semantics CPU min-sum decoder over a CSR parity-check matrix."""

    def __init__(self, M: int, N: int, K: int, row_ptr, col_idx, times: int = 40):
        self.M, self.N, self.K, self.times = int(M), int(N), int(K), int(times)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        assert self.row_ptr.shape == (self.M + 1,)
        self._t = lib().oracle_tables_build(self.M, self.N, self.row_ptr.ctypes.data, self.col_idx.ctypes.data)

    def __del__(self):
        try:
            if self._t:
                lib().oracle_tables_free(self._t)
                self._t = None
        except Exception:
            pass

    def decode_stream(self, post_code: np.ndarray, src_length: int):
        """Coder::decode(..., DecodeCPU) restated: returns (srcCode bytes, iters, hard, post)."""
        L = lib()
        ncw = L.oracle_getCodeSize(self.K, src_length)
        y = np.ascontiguousarray(post_code, dtype=np.float32).reshape(-1)
        assert y.size >= ncw * self.N
        out = np.zeros(src_length + 1, dtype=np.uint8)
        iters = np.zeros(ncw, dtype=np.int32)
        hard = np.zeros((ncw, self.N), dtype=np.uint8)
        post = np.zeros((ncw, self.N), dtype=np.float32)
        L.oracle_decodeCPU(self._t, self.K, self.times, y.ctypes.data, out.ctypes.data, src_length,
                           iters.ctypes.data, hard.ctypes.data, post.ctypes.data)
        return out[:src_length], iters, hard, post

    def decode(self, llr: np.ndarray, threads: int = 0, literal: bool = True, want_post: bool = True,
               want_hard: bool = True):
        """Per-codeword records: (info[ncw, ceil(K/8)], iters[ncw], hard[ncw,N], post[ncw,N])."""
        L = lib()
        y = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, self.N)
        ncw = y.shape[0]
        if threads <= 0:
            threads = os.cpu_count() or 1
        info = np.zeros((ncw, (self.K + 7) // 8), dtype=np.uint8)
        iters = np.zeros(ncw, dtype=np.int32)
        hard = np.zeros((ncw, self.N), dtype=np.uint8) if want_hard else None
        post = np.zeros((ncw, self.N), dtype=np.float32) if want_post else None
        L.oracle_decode_batch(self._t, self.K, self.times, y.ctypes.data, ncw, info.ctypes.data,
                              iters.ctypes.data, hard.ctypes.data if want_hard else None,
                              post.ctypes.data if want_post else None, threads, 1 if literal else 0)
        return info, iters, hard, post


def decode_sp(o: "Oracle", llr: np.ndarray, threads: int = 0):
    """Sum-product restatement (see ldpc_oracle.h): (info, iters, hard, post0, post1)."""
    L = lib()
    y = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, o.N)
    ncw = y.shape[0]
    if threads <= 0:
        threads = os.cpu_count() or 1
    info = np.zeros((ncw, (o.K + 7) // 8), dtype=np.uint8)
    iters = np.zeros(ncw, dtype=np.int32)
    hard = np.zeros((ncw, o.N), dtype=np.uint8)
    p0 = np.zeros((ncw, o.N), dtype=np.float32)
    p1 = np.zeros((ncw, o.N), dtype=np.float32)
    L.oracle_decode_sp_batch(o._t, o.K, o.times, y.ctypes.data, ncw, info.ctypes.data, iters.ctypes.data,
                             hard.ctypes.data, p0.ctypes.data, p1.ctypes.data, threads)
    return info, iters, hard, p0, p1


def decode_tdmp(o: "Oracle", llr: np.ndarray, z: int, threads: int = 0):
    """Layered min-sum restatement (see ldpc_oracle.h): (info, iters, hard, post); layers of z rows."""
    L = lib()
    y = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, o.N)
    ncw = y.shape[0]
    if threads <= 0:
        threads = os.cpu_count() or 1
    info = np.zeros((ncw, (o.K + 7) // 8), dtype=np.uint8)
    iters = np.zeros(ncw, dtype=np.int32)
    hard = np.zeros((ncw, o.N), dtype=np.uint8)
    post = np.zeros((ncw, o.N), dtype=np.float32)
    rc = L.oracle_decode_tdmp_batch(o._t, o.K, o.times, int(z), y.ctypes.data, ncw, info.ctypes.data,
                                    iters.ctypes.data, hard.ctypes.data, post.ctypes.data, threads)
    if rc:
        raise ValueError("layers of %d rows are not column-disjoint" % z)
    return info, iters, hard, post


def decode_fused(o: "Oracle", llr: np.ndarray, z: int, layered: bool, times: int | None = None, threads: int = 0):
    """The reference's fused OpenCL kernels restated (see ldpc_oracle.h): decodeOnceMS (layered=False, cap 120 in the
    reference) or decodeOnceTDMP (layered=True, cap 40).  Returns (info, iters, hard, post)."""
    L = lib()
    y = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, o.N)
    ncw = y.shape[0]
    if threads <= 0:
        threads = os.cpu_count() or 1
    if times is None:
        times = 40 if layered else 120  # decodeCL.c:344, 479
    info = np.zeros((ncw, (o.K + 7) // 8), dtype=np.uint8)
    iters = np.zeros(ncw, dtype=np.int32)
    hard = np.zeros((ncw, o.N), dtype=np.uint8)
    post = np.zeros((ncw, o.N), dtype=np.float32)
    rc = L.oracle_decode_fused_batch(o._t, o.K, int(times), int(z), 1 if layered else 0, y.ctypes.data, ncw,
                                     info.ctypes.data, iters.ctypes.data, hard.ctypes.data, post.ctypes.data, threads)
    if rc:
        raise ValueError("layers of %d rows are not column-disjoint" % z)
    return info, iters, hard, post


def bpsk(bytes_: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(bytes_, dtype=np.uint8)
    out = np.empty(b.size * 8, dtype=np.float32)
    lib().oracle_bpsk(b.ctypes.data, b.size, out.ctypes.data)
    return out
