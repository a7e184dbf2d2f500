"""oracle.ref -- TEST INFRASTRUCTURE ONLY.

ctypes door into oracle/_ref/libmyldpc_ref*.so: the reference's own MyLdpc.cpp compiled
unmodified (oracle/Makefile) against container-only Eigen / OpenCL stand-ins (oracle/shim/).
In libmyldpc_ref{,_O2}.so only Coder's host paths are meaningful: initCheckMatrix, forDecoder's edge tables,
decode(..., DecodeCPU), forEncoder/encode, test.  libmyldpc_refcl.so (opt="cl") additionally compiles the
reference's decodeCL.c unmodified and EXECUTES its OpenCL decode variants on the CPU (oracle/shim/cl_exec.h,
cl_kernels.cpp): RefCoder.decode_cl.  Prebuilt files travel to the GPU box;
/root/reference itself is never read at run time.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import threading

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIBS = {}


def _path(opt: str) -> pathlib.Path:
    return _HERE / "_ref" / {"O0": "libmyldpc_ref.so", "O2": "libmyldpc_ref_O2.so", "cl": "libmyldpc_refcl.so"}[opt]


def available(opt: str = "O0") -> bool:
    return _path(opt).exists()


def lib(opt: str = "O0") -> C.CDLL:
    if opt not in _LIBS:
        L = C.CDLL(str(_path(opt)))
        vp, i = C.c_void_p, C.c_int
        L.ref_coder_new.restype = vp
        L.ref_coder_new.argtypes = [i, i, i]
        L.ref_coder_free.argtypes = [vp]
        for f in ("ref_forDecoder", "ref_getCodeSize", "ref_getPostCodeLength", "ref_getPriorCodeLength"):
            getattr(L, f).argtypes = [vp, i]
        L.ref_forEncoder.argtypes = [vp]
        L.ref_set_times.argtypes = [vp, i]
        L.ref_decode_cpu.argtypes = [vp, vp, vp, i]
        L.ref_encode.argtypes = [vp, vp, vp, i]
        L.ref_nnz.argtypes = [vp]
        L.ref_M.argtypes = [vp]
        L.ref_csr.argtypes = [vp, vp, vp]
        L.ref_edge_tables.argtypes = [vp] * 8
        L.ref_test.argtypes = [vp, vp, vp, i, C.c_float]
        L.ref_addDecodeType.argtypes = [vp, i]
        L.ref_decode.argtypes = [vp, vp, vp, i, i, vp, i]
        L.ref_stepTime.argtypes = [vp, i]
        L.ref_stepTime.restype = C.c_double
        L.ref_z.argtypes = [vp]
        _LIBS[opt] = L
    return _LIBS[opt]


class RefCoder:
    """The reference's `Coder` (MyLdpc.h:104-129), CPU paths only."""

    def __init__(self, K: int, N: int, rate: int, opt: str = "O0"):
        self.L = lib(opt)
        self.K, self.N, self.rate = K, N, rate
        self.h = self.L.ref_coder_new(K, N, rate)
        self.M = self.L.ref_M(self.h)
        self._dec = False
        self._enc = False

    def __del__(self):
        try:
            if self.h:
                self.L.ref_coder_free(self.h)
                self.h = None
        except Exception:
            pass

    def csr(self):
        nnz = self.L.ref_nnz(self.h)
        rp = np.zeros(self.M + 1, dtype=np.int32)
        ci = np.zeros(nnz, dtype=np.int32)
        self.L.ref_csr(self.h, rp.ctypes.data, ci.ctypes.data)
        return rp, ci

    def forDecoder(self, batch: int = 1):
        if not self._dec:
            self.L.ref_forDecoder(self.h, batch)
            self._dec = True

    def edge_tables(self):
        self.forDecoder()
        nnz = self.L.ref_nnz(self.h)
        a = [np.zeros(n, dtype=np.int32) for n in (nnz, nnz, self.M, nnz, self.N, nnz, self.M + 1)]
        self.L.ref_edge_tables(self.h, *[x.ctypes.data for x in a])
        return dict(zip(("hRows", "hCols", "hRowFirstPtr", "hRowNextPtr", "hColFirstPtr", "hColNextPtr", "hRowRange"), a))

    def set_times(self, times: int):
        self.L.ref_set_times(self.h, times)

    def decode_cpu(self, post_code: np.ndarray, src_length: int) -> np.ndarray:
        """Coder::decode(postCode, srcCode, srcLength, DecodeCPU) -> srcCode[:srcLength]."""
        self.forDecoder()
        y = np.ascontiguousarray(post_code, dtype=np.float32).reshape(-1)
        need = self.L.ref_getPostCodeLength(self.h, src_length)
        assert y.size >= need
        out = np.zeros(src_length + 1, dtype=np.uint8)  # the reference may touch srcCode[srcLength]
        self.L.ref_decode_cpu(self.h, y.ctypes.data, out.ctypes.data, src_length)
        return out[:src_length]

    def decode_cl(self, post_code: np.ndarray, src_length: int, de_type: int, batch: int = 1):
        """Coder::decode(postCode, srcCode, srcLength, deType) for the reference's OpenCL variants
        (DecodeMS=1, DecodeSP=2, DecodeTDMP=3, DecodeTDMPCL=4, DecodeMSCL=5), executed on the CPU by
        oracle/shim/cl_exec.h -- needs opt="cl".  Follows Test.cpp's call order: forDecoder(batch),
        addDecodeType(deType), decode.  Returns (srcCode[:srcLength], times) where times holds the "Time="
        value the reference printed for every chunk of `batch` words (batch = 1: its per-word iteration counts;
        empty for the two fused kernels, which print nothing)."""
        assert self.L is lib("cl"), "decode_cl needs RefCoder(..., opt='cl')"
        if not self._dec or getattr(self, "_batch", None) != batch:
            assert not self._dec, "forDecoder was already called with another batch size"
            self.L.ref_forDecoder(self.h, batch)
            self._dec, self._batch, self._types = True, batch, set()
        if de_type not in self._types:
            self.L.ref_addDecodeType(self.h, de_type)
            self._types.add(de_type)
        y = np.ascontiguousarray(post_code, dtype=np.float32).reshape(-1)
        assert y.size >= self.L.ref_getPostCodeLength(self.h, src_length)
        out = np.zeros(src_length + 1, dtype=np.uint8)
        nchunks = (self.L.ref_getCodeSize(self.h, src_length) + batch - 1) // batch
        times = np.zeros(nchunks + 1, dtype=np.int32)
        n = self.L.ref_decode(self.h, y.ctypes.data, out.ctypes.data, src_length, de_type, times.ctypes.data, times.size)
        return out[:src_length], times[:min(n, nchunks)]

    def encode(self, src: np.ndarray) -> np.ndarray:
        """Coder::encode -> priorCode bytes (getPriorCodeLength(srcLength))."""
        if not self._enc:
            self.L.ref_forEncoder(self.h)
            self._enc = True
        s = np.ascontiguousarray(src, dtype=np.uint8)
        n = self.L.ref_getPriorCodeLength(self.h, s.size)
        out = np.zeros(n + 8, dtype=np.uint8)
        self.L.ref_encode(self.h, s.ctypes.data, out.ctypes.data, s.size)
        return out[:n]

    def bpsk(self, prior: np.ndarray) -> np.ndarray:
        """Coder::test with sd = 0: the reference's bit -> +/-1.0 map (noise term is exactly 0)."""
        p = np.ascontiguousarray(prior, dtype=np.uint8)
        out = np.zeros(p.size * 8, dtype=np.float32)
        self.L.ref_test(self.h, p.ctypes.data, out.ctypes.data, p.size, 0.0)
        return out


def decode_cpu_parallel(K: int, N: int, rate: int, llr: np.ndarray, times: int, threads: int, opt: str = "O2") -> np.ndarray:
    """All host threads: one reference Coder per thread, each decoding a contiguous slice of
    codewords with the reference's single-threaded decodeCPU.  Returns info bytes [ncw, K/8]."""
    y = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, N)
    ncw = y.shape[0]
    kb = K // 8
    out = np.zeros((ncw, kb), dtype=np.uint8)
    threads = max(1, min(threads, ncw))
    bounds = [ncw * t // threads for t in range(threads + 1)]

    def work(t):
        a, b = bounds[t], bounds[t + 1]
        if b <= a:
            return
        c = RefCoder(K, N, rate, opt=opt)
        c.set_times(times)
        out[a:b] = c.decode_cpu(y[a:b], (b - a) * kb).reshape(b - a, kb)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out
