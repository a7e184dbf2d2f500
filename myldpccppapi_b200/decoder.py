"""Host-side Python front-end of the C-ABI: a handle class plus a mirror of the reference's
`Coder` decode interface (same method names, argument meaning and size rules as
reference MyLdpc.h:104-129 / MyLdpc.cpp:571-631), so the parity tests read like Test.cpp.

torch is used only for device memory and streams; all compute happens inside
libldpc_b200.so.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import lib as _lib
from .lib import Info, LdpcError, Timing, check

# enum rate_type / decodeType, reference MyLdpc.h:33-39
rate_1_2, rate_2_3_a, rate_2_3_b, rate_3_4_a, rate_3_4_b, rate_5_6 = range(6)
DecodeCPU, DecodeMS, DecodeSP, DecodeTDMP, DecodeTDMPCL, DecodeMSCL = range(6)
RATE_BY_NAME = {"1/2": 0, "2/3A": 1, "2/3B": 2, "3/4A": 3, "3/4B": 4, "5/6": 5}


def wimax_csr(K: int, N: int, rate: int):
    """H of the reference's 802.16e code (Coder::initCheckMatrix) as (row_ptr, col_idx, M)."""
    L = _lib.load()
    M, nnz = C.c_int(), C.c_int()
    check(L.ldpc_b200_wimax_csr(K, N, rate, None, None, C.byref(M), C.byref(nnz)))
    rp = np.zeros(M.value + 1, dtype=np.int32)
    ci = np.zeros(nnz.value, dtype=np.int32)
    check(L.ldpc_b200_wimax_csr(K, N, rate, rp.ctypes.data, ci.ctypes.data, C.byref(M), C.byref(nnz)))
    return rp, ci, M.value


def edge_tables(M: int, N: int, row_ptr, col_idx):
    """(col_ptr, vn_edge, max_row_weight, max_col_weight) as the kernels use them."""
    L = _lib.load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    cp = np.zeros(N + 1, dtype=np.int32)
    ve = np.zeros(max(ci.size, 1), dtype=np.uint32)
    rw, cw = C.c_int(), C.c_int()
    check(L.ldpc_b200_edge_tables(M, N, rp.ctypes.data, ci.ctypes.data, cp.ctypes.data, ve.ctypes.data,
                                  C.byref(rw), C.byref(cw)))
    return cp, ve[:ci.size], rw.value, cw.value


class Decoder:
    """One decoder handle on one GPU (ldpc_b200_create / _destroy)."""

    def __init__(self, M: int, N: int, K: int, row_ptr, col_idx, device: int = 0, max_iter: int = 40,
                 early_termination: bool = True):
        self._L = _lib.load()
        self._h = C.c_void_p()
        rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        if rp.shape != (M + 1,):
            raise ValueError("row_ptr must have M+1 entries")
        check(self._L.ldpc_b200_create(C.byref(self._h), M, N, K, rp.ctypes.data, ci.ctypes.data, device))
        self.M, self.N, self.K, self.device = M, N, K, device
        self.KB, self.NB = (K + 7) // 8, (N + 7) // 8
        check(self._L.ldpc_b200_set_max_iter(self._h, max_iter))
        check(self._L.ldpc_b200_set_early_termination(self._h, 1 if early_termination else 0))

    @classmethod
    def wimax(cls, K: int, N: int, rate: int, device: int = 0, **kw) -> "Decoder":
        rp, ci, M = wimax_csr(K, N, rate)
        self = cls(M, N, K, rp, ci, device=device, **kw)
        self.set_layer_height(N // 24)  # as ldpc_b200_create_wimax does
        return self

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._L.ldpc_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ---------------------------------------------------------------
    def set_max_iter(self, n: int) -> None:
        check(self._L.ldpc_b200_set_max_iter(self._h, n))

    def set_early_termination(self, on: bool) -> None:
        check(self._L.ldpc_b200_set_early_termination(self._h, 1 if on else 0))

    def set_algorithm(self, alg: int) -> None:
        """0 = min-sum (default), 1 = probability-domain sum-product (DecodeSP), 2 = layered min-sum (DecodeTDMP),
        3 / 4 = the reference's fused kernels decodeOnceMS / decodeOnceTDMP with their own arithmetic (exact, any-size kernel)."""
        check(self._L.ldpc_b200_set_algorithm(self._h, alg))

    def set_layer_height(self, z: int) -> None:
        """Rows per layer for the layered decoder (Decoder.wimax sets N/24)."""
        check(self._L.ldpc_b200_set_layer_height(self._h, z))

    def set_path(self, path: int) -> None:
        check(self._L.ldpc_b200_set_path(self._h, path))

    def set_option(self, name: str, value: int) -> None:
        """Host-pipeline experiment switch (ldpc_b200_set_option), e.g. "stream_chunk", "no_streamed"."""
        check(self._L.ldpc_b200_set_option(self._h, name.encode(), int(value)))

    def reserve(self, batch: int) -> None:
        check(self._L.ldpc_b200_reserve(self._h, batch))

    def info(self) -> dict:
        inf = Info()
        check(self._L.ldpc_b200_get_info(self._h, C.byref(inf)))
        return inf.asdict()

    def timing(self, reset: bool = False) -> dict:
        """Phase timers of decode_host accumulated so far (ldpc_b200_get_timing): calls, codewords, wall_s, h2d_s,
        kernel_s, d2h_s."""
        t = Timing()
        check(self._L.ldpc_b200_get_timing(self._h, C.byref(t)))
        if reset:
            check(self._L.ldpc_b200_reset_timing(self._h))
        return t.asdict()

    @property
    def launches(self) -> int:
        return int(self._L.ldpc_b200_launch_count(self._h))

    # -- decode ----------------------------------------------------------------------
    def decode_device(self, llr, want_hard: bool = False, want_post: bool = False, want_iters: bool = True,
                      out: Optional[dict] = None, stream=None):
        """llr: CUDA float32 tensor [ncw, N] on this decoder's device.  Returns a dict of CUDA
        tensors {info[ncw,KB] u8, iters[ncw] i32, hard[ncw,NB] u8, post[ncw,N] f32}.  Async."""
        import torch

        if not llr.is_cuda or llr.dtype != torch.float32 or not llr.is_contiguous():
            raise ValueError("llr must be a contiguous CUDA float32 tensor")
        if llr.device.index != self.device:
            raise ValueError("llr lives on another device than this decoder")
        ncw = llr.numel() // self.N
        dev = llr.device
        out = out if out is not None else {}
        if "info" not in out:
            out["info"] = torch.empty((ncw, self.KB), dtype=torch.uint8, device=dev)
        if want_iters and "iters" not in out:
            out["iters"] = torch.empty((ncw,), dtype=torch.int32, device=dev)
        if want_hard and "hard" not in out:
            out["hard"] = torch.empty((ncw, self.NB), dtype=torch.uint8, device=dev)
        if want_post and "post" not in out:
            out["post"] = torch.empty((ncw, self.N), dtype=torch.float32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        check(self._L.ldpc_b200_decode_device(self._h, llr.data_ptr(), ncw, ptr("info"), ptr("hard"), ptr("iters"),
                                              ptr("post"), C.c_void_p(st.cuda_stream)))
        return out

    # -- encode (Coder::forEncoder / Coder::encode on the device) ----------------------
    def encode_device(self, info):
        """info: CUDA uint8 tensor [ncw, K/8] (bits LSB-first) -> CUDA uint8 tensor [ncw, N/8] of codewords in
        the reference's priorCode layout."""
        import torch

        assert info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()
        ncw = info.shape[0]
        out = torch.empty((ncw, self.N // 8), dtype=torch.uint8, device=info.device)
        st = torch.cuda.current_stream(info.device)
        check(self._L.ldpc_b200_encode_device(self._h, info.data_ptr(), ncw, out.data_ptr(), C.c_void_p(st.cuda_stream)))
        return out

    def encode_host(self, info: np.ndarray) -> np.ndarray:
        info = np.ascontiguousarray(info, dtype=np.uint8).reshape(-1, self.K // 8)
        out = np.empty((info.shape[0], self.N // 8), dtype=np.uint8)
        check(self._L.ldpc_b200_encode_host(self._h, info.ctypes.data, info.shape[0], out.ctypes.data))
        return out

    def decode_host(self, llr: np.ndarray, want_hard: bool = False, want_post: bool = False, out: Optional[dict] = None):
        """llr: host float32 array/tensor [ncw, N] (pinned memory gives full-speed copies).
        Returns numpy arrays {info, iters, hard?, post?}.  Blocking."""
        a = _as_host_array(llr, np.float32)
        ncw = _numel(a) // self.N
        out = out if out is not None else {}
        if "info" not in out:
            out["info"] = np.empty((ncw, self.KB), dtype=np.uint8)
        if "iters" not in out:
            out["iters"] = np.empty((ncw,), dtype=np.int32)
        if want_hard and "hard" not in out:
            out["hard"] = np.empty((ncw, self.NB), dtype=np.uint8)
        if want_post and "post" not in out:
            out["post"] = np.empty((ncw, self.N), dtype=np.float32)
        ptr = lambda k: _host_ptr(out[k]) if k in out else None  # noqa: E731
        check(self._L.ldpc_b200_decode_host(self._h, _host_ptr(a), ncw, ptr("info"), ptr("hard"), ptr("iters"), ptr("post")))
        return out

    def decode_host_packed(self, llr, scale: float = 1.0, want_hard: bool = False, want_post: bool = False, out: Optional[dict] = None):
        """decode_host with the channel values as float16 or int8 (numpy array or pinned torch tensor), widened on the device
        to (float)x * scale: half / a quarter of the host-to-device bytes (ldpc_b200_decode_host_packed)."""
        a, fmt = _packed_host_array(llr)
        ncw = _numel(a) // self.N
        out = out if out is not None else {}
        if "info" not in out:
            out["info"] = np.empty((ncw, self.KB), dtype=np.uint8)
        if "iters" not in out:
            out["iters"] = np.empty((ncw,), dtype=np.int32)
        if want_hard and "hard" not in out:
            out["hard"] = np.empty((ncw, self.NB), dtype=np.uint8)
        if want_post and "post" not in out:
            out["post"] = np.empty((ncw, self.N), dtype=np.float32)
        ptr = lambda k: _host_ptr(out[k]) if k in out else None  # noqa: E731
        check(self._L.ldpc_b200_decode_host_packed(self._h, _host_ptr(a), fmt, C.c_float(scale), ncw, ptr("info"), ptr("hard"),
                                                   ptr("iters"), ptr("post")))
        return out


def _packed_host_array(x):
    """(array or pinned tensor, format code) for decode_host_packed: float16 -> 1, int8 -> 2."""
    if hasattr(x, "data_ptr"):
        import torch

        if x.is_cuda:
            raise ValueError("expected a host tensor")
        fmt = {torch.float16: 1, torch.int8: 2}.get(x.dtype)
        if fmt is None:
            raise ValueError("expected a float16 or int8 tensor, got %s" % x.dtype)
        return x.contiguous(), fmt
    a = np.ascontiguousarray(x)
    fmt = {np.dtype(np.float16): 1, np.dtype(np.int8): 2}.get(a.dtype)
    if fmt is None:
        raise ValueError("expected float16 or int8 data, got %s" % a.dtype)
    return a, fmt


def _as_host_array(x, dtype):
    if hasattr(x, "data_ptr"):  # torch CPU tensor (possibly pinned)
        import torch

        if x.is_cuda:
            raise ValueError("expected a host tensor")
        want = {np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
        if x.dtype != want:
            raise ValueError("expected a %s tensor, got %s" % (want, x.dtype))
        return x.contiguous()
    a = np.asarray(x)
    if a.dtype != dtype:
        if a.dtype.kind != np.dtype(dtype).kind:
            raise ValueError("expected %s data, got %s" % (np.dtype(dtype), a.dtype))
        a = a.astype(dtype)  # float64 -> float32 etc.: an explicit, rounding conversion
    return np.ascontiguousarray(a)


def _numel(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(x.size)


def _host_ptr(x):
    return x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data


def synth_llr(ncw: int, N: int, sigma: float, seed: int, device: int = 0, bits=None, out=None):
    """BPSK + AWGN channel values generated on the GPU (ldpc_b200_synth_llr).  `bits`:
    optional CUDA uint8 tensor [ncw, ceil(N/8)] of packed codeword bits (LSB first)."""
    import torch

    L = _lib.load()
    dev = torch.device("cuda", device)
    if out is None:
        out = torch.empty((ncw, N), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev)
    check(L.ldpc_b200_synth_llr(out.data_ptr(), ncw, N, float(sigma), int(seed) & (2**64 - 1),
                                bits.data_ptr() if bits is not None else None, device, C.c_void_p(st.cuda_stream)))
    return out


class Coder:
    """ctypes binding of the drop-in C++ `class Coder` (include/MyLdpc.h, libmyldpc_b200.so) through its C doorway
    include/MyLdpc_c.h -- same method names, argument meaning and return values as the reference's Coder
    (MyLdpc.h:104-129).  There is ONE implementation of the Coder logic (csrc/mycoder.cpp); this class only marshals.

    coder = Coder(ldpcK, ldpcN, rate); coder.forDecoder(batchSize); coder.addDecodeType(DecodeMS)
    coder.decode(postCode, srcCode, srcLength, DecodeMS)      # -> 0 like the reference

    DecodeSP runs the sum-product kernel, DecodeTDMP / DecodeTDMPCL the layered kernel, DecodeCPU / DecodeMS /
    DecodeMSCL flooding min-sum with Coder::decodeCPU's semantics (DecodeMSCL with the reference kernel's cap of 120);
    there is no CPU path.  Additive: from_csr(), setMaxIter(), setDevices(), setStrictDecodeType(), lastIterations,
    lastAlgorithm, lastStepTimes.
    """

    def __init__(self, ldpcK: int, ldpcN: int, rate: int, device: int = 0, _csr=None):
        self._L = _lib.load_coder()
        self.ldpcK, self.ldpcN = int(ldpcK), int(ldpcN)
        if _csr is None:
            self.ldpcM, self.rate = self.ldpcN - self.ldpcK, int(rate)
            self._c = self._L.myldpc_coder_new(self.ldpcK, self.ldpcN, self.rate)
        else:
            M, rp, ci = _csr
            self.ldpcM, self.rate = int(M), None
            rp = np.ascontiguousarray(rp, dtype=np.int32)
            ci = np.ascontiguousarray(ci, dtype=np.int32)
            self._c = self._L.myldpc_coder_new_csr(self.ldpcM, self.ldpcN, self.ldpcK, rp.ctypes.data, ci.ctypes.data)
        if not self._c:
            raise LdpcError(-4, "Coder construction failed")
        self.device = device
        if device != 0:
            self.setDevices([device])
        self.lastIterations = None

    @classmethod
    def from_csr(cls, M: int, N: int, K: int, row_ptr, col_idx, device: int = 0) -> "Coder":
        return cls(K, N, -1, device=device, _csr=(M, row_ptr, col_idx))

    def close(self) -> None:
        if getattr(self, "_c", None):
            self._L.myldpc_coder_free(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> int:
        if rc != 0:
            raise LdpcError(rc, self.lastError())
        return rc

    @property
    def checkMatrix(self):
        """(row_ptr, col_idx) of the public member checkMatrix (MyLdpc.h:128)."""
        rows, cols, nnz = C.c_int(), C.c_int(), C.c_int()
        self._L.myldpc_checkMatrix(self._c, C.byref(rows), C.byref(cols), C.byref(nnz), None, None)
        rp = np.zeros(rows.value + 1, dtype=np.int32)
        ci = np.zeros(nnz.value, dtype=np.int32)
        self._L.myldpc_checkMatrix(self._c, None, None, None, rp.ctypes.data, ci.ctypes.data)
        return rp, ci

    # reference MyLdpc.cpp:137-165 / :167-305 / :307-552
    def forEncoder(self) -> int:
        return self._check(self._L.myldpc_forEncoder(self._c))

    def forDecoder(self, batchSize: int) -> int:
        return self._check(self._L.myldpc_forDecoder(self._c, int(batchSize)))

    def addDecodeType(self, deType: int) -> int:
        return self._check(self._L.myldpc_addDecodeType(self._c, int(deType)))

    # [B200] additive
    def setMaxIter(self, times: int) -> int:
        return self._check(self._L.myldpc_setMaxIter(self._c, int(times)))

    def setDevices(self, devices) -> int:
        d = np.ascontiguousarray(list(devices), dtype=np.int32)
        return self._check(self._L.myldpc_setDevices(self._c, d.ctypes.data, d.size))

    def setEarlyTermination(self, on: bool) -> int:
        return self._check(self._L.myldpc_setEarlyTermination(self._c, 1 if on else 0))

    def setStrictDecodeType(self, strict: bool) -> int:
        return self._check(self._L.myldpc_setStrictDecodeType(self._c, 1 if strict else 0))

    def setRegisterHostBuffers(self, on: bool) -> int:
        return self._check(self._L.myldpc_setRegisterHostBuffers(self._c, 1 if on else 0))

    def setFusedKernelArithmetic(self, exact: bool) -> int:
        return self._check(self._L.myldpc_setFusedKernelArithmetic(self._c, 1 if exact else 0))

    @property
    def lastAlgorithm(self) -> int:
        return int(self._L.myldpc_lastAlgorithm(self._c))

    def lastError(self) -> str:
        return (self._L.myldpc_lastError(self._c) or b"").decode("utf-8", "replace")

    def lastStepTimes(self) -> dict:
        t = (C.c_double * 4)()
        n = self._L.myldpc_lastStepTimes(self._c, t, 4)
        return dict(zip(("wall_s", "h2d_s", "kernel_s", "d2h_s"), list(t)[:n]))

    # reference MyLdpc.cpp:620-631
    def getCodeSize(self, srcLength: int) -> int:
        return int(self._L.myldpc_getCodeSize(self._c, int(srcLength)))

    def getPostCodeLength(self, srcLength: int) -> int:
        return int(self._L.myldpc_getPostCodeLength(self._c, int(srcLength)))

    def getPriorCodeLength(self, srcLength: int) -> int:
        return int(self._L.myldpc_getPriorCodeLength(self._c, int(srcLength)))

    # reference MyLdpc.cpp:554-569
    def encode(self, srcCode, priorCode, srcLength: int) -> int:
        src = np.ascontiguousarray(srcCode, dtype=np.uint8)
        if priorCode.dtype != np.uint8 or priorCode.size < self.getPriorCodeLength(srcLength):
            raise ValueError("priorCode must be a uint8 array of getPriorCodeLength(srcLength) bytes")
        return self._check(self._L.myldpc_encode(self._c, src.ctypes.data, priorCode.ctypes.data, int(srcLength)))

    # reference MyLdpc.cpp:1061-1078 (rand()-based Box-Muller: seed it with libc srand)
    def test(self, priorCode, postCode, priorCodeLength: int, sd: float) -> int:
        pr = np.ascontiguousarray(priorCode, dtype=np.uint8)
        if postCode.dtype != np.float32 or postCode.size < 8 * priorCodeLength:
            raise ValueError("postCode must be a float32 array of 8 * priorCodeLength values")
        return self._check(self._L.myldpc_test(self._c, pr.ctypes.data, postCode.ctypes.data, int(priorCodeLength), float(sd)))

    # reference MyLdpc.cpp:571-618 (+ :684-784 for the semantics)
    def decode(self, postCode, srcCode, srcLength: int, deType: int = DecodeMS) -> int:
        """postCode: host float32 array / CPU tensor of getPostCodeLength(srcLength) values; srcCode: writable uint8
        buffer of at least srcLength bytes.  Returns 0, or raises LdpcError with the C++ Coder's code and message."""
        y = _as_host_array(postCode, np.float32)
        if _numel(y) < self.getPostCodeLength(srcLength):
            raise ValueError("postCode shorter than getPostCodeLength(srcLength)")
        dst = np.frombuffer(srcCode, dtype=np.uint8) if not isinstance(srcCode, np.ndarray) else srcCode.view(np.uint8)
        if dst.size < srcLength or not dst.flags.c_contiguous:
            raise ValueError("srcCode must be a contiguous buffer of at least srcLength bytes")
        rc = self._L.myldpc_decode(self._c, _host_ptr(y), dst.ctypes.data, int(srcLength), int(deType))
        if rc != 0:
            raise LdpcError(rc, self.lastError())
        n = int(self._L.myldpc_lastCodeSize(self._c))
        self.lastIterations = np.ctypeslib.as_array(self._L.myldpc_lastIterations(self._c), shape=(n,)).copy()
        return 0
