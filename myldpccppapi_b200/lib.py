"""ctypes binding of the C-ABI in include/ldpc_b200.h (libldpc_b200.so).

There is no CPU fallback anywhere in this package: if the shared library is missing it is
built (nvcc, in-tree); if that fails, or a compute call finds no CUDA device, it raises.
"""
from __future__ import annotations

import ctypes as C
import re

from . import _build

OK = 0
PATH_LANE_SMEM, PATH_LANE_GLOBAL, PATH_LANE16, PATH_GROUP, PATH_CLUSTER, PATH_STREAM, PATH_QC, PATH_WARP = 0, 1, 3, 4, 5, 6, 7, 8
PATH_NAMES = {0: "lane_smem", 1: "lane_global", 3: "lane16", 4: "group", 5: "cluster", 6: "stream", 7: "qc", 8: "warp"}


class LdpcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("ldpc_b200 error %d: %s" % (code, msg))
        self.code = code


class Info(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("nnz", C.c_int),
        ("max_row_weight", C.c_int), ("max_col_weight", C.c_int),
        ("max_iter", C.c_int), ("early_termination", C.c_int),
        ("device", C.c_int), ("sm_count", C.c_int),
        ("path", C.c_int), ("threads_per_cta", C.c_int), ("ctas", C.c_int),
        ("codewords_per_cta", C.c_int),
        ("smem_bytes", C.c_size_t), ("workspace_bytes", C.c_size_t), ("table_bytes", C.c_size_t),
        ("kernel_variant", C.c_int), ("et_available", C.c_int),
    ]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["path_name"] = PATH_NAMES.get(d["path"], "?")
        return d


class Timing(C.Structure):
    _fields_ = [("calls", C.c_int64), ("codewords", C.c_int64), ("wall_s", C.c_double), ("h2d_s", C.c_double),
                ("kernel_s", C.c_double), ("d2h_s", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_vp, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/ldpc_b200.h declares
SIGNATURES = {
    "ldpc_b200_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _vp, _i]),
    "ldpc_b200_create_wimax": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "ldpc_b200_destroy": (_i, [_vp]),
    "ldpc_b200_set_max_iter": (_i, [_vp, _i]),
    "ldpc_b200_set_early_termination": (_i, [_vp, _i]),
    "ldpc_b200_set_path": (_i, [_vp, _i]),
    "ldpc_b200_set_algorithm": (_i, [_vp, _i]),
    "ldpc_b200_set_layer_height": (_i, [_vp, _i]),
    "ldpc_b200_set_option": (_i, [_vp, C.c_char_p, C.c_longlong]),
    "ldpc_b200_encoder_init": (_i, [_vp]),
    "ldpc_b200_encode_device": (_i, [_vp, _vp, C.c_int64, _vp, _vp]),
    "ldpc_b200_encode_host": (_i, [_vp, _vp, C.c_int64, _vp]),
    "ldpc_b200_get_info": (_i, [_vp, C.POINTER(Info)]),
    "ldpc_b200_get_csr": (_i, [_vp, _vp, _vp]),
    "ldpc_b200_reserve": (_i, [_vp, _i64]),
    "ldpc_b200_decode_device": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "ldpc_b200_decode_host": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ldpc_b200_decode_host_packed": (_i, [_vp, _vp, _i, _f, _i64, _vp, _vp, _vp, _vp]),
    "ldpc_b200_synth_llr": (_i, [_vp, _i64, _i, _f, _u64, _vp, _i, _vp]),
    "ldpc_b200_wimax_csr": (_i, [_i, _i, _i, _vp, _vp, _pi, _pi]),
    "ldpc_b200_edge_tables": (_i, [_i, _i, _vp, _vp, _vp, _vp, _pi, _pi]),
    "ldpc_b200_probe_smem_bandwidth": (_i, [_i, C.POINTER(C.c_double)]),
    "ldpc_b200_launch_count": (_i64, [_vp]),
    "ldpc_b200_host_alloc": (_vp, [C.c_size_t]),
    "ldpc_b200_host_free": (_i, [_vp]),
    "ldpc_b200_get_timing": (_i, [_vp, C.POINTER(Timing)]),
    "ldpc_b200_reset_timing": (_i, [_vp]),
    "ldpc_b200_last_error": (C.c_char_p, []),
    "ldpc_b200_version": (C.c_char_p, []),
}

_LIB = None


def header_symbols() -> list[str]:
    """Function names declared in include/ldpc_b200.h (used by the export test)."""
    text = (_build.ROOT / "include" / "ldpc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_b200_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load (building if needed) libldpc_b200.so.  Raises if it cannot be had."""
    global _LIB
    if _LIB is None:
        so = _build.build_cuda()
        if not so.exists():
            raise RuntimeError("libldpc_b200.so is missing and could not be built; no CPU fallback exists")
        L = C.CDLL(str(so))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


# name -> (restype, argtypes) of include/MyLdpc_c.h (libmyldpc_b200.so)
_pint = C.POINTER(C.c_int)
CODER_SIGNATURES = {
    "myldpc_coder_new": (_vp, [_i, _i, _i]),
    "myldpc_coder_new_csr": (_vp, [_i, _i, _i, _vp, _vp]),
    "myldpc_coder_free": (None, [_vp]),
    "myldpc_forEncoder": (_i, [_vp]),
    "myldpc_forDecoder": (_i, [_vp, _i]),
    "myldpc_addDecodeType": (_i, [_vp, _i]),
    "myldpc_encode": (_i, [_vp, _vp, _vp, _i]),
    "myldpc_decode": (_i, [_vp, _vp, _vp, _i, _i]),
    "myldpc_test": (_i, [_vp, _vp, _vp, _i, _f]),
    "myldpc_getPriorCodeLength": (_i, [_vp, _i]),
    "myldpc_getPostCodeLength": (_i, [_vp, _i]),
    "myldpc_getCodeSize": (_i, [_vp, _i]),
    "myldpc_checkMatrix": (_i, [_vp, _pint, _pint, _pint, _vp, _vp]),
    "myldpc_setMaxIter": (_i, [_vp, _i]),
    "myldpc_setDevices": (_i, [_vp, _vp, _i]),
    "myldpc_setEarlyTermination": (_i, [_vp, _i]),
    "myldpc_setStrictDecodeType": (_i, [_vp, _i]),
    "myldpc_setFusedKernelArithmetic": (_i, [_vp, _i]),
    "myldpc_setRegisterHostBuffers": (_i, [_vp, _i]),
    "myldpc_lastAlgorithm": (_i, [_vp]),
    "myldpc_lastIterations": (C.POINTER(C.c_int32), [_vp]),
    "myldpc_lastCodeSize": (_i, [_vp]),
    "myldpc_lastError": (C.c_char_p, [_vp]),
    "myldpc_lastStepTimes": (_i, [_vp, C.POINTER(C.c_double), _i]),
}
_CODER_LIB = None


def coder_header_symbols() -> list[str]:
    text = (_build.ROOT / "include" / "MyLdpc_c.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(myldpc_[A-Za-z0-9_]+)\s*\(", text)))


def load_coder() -> C.CDLL:
    """Load (building if needed) libmyldpc_b200.so, the drop-in C++ Coder, through its C doorway."""
    global _CODER_LIB
    if _CODER_LIB is None:
        load()
        so = _build.build_coder()
        if not so.exists():
            raise RuntimeError("libmyldpc_b200.so is missing and could not be built")
        L = C.CDLL(str(so))
        for name, (res, args) in CODER_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _CODER_LIB = L
    return _CODER_LIB


def check(rc: int) -> None:
    if rc != OK:
        raise LdpcError(rc, load().ldpc_b200_last_error().decode("utf-8", "replace"))
