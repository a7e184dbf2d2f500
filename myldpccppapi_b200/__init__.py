"""myldpccppapi_b200 -- B200-native LDPC min-sum decoder behind wing02/MyLdpcCppApi's API.

Scope is the reference's decode hot path only: parity-check-matrix
setup, channel values in, flooding min-sum on sm_100a CUDA kernels, packed hard decisions
out.  The native code lives in csrc/ and is reached through the C-ABI in
include/ldpc_b200.h (libldpc_b200.so); include/MyLdpc.h is the drop-in C++ `Coder`.
"""
from .lib import LdpcError, load  # noqa: F401
from .decoder import (  # noqa: F401
    Coder, Decoder, synth_llr, wimax_csr, edge_tables,
    rate_1_2, rate_2_3_a, rate_2_3_b, rate_3_4_a, rate_3_4_b, rate_5_6,
    DecodeCPU, DecodeMS, DecodeSP, DecodeTDMP, DecodeTDMPCL, DecodeMSCL,
)
from . import codes  # noqa: F401
from . import shard  # noqa: F401
from .shard import bind_to_gpu_numa_node, shard_range, shard_ranges  # noqa: F401

__all__ = ["Coder", "Decoder", "synth_llr", "wimax_csr", "edge_tables", "codes", "LdpcError", "load"]
