"""Two device-resident decodes of one of bench.py's sub-record workloads (ncu target for the DRAM traffic per launch:
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:ldpc_ms --launch-skip 1
--launch-count 1 ... python tools/ncu_traffic_run.py cfg3|cfg4|cfg4_4dB|family_z60); tools/ncu_traffic_json.py turns the
CSV into profiles/ncu_traffic_<workload>.json, which bench.py reads."""
import math
import os
import sys

import torch

wl = sys.argv[1]
if wl == "cfg3":
    os.environ["LDPC_B200_PLACE_EFFORT"] = "48"   # bench.py's PLACE_EFFORT for this workload
import myldpccppapi_b200 as m  # noqa: E402


def sigma_from_ebn0(db, rate):
    return math.sqrt(1.0 / (2.0 * rate * 10.0 ** (db / 10.0)))


if wl == "cfg3":
    M, N, K, rp, ci = m.codes.regular_code()
    dec, ncw, sigma = m.Decoder(M, N, K, rp, ci), 131072, 1.0
elif wl in ("cfg4", "cfg4_4dB"):
    N = 576
    dec, ncw, sigma = m.Decoder.wimax(432, 576, 4), 65536, sigma_from_ebn0(3.5 if wl == "cfg4" else 4.0, 0.75)
    dec.set_option("no_warm", 1)
    dec.reserve(4096)
    dec.set_option("qc_et", 1)   # the kernel bench.py's handle settles on in this regime (ldpc_ms_qcw_kernel)
elif wl == "family_z60":
    N = 1440
    dec, ncw, sigma = m.Decoder.wimax(1080, 1440, 4), 32768, 1.0
else:
    raise SystemExit("unknown workload " + wl)
dec.set_option("no_warm", 1)
x = m.synth_llr(ncw, N, sigma, seed=1)
for _ in range(2):
    out = dec.decode_device(x)
    torch.cuda.synchronize()
print(wl, dec.info()["path_name"], dec.info()["kernel_variant"], ncw, float(out["iters"].float().mean()))
