"""Group-of-warps kernel: one codeword per group against two / three side by side (ldpc_ms_qcm_multi_kernel), at the cap
(sigma 1.0) and in the early-termination regime.  usage: PYTHONPATH=. python tools/qcm_pack_time.py [--zs 36,44,...] [--rates 0,4]"""
import argparse
import os

import numpy as np
import torch

import myldpccppapi_b200 as m

RATES = {0: ("1/2", 1, 2), 1: ("2/3A", 2, 3), 2: ("2/3B", 2, 3), 3: ("3/4A", 3, 4), 4: ("3/4B", 3, 4), 5: ("5/6", 5, 6)}


def timed(dec, x, out, reps=4):
    for _ in range(2):
        dec.decode_device(x, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dec.decode_device(x, out=out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zs", default="28,36,44,52,56,60,68,72,76,84,88,92")
    ap.add_argument("--rates", default="0,4")
    ap.add_argument("--ncw", type=int, default=16384)
    ap.add_argument("--ebn0", default="2.5,3.5", help="comma list; empty = the cap only")
    args = ap.parse_args()
    pts = [float(v) for v in args.ebn0.split(",") if v]
    print("| z | rate | pack | words x threads per CTA | cap: Gbit/s | " + " | ".join("%.1f dB: ms (mean it)" % e for e in pts) + " | identical |")
    print("|---|---|---|---|---|" + "---|" * (len(pts) + 1))
    for z in (int(v) for v in args.zs.split(",")):
        N = 24 * z
        for rate in (int(v) for v in args.rates.split(",")):
            name, num, den = RATES[rate]
            K = N * num // den
            rng = np.random.default_rng(z)
            x_cap = m.synth_llr(args.ncw, N, 1.0, seed=z + rate)
            base = None
            for pack in (1, 2, 3):
                os.environ["LDPC_B200_QCM_PACK"] = str(pack)
                dec = m.Decoder.wimax(K, N, rate, max_iter=40)
                del os.environ["LDPC_B200_QCM_PACK"]
                dec.set_option("qcm_multi_pct", 0)
                out = dec.decode_device(x_cap)
                torch.cuda.synchronize()
                if pack > 1 and dec.info()["kernel_variant"] != 4:
                    print("| %d | %s | %d | does not fit | | | |" % (z, name, pack))
                    continue
                ms = timed(dec, x_cap, out)
                cells, same = [], True
                if base is None:
                    bits = dec.encode_device(torch.from_numpy(m.codes.pack_bits(rng.integers(0, 2, (args.ncw, K)).astype(np.uint8))).cuda())
                    base = {"bits": bits, "cap": {k: v.clone() for k, v in out.items()}, "pts": {}}
                else:
                    same = same and all(torch.equal(out[k], base["cap"][k]) for k in ("info", "iters"))
                for i, e in enumerate(pts):
                    sigma = float(np.sqrt(1.0 / (2.0 * (num / den) * 10.0 ** (e / 10.0))))
                    x = m.synth_llr(args.ncw, N, sigma, seed=500 + i, bits=base["bits"])
                    o = dec.decode_device(x)
                    t = timed(dec, x, o)
                    if i in base["pts"]:
                        same = same and all(torch.equal(o[k], base["pts"][i][k]) for k in ("info", "iters"))
                    else:
                        base["pts"][i] = {k: v.clone() for k, v in o.items()}
                    cells.append("%.3f (%.1f)" % (t, float(o["iters"].float().mean())))
                # geometry: the multi kernel's own (info() reports the single-codeword plan)
                print("| %d | %s | %d | %s | %.2f | %s | %s |" % (z, name, pack, "", args.ncw * K / ms / 1e6, " | ".join(cells), same), flush=True)
                del dec


if __name__ == "__main__":
    main()
