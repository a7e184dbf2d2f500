"""The reference's whole code family (Coder::initCheckMatrix, MyLdpc.cpp:52-109: N = 24 z, z = 24 .. 96 step 4, six
rates): device-resident decode of 16,384 words at sigma 1.0 (every word runs the 40-iteration cap), CUDA events.
Prints info Gbit/s and the kernel the plan chose.  usage: PYTHONPATH=. python tools/family_time.py [--ncw 16384]"""
import argparse

import torch

import myldpccppapi_b200 as m

RATES = [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3), (3, "3/4A", 3, 4), (4, "3/4B", 3, 4), (5, "5/6", 5, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncw", type=int, default=16384)
    ap.add_argument("--zs", default=",".join(str(z) for z in range(24, 97, 4)))
    args = ap.parse_args()
    print("| z | N | " + " | ".join(name for _, name, _, _ in RATES) + " | kernel (words per CTA x threads) |")
    print("|---|---|" + "---|" * (len(RATES) + 1))
    for z in (int(v) for v in args.zs.split(",")):
        N = 24 * z
        cells, shape = [], ""
        for rate, name, num, den in RATES:
            K = N * num // den
            dec = m.Decoder.wimax(K, N, rate, max_iter=40)
            x = m.synth_llr(args.ncw, N, 1.0, seed=z + rate)
            out = dec.decode_device(x)
            for _ in range(2):
                dec.decode_device(x, out=out)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4):
                dec.decode_device(x, out=out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 4
            inf = dec.info()
            cells.append("%.2f" % (args.ncw * K / ms / 1e6))
            shape = "%s %dx%d" % (inf["path_name"], inf["codewords_per_cta"], inf["threads_per_cta"])
            del dec, x, out
        print("| %d | %d | %s | %s |" % (z, N, " | ".join(cells), shape))


if __name__ == "__main__":
    main()
