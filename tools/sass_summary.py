#!/usr/bin/env python3
"""SASS listing summary of the hot kernels (cuobjdump on the built library; no GPU needed):
registers / shared / spills and the opcode histogram of each, written to profiles/r02_sass_kernels.txt (round 1: r01_sass_kernels.txt)."""
import collections, pathlib, re, subprocess, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "myldpccppapi_b200" / "libldpc_b200.so"
WANT = [
    ("cfg1/2 default (lockstep)", r"ldpc_ms_qc_kernel<ldpc_b200::QcProfile34B<24, 8, 12>"),
    ("cfg4 default (warp per codeword)", r"ldpc_ms_qcw_kernel<ldpc_b200::QcwProfile<ldpc_b200::QcwCode34B_24>, false>"),
    ("block sizes without a compiled profile (group of warps per codeword), rate 3/4B", r"ldpc_ms_qcm_kernel<ldpc_b200::QcwCode34B_24, false>"),
    ("... with several codewords per group (z = 36, 44, 68, 72, 76)", r"ldpc_ms_qcm_multi_kernel<ldpc_b200::QcwCode34B_24>"),
    ("DecodeSP on the quasi-cyclic layout (z >= 32), rate 3/4B", r"ldpc_sp_qcm_kernel<ldpc_b200::QcwCode34B_24>"),
    ("generic on-chip, G=8 static profile", r"ldpc_ms_group_kernel<8, 16, true, 384, false, ldpc_b200::ProfileWimax34B576, false>"),
    ("cfg3 default", r"ldpc_ms_group_kernel<1, 8, false, 1024, false, ldpc_b200::ProfileRegular36N8192, true>"),
    ("cfg5 default", r"ldpc_ms_stream_kernel<1024>"),
    ("DecodeTDMP", r"ldpc_tdmp_group_kernel<4, 384>"),
    ("DecodeSP", r"ldpc_sp_group_kernel<8, 20, 384>"),
    ("warp-per-check alternative", r"ldpc_ms_warp_kernel<16>"),
    ("encoder", r"ldpc_encode_kernel"),
]
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)); funcs[cur] = []
    elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
usage = {}
rl = res.splitlines()
for i, line in enumerate(rl):
    m = re.search(r"Function (\S+):", line)
    if m and i + 1 < len(rl):
        usage[demangle(m.group(1))] = rl[i + 1].strip()
out = ["cuobjdump -sass / --dump-resource-usage of myldpccppapi_b200/libldpc_b200.so (sm_100a), tools/sass_summary.py", ""]
for label, pat in WANT:
    for name, ins in funcs.items():
        if pat in name:
            ops = collections.Counter()
            for l in ins:
                t = l.split("*/", 1)[1].strip()
                t = re.sub(r"^@!?U?P\w+\s+", "", t)
                ops[t.split()[0].rstrip(";")] += 1
            fam = collections.Counter()
            for k, v in ops.items():
                fam[k.split(".")[0]] += v
            out.append("%s -- %s" % (label, name))
            out.append("  %s" % usage.get(name, "?"))
            out.append("  %d instructions; " % len(ins) + ", ".join("%s %d" % kv for kv in fam.most_common(16)))
            keys = ["LDS", "STS", "LDG", "STG", "LDGSTS", "LDCU", "LDC", "SHFL", "VOTE", "BAR", "ATOMG", "UTMALDG", "UBLKCP"]
            out.append("  memory / cross-lane: " + ", ".join("%s %d" % (k, fam.get(k, 0)) for k in keys if fam.get(k, 0)))
            out.append("")
            break
(ROOT / "profiles" / "r02_sass_kernels.txt").write_text("\n".join(out) + "\n")
print("\n".join(out))
