#!/usr/bin/env python3
"""Cluster the SASS of an .ncu-rep (first kernel) into runs of equal execution count and print, per run, its share of
the executed instructions and of the stall samples with the top stall reasons.  tools/ncu_blocks.py rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr, data = rows[1], rows[2:]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
blocks, cur = [], None
for r in data:
    try: n = int(r[ia]); s = int(r[isamp])
    except Exception: continue
    st = {k: int(r[hdr.index(k)] or 0) for k in stalls}
    if cur and cur["n"] == n:
        cur["len"] += 1; cur["samples"] += s; cur["last"] = r[isrc].strip()
        for k in stalls: cur["st"][k] += st[k]
    else:
        cur = {"n": n, "len": 1, "samples": s, "first": r[isrc].strip(), "last": r[isrc].strip(), "idx": len(blocks), "st": st}
        blocks.append(cur)
tot = sum(b["n"] * b["len"] for b in blocks); totS = sum(b["samples"] for b in blocks)
print("instructions executed", tot, "samples", totS)
for b in blocks:
    w = b["n"] * b["len"]
    if b["samples"] / totS > thr or w / tot > thr:
        top = sorted(b["st"].items(), key=lambda kv: -kv[1])[:3]
        print("%4d exec %9d len %4d inst %5.1f%% samples %5.1f%% %s | %s ... %s" % (b["idx"], b["n"], b["len"], 100 * w / tot, 100 * b["samples"] / totS, [(k[6:], v) for k, v in top], b["first"][:34], b["last"][:34]))
