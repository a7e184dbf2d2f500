#!/usr/bin/env python3
"""Per-phase (between BAR.SYNCs) breakdown of an ncu source-page CSV: samples, instructions and shared
wavefronts per warp-edge, stall mix and opcode mix.  usage: ncu_phases.py src.csv <warp_edges>"""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1]))); E=float(sys.argv[2]) if len(sys.argv)>2 else 173.0e6
hdr=rows[1]; data=rows[2:]; ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,"# Samples") for r in data)
keys=["stall_barrier","stall_math","stall_mio","stall_short_sb","stall_wait","stall_not_selected","stall_selected","stall_long_sb","stall_no_inst","stall_branch_resolving"]
print("all", {k[6:]:round(100*sum(f(r,k) for r in data)/tot,1) for k in keys})
bars=[0]+[i for i,r in enumerate(data) if 'BAR.SYNC' in r[ix["Source"]]]+[len(data)]
for a,b in zip(bars,bars[1:]):
    s=sum(f(r,"# Samples") for r in data[a:b]); n=sum(f(r,"Instructions Executed") for r in data[a:b]); w=sum(f(r,"L1 Wavefronts Shared") for r in data[a:b])
    if s/tot<0.005: continue
    st={k[6:]:round(100*sum(f(r,k) for r in data[a:b])/max(s,1),1) for k in keys}
    c=collections.Counter()
    for r in data[a:b]:
        src=[x for x in r[ix["Source"]].split() if not x.startswith('@')]
        c[src[0].split('.')[0] if src else '?']+=f(r,"Instructions Executed")
    print("lines %d-%d: samples %.1f%% inst/edge %.2f wf/edge %.2f\n   stalls %s\n   ops %s"%(a,b,100*s/tot,n/E,w/E,st,[(k,round(v/E,2)) for k,v in c.most_common(12)]))
