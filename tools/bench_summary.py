#!/usr/bin/env python3
"""One-screen summary of a bench.py JSON line (tools/bench_summary.py file.json)."""
import json, sys
d = json.load(open(sys.argv[1]))
print("headline: %.3f Gbit/s  %.3f ms/step  frac %.3f  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
if d.get("sustained"): print("sustained: %.3f Gbit/s over %.2f s  %s" % (d["sustained"]["value"], d["sustained"]["seconds"], d["sustained"]["clocks"]))
if d.get("e2e"): e = d["e2e"]; print("e2e: %.3f Gbit/s %.3f ms  phases %s  h2d/gpu %s  per-rank ms %s" % (e["value"], e["ms_per_step"], {k: round(v, 3) for k, v in e["phase_ms_per_step"].items()}, e["h2d_gbs_per_gpu"], e.get("ms_per_step_per_rank")))
if d.get("e2e") and "h2d_ceiling_gbs_per_gpu" in d["e2e"]: print("   h2d ceiling/gpu %s -> bound %.3f ms/step" % (d["e2e"]["h2d_ceiling_gbs_per_gpu"], d["e2e"]["h2d_bound_ms_per_step"]))
if d.get("e2e_plugin") and d["e2e_plugin"].get("registered"): r = d["e2e_plugin"]["registered"]; print("plugin registered: %.3f Gbit/s %.3f ms  first %.2f ms match %s" % (r["value"], r["ms_per_step"], r["first_call_ms"], r["bytes_match_device_path"]))
if d.get("e2e_plugin"): p = d["e2e_plugin"]; print("plugin: %.3f Gbit/s %.3f ms  first %.2f ms  setup %.1f ms  match %s" % (p["value"], p["ms_per_step"], p["first_call_ms"], p["setup_ms"], p["bytes_match_device_path"]))
if d.get("e2e") and d["e2e"].get("packed_input"): print("   packed input:", {k: (round(v["value"], 2), round(v["ms_per_step"], 3)) if "value" in v else v for k, v in d["e2e"]["packed_input"].items()})
for k, v in (d.get("workloads") or {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    print("%s: %.3f Gbit/s  %.3f ms  iters %.2f  frac %.3f (%s, %s)  %s  oracle %s" % (k, v["value"], v["ms_per_step"], v["mean_iterations"], v["roofline"]["frac"], v["roofline"]["bound"], v["roofline"].get("kernel"), v["clocks"], v["gpu_matches_oracle_all_ranks"]))
if (d.get("workloads") or {}).get("cfg4", {}).get("e2e_by_input_format"): print("cfg4 e2e by input format:", {k: (round(v["value"], 2), round(v["ms_per_step"], 3)) for k, v in d["workloads"]["cfg4"]["e2e_by_input_format"].items()})
print("parity_all_ranks:", d.get("parity_all_ranks", {}).get("ok"), " setdevices:", d.get("setdevices"))
print("cpu_baseline:", d.get("cpu_baseline"))
