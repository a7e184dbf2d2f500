"""ncu CSV (one kernel launch, metrics dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum) ->
profiles/ncu_traffic_<workload>.json.  usage: python tools/ncu_traffic_json.py workload codewords file.csv"""
import csv
import json
import pathlib
import sys

wl, ncw, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
head = rows[0]
ix = {k: head.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
last = max(int(r[head.index("ID")]) for r in rows[1:])   # the last captured launch (earlier ones: warm-up decode)
val, kern = {}, None
for r in rows[1:]:
    if int(r[head.index("ID")]) != last:
        continue
    kern = r[ix["Kernel Name"]]
    val[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
out = {"workload": wl, "kernel": kern, "codewords_per_launch": ncw, "dram_bytes_read": val["dram__bytes_read.sum"],
       "dram_bytes_write": val["dram__bytes_write.sum"], "launch_ms_under_ncu": val.get("gpu__time_duration.sum"),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one launch of tools/ncu_traffic_run.py %s (CSV: profiles/r02_ncu_traffic_%s.csv)" % (wl, wl)}
pathlib.Path("profiles/ncu_traffic_%s.json" % wl).write_text(json.dumps(out, indent=1) + "\n")
print(out)
