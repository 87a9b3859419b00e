#!/usr/bin/env python3
"""profiles/traffic.json from an ncu --set full report: mean DRAM bytes (read + write) per captured launch of
each kernel class bench.py reports a roofline for. usage: traffic_from_ncu.py X.ncu-rep"""
import csv, json, os, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki, ri, wi, ti = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ki]
    cls = None
    if name.startswith("void k_trace") and "<0" in name.split("(")[0] or "<(bool)0" in name.split(",")[0]:
        cls = "trace_closest_path"
    elif name.startswith("void k_trace"):
        cls = "trace_any_shadow"
    elif "k_shade" in name.split("(")[0]: cls = "shade"
    elif "k_accumulate" in name.split("(")[0]: cls = "accumulate"
    if not cls: continue
    b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
    if float(r[ti]) < 0.5: continue      # near-empty launches (MIS queues, the last bounce of the previous frame)
    acc.setdefault(cls, []).append(b)
res = {k: sum(v) / len(v) for k, v in acc.items()}
res["_source"] = os.path.basename(sys.argv[1]) + ": mean over the captured launches (bounces 0-1 of one frame), dram__bytes_read.sum + dram__bytes_write.sum"
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
