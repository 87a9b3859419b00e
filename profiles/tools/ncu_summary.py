#!/usr/bin/env python3
"""Per-kernel summary of an ncu --set full report as JSON (run where ncu is: on the GPU box, so that only the small summary
has to travel): first launch of each kernel class of the captured frame, i.e. bounce 0 = the launch bench.py's roofline quotes.
usage: ncu_summary.py X.ncu-rep out.json"""
import csv, json, subprocess, sys
KEYS = {"ms": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
        "l2_sectors_read": "lts__t_sectors_op_read.sum", "l2_sectors_write": "lts__t_sectors_op_write.sum", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "inst": "smsp__inst_executed.sum",
        "l2_bytes": "lts__t_bytes.sum", "l1_bytes": "l1tex__t_bytes.sum",
        "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
def cls_of(name):
    n = name.split("(")[0]
    for key, c in (("k_trace_multi", "trace"), ("k_shade", "shade"), ("k_advance", "advance"), ("k_addlight", "addlight"), ("k_film_add", "film_add"),
                   ("k_gen_camera", "gen_camera"), ("k_compact_hits", "compact_hits")):
        if key in n: return c
    return None
res, order = {}, []
for r in rows[2:]:
    c = cls_of(r[ki])
    if not c: continue
    d = {"kernel": r[ki].split("(")[0]}
    for k, m in KEYS.items():
        if m in h:
            i = h.index(m)
            try: v = float(r[i].replace(",", ""))
            except ValueError: continue
            d[k] = v * SCALE.get(units[i], 1.0)
    if "dram_read" in d: d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    if "l2_sectors_read" in d and d.get("ms"): d["l2_gbs"] = 32.0 * (d["l2_sectors_read"] + d["l2_sectors_write"]) / (d["ms"] * 1e-3) / 1e9
    if d.get("l2_bytes") and d.get("ms"): d["l2_gbs"] = d["l2_bytes"] / (d["ms"] * 1e-3) / 1e9
    if d.get("l1_bytes") and d.get("ms"): d["l1_gbs"] = d["l1_bytes"] / (d["ms"] * 1e-3) / 1e9
    if d.get("dram_bytes") and d.get("ms"): d["dram_gbs"] = d["dram_bytes"] / (d["ms"] * 1e-3) / 1e9
    order.append((c, d))
for c, d in order:                      # first launch of each class = bounce 0; later ones kept as <class>_b1 ...
    k, n = c, 1
    while k in res: k = "%s_b%d" % (c, n); n += 1
    res[k] = d
res["_source"] = sys.argv[1].split("/")[-1] + ": ncu --set full --clock-control none, one steady-state frame on one stream (SPT_LANES=1); <class> = bounce 0, <class>_bN = later launches in order"
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: {q: (round(v, 3) if isinstance(v, float) else v) for q, v in d.items() if q in ("ms", "dram_gbs", "l2_gbs", "l1_gbs", "issue_active_pct", "lanes_per_inst", "fma_pipe_pct")} for k, d in res.items() if isinstance(d, dict)}, indent=0))
