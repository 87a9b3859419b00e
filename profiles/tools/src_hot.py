#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample shares from an ncu report.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; src_hot.py src.csv [kernel-substring] [topN]"""
import csv, sys, collections
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
secs = []; cur = None; path = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur = {"fn": r[1], "file": path, "rows": [], "hdr": None}; secs.append(cur); continue
    if r[0] == "Line No": cur["hdr"] = r; continue
    if cur is not None and cur["hdr"] is not None: cur["rows"].append(r)
by_fn = collections.defaultdict(list)
for s in secs:
    by_fn[s["fn"]].append(s)
for fn, ss in by_fn.items():
    if want not in fn: continue
    lines = []
    for s in ss:
        h = s["hdr"]; ie = h.index("Instructions Executed"); sm = h.index("# Samples")
        for r in s["rows"]:
            if r[0] != "" and r[ie].isdigit():
                lines.append((int(r[ie]), int(r[sm]), s["file"], r[0], r[1].strip()[:100]))
    tot = sum(l[0] for l in lines) or 1; ts = sum(l[1] for l in lines) or 1
    print("==", fn[:60], "warp-inst", tot, "samples", ts)
    perfile = collections.Counter()
    for l in lines: perfile[l[2]] += l[0]
    print("   per file:", {k: "%.1f%%" % (100 * v / tot) for k, v in perfile.items()})
    for l in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * l[0] / tot, 100 * l[1] / ts, l[2], l[3], l[4]))
