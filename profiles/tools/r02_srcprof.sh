#!/bin/bash
# per-source-line instruction / stall shares of k_shade (bounce 0) and the bounce-1 trace launch
mkdir -p gpurun_out
SPT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_shade -s 8 -c 1 -f -o /tmp/prof_shade python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade.log 2>&1
ncu -i /tmp/prof_shade.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_shade.csv 2>/dev/null
python profiles/tools/src_hot.py /tmp/src_shade.csv k_shade 70 > gpurun_out/r02_src_hot_shade.txt 2>&1
SPT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_trace_multi -s 29 -c 1 -f -o /tmp/prof_trace python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1
ncu -i /tmp/prof_trace.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_trace.csv 2>/dev/null
python profiles/tools/src_hot.py /tmp/src_trace.csv k_trace 60 > gpurun_out/r02_src_hot_trace.txt 2>&1
head -40 gpurun_out/r02_src_hot_shade.txt
