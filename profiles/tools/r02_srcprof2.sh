#!/bin/bash
# per-source-line instruction / stall shares of one kernel's launch: r02_srcprof2.sh <kernel regex> <skip> <tag> [workload]
mkdir -p gpurun_out
K=$1; S=$2; TAG=$3; WL=${4:-killeroo_path}
SPT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o /tmp/prof_$TAG python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$TAG.csv 2>/dev/null
python profiles/tools/src_hot.py /tmp/src_$TAG.csv "" 60 > gpurun_out/r02_src_hot_$TAG.txt 2>&1
python profiles/tools/ncu_key_metrics.py /tmp/prof_$TAG.ncu-rep > gpurun_out/r02_ncu_metrics_$TAG.txt 2>&1
head -45 gpurun_out/r02_src_hot_$TAG.txt | cut -c1-170
