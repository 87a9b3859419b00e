#!/bin/bash
# --set full capture of one steady-state frame (one lane, so the launch order is the plain wavefront order) + launch list
mkdir -p gpurun_out
TAG=r02_v6
SPT_LANES=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
SPT_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
SPT_LANES=1 timeout 1200 ncu --set full --clock-control none --import-source on -s 135 -c 34 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
