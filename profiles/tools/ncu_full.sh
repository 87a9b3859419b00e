#!/bin/bash
# One --set full capture of each kernel of a steady-state frame (bounces 0-1), after a plain run exited 0.
# Captures run with SPT_LANES=1 (the frame as ONE wave on one stream) so the launch sequence is the plain
# K1, (K2, compaction, K5, K3, K2, K6) x bounces, K7 and -s lands on the same kernels every time.
# usage: ncu_full.sh <tag> [skip] [count]
TAG=$1; SKIP=${2:-80}; CNT=${3:-14}
mkdir -p gpurun_out
SPT_LANES=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
SPT_LANES=1 ncu --set full --clock-control none --import-source on -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
SPT_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/bench_$TAG.json
