#!/bin/bash
# --set full capture of the launches of ONE kernel (name regex) on a given bench workload, after a plain run exited 0;
# plus the launch list of one frame. SPT_LANES=1: one stream, plain wavefront order.
# usage: ncu_kernel.sh <tag> <workload> <kernel regex> <launch-skip> <launch-count> [list-skip] [list-count]
TAG=$1; WL=$2; KREG=$3; SKIP=$4; CNT=$5; LSKIP=${6:-0}; LCNT=${7:-200}
mkdir -p gpurun_out
SPT_LANES=1 python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
SPT_LANES=1 ncu --set full --clock-control none --import-source on -k regex:$KREG -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
SPT_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -s $LSKIP -c $LCNT --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/ncu_$TAG.log
