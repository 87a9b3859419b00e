#!/bin/bash
# --set full capture of the first 16 product kernels of a steady-state frame (bounces 0-2, one stream) + launch list; only the
# small summaries travel back
mkdir -p gpurun_out
TAG=${1:-r02_v9}
SPT_LANES=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
SPT_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 131 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
SPT_LANES=1 timeout 900 ncu --set full --metrics lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none -k regex:^k_ -s 131 -c 16 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
python profiles/tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_summary_$TAG.json > gpurun_out/ncu_summary_$TAG.txt 2>&1
python profiles/tools/ncu_key_metrics.py /tmp/prof_$TAG.ncu-rep > gpurun_out/ncu_metrics_$TAG.txt 2>&1
ls -la /tmp/prof_$TAG.ncu-rep; sz=$(stat -c %s /tmp/prof_$TAG.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
tail -3 gpurun_out/ncu_$TAG.log; head -c 1500 gpurun_out/ncu_summary_$TAG.txt
