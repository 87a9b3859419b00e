#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_v17.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_v17.log
python bench.py > gpurun_out/r02_bench_v17.json 2> gpurun_out/r02_bench_v17.err
python bench.py --workload metal_path > gpurun_out/r02_bench_metal_v17.json 2> gpurun_out/r02_bench_metal_v17.err
SPT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_shade -s 48 -c 1 -f -o /tmp/prof_shade python bench.py --workload metal_path --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade_metal.log 2>&1
ncu -i /tmp/prof_shade.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_shade.csv 2>/dev/null
python profiles/tools/src_hot.py /tmp/src_shade.csv k_shade 90 > gpurun_out/r02_src_hot_shade_metal.txt 2>&1
python profiles/tools/ncu_key_metrics.py /tmp/prof_shade.ncu-rep > gpurun_out/r02_ncu_metrics_shade_metal.txt 2>&1
head -30 gpurun_out/r02_src_hot_shade_metal.txt
