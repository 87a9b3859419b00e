#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s > gpurun_out/r02_gpu_tests_v18.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_v18.log
log=gpurun_out/r02_pow_sincos.log
: > $log
for wl in killeroo_path metal_path bunny_path killeroo_direct bunny_shipped ssenv_path; do
  python profiles/tools/quick_ranks.py $wl >> $log 2>&1
done
cat $log
