#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests_v9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v9.log
tail -4 gpurun_out/r02_gpu_tests_v9.log; grep "fast traversal\|merl" gpurun_out/r02_gpu_tests_v9.log | head
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_v9.json 2> gpurun_out/r02_bench_v9.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_v9.err
python -c "
import json; b=json.load(open('gpurun_out/r02_bench_v9.json')); print(b['value'], b['ms_per_step'], b['fast_mode'])"
bash profiles/tools/r02_ncu2.sh r02_v9
