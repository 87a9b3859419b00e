#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v3.log
tail -4 gpurun_out/r02_gpu_tests_v3.log
python bench.py --steps 10 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_v3.err
python profiles/tools/quick_ranks.py > gpurun_out/r02_ranks_v3.log 2>&1; cat gpurun_out/r02_ranks_v3.log
