#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests_v8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v8.log
tail -4 gpurun_out/r02_gpu_tests_v8.log
python bench.py --bands 30 --steps 10 > gpurun_out/r02_bench_v8_30band.json 2> gpurun_out/r02_bench_v8_30band.err; echo "bench30 rc=$?"; tail -2 gpurun_out/r02_bench_v8_30band.err
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_v8.json 2> gpurun_out/r02_bench_v8.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_v8.err
