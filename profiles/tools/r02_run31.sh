#!/bin/bash
# full GPU suite + default bench with the alternating-lane pipelined frames (v25)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_v25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v25.log
tail -4 gpurun_out/r02_gpu_tests_v25.log
python bench.py > gpurun_out/r02_bench_v25.json 2> gpurun_out/r02_bench_v25.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_v25.err; cut -c1-400 gpurun_out/r02_bench_v25.json
