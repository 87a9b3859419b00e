#!/bin/bash
# final code of the round: full GPU suite, smoke, default bench (v27)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_v27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v27.log
tail -3 gpurun_out/r02_gpu_tests_v27.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v27.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_v27.log
python bench.py > gpurun_out/r02_bench_v27.json 2> gpurun_out/r02_bench_v27.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_v27.err; cut -c1-300 gpurun_out/r02_bench_v27.json
