#!/bin/bash
# first GPU session of round 2: parity tests, bench line, A/B of the trace-kernel knobs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gpu_tests_v1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v1.log
tail -5 gpurun_out/r02_gpu_tests_v1.log
python bench.py --steps 10 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; echo "bench rc=$?"
for cfg in "" "SPT_STACK_SMEM=0" "SPT_STACK_SMEM=16" "SPT_MERGE_TRACE=0" "SPT_FETCH_THRESHOLD=8" "SPT_FETCH_THRESHOLD=20" "SPT_FETCH_THRESHOLD=28"; do
  env $cfg python profiles/tools/quick_ranks.py >> gpurun_out/r02_ranks_v1.log 2>&1
done
cat gpurun_out/r02_ranks_v1.log
env python profiles/tools/quick_ranks.py synth_1m >> gpurun_out/r02_ranks_v1.log 2>&1
