#!/bin/bash
mkdir -p gpurun_out
for cfg in "SPT_TRACE_SPEC=0" "SPT_TRACE_SPEC=1"; do
  echo "== $cfg" >> gpurun_out/r02_trace_spec.log
  env $cfg python profiles/tools/quick_ranks.py >> gpurun_out/r02_trace_spec.log 2>&1
  env $cfg python profiles/tools/quick_ranks.py synth_1m >> gpurun_out/r02_trace_spec.log 2>&1
done
cat gpurun_out/r02_trace_spec.log
SPT_TRACE_SPEC=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "first_hit or secondary or baseline or render_matches" > gpurun_out/r02_spec_tests.log 2>&1; tail -3 gpurun_out/r02_spec_tests.log
