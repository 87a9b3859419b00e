#!/bin/bash
# second GPU session: warp modes of the trace kernel; the two heavy-tailed image checks with the reference-vs-reference floor
mkdir -p gpurun_out
for cfg in "SPT_TRACE_MODE=0" "SPT_TRACE_MODE=1" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=4" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=8" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=12" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=16" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=8 SPT_STACK_SMEM=8" "SPT_TRACE_MODE=0 SPT_LANES=1" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=8 SPT_LANES=1"; do
  echo "== $cfg" >> gpurun_out/r02_ranks_v2.log
  env $cfg python profiles/tools/quick_ranks.py >> gpurun_out/r02_ranks_v2.log 2>&1
done
for cfg in "SPT_TRACE_MODE=0" "SPT_TRACE_MODE=2 SPT_LEAF_BATCH=8"; do
  echo "== $cfg" >> gpurun_out/r02_ranks_v2.log
  env $cfg python profiles/tools/quick_ranks.py synth_1m >> gpurun_out/r02_ranks_v2.log 2>&1
done
cat gpurun_out/r02_ranks_v2.log
python -m pytest tests/test_image_parity.py -m gpu -x -q -s -k "metal_shipped_small or bunny_measured_small" > gpurun_out/r02_image_floor_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_image_floor_v2.log
grep -v "^$" gpurun_out/r02_image_floor_v2.log | tail -30
