#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "SPT_FORCE_LANES=1" "SPT_FORCE_LANES=3" "SPT_FORCE_LANES=4"; do
  echo "== $cfg" >> gpurun_out/r02_lanes.log
  env $cfg python profiles/tools/quick_ranks.py >> gpurun_out/r02_lanes.log 2>&1
done
cat gpurun_out/r02_lanes.log
