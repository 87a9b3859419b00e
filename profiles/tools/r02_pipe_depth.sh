#!/bin/bash
# frames in flight x lanes the one-wave frames rotate over, on rank 0's tile set of 8 / 4 / 2 GPUs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "in_flight or tile_sets" 2>&1 | tail -3 | tee gpurun_out/r02_pipe_depth.log
for cfg in "1 2" "2 3" "3 4" "2 4" "3 2" "1 4"; do set -- $cfg
  echo "== PIPE_DEPTH=$1 (frames in flight = $(( $1 + 1 ))) SPT_PIPE_LANES=$2"
  PIPE_RANKS=8,4,2 PIPE_DEPTH=$1 SPT_PIPE_LANES=$2 python profiles/tools/pipelined_ranks.py
done 2>&1 | tee -a gpurun_out/r02_pipe_depth.log
