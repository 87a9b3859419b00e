#!/bin/bash
mkdir -p gpurun_out
for t in "" tr9 tr10 tr6; do
  if [ -z "$t" ]; then lib=pbrt_v2_spectral_b200/libspt.so; else lib=variants/$t/libspt.so; fi
  echo "== $t" >> gpurun_out/r02_trace_occupancy.log
  SPT_LIB=$PWD/$lib python profiles/tools/quick_ranks.py >> gpurun_out/r02_trace_occupancy.log 2>&1
done
cat gpurun_out/r02_trace_occupancy.log
