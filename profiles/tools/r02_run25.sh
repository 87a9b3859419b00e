#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r02_early_dims.log
: > $log
for t in "" early "" early; do
  if [ -z "$t" ]; then lib=pbrt_v2_spectral_b200/libspt.so; else lib=variants/$t/libspt.so; fi
  echo "== $t" >> $log
  SPT_LIB=$PWD/$lib python profiles/tools/quick_ranks.py >> $log 2>&1
done
SPT_LIB=$PWD/variants/early/libspt.so python profiles/tools/quick_ranks.py metal_path >> $log 2>&1
cat $log
SPT_TIMING=1 python profiles/tools/e2e_breakdown.py 2>&1 | tail -9
