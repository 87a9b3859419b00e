#!/bin/bash
# k_shade: warps of a block re-aligned once per vertex (__syncthreads at the loop top) so they share instruction-cache lines
mkdir -p gpurun_out
log=gpurun_out/r02_shade_sync.log
: > $log
for t in "" s2_128 s2_256 s2_512 sy512 ""; do
  if [ -z "$t" ]; then lib=pbrt_v2_spectral_b200/libspt.so; else lib=variants/$t/libspt.so; fi
  echo "== $t" >> $log
  SPT_LIB=$PWD/$lib python profiles/tools/quick_ranks.py >> $log 2>&1
done
cat $log
