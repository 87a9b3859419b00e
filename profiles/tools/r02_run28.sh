#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r02_advance_bisect2.log
: > $log
for t in forig ch8 ch2 ""; do
  if [ -z "$t" ]; then lib=pbrt_v2_spectral_b200/libspt.so; else lib=variants/$t/libspt.so; fi
  echo "== $t" >> $log
  for lanes in 1 2; do
  SPT_LIB=$PWD/$lib SPT_LANES=$lanes python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']; print('lanes $lanes', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, {k:round(r[k]['avg_launch_ms'],3) for k in ('advance','addlight')})" >> $log
  done
done
cat $log
