#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r02_advance_fold.log
: > $log
for t in "" advbr "" advbr; do
  if [ -z "$t" ]; then lib=pbrt_v2_spectral_b200/libspt.so; else lib=variants/$t/libspt.so; fi
  echo "== $t" >> $log
  SPT_LIB=$PWD/$lib python profiles/tools/quick_ranks.py >> $log 2>&1
  SPT_LIB=$PWD/$lib SPT_LANES=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']; print(d['kernel_ms_per_step'], {k:round(r[k]['avg_launch_ms'],3) for k in ('advance','addlight')})" >> $log
done
cat $log
