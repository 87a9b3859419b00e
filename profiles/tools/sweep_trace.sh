#!/bin/bash
# bench.py under each traversal variant / leaf_wait setting; prints the per-class kernel times
# usage: sweep_trace.sh "<variant> <leaf_wait>" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  SPT_TRACE_VARIANT=$1 SPT_LEAF_WAIT=$2 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('variant $1 leaf_wait $2: total %.1f ms  path %.2f shadow %.2f mis %.2f shade %.2f acc %.2f  e2e %.1f ms chk %.0f' % (d['ms_per_step'],k['trace_closest_path'],k['trace_any_shadow'],k['trace_closest_mis'],k['shade'],k['accumulate'],d['e2e']['ms_per_step'],d['image_checksum']))"
done | tee -a gpurun_out/sweep_trace.txt
