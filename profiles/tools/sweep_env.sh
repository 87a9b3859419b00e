#!/bin/bash
# usage: sweep_env.sh VAR v1 v2 ... : bench.py under each value of an environment knob
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$VAR=$v: total %.1f ms  gen %.2f path %.2f shadow %.2f mis %.2f shade %.2f acc %.2f film %.2f e2e %.1f ms chk %.0f' % (d['ms_per_step'],k['gen_camera'],k['trace_closest_path'],k['trace_any_shadow'],k['trace_closest_mis'],k['shade'],k['accumulate'],k['film_add'],d['e2e']['ms_per_step'],d['image_checksum']))"
done | tee -a gpurun_out/sweep_env.txt
