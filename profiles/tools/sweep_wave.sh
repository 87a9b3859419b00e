#!/bin/bash
# bench.py at several wave sizes (pixels per wave); prints per-class kernel times
mkdir -p gpurun_out
for wp in "$@"; do
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --wave-pixels $wp 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('wave_pixels $wp: total %.1f ms  path %.2f shadow %.2f mis %.2f shade %.2f acc %.2f film %.2f e2e %.1f ms launches %d' % (d['ms_per_step'],k['trace_closest_path'],k['trace_any_shadow'],k['trace_closest_mis'],k['shade'],k['accumulate'],k['film_add'],d['e2e']['ms_per_step'],d['gpu_launches']))"
done | tee -a gpurun_out/sweep_wave.txt
