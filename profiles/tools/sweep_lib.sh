#!/bin/bash
# usage: sweep_lib.sh lib1.so lib2.so ... : bench.py (serialized kernel times) under each build of libspt.so
for lib in "$@"; do
  SPT_LIB=$PWD/$lib timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$lib: total %.2f ms  gen %.2f path %.2f shadow %.2f mis %.2f shade %.2f acc %.2f film %.2f e2e %.1f ms chk %.0f' % (d['ms_per_step'],k['gen_camera'],k['trace_closest_path'],k['trace_any_shadow'],k['trace_closest_mis'],k['shade'],k['accumulate'],k['film_add'],d['e2e']['ms_per_step'],d['image_checksum']))"
done | tee -a gpurun_out/sweep_lib.txt
