#!/bin/bash
# final single-GPU measurements of the round: ncu summary of a steady-state frame, the other workloads, the 30-band build
mkdir -p gpurun_out
bash profiles/tools/r02_ncu2.sh r02_v22 > gpurun_out/r02_ncu2_v22.out 2>&1
tail -c 600 gpurun_out/r02_ncu2_v22.out
for wl in metal_path ssenv_path bunny_path killeroo_direct bunny_shipped synth_1m; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02_bench_${wl}_v22.json 2> gpurun_out/r02_bench_${wl}_v22.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r02_bench_${wl}_v22.json')); print('$wl', round(d['ms_per_step'],2), round(d['value'],1), round(d['e2e']['value'],1))
except Exception as e: print('$wl failed', e)"
done
python bench.py --bands 30 > gpurun_out/r02_bench_30band_v22.json 2> gpurun_out/r02_bench_30band_v22.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_30band_v22.json')); print('30band', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline'])"
