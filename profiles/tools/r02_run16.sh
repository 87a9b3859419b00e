#!/bin/bash
# the other BASELINE configs at full size with the round-2 kernels (no CPU leg), smoke(), and the 30-band / default lines again
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
for w in metal_path ssenv_path bunny_shipped killeroo_direct bunny_path synth_1m; do
  python bench.py --workload $w --steps 3 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
  python -c "
import json,sys; b=json.load(open('gpurun_out/r02_bench_$w.json')); print('$w', round(b['value'],1), 'Msamples/s', round(b['ms_per_step'],2), 'ms  e2e', round(b['e2e']['value'],1), ' fast', round(b['fast_mode']['value'],1) if b.get('fast_mode') else None, {k: round(v,1) for k,v in b['kernel_ms_per_step'].items()})"
done
