#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 > gpurun_out/r02_bench_v13.json 2> gpurun_out/r02_bench_v13.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_v13.err
python -c "
import json; b=json.load(open('gpurun_out/r02_bench_v13.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['fast_mode']['ms_per_step'], b['cpu_baseline']); print(json.dumps(b['roofline'])[:1500])"
python profiles/tools/rank_balance.py > gpurun_out/r02_rank_balance.log 2>&1; cat gpurun_out/r02_rank_balance.log
