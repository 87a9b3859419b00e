#!/bin/bash
# 8-GPU box: the 1 / 2 / 4 / 8 scaling series of config 1 back to back (as the driver runs it), plus the multi-GPU tests
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v23_1gpu.json 2> gpurun_out/r02_bench_v23_1gpu.err; echo "bench1 rc=$?"
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v23_${n}gpu.json 2> gpurun_out/r02_bench_v23_${n}gpu.err; echo "bench$n rc=$?"
done
python -m pytest tests/test_multi_gpu.py tests/test_dropin.py -m gpu -q > gpurun_out/r02_multi_tests_v23.log 2>&1; tail -1 gpurun_out/r02_multi_tests_v23.log
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/r02_bench_v23_%dgpu.json"%n))
        if n==1: base=d["value"]
        print(n, round(d["ms_per_step"],3), round(d["value"],1), "eff", round(d["value"]/n/base,3), "e2e", round(d["e2e"]["value"],1), d.get("film_check"))
    except Exception as e: print(n, "failed", e)
PY
