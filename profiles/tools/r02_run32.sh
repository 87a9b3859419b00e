#!/bin/bash
# N GPUs (first argument): bench with three frames in flight on rotating lanes
n=${1:-2}; tag=${2:-v26}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02_bench_${tag}_${n}gpu.json 2> gpurun_out/r02_bench_${tag}_${n}gpu.err; echo "bench$n rc=$?"
grep "rank" gpurun_out/r02_bench_${tag}_${n}gpu.err | head -8; cut -c1-200 gpurun_out/r02_bench_${tag}_${n}gpu.json
python - <<P
import json
j=json.loads(open("gpurun_out/r02_bench_${tag}_${n}gpu.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["film_check"], j["e2e"]["value"])
P
