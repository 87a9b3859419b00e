#!/bin/bash
# 8 GPUs: the default bench first (as the driver runs it on a fresh box), then the round's earlier protocol, then the default again
n=${1:-8}
mkdir -p gpurun_out
i=0
for cfg in "0 1" "1 0" "0 1"; do set -- $cfg; i=$((i+1))
echo "== run $i: --frames-ahead $1 (0 = default), SPT_PIPE_MODE=$2"
SPT_PIPE_MODE=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$i bench.py --gpus $n --steps 20 --warmup 3 --frames-ahead $1 > gpurun_out/r02_bench_v26_${n}gpu_run$i.json 2> gpurun_out/r02_bench_v26_${n}gpu_run$i.err; echo "rc=$?"
grep "rank" gpurun_out/r02_bench_v26_${n}gpu_run$i.err | head -8
python - <<P
import json
j=json.loads(open("gpurun_out/r02_bench_v26_${n}gpu_run$i.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["film_check"], j["e2e"]["value"])
P
done 2>&1 | tee gpurun_out/r02_ahead_modes_${n}gpu.log
