#!/bin/bash
# N GPUs: frames ahead x frame layout (SPT_PIPE_MODE=0: two half-size waves per frame; 1: one wave, lanes rotate)
n=${1:-2}
mkdir -p gpurun_out
for cfg in "1 1" "2 0" "1 0" "3 1" "2 1"; do set -- $cfg
echo "== frames ahead $1, SPT_PIPE_MODE=$2"
SPT_PIPE_MODE=$2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 3 --frames-ahead $1 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; echo "rc=$?"
grep "rank" gpurun_out/tmp.err | head -8
python - <<P
import json
j=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["film_check"]["ok"], j["e2e"]["value"])
P
done 2>&1 | tee gpurun_out/r02_ahead_modes_${n}gpu.log
