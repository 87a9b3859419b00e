#!/bin/bash
# N-GPU box: bench at every power of two up to the box's GPU count (frames pipelined over three film buffers)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v24_${n}gpu.json 2> gpurun_out/r02_bench_v24_${n}gpu.err; echo "bench$n rc=$?"
  grep "rank" gpurun_out/r02_bench_v24_${n}gpu.err | sort | head -8
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v24_${n}gpu.json')); print($n, round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('film_check'))"
done
