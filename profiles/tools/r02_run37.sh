#!/bin/bash
# N = 2 with the final bench (one frame ahead where a rank's frame is two waves)
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_v29_2gpu.json 2> gpurun_out/r02_bench_v29_2gpu.err; echo "bench2 rc=$?"
grep "rank" gpurun_out/r02_bench_v29_2gpu.err | head -3
python -c "
import json
j=json.loads(open('gpurun_out/r02_bench_v29_2gpu.json').read().strip().splitlines()[-1])
print(j['n_gpus'], j['value'], j['ms_per_step'], j['film_check'], j['e2e']['value'], j['config']['pipelining'][:80])"
