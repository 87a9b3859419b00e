#!/bin/bash
# bench.py with the frame loop moved into multi.pipelined_frames: N = 1 (short) and N = 2
mkdir -p gpurun_out
timeout 100 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_v28.json 2> gpurun_out/r02_bench_v28.err; echo "bench1 rc=$?"; tail -1 gpurun_out/r02_bench_v28.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_v28_2gpu.json 2> gpurun_out/r02_bench_v28_2gpu.err; echo "bench2 rc=$?"
grep "rank" gpurun_out/r02_bench_v28_2gpu.err | head -3
python - <<P
import json
for f in ("gpurun_out/r02_bench_v28.json", "gpurun_out/r02_bench_v28_2gpu.json"):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(j["n_gpus"], j["value"], j["ms_per_step"], j["film_check"], j["e2e"]["value"], j["gpu_launches"])
P
