#!/bin/bash
# BASELINE config 5 at full size on 8 GPUs (strong scaling of a 29 s frame)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --workload synth_10m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_synth10m_8gpu.json 2> gpurun_out/r02_bench_synth10m_8gpu.err; echo "bench rc=$?"
grep "rank\|bench\]" gpurun_out/r02_bench_synth10m_8gpu.err | tail -10; cut -c1-300 gpurun_out/r02_bench_synth10m_8gpu.json
