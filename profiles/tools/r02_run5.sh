#!/bin/bash
# 8 GPUs: bench N=8 and N=4, multi-device test with 4 devices
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02_bench_v5_${n}gpu.json 2> gpurun_out/r02_bench_v5_${n}gpu.err; echo "bench$n rc=$?"
grep "rank" gpurun_out/r02_bench_v5_${n}gpu.err | head -8; cut -c1-200 gpurun_out/r02_bench_v5_${n}gpu.json
done
python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "multi_render" > gpurun_out/r02_multi_tests_v5.log 2>&1; tail -3 gpurun_out/r02_multi_tests_v5.log
