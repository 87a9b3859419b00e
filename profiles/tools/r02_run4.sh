#!/bin/bash
# 2 GPUs: multi-device tests + bench N=2 (film shared over peer memory)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo2.log 2>&1
python -m pytest tests/test_multi_gpu.py tests/test_dropin.py -m gpu -x -q -s > gpurun_out/r02_multi_tests_v4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_multi_tests_v4.log
tail -15 gpurun_out/r02_multi_tests_v4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_v4_2gpu.json 2> gpurun_out/r02_bench_v4_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02_bench_v4_2gpu.err; cat gpurun_out/r02_bench_v4_2gpu.json | cut -c1-600
