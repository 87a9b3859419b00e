#!/bin/bash
# BASELINE config 5 at full size on one GPU: 10 M triangles built through the reference's API on the box, 3840x2160, 1024 spp
mkdir -p gpurun_out
nproc; free -g | head -2
python bench.py --workload synth_10m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_synth10m.json 2> gpurun_out/r02_bench_synth10m.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench_synth10m.err; cut -c1-400 gpurun_out/r02_bench_synth10m.json
