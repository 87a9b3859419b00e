#!/bin/bash
mkdir -p gpurun_out
for cfg in "SPT_BIN_RAYS=0" "SPT_BIN_RAYS=1" "SPT_BIN_RAYS=2"; do
  echo "== $cfg" >> gpurun_out/r02_bin_rays.log
  env $cfg python profiles/tools/quick_ranks.py >> gpurun_out/r02_bin_rays.log 2>&1
done
cat gpurun_out/r02_bin_rays.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_v11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_v11.log
tail -4 gpurun_out/r02_gpu_tests_v11.log
# DRAM traffic of the traversal kernel on the 10 M-triangle scene (HBM-resident BVH): same scene at 1920x1080, 64 spp
python bench.py --workload synth_10m_1080p64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_synth10m_1080p64.json 2> gpurun_out/r02_bench_synth10m_1080p64.err; echo "bench synth rc=$?"
SPT_LANES=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_trace_multi -s 30 -c 8 --csv --log-file gpurun_out/r02_ncu_trace_synth10m.csv python bench.py --workload synth_10m_1080p64 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 30 gpurun_out/r02_ncu_trace_synth10m.csv | cut -c1-300
