#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r02_shade256.log
: > $log
for wl in killeroo_path metal_path ssenv_path bunny_path killeroo_direct bunny_shipped; do
  python profiles/tools/quick_ranks.py $wl >> $log 2>&1
done
python profiles/tools/e2e_breakdown.py >> $log 2>&1
cat $log
