#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r02_shade512.log
: > $log
for t in sy512 sy384; do
echo "== $t" >> $log
for wl in metal_path bunny_shipped bunny_path killeroo_direct; do
  SPT_LIB=$PWD/variants/$t/libspt.so python profiles/tools/quick_ranks.py $wl >> $log 2>&1
done
done
cat $log
