#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over the smoke path: trace + shade_samples + a small render on the tiny scene
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02_sanitizer_$tool.log
  tail -4 gpurun_out/r02_sanitizer_$tool.log
done
