#!/bin/bash
# compute-sanitizer over the time-sliced step kernel (memcheck, synccheck, racecheck); logs -> gpurun_out/<tag>_sanitizer_*.log
# usage (on the GPU box): bash scripts/sanitize.sh <tag> [per-tool timeout s]
tag=${1:-r02}; lim=${2:-420}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  n=8; extra=""; [ $tool = racecheck ] && { n=5; extra="quick"; }
  timeout $lim compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_workload.py $n 3 8 $extra > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/${tag}_sanitizer_$tool.log
  tail -4 gpurun_out/${tag}_sanitizer_$tool.log
done
