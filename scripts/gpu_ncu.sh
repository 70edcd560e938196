#!/bin/bash
# ncu --set full capture (with source correlation) of ONE env.step launch of the contract bench's workload.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:cloth_step_kernel --launch-count 1 -f \
    -o $out/${tag}_step python scripts/quick_bench.py 4096 benchphases f32 > $out/${tag}_ncu.log 2>&1
tail -4 $out/${tag}_ncu.log
ncu -i $out/${tag}_step.ncu-rep --page raw --csv > $out/${tag}_step_raw.csv 2>/dev/null
ls -la $out/${tag}_step*
