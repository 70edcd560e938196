#!/bin/bash
# Round-end measurements on one B200 (run through gpurun): contract bench, ncu launch list of the same command, and one
# `ncu --set full` capture of a real env.step launch.  Usage: bash scripts/profile_round.sh r01i
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json | head -c 300; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-pairs > $out/${tag}_bench_under_ncu.log 2>&1
# index (among cloth_step_kernel launches) of the last launch that ran longer than 100 ms: a timed env.step
idx=$(python - <<PY
import csv
rows = list(csv.reader(open("$out/${tag}_launches.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
n = -1; last = -1
for r in rows[h + 1:]:
    if len(r) != len(hdr) or "cloth_step_kernel" not in r[k]:
        continue
    n += 1
    t = float(r[v].replace(",", "")); t = t / 1e6 if r[u] == "ns" else (t / 1e3 if r[u] in ("us", "usecond") else t)
    if t > 100.0:
        last = n
print(last)
PY
)
echo "full capture of cloth_step_kernel launch #$idx"
ncu --set full --clock-control none --import-source on -k regex:cloth_step_kernel --launch-skip $idx --launch-count 1 -f -o $out/${tag}_step_kernel \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-pairs > $out/${tag}_full_capture.log 2>&1
ncu -i $out/${tag}_step_kernel.ncu-rep --page raw --csv > $out/${tag}_step_kernel_raw.csv 2>/dev/null
ls -la $out | grep ${tag}
# second source for the shared-memory roofline denominator: ncu's own wavefront counter on the LDS.128 microbenchmark
ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.per_second,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.max \
    --clock-control none -k regex:smem_bw_kernel --csv --log-file $out/${tag}_smem_microbench_ncu.csv python - > $out/${tag}_smem_microbench.log 2>&1 <<PY
import ctypes as C, torch
from gym_cloth_b200 import lib as L
torch.zeros(1, device="cuda")
lib = L.lib(); v = C.c_double(0)
lib.clothb200_bench_smem_bandwidth(2000, C.byref(v), None)
print("clothb200_bench_smem_bandwidth: %.1f GB/s" % v.value)
PY
tail -3 $out/${tag}_smem_microbench.log
