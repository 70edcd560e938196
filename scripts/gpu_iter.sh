#!/bin/bash
# One kernel-iteration measurement on the GPU box: parity tests, phase profile on the bench workload, contract bench (no extras).
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slicing.py tests/test_gpu_edge.py -x -q > $out/${tag}_tests.txt 2>&1; tail -3 $out/${tag}_tests.txt
python scripts/quick_bench.py 4096 benchphases f32 > $out/${tag}_phases.txt 2>&1; head -16 $out/${tag}_phases.txt
python bench.py --no-cpu-baseline --no-extras > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print("env-steps/s %.0f  substeps/s %.3e  ms/step %.1f  e2e %.0f  smem frac %.3f" % (d["value"], d["substeps_per_s"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
print(d["launch_balance"])
PY
