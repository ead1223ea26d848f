#!/bin/bash
# After the masked-attention change: the whole GPU suite and the cfg 5 (Large multiple-choice scoring) line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_final2.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_final2.log
timeout 300 python bench.py --config large --grid 18 32 --workload mc_scoring --steps 5 --no-cpu-baseline > gpurun_out/bench_r01q2_cfg5_mc_scoring.json 2> gpurun_out/bench_cfg5b.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01q2_cfg5_mc_scoring.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d["unit"], "ms", round(d["ms_per_step"], 2), "whole", round(d["whole_step_frac_of_sustained_peak"], 3))
for k, v in list(d["kernels"].items())[:5]:
    print("   ", k, round(v["ms_per_step"], 2), v["launches_per_step"], v.get("tflops") and round(v["tflops"]))
PY
