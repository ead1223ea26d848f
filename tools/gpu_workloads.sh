#!/bin/bash
# Secondary BASELINE.json configs (not the driver's default bench line): cfg 4 (Large 18x32 embed_video),
# cfg 5 (Large multiple-choice scoring), forward half of cfg 3 (Base pretraining forward + loss).
set -u
mkdir -p gpurun_out
TAG="${1:-r01}"
timeout 900 python bench.py --config large --grid 18 32 --clips 16 --steps 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg4_large_18x32.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"
timeout 900 python bench.py --config large --grid 18 32 --workload mc_scoring --steps 5 > gpurun_out/bench_${TAG}_cfg5_mc_scoring.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python bench.py --config base --workload pretrain_fwd --steps 5 > gpurun_out/bench_${TAG}_cfg3_pretrain_fwd.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"
tail -n 3 gpurun_out/bench_cfg*.err
for f in gpurun_out/bench_${TAG}_cfg*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["metric"], round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
      "whole-step frac", round(d["whole_step_frac_of_sustained_peak"], 3), "gemm frac", round(d["roofline"]["frac"], 3))
for k, v in list(d["kernels"].items())[:7]:
    print("   ", k, round(v["ms_per_step"], 2), v["launches_per_step"], v["tflops"] and round(v["tflops"]))
PY
done
