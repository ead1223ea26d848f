#!/bin/bash
# End-of-round check on one B200: the whole GPU suite, smoke(), the default bench line and the train-step line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_final.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --workload pretrain_step --steps 5 --no-cpu-baseline > gpurun_out/bench_final_train.json 2> gpurun_out/bench_final_train.err; echo "train rc=$?"
for f in gpurun_out/bench_final.json gpurun_out/bench_final_train.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 2), d["unit"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "gemm frac", round(d["roofline"]["frac"], 3),
      "whole", round(d["whole_step_frac_of_sustained_peak"], 3), d["clocks"], d.get("cpu_baseline"))
PY
done
