#!/bin/bash
# Round-1q GPU visit 2: whole GPU suite, kernel A/B probe, headline bench, train-step bench, ncu launch lists,
# one full capture of the short-sequence attention kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all_q2.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/t_all_q2.log
timeout 300 python tools/r01q_probe.py --reps 15 > gpurun_out/r01q_probe2.jsonl 2> gpurun_out/r01q_probe2.err; echo "probe rc=$?"
grep -v '"wgrad"' gpurun_out/r01q_probe2.jsonl
timeout 400 python bench.py > gpurun_out/bench_r01q.json 2> gpurun_out/bench_r01q.err; echo "bench rc=$?"
timeout 300 python bench.py --workload pretrain_step --steps 5 --no-cpu-baseline > gpurun_out/bench_r01q_cfg3_pretrain_step.json 2> gpurun_out/bench_q2_train.err; echo "train rc=$?"
for f in gpurun_out/bench_r01q.json gpurun_out/bench_r01q_cfg3_pretrain_step.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 2), d["unit"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"],
      "gemm frac", round(d["roofline"]["frac"], 3), "whole", round(d["whole_step_frac_of_sustained_peak"], 3), d["clocks"])
for k, v in list(d["kernels"].items())[:12]:
    print("   ", k, round(v["ms_per_step"], 2), v["launches_per_step"], v.get("tflops") and round(v["tflops"]))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv \
  --log-file gpurun_out/launches_r01q.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r01q.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 3200 --csv \
  --log-file gpurun_out/launches_train_r01q.csv python bench.py --workload pretrain_step --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_r01q.log 2>&1; echo "train launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 2 -c 1 \
  -o gpurun_out/prof_attn_short_r01q -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_short_r01q.log 2>&1; echo "short attention capture rc=$?"
ls -la gpurun_out | tail -20
