#!/bin/bash
# Round-1q GPU visit: parity of the new paths, kernel-level A/B, train-step A/B (same box, back to back).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_training_gpu.py -m gpu -q -x > gpurun_out/t_q1.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/t_q1.log
timeout 240 python tools/r01q_probe.py --reps 15 > gpurun_out/r01q_probe.jsonl 2> gpurun_out/r01q_probe.err; echo "probe rc=$?"
cat gpurun_out/r01q_probe.jsonl
timeout 300 python bench.py --workload pretrain_step --steps 5 --no-cpu-baseline > gpurun_out/bench_q_train_new.json 2> gpurun_out/bench_q_train_new.err; echo "train new rc=$?"
MR_WGRAD_PLAN=legacy MR_FUSED_DGELU=0 timeout 300 python bench.py --workload pretrain_step --steps 5 --no-cpu-baseline > gpurun_out/bench_q_train_old.json 2> gpurun_out/bench_q_train_old.err; echo "train old rc=$?"
for f in gpurun_out/bench_q_train_new.json gpurun_out/bench_q_train_old.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 2), d["unit"], "ms", round(d["ms_per_step"], 2), "launches", d["gpu_launches"])
for k, v in list(d["kernels"].items())[:9]:
    print("   ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
PY
done
