#!/bin/bash
# 2-GPU visit: NCCL tests (packed all-gather loss, overlapped gradient all-reduce), train-step bench at N = 2 with the
# bucketed all-reduce under the backward pass and with the single blocking all-reduce.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q > gpurun_out/t_dist_n2.log 2>&1; echo "dist tests rc=$?"; tail -n 6 gpurun_out/t_dist_n2.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29511 --workload pretrain_step --steps 8 --no-cpu-baseline > gpurun_out/bench_r01q_cfg3_pretrain_step_n2.json 2> gpurun_out/bench_n2_a.err; echo "train n2 overlap rc=$?"
MR_OVERLAP_ALLREDUCE=0 run 29512 --workload pretrain_step --steps 8 --no-cpu-baseline > gpurun_out/bench_r01q_cfg3_pretrain_step_n2_blocking.json 2> gpurun_out/bench_n2_b.err; echo "train n2 blocking rc=$?"
tail -n 3 gpurun_out/bench_n2_*.err
for f in gpurun_out/bench_r01q_cfg3_pretrain_step_n2.json gpurun_out/bench_r01q_cfg3_pretrain_step_n2_blocking.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 2), d["unit"], "ms", round(d["ms_per_step"], 2), "n", d["n_gpus"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
