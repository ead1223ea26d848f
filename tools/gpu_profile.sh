#!/bin/bash
# ncu evidence for one round: (1) per-launch durations of one bench step, (2) a full capture of the top GEMM
# and attention kernels.  Numbers printed by bench.py under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
TAG="${1:-r01}"
# 3 warm-up steps + timed steps come first; skip far enough to land inside steady state
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 4 \
  -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_${TAG}.log 2>&1
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 12 -c 2 \
  -o gpurun_out/prof_attn_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_${TAG}.log 2>&1
echo "attention capture rc=$?"
ls -la gpurun_out/
