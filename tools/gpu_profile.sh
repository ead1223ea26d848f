#!/bin/bash
# ncu evidence for one round: (1) per-launch durations of one bench step, (2) full captures of the top GEMM
# and attention kernels.  Numbers printed by bench.py under ncu are NOT bench values.
#   usage: tools/gpu_profile.sh TAG [what...]   what in {launches gemm attn}; default all
set -u
mkdir -p gpurun_out
TAG="${1:-r01}"; shift || true
WHAT="${*:-launches gemm attn}"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
for w in $WHAT; do
  case $w in
    launches)
      # 3 warm-up steps come first (277 launches each); skip far enough to land inside steady state
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv \
        --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_bench_${TAG}.log 2>&1
      echo "launch list rc=$?" ;;
    gemm)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 4 \
        -o gpurun_out/prof_gemm_${TAG} -f $BENCH > gpurun_out/ncu_gemm_${TAG}.log 2>&1
      echo "gemm capture rc=$?" ;;
    attn)
      # launch 3 = a vision layer (unmasked), launch 13.. = joint layers (masked) of the first warm-up step
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 \
        -o gpurun_out/prof_attn_vis_${TAG} -f $BENCH > gpurun_out/ncu_attn_${TAG}.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 14 -c 1 \
        -o gpurun_out/prof_attn_joint_${TAG} -f $BENCH >> gpurun_out/ncu_attn_${TAG}.log 2>&1
      echo "attention capture rc=$?" ;;
  esac
done
ls -la gpurun_out/
