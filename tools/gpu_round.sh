#!/bin/bash
# One GPU-box visit: kernel parity, model parity, bench.  Each stage under its own timeout so that a hang
# in one kernel cannot eat the box.  Everything worth reading lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
STAGES="${STAGES:-kernels model bench}"
for s in $STAGES; do
  case $s in
    gemm)    timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" ;;
    kernels) timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" ;;
    model)   timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model rc=$?" ;;
    all)     timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" ;;
    bench)   timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" ;;
    probe)   timeout 600 python tools/gemm_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" ;;
  esac
done
tail -n 30 gpurun_out/t_*.log gpurun_out/bench.log gpurun_out/probe.log 2>/dev/null; true
