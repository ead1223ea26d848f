#!/bin/bash
# Two-pair cluster GEMM (algo_pair = 3): parity against the single-CTA / pair kernels, then the layer A/B.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cta_pairs" > gpurun_out/t_quad.log 2>&1; rc=$?; echo "quad tests rc=$rc"
tail -n 8 gpurun_out/t_quad.log
if [ $rc -eq 0 ]; then
  timeout 200 python tools/gemm_ab.py 768 246784 --quad > gpurun_out/gemm_ab_quad.log 2>&1; echo "ab rc=$?"; cat gpurun_out/gemm_ab_quad.log
fi
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader
