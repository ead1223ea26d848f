#!/bin/bash
# compute-sanitizer over the tests of the kernels added in r01q (short-sequence attention forward / backward, D prep,
# rotary_bwd, fused GELU backward epilogue, forced-tile wgrad): memcheck, then racecheck on the shared-memory kernels.
set -u
mkdir -p gpurun_out
K='short_sequences or fused_gelu or forced_tile or (attention_bwd and not many_blocks) or rotary'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -m gpu -q -x -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
K2='short_sequences or (attention_bwd and not many_blocks)'
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -m gpu -q -x -k "$K2" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
