#!/bin/bash
# Round-1q GPU visit 3: the full-size parity / property tests, smoke(), secondary BASELINE configs with the r01q kernels.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q --durations=5 > gpurun_out/t_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -n 12 gpurun_out/t_fullsize.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
bash tools/gpu_workloads.sh r01q
