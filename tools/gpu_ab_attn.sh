#!/bin/bash
# attention_tc A/B: in-tree build against gpurun_ab/lib_old.so (same box, alternating), after the parity tests.
set -u
mkdir -p gpurun_out
cp merlot_reserve_b200/lib/libmreserve_b200.so /tmp/main.so; cp gpurun_ab/lib_new.so merlot_reserve_b200/lib/libmreserve_b200.so; timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_backward_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x -k "attention or embed_video or multiple" > gpurun_out/t_ab_attn.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_ab_attn.log
LIB=merlot_reserve_b200/lib/libmreserve_b200.so
true
: > gpurun_out/ab_attn_probe.jsonl
for v in new old new old; do
  if [ $v = new ]; then cp gpurun_ab/lib_new.so $LIB; else cp /tmp/main.so $LIB; fi
  echo "== $v" | tee -a gpurun_out/ab_attn_probe.jsonl
  timeout 120 python tools/attn_probe.py --reps 15 2>&1 | grep -v audio | tee -a gpurun_out/ab_attn_probe.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['shape'], d['T'], d['us_median'], d['tflops'])
    except Exception: print(l.strip()[:100])
"
done
cp /tmp/main.so $LIB
