#!/bin/bash
python tools/attn_probe.py --only base_ 2>&1 | grep -o '"shape": "[a-z_0-9]*"\|"us_median": [0-9.]*' | paste - - 
python tools/attn_probe.py --only large_mc_joint 2>&1 | grep -o '"shape": "[a-z_0-9]*"\|"us_median": [0-9.]*' | paste - -
python tools/attn_bwd_probe.py --tc-only 2>&1 | grep tcgen05 | grep -o '"shape": "[a-z_0-9]*"\|"us_entry_median": [0-9.]*' | paste - -
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), 'seg/s', round(d['ms_per_step'],2), 'ms', 'gemm TF', round(d['roofline']['achieved'],1), d['clocks'])"
