"""Runs N eager Base train steps (b = 8) -- the command profiled for profiles/launches_train_*.csv:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python tools/train_steps.py 3
(cudaProfilerStart / Stop bracket the LAST step, so exactly one warm step is captured whatever its launch count)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from merlot_reserve_b200.synthetic import load_config, synthetic_params, synthetic_pretrain_batch  # noqa: E402
from merlot_reserve_b200.training import PretrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = load_config("base")
step = PretrainStep(cfg, synthetic_params(cfg, seed=0))
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic_pretrain_batch(cfg, 8, seed=1).items()}
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            torch.cuda.cudart().cudaProfilerStart()
        loss, _ = step.train_step(batch)
        torch.cuda.synchronize()
        if i == n - 1:
            torch.cuda.cudart().cudaProfilerStop()
        print("step", i, float(loss), flush=True)
