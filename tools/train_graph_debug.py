"""Eager vs CUDA-graph train_step, state after every step (debug aid for capture_train_step)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from merlot_reserve_b200.synthetic import synthetic_params, synthetic_pretrain_batch, tiny_config
from merlot_reserve_b200.training import PretrainStep

cfg = tiny_config()
opt = {"learning_rate": 2e-4, "num_train_steps": 100, "num_warmup_steps": 0, "weight_decay_rate": 0.1, "beta_2": 0.98,
       "eps": 1e-6, "final_lr_scale": 1.0}
batches = [synthetic_pretrain_batch(cfg, 2, seed=40 + i) for i in range(3)]
eager = PretrainStep(cfg, synthetic_params(cfg, seed=0, weight_scale=4.0))
graphed = PretrainStep(cfg, synthetic_params(cfg, seed=0, weight_scale=4.0))
run = graphed.capture_train_step(batches[0], opt)
for i, bt in enumerate(batches):
    le = float(eager.train_step(bt, opt)[0])
    lg = float(run(bt)[0])
    torch.cuda.synchronize()
    d = {k: float((eager._state[k].float() - graphed._state[k].float()).abs().max()) for k in ("p", "m", "v", "g16")}
    print(i, "loss", le, lg, "state diff", d, "scales", eager.model._p.contrastive_scales, graphed.model._p.contrastive_scales)
    n = eager._state["n"]
    print("   g tail", eager._state["g16"][n:n + 3].float().cpu().numpy(), graphed._state["g16"][n:n + 3].float().cpu().numpy())
