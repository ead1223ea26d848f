"""Same-box A/B of the Base train step (b = 8 per GPU, CUDA-graph replay) over module switches of
merlot_reserve_b200.training.  Alternates the variants (A B A B) so clock drift hits both; one JSON line per variant.
    python tools/train_ab.py fused_tail                                  # 1 GPU: _FUSED_QKV_TAIL on / off
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/train_ab.py native_collectives                              # library collectives vs torch.distributed"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from merlot_reserve_b200 import training  # noqa: E402
from merlot_reserve_b200.synthetic import load_config, synthetic_params, synthetic_pretrain_batch  # noqa: E402

SWITCH = {"fused_tail": "_FUSED_QKV_TAIL", "native_collectives": "_NATIVE_COLLECTIVES"}[sys.argv[1]]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
cfg = load_config("base")
params = synthetic_params(cfg, seed=0)
batch = {k: torch.from_numpy(v).cuda() for k, v in synthetic_pretrain_batch(cfg, 8, seed=1 + rank).items()}
runs = {}
for value in (True, False):
    setattr(training, SWITCH, value)
    step = training.PretrainStep(cfg, params)
    runs[value] = step.capture_train_step(batch)
    for _ in range(3):
        runs[value]()
times = {True: [], False: []}
for rep in range(2):
    for value in (True, False):
        run = runs[value]
        run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            run()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[value].append(round(float(t), 3))
if rank == 0:
    for value in (True, False):
        print(json.dumps({"switch": SWITCH, "value": value, "n_gpus": world, "ms_per_step": times[value]}), flush=True)
if world > 1:
    dist.destroy_process_group()
