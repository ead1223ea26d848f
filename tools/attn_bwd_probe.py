"""Times mr_attention_bwd (whole entry: memset + prep + main kernel + dQ convert) and its kernels on the training
shapes.  python tools/attn_bwd_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merlot_reserve_b200 import _lib as L  # noqa: E402

SHAPES = [("pretrain_vision_b8", 128, 241, 768, False), ("pretrain_joint_b8", 48, 640, 768, True),
          ("pretrain_audio_b8", 384, 31, 768, False), ("pretrain_span_b8", 384, 16, 768, True)]
if "--blocks-per-cta" in sys.argv:
    # equal total work (6 144 blocks of 128 x 128, ~30 848 rows), 2 / 4 / 8 / 16 query blocks per CTA: separates the fixed
    # cost of a CTA (launch, TMEM allocation, cold loads, dK / dV write-out) from the steady-state time per block
    SHAPES = [("blocks2", 128, 241, 768, False), ("blocks4", 32, 497, 768, False), ("blocks8", 8, 1009, 768, False),
              ("blocks16", 2, 2033, 768, False)]
FINISH = "--finish" in sys.argv      # the train step's call: qkv-epilogue backward + bias gradient in the tail pass
ALGOS = [2] if "--blocks-per-cta" in sys.argv or "--tc-only" in sys.argv else [1, 2]
torch.cuda.set_device(0)
L.require_b200(0)
for name, B, T, H, masked in SHAPES:
    qkv = torch.randn(B * T, 3 * H, device="cuda", dtype=torch.bfloat16)
    qkv[:, :H] *= 0.125
    dctx = torch.randn(B * T, H, device="cuda", dtype=torch.bfloat16)
    valid = group = None
    if masked:
        valid = torch.ones(B, T, dtype=torch.uint8, device="cuda")
        valid[:, T - T // 8:] = 0
        group = (torch.arange(T, device="cuda")[None] >= T // 2).to(torch.int32).expand(B, T).contiguous().view(-1)
        valid = valid.view(-1)
    lse = torch.empty(B, H // 64, T, device="cuda")
    extra = dict(rot=torch.randn(T, 32, device="cuda"), q_scale=0.125, bias_grad=True) if FINISH else {}
    ctx = L.attention(qkv, B, T, H, valid=valid, group=group, lse=lse)
    for algo in (ALGOS if T > 64 else [1]):
        for _ in range(3):
            L.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid, group=group, algo=algo, **extra)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            a.record()
            L.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid, group=group, algo=algo, **extra)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        med = ts[len(ts) // 2]
        print(json.dumps({"shape": name, "B": B, "T": T, "H": H, "masked": masked, "algo": {1: "mma_sync", 2: "tcgen05"}[algo],
                          "us_entry_median": round(med * 1e3, 1), "tflops_entry": round(10.0 * B * T * T * H / med / 1e9, 1)}),
              flush=True)
