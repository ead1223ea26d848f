"""Times mr_attention_fwd alone on the attention shapes of the BASELINE configs (CUDA events, L2 flushed between
launches by cycling through input buffers larger than L2).  One JSON line per shape; run on the GPU box:
    python tools/attn_probe.py [--reps 20]
Environment knobs of the library (e.g. MR_ATTN_PINGPONG=0) apply, so A/B runs are two invocations."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merlot_reserve_b200 import _lib as L  # noqa: E402

SHAPES = [
    # name, B, T, H, masked
    ("base_vision_x64", 1024, 241, 768, False),
    ("base_joint_x64", 64, 1120, 768, True),
    ("base_audio_x64", 3072, 31, 768, False),
    ("large_vision_18x32_x16", 256, 577, 1024, False),
    ("large_joint_18x32_x16", 16, 2464, 1024, True),
    ("pretrain_joint_b8", 48, 640, 768, True),
    ("large_mc_joint_b128", 512, 288, 1024, True),     # cfg 5: 128 examples x 4 candidates, T = 144 + 144
    ("large_mc_vision_b128", 128, 577, 1024, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--groups", action="store_true", help="masked shapes also carry group ids (pretraining joint)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    L.require_b200(0)
    for name, B, T, H, masked in SHAPES:
        if args.only and args.only not in name:
            continue
        nbuf = max(2, int(300e6 // (B * T * 3 * H * 2)) + 1)
        qkvs = [torch.randn(B * T, 3 * H, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        for t in qkvs:
            t[:, :H] *= 0.125                      # the QKV GEMM epilogue hands attention q / sqrt(64)
        out = torch.empty(B * T, H, device="cuda", dtype=torch.bfloat16)
        valid = group = None
        if masked:
            valid = torch.ones(B, T, dtype=torch.uint8, device="cuda")
            valid[:, 13:160] = 0          # embed_video layout: ~12 text tokens, then padding up to L = 160
            valid = valid.reshape(-1)
            if args.groups:      # pretraining: video_src ids (here: two halves of the sequence)
                group = (torch.arange(T, device="cuda")[None] >= T // 2).to(torch.int32).expand(B, T).contiguous().view(-1)
        for i in range(3):
            L.attention(qkvs[i % nbuf], B, T, H, valid=valid, group=group, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for i, (a, b) in enumerate(ev):
            a.record()
            L.attention(qkvs[i % nbuf], B, T, H, valid=valid, group=group, out=out)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        med = ts[len(ts) // 2]
        flops = 4.0 * B * T * T * H
        print(json.dumps({"shape": name, "B": B, "T": T, "H": H, "masked": masked, "us_median": round(med * 1e3, 1),
                          "us_min": round(ts[0] * 1e3, 1), "tflops": round(flops / med / 1e9, 1),
                          "stagger": os.environ.get("MR_ATTN_STAGGER", "default")}), flush=True)
        del qkvs


if __name__ == "__main__":
    main()
