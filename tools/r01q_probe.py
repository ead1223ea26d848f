"""Same-box A/B of the round-1q changes, kernel level (CUDA events, inputs rotated through > L2 of buffers):
  * short-sequence attention: one-tile two-warp kernel (algo AUTO) against the 64 x 64 kernel (algo MMA_SYNC)
  * wgrad: `wgrad_plan` (256-wide tiles, few K parts) against the round-1 rule (128-wide tiles, rows // 2048 parts)
  * GELU backward fused into the dgrad epilogue against dgrad + mr_gelu
One JSON line per measurement:  python tools/r01q_probe.py [--reps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merlot_reserve_b200 import _lib as L  # noqa: E402

BF = torch.bfloat16


def timed(fn, reps, nbuf):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, b) in enumerate(ev):
        a.record()
        fn(i % nbuf)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return round(ts[len(ts) // 2] * 1e3, 1), round(ts[0] * 1e3, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    L.require_b200(0)
    dev = "cuda"

    for name, B, T, H in [("base_audio_x64", 3072, 31, 768), ("pretrain_audio_b8", 384, 31, 768), ("pretrain_span_b8", 384, 16, 768),
                          ("large_audio_x16", 768, 31, 1024)]:
        nbuf = max(2, int(300e6 // (B * T * 3 * H * 2)) + 1)
        qkvs = [torch.randn(B * T, 3 * H, device=dev, dtype=BF) for _ in range(nbuf)]
        out = torch.empty(B * T, H, device=dev, dtype=BF)
        res = {}
        for label, algo in (("one_tile", 0), ("tile64", 1), ("one_tile_again", 0)):
            res[label] = timed(lambda i: L.attention(qkvs[i], B, T, H, out=out, algo=algo), args.reps, nbuf)
        print(json.dumps({"probe": "attention_short", "shape": name, "B": B, "T": T, "H": H, "us_median_min": res,
                          "hbm_bytes": B * T * 4 * H * 2}), flush=True)
        del qkvs

    for name, o, i, rows in [("mlp_up", 3072, 768, 30848), ("mlp_down", 768, 3072, 30848), ("qkv", 2304, 768, 30848),
                             ("attn_out", 768, 768, 30848), ("audio_up", 3072, 768, 11904), ("audio_out", 768, 768, 11904),
                             ("large_up", 4096, 1024, 73856), ("large_out", 1024, 1024, 73856)]:
        nbuf = max(2, int(300e6 // (rows * (o + i) * 2)) + 1)
        dys = [torch.randn(rows, o, device=dev, dtype=BF) for _ in range(nbuf)]
        xs = [torch.randn(rows, i, device=dev, dtype=BF) for _ in range(nbuf)]
        legacy = int(max(1, min(16, rows // 2048)))
        bn, s = L.wgrad_plan(o, i, rows)
        res = {}
        variants = [("plan", bn, s), ("legacy", 0, legacy), ("plan_again", bn, s)]
        if i % 256 == 0:
            variants += [("t256_s%d" % k, 256, k) for k in (2, 3, 4, 6, 8) if (256, k) != (bn, s)]
        for label, t, k in variants:
            res[label] = timed(lambda j: L.gemm_bwd(dys[j], xs[j], trans_a=True, trans_b=True, split_k=k, tile_n=t), args.reps, nbuf)
        flops = 2.0 * o * i * rows
        print(json.dumps({"probe": "wgrad", "shape": name, "out": o, "in": i, "rows": rows, "plan": [bn, s], "legacy_split": legacy,
                          "us_median_min": res, "tflops_plan": round(flops / res["plan"][0] / 1e6, 1),
                          "tflops_legacy": round(flops / res["legacy"][0] / 1e6, 1)}), flush=True)
        del dys, xs

    # attention backward: short sequences (one-block kernel, algo AUTO, against the 64 x 64 kernel, algo MMA_SYNC) and the
    # tcgen05 shapes (whole entry: memset + D = rowsum(dO . O) prep + main kernel + dQ convert)
    for name, B, T, H, algos in [("pretrain_audio_b8", 384, 31, 768, (0, 1, 0)), ("pretrain_span_b8", 384, 16, 768, (0, 1, 0)),
                                 ("pretrain_vision_b8", 128, 241, 768, (0,)), ("pretrain_joint_b8", 48, 640, 768, (0,))]:
        nbuf = 3
        qkvs = [torch.randn(B * T, 3 * H, device=dev, dtype=BF) * 0.5 for _ in range(nbuf)]
        dctx = torch.randn(B * T, H, device=dev, dtype=BF)
        lse = torch.empty(B, H // 64, T, device=dev)
        ctxs = [L.attention(q, B, T, H, lse=lse) for q in qkvs]
        res = {}
        for n, algo in enumerate(algos):
            res[f"algo{algo}_{n}"] = timed(lambda i: L.attention_bwd(qkvs[i], dctx, ctxs[i], lse, B, T, H, algo=algo), args.reps, nbuf)
        print(json.dumps({"probe": "attention_bwd", "shape": name, "B": B, "T": T, "H": H, "us_median_min": res}), flush=True)

    for name, rows, H, rr in [("rotary_bwd_vision_b8", 30848, 768, 241), ("rotary_bwd_joint_b8", 30720, 768, 30720)]:
        dq = [torch.randn(rows, 3 * H, device=dev, dtype=BF) for _ in range(3)]
        rot = torch.randn(rr, 32, device=dev)
        t = timed(lambda i: L.rotary_bwd(dq[i], rot, H), args.reps, 3)
        print(json.dumps({"probe": "rotary_bwd", "shape": name, "rows": rows, "us_median_min": t,
                          "touched_bytes": rows * (H + H // 2) * 2 * 2}), flush=True)

    for name, rows, H in [("base_b8_vision", 30848, 768), ("base_b8_audio", 11904, 768)]:
        nbuf = 3
        dxs = [torch.randn(rows, H, device=dev, dtype=BF) for _ in range(nbuf)]
        ups = [torch.randn(rows, 4 * H, device=dev, dtype=BF) for _ in range(nbuf)]
        w2 = (torch.randn(H, 4 * H, device=dev) * 0.03).to(BF)
        res = {}
        for label in ("fused", "two_pass", "fused_again"):
            if label.startswith("fused"):
                fn = lambda j: L.gemm_bwd(dxs[j], w2, trans_b=True, out_f32=False, dgelu_pre=ups[j])
            else:
                fn = lambda j: L.gelu(ups[j], L.gemm_bwd(dxs[j], w2, trans_b=True, out_f32=False))
            res[label] = timed(fn, args.reps, nbuf)
        print(json.dumps({"probe": "dgelu", "shape": name, "rows": rows, "H": H, "us_median_min": res}), flush=True)


if __name__ == "__main__":
    main()
