"""Times the individual kernels at the BASELINE config-2 shapes (Base, 64 clips) with CUDA events.
Development probe, not the benchmark: prints achieved TFLOP/s / GB/s per kernel and cuBLAS beside it."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from merlot_reserve_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def main():
    L.require_b200(0)
    dev = torch.device("cuda:0")
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    res = []
    shapes = [("vision", 246784), ("audio", 95232), ("joint", 71680), ("1clip-vision", 3856)]
    if "--attn-only" in sys.argv:
        shapes = []
    for name, M in shapes:
        a = torch.randn(M, H, device=dev).to(torch.bfloat16)
        a4 = torch.randn(M, 4 * H, device=dev).to(torch.bfloat16)
        x = torch.randn(M, H, device=dev).to(torch.bfloat16)
        rot = torch.randn(241, 32, device=dev)
        for tag, (N, K) in {"qkv": (3 * H, H), "out": (H, H), "up": (4 * H, H), "down": (H, 4 * H)}.items():
            w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
            b = torch.randn(N, device=dev)
            src = a4 if K == 4 * H else a
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            if tag == "qkv":
                fn = lambda: L.gemm(src, w, bias=b, rot=rot, hidden=H, q_scale=0.125, out=out)
            elif tag == "up":
                fn = lambda: L.gemm(src, w, bias=b, gelu=True, out=out)
            elif tag in ("out", "down"):
                fn = lambda: L.gemm(src, w, residual=x, out=x)
            ms = timeit(fn)
            ms_plain = timeit(lambda: L.gemm(src, w, out=out))
            ms_cublas = timeit(lambda: torch.matmul(src, w.t()))
            fl = 2.0 * M * N * K
            res.append(dict(kernel=f"gemm_{tag}", rows=name, M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9,
                            plain_tflops=fl / ms_plain / 1e9, cublas_tflops=fl / ms_cublas / 1e9))
            print(json.dumps(res[-1]), flush=True)
        del a4
        ln_g, ln_b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
        ms = timeit(lambda: L.layernorm(a, ln_g, ln_b, out=x))
        res.append(dict(kernel="layernorm", rows=name, M=M, ms=ms, gbs=4.0 * M * H / ms / 1e6))
        print(json.dumps(res[-1]), flush=True)
    for name, (B, T) in {"vision": (1024, 241), "audio": (3072, 31), "joint": (64, 1120), "large-vision": (256, 577),
                         "large-joint16": (16, 2464)}.items():
        qkv = torch.randn(B * T, 3 * H, device=dev).to(torch.bfloat16)
        out = torch.empty(B * T, H, device=dev, dtype=torch.bfloat16)
        fl = 4.0 * B * T * T * H
        for algo in (1, 2):
            ms = timeit(lambda: L.attention(qkv, B, T, H, out=out, algo=algo))
            res.append(dict(kernel="attention", algo=algo, rows=name, B=B, T=T, ms=ms, tflops=fl / ms / 1e9))
            print(json.dumps(res[-1]), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/probe_H{H}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
