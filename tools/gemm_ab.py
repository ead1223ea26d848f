"""A/B probe (GPU box): one transformer layer's four GEMMs (LN-folded qkv, out+residual+stats, LN-folded
up+GELU, down+residual+stats) at the vision shape, CTA pairs vs single CTA, alternating blocks under sustained
load; cuBLAS on the same four shapes beside it.  Development probe, not the benchmark."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from merlot_reserve_b200 import _lib as L  # noqa: E402

L.require_b200(0)
dev = torch.device("cuda:0")
H = int(sys.argv[1]) if len(sys.argv) > 1 else 768
M = int(sys.argv[2]) if len(sys.argv) > 2 else 246784
x = torch.randn(M, H, device=dev).to(torch.bfloat16)
qkv = torch.empty(M, 3 * H, device=dev, dtype=torch.bfloat16)
ctx = torch.randn(M, H, device=dev).to(torch.bfloat16)
u = torch.empty(M, 4 * H, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(H // 64, M, 2, device=dev)
stats[0, :, 0], stats[0, :, 1] = x.float().sum(-1), (x.float() ** 2).sum(-1)
rot = torch.randn(241, 32, device=dev)
mk = lambda n, k: (torch.randn(n, k, device=dev) * 0.02).to(torch.bfloat16)
wqkv, wo, w1, w2 = mk(3 * H, H), mk(H, H), mk(4 * H, H), mk(H, 4 * H)
bq, b1 = torch.randn(3 * H, device=dev), torch.randn(4 * H, device=dev)
sq, s1 = wqkv.float().sum(1).contiguous(), w1.float().sum(1).contiguous()
flops = 2.0 * M * H * H * 12


def layer(pair):
    L.gemm(x, wqkv, bias=bq, rot=rot, hidden=H, q_scale=0.125, out=qkv, ln_stats=stats, ln_colsum=sq, pair=pair)
    L.gemm(ctx, wo, residual=x, out=x, stats_out=stats, pair=pair)
    L.gemm(x, w1, bias=b1, gelu=True, out=u, ln_stats=stats, ln_colsum=s1, pair=pair)
    L.gemm(u, w2, residual=x, out=x, stats_out=stats, pair=pair)


def cublas():
    torch.matmul(x, wqkv.t(), out=qkv)
    torch.matmul(ctx, wo.t())
    torch.matmul(x, w1.t(), out=u)
    torch.matmul(u, w2.t())


def block(fn, n=12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    return e0, e1, n


for _ in range(3):
    layer(1), layer(2), cublas()
res = {"pair": [], "single": [], "cublas": []}
blocks = []
for r in range(6):
    blocks.append(("pair", block(lambda: layer(1))))
    blocks.append(("single", block(lambda: layer(2))))
    blocks.append(("cublas", block(cublas)))
torch.cuda.synchronize()
for k, (e0, e1, n) in blocks:
    res[k].append(flops * n / e0.elapsed_time(e1) / 1e9)
for k, v in res.items():
    print(f"{k:7s} TFLOP/s per block: " + " ".join(f"{t:7.1f}" for t in v) + f"   median {sorted(v)[len(v) // 2]:.1f}")
# per-GEMM timings
def one(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1, _ = block(fn, n)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for pair in (1, 2):
    t = [one(lambda: L.gemm(x, wqkv, bias=bq, rot=rot, hidden=H, q_scale=0.125, out=qkv, ln_stats=stats, ln_colsum=sq, pair=pair)),
         one(lambda: L.gemm(ctx, wo, residual=x, out=x, stats_out=stats, pair=pair)),
         one(lambda: L.gemm(x, w1, bias=b1, gelu=True, out=u, ln_stats=stats, ln_colsum=s1, pair=pair)),
         one(lambda: L.gemm(u, w2, residual=x, out=x, stats_out=stats, pair=pair))]
    fl = [2.0 * M * 3 * H * H, 2.0 * M * H * H, 2.0 * M * 4 * H * H, 2.0 * M * 4 * H * H]
    print(f"pair={pair}: " + " ".join(f"{n} {ms:.3f} ms {f / ms / 1e9:.0f} TF |" for n, ms, f in zip(("qkv", "out", "up", "down"), t, fl)))
t = [one(lambda: torch.matmul(x, wqkv.t(), out=qkv)), one(lambda: torch.matmul(ctx, wo.t())),
     one(lambda: torch.matmul(x, w1.t(), out=u)), one(lambda: torch.matmul(u, w2.t()))]
print("cublas: " + " ".join(f"{n} {ms:.3f} ms {f / ms / 1e9:.0f} TF |" for n, ms, f in zip(("qkv", "out", "up", "down"), t, fl)))
