"""Backward kernels (include/mreserve_b200.h, "Backward entries") against torch autograd of the same op in fp32
on bf16-exact inputs.  The reference differentiates its forward with jax.value_and_grad
(pretrain/pretrain_model.py:306-321); the op-level gradients tested here are what that expands to."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {err.max().item():.4g}"


@pytest.mark.parametrize("rows,H", [(1000, 768), (333, 1024), (50, 128), (4099, 256)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_bwd(cuda, lib, rows, H, with_res):
    g = torch.Generator().manual_seed(rows + H)
    x = _bf(torch.randn(rows, H, generator=g) * 2 + torch.randn(rows, 1, generator=g))
    dy = _bf(torch.randn(rows, H, generator=g))
    gamma = 1 + 0.1 * torch.randn(H, generator=g)
    beta = 0.1 * torch.randn(H, generator=g)
    dres = _bf(torch.randn(rows, H, generator=g)) if with_res else None
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    mu = xr.mean(-1, keepdim=True)
    var = (xr * xr).mean(-1, keepdim=True) - mu * mu
    y = (xr - mu) * torch.rsqrt(var + 1e-5) * gr + br
    y.backward(dy.float())
    dx, dg, db = lib.layernorm_bwd(dy.to(cuda), x.to(cuda), gamma.to(cuda), dres=None if dres is None else dres.to(cuda))
    ref_dx = xr.grad + (dres.float() if with_res else 0)
    _close(dx, ref_dx, 3e-2, 2e-2, "dx")
    _close(dg, gr.grad, 2e-2 * math.sqrt(rows / 100), 1e-3, "dgamma")
    _close(db, br.grad, 2e-2 * math.sqrt(rows / 100), 1e-3, "dbeta")


def test_gelu_colsum_rotary_bwd(cuda, lib):
    g = torch.Generator().manual_seed(3)
    p = _bf(torch.randn(300, 512, generator=g) * 2)
    d = _bf(torch.randn(300, 512, generator=g))
    pr = p.float().requires_grad_(True)
    u = pr * torch.sigmoid(1.702 * pr)
    u.backward(d.float())
    _close(lib.gelu(p.to(cuda)), u.detach(), 2e-2, 1e-2, "gelu fwd")
    _close(lib.gelu(p.to(cuda), d.to(cuda)), pr.grad, 2e-2, 1e-2, "gelu bwd")
    x = _bf(torch.randn(2077, 776, generator=g))
    _close(lib.colsum(x.to(cuda)), x.float().sum(0), 5e-2, 1e-3, "colsum")
    H, M = 128, 77
    dq = _bf(torch.randn(M, 3 * H, generator=g))
    rot = torch.rand(7, 32, generator=g) * 2 - 1
    ref = dq.float().view(M, 3, H // 64, 64).clone()
    sc = rot[torch.arange(M) % 7]
    ref[:, :2, :, :32] *= sc[:, None, None, :]
    ref[:, 0] *= 0.125
    got = lib.rotary_bwd(dq.to(cuda).clone(), rot.to(cuda), H)
    _close(got, ref.view(M, 3 * H), 1e-2, 1e-2, "rotary bwd")


@pytest.mark.parametrize("n,n_all,pos_off", [(96, 384, 96), (24, 40, 8), (128, 128, 0)])
def test_contrastive_lse_bwd(cuda, lib, n, n_all, pos_off):
    H = 256
    g = torch.Generator().manual_seed(n)
    x = _bf(torch.nn.functional.normalize(torch.randn(n, H, generator=g), dim=-1) * 4.4)
    y = _bf(torch.nn.functional.normalize(torch.randn(n_all, H, generator=g), dim=-1) * 4.4)
    xr, yr = x.float().requires_grad_(True), y.float().requires_grad_(True)
    logits = xr @ yr.t()
    numer = (xr * yr[pos_off:pos_off + n]).sum(-1)
    loss = (torch.logsumexp(logits, -1) - numer).mean() / 2.0
    loss.backward()
    lse, _ = lib.contrastive_lse(x.to(cuda), y.to(cuda), y[pos_off:pos_off + n].contiguous().to(cuda))
    dx, dy = lib.contrastive_lse_bwd(x.to(cuda), y.to(cuda), lse, pos_off, 1.0 / (2 * n))
    _close(dx, xr.grad, 2e-3, 2e-2, "dx")
    _close(dy, yr.grad, 2e-3, 2e-2, "dy_all")


def test_unit_normalize_bwd(cuda, lib):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(40, 256, generator=g)
    idx = torch.tensor([5, 7, 39, 0, 17], dtype=torch.int32)
    dy = _bf(torch.randn(5, 256, generator=g))
    xr = x.clone().requires_grad_(True)
    sel = xr[idx.long()]
    y = 3.0 * sel / torch.sqrt((sel * sel).sum(-1, keepdim=True) + 1e-5)
    y.backward(dy.float())
    got = lib.unit_normalize_bwd(x.to(cuda), dy.to(cuda), scale=3.0, row_idx=idx.to(cuda))
    _close(got, xr.grad, 1e-4, 1e-3, "unit_normalize bwd (gathered)")
    xb = _bf(x)
    xr = xb.float().requires_grad_(True)
    y = xr / torch.sqrt((xr * xr).sum(-1, keepdim=True) + 1e-5)
    dyf = torch.randn(40, 256, generator=g)
    y.backward(dyf)
    _close(lib.unit_normalize_bwd(xb.to(cuda), dyf.to(cuda)), xr.grad, 1e-4, 1e-3, "unit_normalize bwd")
