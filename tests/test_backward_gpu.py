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
    lse, _ = lib.contrastive_lse(x.to(cuda), y.to(cuda), pos_off)
    dx, dy = lib.contrastive_lse_bwd(x.to(cuda), y.to(cuda), lse, pos_off, 1.0 / (2 * n))
    _close(dx, xr.grad, 2e-3, 2e-2, "dx")
    _close(dy, yr.grad, 2e-3, 2e-2, "dy_all")
    # rows written rank-major into a packed [world, S, H] buffer (the layout of the step's single reduce-scatter)
    if n_all % n == 0:
        world, S, off = n_all // n, n + 16, 8
        packed = torch.full((world, S, H), 7.0, device=cuda)
        lib.contrastive_lse_bwd(x.to(cuda), y.to(cuda), lse, pos_off, 1.0 / (2 * n), dy_out=packed, dy_seq_in=n,
                                dy_seq_out=S, dy_row_off=off)
        assert torch.equal(packed[:, off:off + n].reshape(n_all, H), dy)
        assert (packed[:, :off] == 7.0).all() and (packed[:, off + n:] == 7.0).all()


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


def _attn_ref_autograd(qkv, dctx, B, T, H, valid=None, group=None):
    """fp32 torch restatement of apply_attention (modeling.py:188-200) with the reference's 0 / -1e10 bias, and its
    autograd gradient w.r.t. qkv (q is already scaled by 1/8 and rotary-scaled in qkv, as the kernels get it)."""
    nh = H // 64
    x = qkv.float().cpu().clone().requires_grad_(True)
    q, k, v = x.view(B, T, 3, nh, 64).unbind(2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k)
    if valid is not None or group is not None:
        m = torch.ones(B, T, T, dtype=torch.bool)
        if valid is not None:
            vv = valid.cpu().bool().view(B, T)
            m &= vv[:, :, None] & vv[:, None, :]
        if group is not None:
            gg = group.cpu().view(B, T)
            m &= gg[:, :, None] == gg[:, None, :]
        s = s + torch.where(m, 0.0, -1e10)[:, None]
    w = torch.softmax(s, -1)
    ctx = torch.einsum("bhqk,bkhd->bqhd", w, v).reshape(B * T, H)
    ctx.backward(dctx.float().cpu())
    return ctx.detach(), x.grad, torch.logsumexp(s, -1).detach()


@pytest.mark.parametrize("B,T,H,masked", [(3, 241, 128, False), (5, 31, 128, False), (4, 16, 128, True), (2, 640, 128, True),
                                          (2, 200, 256, True), (1, 64, 64, False), (6, 31, 768, True), (3, 32, 256, True),
                                          (400, 31, 768, False), (2, 17, 64, False)])
def test_attention_bwd(cuda, lib, B, T, H, masked):
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g) * 0.8)
    qkv[:, :H] *= 0.35                                   # q carries the 1/sqrt(d) scale already
    dctx = _bf(torch.randn(B * T, H, generator=g))
    valid = group = None
    if masked:
        valid = torch.ones(B, T, dtype=torch.uint8)
        for b in range(B):
            valid[b, T - 3 - 7 * b:] = 0                 # right padding, different per sequence
        group = (torch.arange(T)[None] >= (T // 2 + torch.arange(B)[:, None])).to(torch.int32) + 1
        group[valid == 0] = -1
        dctx.view(B, T, H)[valid == 0] = 0               # outputs of padded rows are never used downstream
    ref_ctx, ref_dqkv, ref_lse = _attn_ref_autograd(qkv, dctx, B, T, H, valid, group)
    dev = lambda t: None if t is None else t.to(cuda).contiguous().view(-1)
    # algo 0 (AUTO) at T <= 32 = the one-block two-warp kernels (forward and backward); 1 = the 64 x 64 mma.sync kernels
    for algo in ([1, 2] if T > 64 else ([0, 1] if T <= 32 else [1])):
        lse = torch.empty(B, H // 64, T, device=cuda)
        ctx = lib.attention(qkv.to(cuda), B, T, H, valid=dev(valid), group=dev(group), algo=algo, lse=lse)
        _close(ctx, ref_ctx, 2e-2, 2e-2, f"ctx algo {algo}")
        ok = torch.ones(B, T, dtype=torch.bool) if valid is None else valid.bool()
        _close(lse.cpu().permute(0, 2, 1)[ok], ref_lse.permute(0, 2, 1)[ok], 2e-3, 1e-3, f"lse algo {algo}")
        scale = ref_dqkv.abs().max().item()
        for balgo in ([1, 2] if T > 64 else ([0, 1] if T <= 32 else [1])):   # mma.sync kernels / tcgen05 + TMEM kernel
            dqkv = lib.attention_bwd(qkv.to(cuda), dctx.to(cuda), ctx, lse, B, T, H, valid=dev(valid), group=dev(group),
                                     algo=balgo)
            _close(dqkv, ref_dqkv, 2e-2 * scale + 1e-3, 3e-2, f"dqkv fwd algo {algo} bwd algo {balgo}")


def test_attention_bwd_tc_many_blocks_and_uniform_rows(cuda, lib):
    """The tcgen05 backward on the training shapes' structure: several 128-key CTAs per sequence accumulating dQ
    through global reductions, 12 heads, a padded tail, and invalid QUERY rows that carry a non-zero dctx (their
    uniform attention still sends gradient to v, and dS = P (dP - D) to q and k, as in the mma.sync kernel)."""
    B, T, H = 3, 300, 768
    g = torch.Generator().manual_seed(17)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g) * 0.8)
    qkv[:, :H] *= 0.35
    dctx = _bf(torch.randn(B * T, H, generator=g))
    valid = torch.ones(B, T, dtype=torch.uint8)
    valid[0, 40:90] = 0
    valid[1, 250:] = 0
    dev = lambda t: t.to(cuda).contiguous().view(-1)
    lse = torch.empty(B, H // 64, T, device=cuda)
    ctx = lib.attention(qkv.to(cuda), B, T, H, valid=dev(valid), lse=lse)
    a = lib.attention_bwd(qkv.to(cuda), dctx.to(cuda), ctx, lse, B, T, H, valid=dev(valid), algo=1)
    b = lib.attention_bwd(qkv.to(cuda), dctx.to(cuda), ctx, lse, B, T, H, valid=dev(valid), algo=2)
    scale = a.float().abs().max().item()
    _close(b, a.float().cpu(), 2e-2 * scale + 1e-3, 3e-2, "tcgen05 vs mma.sync backward")
    _, ref_dqkv, _ = _attn_ref_autograd(qkv, dctx, B, T, H, valid, None)
    _close(b, ref_dqkv, 2e-2 * ref_dqkv.abs().max().item() + 1e-3, 3e-2, "tcgen05 backward vs autograd")


@pytest.mark.parametrize("B,T,H", [(7, 300, 768), (40, 200, 256), (33, 130, 128), (2, 1100, 768)])
def test_attention_bwd_tc_persistent_walk(cuda, lib, B, T, H):
    """The tcgen05 backward is persistent: one CTA per SM walks several (sequence, head, key block) items, with the
    K / V tiles, every barrier phase and the dK / dV write-out carried across item boundaries.  Shapes with 1.7, 2.2,
    0.9 and 1.5 items per SM (odd remainders, one-block and nine-block items), key validity + group ids + invalid
    query rows, against the mma.sync kernel on the same inputs."""
    g = torch.Generator().manual_seed(B * 7 + T)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g) * 0.8)
    qkv[:, :H] *= 0.35
    dctx = _bf(torch.randn(B * T, H, generator=g))
    valid = (torch.rand(B, T, generator=g) > 0.2).to(torch.uint8)
    valid[:, 0] = 1
    valid[B // 2, T // 3:] = 0
    group = (torch.arange(T)[None] >= torch.randint(1, T, (B, 1), generator=g)).to(torch.int32)
    dev = lambda t: t.to(cuda).contiguous().view(-1)
    lse = torch.empty(B, H // 64, T, device=cuda)
    ctx = lib.attention(qkv.to(cuda), B, T, H, valid=dev(valid), group=dev(group), lse=lse)
    a = lib.attention_bwd(qkv.to(cuda), dctx.to(cuda), ctx, lse, B, T, H, valid=dev(valid), group=dev(group), algo=1)
    for rep in range(2):                     # twice: nothing may be left behind in the workspace / barriers
        b = lib.attention_bwd(qkv.to(cuda), dctx.to(cuda), ctx, lse, B, T, H, valid=dev(valid), group=dev(group), algo=2)
        scale = a.float().abs().max().item()
        _close(b, a.float().cpu(), 2e-2 * scale + 1e-3, 3e-2, f"persistent tcgen05 vs mma.sync backward (run {rep})")


@pytest.mark.parametrize("B,T,H,per_row_rot", [(3, 300, 768, False), (5, 31, 128, False), (4, 50, 128, True), (2, 130, 256, True)])
def test_attention_bwd_finish_matches_the_three_pass_tail(cuda, lib, B, T, H, per_row_rot):
    """MrAttentionDesc.finish: the backward of the qkv epilogue (rotary scale, 1/sqrt(d)) and the qkv bias gradient
    folded into the dQ conversion pass, against mr_attention_bwd -> mr_rotary_bwd -> mr_colsum on the same inputs.
    Covers the tcgen05 path (fp32 dQ accumulator), the mma.sync 64-tile path and the one-block path (dQ in place),
    a per-position table (vision / audio: row r uses r % T) and a per-row table (joint)."""
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g) * 0.8).to(cuda)
    qkv[:, :H] *= 0.35
    dctx = _bf(torch.randn(B * T, H, generator=g)).to(cuda)
    valid = torch.ones(B, T, dtype=torch.uint8)
    valid[0, T // 2:] = 0
    valid = valid.to(cuda).view(-1)
    rot = (torch.randn(B * T if per_row_rot else T, 32, generator=g)).to(cuda)
    lse = torch.empty(B, H // 64, T, device=cuda)
    ctx = lib.attention(qkv, B, T, H, valid=valid, lse=lse)
    ref = lib.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid)
    plain = ref.clone()
    lib.rotary_bwd(ref, rot, H, q_scale=0.125)
    ref_bias = lib.colsum(ref)
    got, got_bias = lib.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid, rot=rot, q_scale=0.125, bias_grad=True)
    scale = ref.float().abs().max().item()
    # dk / dv go through identical arithmetic; dq is rounded once (fp32 accumulator x scale) instead of twice
    assert torch.equal(got[:, 2 * H:], ref[:, 2 * H:]) and torch.equal(got[:, 2 * H:], plain[:, 2 * H:])
    _close(got[:, H:2 * H], ref[:, H:2 * H].float().cpu(), 1e-6, 2.0 ** -7, "dk")
    _close(got[:, :H], ref[:, :H].float().cpu(), 1e-3 * scale, 2.0 ** -6, "dq")
    _close(got_bias, ref_bias.cpu(), 2e-3 * scale * (B * T) ** 0.5, 2e-2, "qkv bias gradient")
    # each flag alone
    only_scale = lib.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid, q_scale=0.125)
    _close(only_scale[:, :H], plain[:, :H].float().cpu() * 0.125, 1e-3 * scale, 2.0 ** -6, "q_scale only")
    assert torch.equal(only_scale[:, H:], plain[:, H:])
    same, bias_plain = lib.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=valid, bias_grad=True)
    assert torch.equal(same, plain)
    _close(bias_plain, lib.colsum(plain).cpu(), 2e-3 * plain.float().abs().max().item() * (B * T) ** 0.5, 2e-2, "bias only")


def test_transformer_layer_backward_composition(cuda, lib):
    """One pre-LN transformer layer (modeling.py:259-280) forward + backward composed from the C-ABI kernels
    (unfused LayerNorm path), against torch autograd of the fp32 oracle layer: dx and every parameter gradient."""
    from merlot_reserve_b200.synthetic import _encoder_params
    from oracle import reserve_oracle as O
    import numpy as np
    H, B, T = 128, 6, 50
    rng = np.random.default_rng(0)
    lp = _encoder_params(rng, H, 1, False, 4.0)["layer_00"]
    g = torch.Generator().manual_seed(0)
    x0 = _bf(torch.randn(B, T, H, generator=g))
    dout = _bf(torch.randn(B, T, H, generator=g) * 0.1)
    coords = torch.rand(T, 1, generator=g) * 2 - 1
    valid = torch.ones(B, T, dtype=torch.bool)
    valid[:, T - 4:] = False
    dout[~valid] = 0

    # ---- oracle + autograd ----
    def leaf(a):
        return O._t(a).clone().requires_grad_(True)
    P = {"pre_attn_ln": {k: leaf(v) for k, v in lp["pre_attn_ln"].items()},
         "pre_mlp_ln": {k: leaf(v) for k, v in lp["pre_mlp_ln"].items()},
         "attention_layer": {"qkv": {k: leaf(v) for k, v in lp["attention_layer"]["qkv"].items()},
                             "attn_proj": {"kernel": leaf(lp["attention_layer"]["attn_proj"]["kernel"])}},
         "mlp_layer": {"intermediate": {k: leaf(v) for k, v in lp["mlp_layer"]["intermediate"].items()},
                       "out": {"kernel": leaf(lp["mlp_layer"]["out"]["kernel"])}}}
    xr = x0.float().clone().requires_grad_(True)
    bias = torch.where(valid[:, None, :, None] & valid[:, None, None, :], 0.0, -1e10)
    y = O.transformer_layer(xr, P, O.rotary_sinusoids(coords), bias)
    y.backward(dout.float())

    # ---- kernels ----
    from merlot_reserve_b200.modeling import _w, _f
    M = B * T
    a, m = lp["attention_layer"], lp["mlp_layer"]
    wqkv, bqkv, wo = _w(a["qkv"]["kernel"], H), _f(a["qkv"]["bias"]), _w(a["attn_proj"]["kernel"], H)
    w1, b1, w2 = _w(m["intermediate"]["kernel"], H), _f(m["intermediate"]["bias"]), _w(m["out"]["kernel"], 4 * H)
    g1, be1 = _f(lp["pre_attn_ln"]["scale"]), _f(lp["pre_attn_ln"]["bias"])
    g2, be2 = _f(lp["pre_mlp_ln"]["scale"]), _f(lp["pre_mlp_ln"]["bias"])
    rot = lib.rotary_table(coords.to(cuda).contiguous())
    vld = valid.to(torch.uint8).to(cuda).view(-1)
    x = x0.to(cuda).view(M, H)
    # forward, keeping what the backward needs
    h1 = lib.layernorm(x, g1, be1)
    qkv = lib.gemm(h1, wqkv, bias=bqkv, rot=rot, hidden=H, q_scale=0.125)
    lse = torch.empty(B, H // 64, T, device=cuda)
    ctx = lib.attention(qkv, B, T, H, valid=vld, lse=lse)
    xm = lib.gemm(ctx, wo, residual=x)
    h2 = lib.layernorm(xm, g2, be2)
    up = lib.gemm(h2, w1, bias=b1)
    u = lib.gelu(up)
    yk = lib.gemm(u, w2, residual=xm)
    _close(yk.view(B, T, H)[valid], y.detach()[valid], 6e-2, 3e-2, "layer forward")
    # backward
    dy = dout.to(cuda).view(M, H)
    du = lib.gemm_bwd(dy, w2, trans_b=True, out_f32=False)                 # [M, 4H]
    dw2 = lib.gemm_bwd(dy, u, trans_a=True, trans_b=True, split_k=2)       # packed layout [H, 4H]
    dup = lib.gelu(up, du)
    dh2 = lib.gemm_bwd(dup, w1, trans_b=True, out_f32=False)
    dw1 = lib.gemm_bwd(dup, h2, trans_a=True, trans_b=True)
    db1 = lib.colsum(dup)
    dxm, dg2, dbe2 = lib.layernorm_bwd(dh2, xm, g2, dres=dy)
    dctx = lib.gemm_bwd(dxm, wo, trans_b=True, out_f32=False)
    dwo = lib.gemm_bwd(dxm, ctx, trans_a=True, trans_b=True)
    dqkv = lib.attention_bwd(qkv, dctx, ctx, lse, B, T, H, valid=vld)
    lib.rotary_bwd(dqkv, rot, H)
    dh1 = lib.gemm_bwd(dqkv, wqkv, trans_b=True, out_f32=False)
    dwqkv = lib.gemm_bwd(dqkv, h1, trans_a=True, trans_b=True)
    dbqkv = lib.colsum(dqkv)
    dx, dg1, dbe1 = lib.layernorm_bwd(dh1, x, g1, dres=dxm)

    def chk(got, ref, what, tol=4e-2):
        ref = ref.detach().float()
        s = ref.abs().max().item() + 1e-6
        err = (got.float().cpu() - ref).abs().max().item()
        assert err <= tol * s, f"{what}: max err {err:.4g} vs scale {s:.4g}"
    chk(dx.view(B, T, H)[valid], xr.grad[valid], "dx")
    # packed weights are [out, in]: compare with the flax [in, out] gradients transposed
    chk(dw2, P["mlp_layer"]["out"]["kernel"].grad.t(), "dW2")
    chk(dw1, P["mlp_layer"]["intermediate"]["kernel"].grad.t(), "dW1")
    chk(db1, P["mlp_layer"]["intermediate"]["bias"].grad, "db1")
    chk(dwo, P["attention_layer"]["attn_proj"]["kernel"].grad.reshape(H, H).t(), "dWo")
    chk(dwqkv, P["attention_layer"]["qkv"]["kernel"].grad.reshape(H, 3 * H).t(), "dWqkv")
    chk(dbqkv, P["attention_layer"]["qkv"]["bias"].grad.reshape(-1), "dbqkv")
    chk(dg2, P["pre_mlp_ln"]["scale"].grad, "dgamma2")
    chk(dbe2, P["pre_mlp_ln"]["bias"].grad, "dbeta2")
    chk(dg1, P["pre_attn_ln"]["scale"].grad, "dgamma1")
    chk(dbe1, P["pre_attn_ln"]["bias"].grad, "dbeta1")


def test_glue_backward_kernels(cuda, lib):
    """segment_pool / embed_assemble / attention-pool backward against torch autograd of the same gather-scatter ops."""
    g = torch.Generator().manual_seed(21)
    # one_hot_pool backward
    B, Lq, H, S = 3, 40, 64, 6
    dp = (torch.rand(B, Lq, generator=g) > 0.5)
    ix = torch.randint(-1, S + 1, (B, Lq), generator=g).to(torch.int32)
    dpooled = torch.randn(B, S, H, generator=g)
    v = torch.randn(B, Lq, H, generator=g).requires_grad_(True)
    oh = ((ix[..., None] == torch.arange(S)) & dp[..., None]).float()
    torch.einsum("bls,blh->bsh", oh, v).backward(dpooled)
    got = lib.segment_pool_bwd(dp.to(torch.uint8).to(cuda), ix.to(cuda), dpooled.to(cuda), Lq)
    _close(got, v.grad, 1e-2, 1e-2, "segment_pool bwd")

    # attention pool backward (2x2 windows, mode 1) vs autograd of the flax-MHA restatement
    Bp, gh, gw, Hp = 2, 4, 6, 128
    T, nw, nh = gh * gw + 1, (gh // 2) * (gw // 2), Hp // 64
    q = _bf(torch.randn(Bp * nw, Hp, generator=g))
    kv = _bf(torch.randn(Bp * T, 2 * Hp, generator=g))
    dctx = _bf(torch.randn(Bp * nw, Hp, generator=g))
    qr, kvr = q.float().requires_grad_(True), kv.float().requires_grad_(True)
    idx = torch.tensor([[1 + (2 * (w // (gw // 2)) + (j >> 1)) * gw + 2 * (w % (gw // 2)) + (j & 1) for j in range(4)]
                        for w in range(nw)])
    kvw = kvr.view(Bp, T, 2, nh, 64)[:, idx]                               # [B, nw, 4, 2, nh, 64]
    sc = torch.einsum("bwhd,bwjhd->bwhj", qr.view(Bp, nw, nh, 64) / 8.0, kvw[:, :, :, 0])
    ctx = torch.einsum("bwhj,bwjhd->bwhd", torch.softmax(sc, -1), kvw[:, :, :, 1]).reshape(Bp * nw, Hp)
    ctx.backward(dctx.float())
    dq, dkv = lib.window_attention_bwd(q.to(cuda), kv.to(cuda), dctx.to(cuda), Bp, T, Hp, 1, 4, nw, gh, gw)
    _close(dq, qr.grad, 2e-2, 2e-2, "pool dq")
    _close(dkv, kvr.grad, 2e-2, 2e-2, "pool dkv")
    # window mean backward
    dq_in = _bf(torch.randn(Bp * nw, Hp, generator=g))
    dx = _bf(torch.randn(Bp * T, Hp, generator=g))
    ref = dx.float().view(Bp, T, Hp).clone()
    ref[:, idx.reshape(-1)] += (dq_in.float().view(Bp, nw, 1, Hp) / 4).expand(-1, -1, 4, -1).reshape(Bp, nw * 4, Hp)
    got = dx.to(cuda).clone()
    lib.window_mean_bwd(dq_in.to(cuda), got, Bp, T, Hp, 1, 4, nw, gh, gw)
    _close(got, ref.view(-1, Hp), 2e-2, 1e-2, "window mean bwd")

    # joint assembly backward: rows scatter to the embedding table / audio spans / (shared) vision rows
    Hh, Ltok, n_vis, n_aud, vocab = 64, 8, 6, 4, 50
    Bv, G, tile = 2, 2, 3
    R = Bv * G * tile
    toks = torch.randint(6, vocab, (R, Ltok), generator=g).to(torch.int32)
    toks[:, 2:5] = 5
    ptr = torch.randint(0, n_aud, (R, Ltok), generator=g).to(torch.int32)
    dxj = _bf(torch.randn(R, Ltok + n_vis, Hh, generator=g))
    d_emb = torch.zeros(vocab, Hh, device=cuda)
    d_aud = torch.zeros(Bv, n_aud, 6, Hh, device=cuda)
    d_vis = torch.zeros(Bv * G, n_vis, Hh, device=cuda)
    lib.embed_assemble_bwd(toks.to(cuda), dxj.to(cuda).view(-1, Hh), H=Hh, vocab=vocab, d_emb=d_emb, aud_ptr=ptr.to(cuda),
                           d_audio=d_aud, d_vision=d_vis, n_aud=n_aud, n_vis=n_vis, T=Ltok + n_vis, aud_div=G * tile,
                           vis_group=G, vis_tile=tile)
    r_emb, r_aud, r_vis = torch.zeros(vocab, Hh), torch.zeros(Bv, n_aud, 6, Hh), torch.zeros(Bv * G, n_vis, Hh)
    for r in range(R):
        for t in range(Ltok):
            if toks[r, t] == 5:
                r_aud[r // (G * tile), ptr[r, t], t - 2] += dxj[r, t].float()
            else:
                r_emb[toks[r, t]] += dxj[r, t].float()
        r_vis[(r // (G * tile)) * G + r % G] += dxj[r, Ltok:].float()
    _close(d_emb, r_emb, 1e-5, 1e-5, "d_emb")
    _close(d_aud, r_aud, 1e-5, 1e-5, "d_audio")
    _close(d_vis, r_vis, 1e-5, 1e-5, "d_vision")
