"""Per-entry-point parity of the C ABI (include/mreserve_b200.h) against plain torch fp32 restatements
of the same op, on seeded inputs.  Tolerances: inputs are exact bf16 values, accumulation is fp32, so the
only error allowed is the final bf16 rounding (rel 2^-8) plus accumulation-order noise."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err {err.max().item():.4g} "
                           f"at {np.unravel_index(int(err.argmax()), tuple(err.shape))}, ref there "
                           f"{ref.flatten()[int(err.argmax())].item():.4g}")


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 768), (256, 768, 768), (1000, 768, 768), (3856, 2304, 768), (241, 3072, 768),
    (300, 768, 3072), (77, 1024, 1024), (129, 1536, 192), (513, 4096, 1024), (64, 128, 128), (5, 8, 64),
    (200, 264, 72),
]


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(cuda, lib, M, N, K, algo):
    if algo == 1 and M * N * K > 3e9:
        pytest.skip("debug kernel too slow")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    ref = a.float() @ w.float().t()
    out = lib.gemm(a, w, algo=algo)
    _close(out, ref, 2e-2, 1e-2, f"gemm {M}x{N}x{K} algo{algo}")
    out32 = lib.gemm(a, w, out_f32=True, algo=algo)
    _close(out32, ref, 2e-3, 1e-3, f"gemm f32 {M}x{N}x{K} algo{algo}")


@pytest.mark.parametrize("M,N,K", [(1000, 3072, 768), (260, 4096, 1024), (130, 256, 64)])
def test_gemm_bias_gelu(cuda, lib, M, N, K):
    g = torch.Generator().manual_seed(1)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    pre = a.float() @ w.float().t() + b
    _close(lib.gemm(a, w, bias=b), pre, 2e-2, 1e-2, "bias")
    _close(lib.gemm(a, w, bias=b, gelu=True), pre * torch.sigmoid(1.702 * pre), 2e-2, 1e-2, "gelu")


@pytest.mark.parametrize("M,N,K,pair", [(1000, 3072, 768, 0), (40000, 3072, 768, 1), (260, 4096, 1024, 2), (130, 200, 72, 0)])
def test_gemm_bias_gelu_dual_output(cuda, lib, M, N, K, pair):
    """MR_EPI_GELU_DUAL: the activation AND the bf16 pre-activation leave the same epilogue (single-CTA, CTA-pair and
    128-wide tiles; ragged M / N / K): bit-identical to the two single-output GEMMs."""
    g = torch.Generator().manual_seed(M)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    pre = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=cuda)
    act = lib.gemm(a, w, bias=b, gelu=True, pre_out=pre, pair=pair)
    assert torch.equal(pre, lib.gemm(a, w, bias=b, pair=pair))
    assert torch.equal(act, lib.gemm(a, w, bias=b, gelu=True, pair=pair))
    ref = a.float() @ w.float().t() + b
    _close(pre, ref, 2e-2, 1e-2, "pre-activation")
    _close(act, ref * torch.sigmoid(1.702 * ref), 2e-2, 1e-2, "activation")


@pytest.mark.parametrize("M,N,K", [(1000, 768, 3072), (300, 1024, 1024), (128, 256, 64)])
def test_gemm_residual_inplace(cuda, lib, M, N, K):
    g = torch.Generator().manual_seed(2)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    x = _bf(torch.randn(M, N, generator=g)).to(cuda)
    ref = x.float() + a.float() @ w.float().t()
    out = lib.gemm(a, w, residual=x)
    _close(out, ref, 3e-2, 1e-2, "residual")
    x2 = x.clone()
    lib.gemm(a, w, residual=x2, out=x2)  # in place, as the encoder uses it
    assert torch.equal(x2, out)


@pytest.mark.parametrize("H,T,B", [(768, 241, 5), (1024, 31, 9), (128, 16, 7)])
def test_gemm_rotary_qkv(cuda, lib, H, T, B):
    g = torch.Generator().manual_seed(3)
    M = B * T
    a = _bf(torch.randn(M, H, generator=g)).to(cuda)
    w = _bf(torch.randn(3 * H, H, generator=g) * 0.05).to(cuda)
    b = torch.randn(3 * H, generator=g).to(cuda)
    rot = torch.randn(T, 32, generator=g).to(cuda)
    pre = (a.float() @ w.float().t() + b).view(B, T, 3, H // 64, 64).clone()
    pre[:, :, :2, :, :32] *= rot.view(1, T, 1, 1, 32)
    pre[:, :, 0] *= 0.125
    out = lib.gemm(a, w, bias=b, rot=rot, hidden=H, q_scale=0.125)
    _close(out, pre.view(M, 3 * H), 3e-2, 1e-2, "rotary qkv")
    # per-row table (joint encoder): rot_rows == M
    rot2 = torch.randn(M, 32, generator=g).to(cuda)
    pre2 = (a.float() @ w.float().t() + b).view(M, 3, H // 64, 64).clone()
    pre2[:, :2, :, :32] *= rot2.view(M, 1, 1, 32)
    pre2[:, 0] *= 0.125
    out2 = lib.gemm(a, w, bias=b, rot=rot2, hidden=H, q_scale=0.125)
    _close(out2, pre2.view(M, 3 * H), 3e-2, 1e-2, "rotary qkv per-row")


def test_gemm_row_map_and_strided_a(cuda, lib):
    g = torch.Generator().manual_seed(4)
    B, hw, K, N = 6, 24, 768, 256
    a = _bf(torch.randn(B * hw, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    out = torch.full((B * (hw + 1), N), 7.0, dtype=torch.bfloat16, device=cuda)
    lib.gemm(a, w, bias=b, out=out, seq_in=hw, seq_out=hw + 1, row_off=1)
    ref = (a.float() @ w.float().t() + b).view(B, hw, N)
    o = out.view(B, hw + 1, N)
    _close(o[:, 1:], ref, 2e-2, 1e-2, "row map")
    assert (o[:, 0] == 7.0).all(), "CLS rows must be untouched"
    # strided A: row 0 of every [T, K] block
    T = 5
    big = _bf(torch.randn(B * T, K, generator=g)).to(cuda)
    out2 = lib.gemm(big, w, bias=b, M=B, lda=T * K)
    _close(out2, big.view(B, T, K)[:, 0].float() @ w.float().t() + b, 2e-2, 1e-2, "strided A")


def test_gemm_large_throughput_shape(cuda, lib):
    """Many tiles per CTA (persistent loop, both TMEM accumulator stages, ring wrap-around)."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 148 * 128 * 3 + 77, 768, 768
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    out = lib.gemm(a, w)
    ref = (a.float() @ w.float().t())
    _close(out, ref, 2e-2, 1e-2, "large gemm")


def _attn_ref(qkv, B, T, H, valid=None, group=None):
    nh = H // 64
    q, k, v = qkv.float().view(B, T, 3, nh, 64).unbind(2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k)
    if valid is not None or group is not None:
        m = torch.ones(B, T, T, dtype=torch.bool, device=qkv.device)
        if valid is not None:
            vb = valid.view(B, T).bool()
            m &= vb[:, :, None] & vb[:, None, :]
        if group is not None:
            gg = group.view(B, T)
            m &= gg[:, :, None] == gg[:, None, :]
        s = s + torch.where(m, 0.0, -1e10)[:, None]
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(B * T, H)


@pytest.mark.parametrize("B,T,H", [(3, 241, 768), (5, 31, 768), (7, 16, 128), (2, 577, 1024), (1, 1120, 768),
                                   (2, 64, 128), (2, 65, 128), (1, 640, 256)])
@pytest.mark.parametrize("algo", [1, 2], ids=["mma_sync", "tcgen05"])
def test_attention_unmasked(cuda, lib, B, T, H, algo):
    g = torch.Generator().manual_seed(T)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35  # keep the softmax away from one-hot
    out = lib.attention(qkv, B, T, H, algo=algo)
    _close(out, _attn_ref(qkv, B, T, H), 2e-2, 2e-2, f"attention T={T}")


def test_attention_lazy_rescale(cuda, lib):
    """Scores that grow block after block force the tcgen05 kernel through its O-correction path."""
    B, T, H = 2, 640, 128
    g = torch.Generator().manual_seed(0)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    v3 = qkv.view(B, T, 3, H)
    v3[:, :, 0] = _bf(torch.ones(B, T, H) * 0.25).to(cuda)
    ramp = torch.linspace(-1.0, 1.0, T).view(1, T, 1).expand(B, T, H)       # later keys score much higher
    v3[:, :, 1] = _bf(ramp * 1.5 + 0.05 * torch.randn(B, T, H, generator=g)).to(cuda)
    ref = _attn_ref(qkv, B, T, H)
    for algo in (1, 2):
        _close(lib.attention(qkv, B, T, H, algo=algo), ref, 2e-2, 2e-2, f"lazy rescale algo {algo}")


@pytest.mark.parametrize("B,T,H", [(4, 16, 128), (3, 200, 128), (2, 640, 256)])
@pytest.mark.parametrize("algo", [1, 2], ids=["mma_sync", "tcgen05"])
def test_attention_masked(cuda, lib, B, T, H, algo):
    g = torch.Generator().manual_seed(T + 1)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    valid = (torch.rand(B, T, generator=g) > 0.3).to(torch.uint8)
    valid[:, 0] = 1
    valid[0, T // 2:] = 0  # a long padded tail, whole key blocks masked
    group = torch.randint(0, 3, (B, T), generator=g).to(torch.int32)
    valid, group = valid.to(cuda).contiguous(), group.to(cuda).contiguous()
    _close(lib.attention(qkv, B, T, H, valid=valid.view(-1), algo=algo), _attn_ref(qkv, B, T, H, valid), 2e-2, 2e-2, "valid")
    _close(lib.attention(qkv, B, T, H, valid=valid.view(-1), group=group.view(-1), algo=algo),
           _attn_ref(qkv, B, T, H, valid, group), 2e-2, 2e-2, "valid+group")


@pytest.mark.parametrize("B,T,H,L", [(2, 1120, 128, 160), (3, 400, 128, 130), (2, 288, 256, 144)])
def test_attention_joint_style_mask(cuda, lib, B, T, H, L):
    """Joint-encoder layout: padding only inside the first L (text) positions, every later (vision) key valid, so
    most key blocks take the mask-free fast path while invalid QUERY rows (uniform attention) cross them."""
    g = torch.Generator().manual_seed(T + L)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    valid = torch.ones(B, T, dtype=torch.uint8)
    valid[:, 40:L] = (torch.rand(B, L - 40, generator=g) > 0.5).to(torch.uint8)
    valid[1, 20:L] = 0
    valid = valid.to(cuda).contiguous()
    ref = _attn_ref(qkv, B, T, H, valid)
    for algo in (1, 2):
        _close(lib.attention(qkv, B, T, H, valid=valid.view(-1), algo=algo), ref, 2e-2, 2e-2, f"joint mask algo {algo}")


@pytest.mark.parametrize("B,T,H", [(40, 288, 256), (3, 96, 128), (5, 170, 128), (2, 192, 768), (130, 288, 1024)])
def test_attention_96_key_blocks(cuda, lib, B, T, H):
    """Sequences that 96-key blocks pad at least 10 % less than 128-key blocks (the multiple-choice joint length
    T = 144 + 144 = 288 = 3 x 96) run the kKV = 96 instantiation: unmasked, key validity (joint layout with invalid query
    rows), validity + group ids, more items than SMs, and the ring wrap-around."""
    g = torch.Generator().manual_seed(T + B)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    _close(lib.attention(qkv, B, T, H, algo=2), _attn_ref(qkv, B, T, H), 2e-2, 2e-2, f"96-key unmasked T={T}")
    valid = torch.ones(B, T, dtype=torch.uint8)
    valid[:, 13:T // 2] = (torch.rand(B, T // 2 - 13, generator=g) > 0.6).to(torch.uint8)
    valid[0, 5:T // 2] = 0
    group = (torch.arange(T)[None] >= (T * 2) // 3).to(torch.int32).expand(B, T).contiguous()
    valid, group = valid.to(cuda).contiguous(), group.to(cuda).contiguous()
    _close(lib.attention(qkv, B, T, H, valid=valid.view(-1), algo=2), _attn_ref(qkv, B, T, H, valid), 2e-2, 2e-2, "96-key valid")
    _close(lib.attention(qkv, B, T, H, valid=valid.view(-1), group=group.view(-1), algo=2),
           _attn_ref(qkv, B, T, H, valid, group), 2e-2, 2e-2, "96-key valid+group")
    lse = torch.empty(B, H // 64, T, device=cuda)
    lib.attention(qkv, B, T, H, algo=2, lse=lse)
    q, k, _ = qkv.float().view(B, T, 3, H // 64, 64).unbind(2)
    ref_lse = torch.logsumexp(torch.einsum("bqhd,bkhd->bhqk", q, k), -1)
    _close(lse, ref_lse, 2e-2, 2e-2, "96-key lse")


@pytest.mark.parametrize("B,T,H", [(3, 241, 128), (130, 288, 128), (2, 1120, 768), (40, 577, 256), (5, 65, 128), (1, 2464, 128)])
def test_attention_two_tile_and_three_stream_kernels_agree(cuda, lib, B, T, H):
    """The two tensor-memory forward kernels (two 128-query tiles per SM on 128- / 96-key blocks; three independent
    tile streams on 64-key blocks) forced on the same inputs, unmasked and with key validity (long masked tails,
    invalid query rows), output and row log-sum-exp, against the reference and against each other.  Shapes with
    0.05 .. 3.5 tiles per stream, 2 .. 39 key blocks per tile, a 65-key sequence (one key in the second block)."""
    g = torch.Generator().manual_seed(T * 3 + B)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    valid = (torch.rand(B, T, generator=g) > 0.3).to(torch.uint8)
    valid[:, 0] = 1
    valid[0, T // 2:] = 0
    valid = valid.to(cuda).contiguous()
    for v in (None, valid):
        ref = _attn_ref(qkv, B, T, H, v) if B * T * T * (H // 64) < 4e8 else None
        outs, lses = [], []
        for algo in (lib.ATTN_TCGEN05_2TILE, lib.ATTN_TCGEN05_3STREAM):
            lse = torch.empty(B, H // 64, T, device=cuda)
            outs.append(lib.attention(qkv, B, T, H, valid=None if v is None else v.view(-1), algo=algo, lse=lse))
            lses.append(lse)
            if ref is not None:
                _close(outs[-1], ref, 2e-2, 2e-2, f"algo {algo} vs reference")
        _close(outs[1], outs[0], 2e-2, 2e-2, "three-stream vs two-tile")
        _close(lses[1], lses[0], 2e-2, 2e-2, "lse three-stream vs two-tile")


def test_attention_many_items_persistent(cuda, lib):
    """More work items than SMs, odd and even numbers of query tiles: exercises the persistent loop, the Q double
    buffer, the K/V ring wrap-around and pairs whose second tile is absent."""
    g = torch.Generator().manual_seed(99)
    for B, T, H in [(40, 241, 768), (9, 300, 256), (3, 1120, 768)]:
        qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
        qkv.view(B * T, 3, H)[:, 0] *= 0.35
        _close(lib.attention(qkv, B, T, H, algo=2), _attn_ref(qkv, B, T, H), 2e-2, 2e-2, f"persistent T={T}")


@pytest.mark.parametrize("rows,H", [(1000, 768), (333, 1024), (50, 128)])
def test_layernorm(cuda, lib, rows, H):
    g = torch.Generator().manual_seed(6)
    x = _bf(torch.randn(rows, H, generator=g) * 3 + 0.5).to(cuda)
    gm = (1 + 0.1 * torch.randn(H, generator=g)).to(cuda)
    bt = (0.1 * torch.randn(H, generator=g)).to(cuda)
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = (xf * xf).mean(-1, keepdim=True) - mean * mean
    ref = (xf - mean) * (torch.rsqrt(var + 1e-5) * gm) + bt
    _close(lib.layernorm(x, gm, bt), ref, 2e-2, 1e-2, "layernorm")
    # row map: first 3 rows of every block of 10
    nb = rows // 10
    sub = lib.layernorm(x, gm, bt, rows=nb * 3, seq_in=10, seq_out=3)
    _close(sub, ref[: nb * 10].view(nb, 10, H)[:, :3].reshape(nb * 3, H), 2e-2, 1e-2, "layernorm rowmap")


def test_cast_pad_fill_cls(cuda, lib):
    g = torch.Generator().manual_seed(7)
    x = torch.rand(90, 130, generator=g).to(cuda)
    y = lib.cast_pad(x, 90, 130, 192)
    assert torch.equal(y[:, :130], x.to(torch.bfloat16)) and (y[:, 130:] == 0).all()
    x2 = torch.rand(77, 768, generator=g).to(cuda)
    assert torch.equal(lib.cast_pad(x2, 77, 768, 768), x2.to(torch.bfloat16))
    buf = torch.zeros(4 * 5, 128, dtype=torch.bfloat16, device=cuda)
    cls = torch.randn(128, generator=g).to(cuda)
    lib.fill_cls(buf, cls, 4, 5, 128)
    assert torch.equal(buf.view(4, 5, 128)[:, 0], cls.to(torch.bfloat16)[None].expand(4, -1))
    assert (buf.view(4, 5, 128)[:, 1:] == 0).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_window_pool(cuda, lib, mode):
    g = torch.Generator().manual_seed(8)
    H, B = 128, 3
    if mode == 0:
        win, n_win, gh, gw = 5, 6, 0, 0
        T = 31
        idx = torch.arange(30).view(6, 5)
    else:
        gh, gw = 4, 6
        win, n_win, T = 4, 6, 25
        idx = torch.arange(24).view(2, 2, 3, 2).permute(0, 2, 1, 3).reshape(6, 4)
    x = _bf(torch.randn(B * T, H, generator=g)).to(cuda)
    xs = x.view(B, T, H)[:, 1:].float()
    windows = xs[:, idx.to(cuda)]                       # [B, n_win, win, H]
    _close(lib.window_mean(x, B, T, H, mode, win, n_win, gh, gw), windows.mean(2).reshape(B * n_win, H),
           1e-2, 1e-2, "window mean")
    q = _bf(torch.randn(B * n_win, H, generator=g)).to(cuda)
    kv = _bf(torch.randn(B * T, 2 * H, generator=g)).to(cuda)
    kvw = kv.view(B, T, 2 * H)[:, 1:].float()[:, idx.to(cuda)]   # [B, n_win, win, 2H]
    k, v = kvw[..., :H].reshape(B, n_win, win, H // 64, 64), kvw[..., H:].reshape(B, n_win, win, H // 64, 64)
    s = torch.einsum("bnhd,bnwhd->bnhw", q.float().view(B, n_win, H // 64, 64) / 8, k)
    ref = torch.einsum("bnhw,bnwhd->bnhd", torch.softmax(s, -1), v).reshape(B * n_win, H)
    _close(lib.window_attention(q, kv, B, T, H, mode, win, n_win, gh, gw), ref, 2e-2, 2e-2, "window attention")


def test_unit_normalize_segment_pool_rowdot(cuda, lib):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(37, 768, generator=g).to(cuda)
    ref = x / torch.sqrt((x * x).sum(-1, keepdim=True) + 1e-5)
    _close(lib.unit_normalize(x), ref, 1e-6, 1e-5, "unit_normalize f32")
    _close(lib.unit_normalize(_bf(x), scale=2.0), 2 * _bf(x).float() / torch.sqrt((_bf(x).float() ** 2).sum(-1, keepdim=True) + 1e-5),
           1e-6, 1e-5, "unit_normalize bf16 in")
    B, L, H, S = 3, 50, 128, 7
    v = torch.randn(B, L, H, generator=g).to(cuda)
    do_pool = (torch.rand(B, L, generator=g) > 0.6).to(torch.uint8).to(cuda)
    idx = torch.randint(0, S + 2, (B, L), generator=g).to(torch.int32).to(cuda)   # some out of range
    pooled, count = lib.segment_pool(do_pool, idx, v, S)
    oh = ((idx[..., None] == torch.arange(S, device=cuda)) & do_pool.bool()[..., None]).float()
    _close(pooled, torch.einsum("bls,blh->bsh", oh, v), 1e-5, 1e-5, "segment pool")
    assert torch.equal(count, oh.sum(1))
    w = torch.randn(H, generator=g).to(cuda)
    xb = _bf(v.view(B * L, H))
    rows = torch.tensor([3, 0, 77, 149], dtype=torch.int32, device=cuda)
    _close(lib.row_dot(xb, w, rows), xb.float()[rows.long()] @ w, 1e-4, 1e-4, "row dot")


@pytest.mark.parametrize("n,n_all,H,pos_off", [
    (96, 384, 768, 0),        # one row tile, 3 column tiles (cluster of 3)
    (128, 128, 768, 0),       # single tile, cluster of 1
    (384, 3072, 768, 1152),   # cfg 3 stuff_to_span on 8 ranks: 3 row tiles x 24 column tiles over a cluster of 8
    (96, 776, 1024, 680),     # ragged: n_all not a multiple of 128 (and of 8), Large hidden size, diagonal in the tail tile
    (1, 1, 64, 0),            # the smallest legal problem
    (130, 1100, 264, 3),      # ragged rows, K not a multiple of 64, diagonal crossing a tile boundary
])
def test_contrastive_lse(cuda, lib, n, n_all, H, pos_off):
    """mr_contrastive_lse_fwd (fused tcgen05 logits + online LSE + diagonal positive) against fp32 torch."""
    g = torch.Generator().manual_seed(10 + n)
    x = _bf(torch.nn.functional.normalize(torch.randn(n, H, generator=g), dim=-1) * 4.4).to(cuda)
    y = _bf(torch.nn.functional.normalize(torch.randn(n_all, H, generator=g), dim=-1) * 4.4).to(cuda)
    y[pos_off:pos_off + n] = _bf(0.7 * x.float() + 0.3 * y[pos_off:pos_off + n].float())     # positives that stand out
    lse, numer = lib.contrastive_lse(x, y, pos_off)
    logits = x.float() @ y.float().t()
    _close(lse, torch.logsumexp(logits, -1), 1e-3, 1e-4, "lse")
    _close(numer, (x.float() * y[pos_off:pos_off + n].float()).sum(-1), 1e-3, 1e-4, "numer")


def test_rotary_table_and_assemble(cuda, lib):
    import sys
    from oracle import reserve_oracle as O
    g = torch.Generator().manual_seed(11)
    for nd in (1, 2, 4):
        c = (torch.rand(40, nd, generator=g) * 2 - 1)
        _close(lib.rotary_table(c.to(cuda).contiguous()), O.rotary_scale_table(c).to(cuda), 2e-5, 1e-5, f"rotary table nd={nd}")
    # joint assembly against the oracle's prepare_multimodal_inputs
    from merlot_reserve_b200.synthetic import tiny_config, synthetic_params, synthetic_clip
    cfg = tiny_config()
    H = cfg["model"]["hidden_size"]
    params = synthetic_params(cfg)
    orc = O.OracleMerlotReserve(cfg, params)
    clip = synthetic_clip(cfg, nseg=4)
    tokens = torch.from_numpy(clip["tokens"])[None]
    sub = torch.from_numpy(clip["subseg_idxs"])[None]
    vis = torch.randn(1, 4 * 6, H, generator=g)
    aud = torch.randn(1, 12, 6, H, generator=g)
    mm = orc.prepare_multimodal_inputs(tokens, token_segment_idx=torch.div(sub, 3, rounding_mode="floor"),
                                       vision_input=_bf(vis).float(), audio_spans=_bf(aud).float(),
                                       audio_pointers=sub, padding_len=200)
    emb = _bf(torch.from_numpy(params["token_encoder"]["Embed_0"]["embedding"])).to(cuda)
    x, rot, valid = lib.embed_assemble(tokens.int().to(cuda), emb, H=H,
                                       seg_idx=torch.div(sub, 3, rounding_mode="floor").int().to(cuda),
                                       aud_ptr=sub.int().to(cuda), audio=_bf(aud).to(cuda), vision=_bf(vis).to(cuda),
                                       n_aud=12, n_vis=24, pool_h=2, pool_w=3, T=200)
    ref_x = mm["x"][0].clone()
    tok_rows = (tokens[0] != 5)
    ref_x[:160][tok_rows] = _bf(ref_x[:160][tok_rows]).float()      # embedding table is bf16 on the device
    _close(x, ref_x.to(cuda), 1e-6, 1e-6, "assembled x")
    _close(rot, O.rotary_scale_table(mm["rotary_coords"][0]).to(cuda), 3e-5, 1e-5, "assembled rot")
    assert torch.equal(valid.cpu().bool(), mm["is_valid"][0])


@pytest.mark.parametrize("M,N,K", [(96, 12, 128), (50, 100, 64), (300, 188, 768)])
def test_gemm_ragged_n(cuda, lib, M, N, K):
    """N need not be a multiple of 8 (contrastive logits against 12 b / 188 b targets): only ldd is padded."""
    g = torch.Generator().manual_seed(N)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    ld = (N + 7) // 8 * 8
    out = torch.full((M, ld), 7.0, device=cuda)
    lib.gemm(a, w, out=out[:, :N] if ld == N else out)
    d = lib.GemmDesc(M=M, N=N, K=K, lda=K, ldd=ld, epilogue=lib.EPI_OUT_F32, algo=0, q_scale=1.0)
    lib._call("mr_gemm_epilogue", d, [a, w, None, None, None], [out])
    _close(out[:, :N], a.float() @ w.float().t(), 2e-3, 1e-3, "ragged N")
    assert (out[:, N:] == 7.0).all()          # the padding columns are never written


def test_pretrain_row_kernels(cuda, lib):
    """The additions of the pretraining forward: gathered unit_normalize, accumulating segment pool, the source
    row maps of embed_assemble, the scalar loss reduction (pretrain_model.py:97-135,176-230,295-300)."""
    g = torch.Generator().manual_seed(12)
    x = torch.randn(40, 256, generator=g).to(cuda)
    idx = torch.tensor([5, 5, 39, 0, 17], dtype=torch.int32, device=cuda)
    ref = x[idx.long()]
    ref = 3.0 * ref / torch.sqrt((ref * ref).sum(-1, keepdim=True) + 1e-5)
    got = lib.unit_normalize(x, scale=3.0, out_f32=False, row_idx=idx)
    assert got.dtype == torch.bfloat16 and got.shape == (5, 256)
    _close(got, ref, 1e-2, 1e-2, "gathered unit_normalize")

    B, L, H, S = 2, 40, 128, 5
    v1, v2 = torch.randn(B, L, H, generator=g).to(cuda), torch.randn(B, L, H, generator=g).to(cuda)
    dp1 = (torch.rand(B, L, generator=g) > 0.5).to(torch.uint8).to(cuda)
    dp2 = (torch.rand(B, L, generator=g) > 0.5).to(torch.uint8).to(cuda)
    ix = torch.randint(0, S, (B, L), generator=g).to(torch.int32).to(cuda)
    acc = lib.segment_pool(dp1, ix, v1, S)
    lib.segment_pool(dp2, ix, v2, S, accumulate_into=acc)
    oh = lambda dp: ((ix[..., None] == torch.arange(S, device=cuda)) & dp.bool()[..., None]).float()
    _close(acc[0], torch.einsum("bls,blh->bsh", oh(dp1), v1) + torch.einsum("bls,blh->bsh", oh(dp2), v2), 1e-5, 1e-5, "acc pool")
    assert torch.equal(acc[1], oh(dp1).sum(1) + oh(dp2).sum(1))

    n = 300
    lse, numer = torch.randn(n, generator=g).to(cuda) + 5, torch.randn(n, generator=g).to(cuda)
    src = torch.randint(-1, 3, (n,), generator=g).to(torch.int32).to(cuda)
    red = lib.contrastive_reduce(lse, numer, src).cpu()
    dlt = (lse - numer).cpu()
    assert abs(float(red[0]) - float(dlt.mean())) < 1e-5
    for s in range(3):
        m = (src.cpu() == s).float()
        assert abs(float(red[1 + s]) - float((dlt * m).sum() / (m.sum() + 1e-5))) < 1e-4
    assert abs(float(lib.contrastive_reduce(lse, numer)[0]) - float(dlt.mean())) < 1e-5

    # source row maps: sequence r reads audio row r // aud_div and vision row (r // (G n)) G + r % G
    Hh, Ltok, n_vis, n_aud = 64, 8, 6, 4
    Bv, G, tile = 2, 2, 3
    R = Bv * G * tile
    emb = _bf(torch.randn(50, Hh, generator=g)).to(cuda)
    vis = _bf(torch.randn(Bv * G, n_vis, Hh, generator=g)).to(cuda)
    aud = _bf(torch.randn(Bv, n_aud, 6, Hh, generator=g)).to(cuda)
    toks = torch.randint(6, 50, (R, Ltok), generator=g).to(torch.int32)
    toks[:, 2:5] = 5
    ptr = torch.randint(0, n_aud, (R, Ltok), generator=g).to(torch.int32)
    xo, _, _ = lib.embed_assemble(toks.to(cuda), emb, H=Hh, seg_idx=torch.zeros_like(toks).to(cuda), aud_ptr=ptr.to(cuda),
                                  audio=aud, vision=vis, n_aud=n_aud, n_vis=n_vis, pool_h=2, pool_w=3, T=Ltok + n_vis,
                                  aud_div=G * tile, vis_group=G, vis_tile=tile)
    xo = xo.view(R, Ltok + n_vis, Hh)
    for r in range(R):
        vb = (r // (G * tile)) * G + r % G
        assert torch.equal(xo[r, Ltok:], vis[vb])
        for t in range(2, 5):
            assert torch.equal(xo[r, t], aud[r // (G * tile), ptr[r, t], t - 2])


@pytest.mark.parametrize("M,H,N,mode", [(1000, 768, 2304, "rotary"), (300, 1024, 4096, "gelu"), (77, 128, 384, "plain"),
                                        (4000, 768, 3072, "gelu")])
def test_gemm_layernorm_fold(cuda, lib, M, H, N, mode):
    """MR_EPI_LN_IN / MR_EPI_STATS_OUT: Dense(LayerNorm(x)) with the LayerNorm folded into the GEMM epilogue, fed
    by row statistics that (a) the LayerNorm kernel and (b) a residual GEMM emit for their own bf16 outputs.
    Reference: flax LayerNorm in fp32 (var = E[x^2] - E[x]^2, eps 1e-5) then the fp32 matmul (modeling.py:272-279)."""
    g = torch.Generator().manual_seed(M + N)
    gamma = (1 + 0.1 * torch.randn(H, generator=g))
    beta = 0.1 * torch.randn(H, generator=g)
    W = torch.randn(H, N, generator=g) * 0.05          # flax layout [in, out]
    b = 0.1 * torch.randn(N, generator=g)
    wt = (W * gamma[:, None]).t().contiguous().to(torch.bfloat16)
    colsum = wt.float().sum(1).to(cuda)
    bias = (beta @ W + b).to(cuda)

    def ref_of(x_bf16):
        x = x_bf16.float().cpu()
        mu = x.mean(-1, keepdim=True)
        var = (x * x).mean(-1, keepdim=True) - mu * mu
        h = (x - mu) * torch.rsqrt(var + 1e-5) * gamma + beta
        return h @ W + b

    # (a) x comes out of a LayerNorm kernel that also emits the statistics; rows carry a large common offset
    x0 = _bf(torch.randn(M, H, generator=g) * 2 + 3 * torch.randn(M, 1, generator=g)).to(cuda)
    g0, b0 = (1 + 0.1 * torch.randn(H, generator=g)).to(cuda), (0.5 + 0.1 * torch.randn(H, generator=g)).to(cuda)
    stats = torch.full((H // 64, M, 2), float("nan"), device=cuda)
    x = lib.layernorm(x0, g0, b0, stats_out=stats)
    xs = x.float()
    _close(stats.sum(0)[:, 0], xs.sum(-1), 1e-2, 1e-4, "LN stats sum")
    _close(stats.sum(0)[:, 1], (xs * xs).sum(-1), 1e-2, 1e-4, "LN stats sumsq")
    kw = {}
    ref = ref_of(x)
    if mode == "rotary":
        hid = N // 3
        rot = (torch.rand(7, 32, generator=g) * 2 - 1).to(cuda)
        kw = dict(rot=rot, hidden=hid, q_scale=0.125)
        r = ref.clone().view(M, 3, hid // 64, 64)
        sc = rot.cpu()[torch.arange(M) % 7]
        r[:, :2, :, :32] *= sc[:, None, None, :]
        r[:, 0] *= 0.125
        ref = r.view(M, N)
    elif mode == "gelu":
        kw = dict(gelu=True)
        ref = ref * torch.sigmoid(1.702 * ref)
    got = lib.gemm(x, wt.to(cuda), bias=bias, ln_stats=stats, ln_colsum=colsum, **kw)
    _close(got.cpu(), ref, 3e-2, 2e-2, f"LN-folded gemm ({mode}) on LayerNorm-kernel statistics")

    # (b) x comes out of a residual GEMM that emits the statistics of its rounded output
    if N % 64 == 0:
        a = _bf(torch.randn(M, 256, generator=g)).to(cuda)
        w2 = _bf(torch.randn(H, 256, generator=g) * 0.1).to(cuda)
        stats2 = torch.full((H // 64, M, 2), float("nan"), device=cuda)
        x2 = x.clone()
        lib.gemm(a, w2, residual=x2, out=x2, stats_out=stats2)
        xs = x2.float()
        _close(x2, x.float() + a.float() @ w2.float().t(), 5e-2, 2e-2, "residual gemm")
        _close(stats2.sum(0)[:, 0], xs.sum(-1), 1e-2, 1e-4, "gemm stats sum")
        _close(stats2.sum(0)[:, 1], (xs * xs).sum(-1), 1e-2, 1e-4, "gemm stats sumsq")
        got = lib.gemm(x2, wt.to(cuda), bias=bias, ln_stats=stats2, ln_colsum=colsum)
        _close(got.cpu(), ref_of(x2), 3e-2, 2e-2, "LN-folded gemm on GEMM statistics")


@pytest.mark.parametrize("M", [19000, 19000 + 100, 19000 + 200, 256 * 74])
@pytest.mark.parametrize("mode", ["plain", "bias", "gelu", "residual_stats", "rotary_ln"])
def test_gemm_cta_pairs(cuda, lib, M, mode):
    """cta_group::2 path (256 x 256 pair tiles, leader-issued MMAs over both CTAs' shared memory) against the
    single-CTA path and the fp32 reference, for every epilogue of the transformer layer; M tails that leave the
    second CTA of the last pair partially or completely out of range."""
    H = 768
    g = torch.Generator().manual_seed(M)
    a = _bf(torch.randn(M, H, generator=g)).to(cuda)
    N = {"plain": 768, "bias": 1024, "gelu": 3072, "residual_stats": 768, "rotary_ln": 2304}[mode]
    w = _bf(torch.randn(N, H, generator=g) * 0.05).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    outs = {}
    for pair in (1, 2):
        if mode == "plain":
            outs[pair] = lib.gemm(a, w, pair=pair)
        elif mode == "bias":
            outs[pair] = lib.gemm(a, w, bias=b, pair=pair)
        elif mode == "gelu":
            outs[pair] = lib.gemm(a, w, bias=b, gelu=True, pair=pair)
        elif mode == "residual_stats":
            x = _bf(torch.randn(M, N, generator=torch.Generator().manual_seed(1))).to(cuda)
            st = torch.zeros((N // 64, M, 2), device=cuda)
            lib.gemm(a, w, residual=x, out=x, stats_out=st, pair=pair)
            outs[pair] = (x, st)
        else:
            st = torch.zeros((H // 64, M, 2), device=cuda)
            xs = a.float()
            st[0, :, 0], st[0, :, 1] = xs.sum(-1), (xs * xs).sum(-1)
            rot = (torch.rand(241, 32, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(cuda)
            outs[pair] = lib.gemm(a, w, bias=b, rot=rot, hidden=H, q_scale=0.125, ln_stats=st,
                                  ln_colsum=w.float().sum(1).contiguous(), pair=pair)
    if mode == "residual_stats":
        assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])
        x0 = _bf(torch.randn(M, N, generator=torch.Generator().manual_seed(1))).to(cuda).float()
        _close(outs[1][0], x0 + a.float() @ w.float().t(), 5e-2, 2e-2, "pair residual")
    else:
        assert torch.equal(outs[1], outs[2]), f"{mode}: pair and single-CTA results differ"
        if mode in ("plain", "bias", "gelu"):
            ref = a.float() @ w.float().t() + (b if mode != "plain" else 0)
            if mode == "gelu":
                ref = ref * torch.sigmoid(1.702 * ref)
            _close(outs[1], ref, 3e-2, 2e-2, f"pair {mode}")


@pytest.mark.parametrize("M,N,K", [(1000, 768, 3072), (333, 200, 136), (4096, 768, 768), (128, 256, 64)])
def test_gemm_dgrad_transposed_b(cuda, lib, M, N, K):
    """dgrad form: D[M, N] = A[M, K] . Bt[K, N] with Bt consumed MN-major straight from its [K, N] layout
    (dX = dY . W on the forward's packed [N_f, K_f] weights, no transposed copy)."""
    g = torch.Generator().manual_seed(M + K)
    a = _bf(torch.randn(M, K, generator=g)).to(cuda)
    bt = _bf(torch.randn(K, N, generator=g) * 0.05).to(cuda)
    ref = a.float() @ bt.float()
    _close(lib.gemm_bwd(a, bt, trans_b=True), ref, 2e-3, 1e-3, "dgrad f32")
    _close(lib.gemm_bwd(a, bt, trans_b=True, out_f32=False), ref, 3e-2, 1e-2, "dgrad bf16 (staged)")


@pytest.mark.parametrize("M,N,K,split", [(768, 3072, 5000, 0), (768, 3072, 5000, 4), (2304, 768, 30848, 8), (136, 200, 333, 0),
                                         (136, 200, 333, 3), (128, 256, 64, 0), (3072, 768, 777 * 8, 5)])
def test_gemm_wgrad_both_transposed(cuda, lib, M, N, K, split):
    """wgrad form: D[M, N] = At[K, M]^T . Bt[K, N] (dW = dY^T . X, reduction over the token rows), both operands
    MN-major, optional split-K with a deterministic second-stage sum."""
    g = torch.Generator().manual_seed(M + K + split)
    at = _bf(torch.randn(K, M, generator=g)).to(cuda)
    bt = _bf(torch.randn(K, N, generator=g) * 0.05).to(cuda)
    ref = at.float().t() @ bt.float()
    got = lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split)
    _close(got, ref, 5e-3 * math.sqrt(K / 64), 2e-3, f"wgrad split={split}")
    if split:
        assert torch.equal(got, lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split))   # deterministic
        _close(lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split, out_f32=False), ref,
               5e-2 * math.sqrt(K / 64), 1e-2, "wgrad bf16 out")


def test_gemm_transposed_a_only(cuda, lib):
    g = torch.Generator().manual_seed(5)
    K, M, N = 520, 264, 392
    at = _bf(torch.randn(K, M, generator=g)).to(cuda)
    w = _bf(torch.randn(N, K, generator=g) * 0.05).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    _close(lib.gemm_bwd(at, w, trans_a=True, bias=b), at.float().t() @ w.float().t() + b, 2e-3, 1e-3, "A^T . W^T + b")


@pytest.mark.parametrize("M,N,K,tile_n,split", [(3072, 768, 30848, 256, 2), (768, 3072, 30848, 256, 2), (2304, 768, 30848, 256, 5),
                                                (768, 768, 30720, 256, 8), (768, 768, 11904, 128, 4), (520, 512, 1000, 256, 3),
                                                (520, 512, 1000, 256, 0), (520, 512, 1000, 128, 0)])
def test_gemm_wgrad_forced_tile(cuda, lib, M, N, K, tile_n, split):
    """MrGemmDesc.tile_n: the wgrad plans of `wgrad_plan` (256-wide tiles with few K parts) against fp32 matmul and
    against the library's own tile choice (same partial order per element => same bits when the split is equal)."""
    g = torch.Generator().manual_seed(M + K + split + tile_n)
    at = _bf(torch.randn(K, M, generator=g)).to(cuda)
    bt = _bf(torch.randn(K, N, generator=g) * 0.05).to(cuda)
    ref = at.float().t() @ bt.float()
    got = lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split, tile_n=tile_n)
    _close(got, ref, 5e-3 * math.sqrt(K / 64), 2e-3, f"wgrad tile_n={tile_n} split={split}")
    assert torch.equal(got, lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split, tile_n=tile_n))
    auto = lib.gemm_bwd(at, bt, trans_a=True, trans_b=True, split_k=split)
    _close(got, auto, 1e-3 * math.sqrt(K / 64), 1e-3, "forced tile vs auto tile")


def test_gemm_tile_n_rejects_bad_width(cuda, lib):
    a = _bf(torch.randn(64, 128)).to(cuda)
    b = _bf(torch.randn(64, 200)).to(cuda)
    with pytest.raises(lib.MrError):
        lib.gemm_bwd(a, b, trans_a=True, trans_b=True, tile_n=256)      # N % 256 != 0
    with pytest.raises(lib.MrError):
        lib.gemm_bwd(a, b, trans_a=True, trans_b=True, tile_n=64)


@pytest.mark.parametrize("M,N,K", [(1000, 3072, 768), (333, 264, 136), (4096, 3072, 768), (30848, 3072, 768)])
def test_gemm_dgrad_fused_gelu_backward(cuda, lib, M, N, K):
    """MR_EPI_DGELU: d(pre) = (dY . W) * gelu'(pre) in the dgrad epilogue == dgrad followed by mr_gelu(mode 1),
    and == autograd of x * sigmoid(1.702 x) (modeling.py:240)."""
    g = torch.Generator().manual_seed(M + N)
    dy = _bf(torch.randn(M, K, generator=g)).to(cuda)
    wt = _bf(torch.randn(K, N, generator=g) * 0.05).to(cuda)            # packed [out = K_f, in = N_f] weight
    pre = _bf(torch.randn(M, N, generator=g) * 2.0).to(cuda)
    fused = lib.gemm_bwd(dy, wt, trans_b=True, out_f32=False, dgelu_pre=pre)
    two_pass = lib.gelu(pre, lib.gemm_bwd(dy, wt, trans_b=True, out_f32=False))
    p = pre.float().requires_grad_(True)
    (p * torch.sigmoid(1.702 * p)).backward(dy.float() @ wt.float())
    _close(fused, p.grad, 3e-2, 2e-2, "fused dgelu vs autograd")
    _close(fused, two_pass.float(), 3e-2, 2e-2, "fused dgelu vs dgrad + gelu kernel")


@pytest.mark.parametrize("B,T,H", [(5, 31, 768), (7, 16, 128), (3, 32, 256), (4, 17, 128), (2, 1, 128), (300, 31, 768)])
def test_attention_short_sequences_single_tile(cuda, lib, B, T, H):
    """T <= 32 under MR_ATTN_ALGO_AUTO runs the one-tile, two-warp instantiation (audio T = 31, span T = 16):
    against the fp32 reference, with validity + group masks, and its log-sum-exp against the 64 x 64 kernel's."""
    g = torch.Generator().manual_seed(T * 7 + B)
    qkv = _bf(torch.randn(B * T, 3 * H, generator=g)).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    _close(lib.attention(qkv, B, T, H, algo=0), _attn_ref(qkv, B, T, H), 2e-2, 2e-2, f"short T={T}")
    valid = (torch.rand(B, T, generator=g) > 0.3).to(torch.uint8)
    valid[:, 0] = 1
    valid[0, T // 2 + 1:] = 0
    group = torch.randint(0, 3, (B, T), generator=g).to(torch.int32)
    valid, group = valid.to(cuda).contiguous(), group.to(cuda).contiguous()
    lse0 = torch.empty((B, H // 64, T), dtype=torch.float32, device=cuda)
    lse1 = torch.empty_like(lse0)
    o0 = lib.attention(qkv, B, T, H, valid=valid.view(-1), group=group.view(-1), algo=0, lse=lse0)
    o1 = lib.attention(qkv, B, T, H, valid=valid.view(-1), group=group.view(-1), algo=1, lse=lse1)
    _close(o0, _attn_ref(qkv, B, T, H, valid, group), 2e-2, 2e-2, "short valid+group")
    _close(o0, o1.float(), 1e-2, 1e-2, "one-tile vs 64 x 64 kernel")
    _close(lse0, lse1, 1e-4, 1e-5, "lse")


@pytest.mark.parametrize("B,gh,gw", [(3, 12, 20), (1, 18, 32), (2, 2, 2), (0, 12, 20)])
def test_patchify_u8_is_bit_exact(cuda, lib, B, gh, gw):
    """mr_patchify_u8 (convert_image_dtype + space_to_depth(16) + cast to bf16 on the device) against the numpy
    restatement of preprocess.py:464-481 followed by the fp32 -> bf16 cast the model applies: identical bits."""
    from oracle.reserve_oracle import frames_to_patches
    rng = np.random.default_rng(B * 100 + gh)
    frames = rng.integers(0, 256, size=(B, gh * 16, gw * 16, 3), dtype=np.uint8)
    got = lib.patchify_u8(torch.from_numpy(frames).to(cuda), gh, gw)
    assert got.shape == (B * gh * gw, 768) and got.dtype == torch.bfloat16
    if B == 0:
        return
    ref32 = torch.from_numpy(frames_to_patches(frames, gh, gw)).reshape(B * gh * gw, 768)
    assert torch.equal(got.cpu(), ref32.to(torch.bfloat16))
    # and identical to the existing path: fp32 patches through mr_cast_pad
    via_cast = lib.cast_pad(ref32.to(cuda), B * gh * gw, 768, 768)
    assert torch.equal(got, via_cast)
