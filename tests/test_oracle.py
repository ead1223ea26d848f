"""CPU tests of the oracle (test infrastructure): the algebraic identities the CUDA design relies on
(SURVEY.md §8c i-v), the committed golden vectors, and the synthetic generators.  No GPU needed."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from merlot_reserve_b200.synthetic import (count_params, load_config, synthetic_clip, synthetic_params,
                                           synthetic_spans, tiny_config)
from oracle import reserve_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_rotary_quirk_equals_scale_table():
    """apply_rotary (modeling.py:116-144) == per-dim scale (sin-cos, sin+cos) on the first 32 dims."""
    g = torch.Generator().manual_seed(0)
    for nd in (1, 2, 4):
        coords = torch.rand(3, 11, nd, generator=g) * 2 - 1
        qk = torch.randn(3, 11, 6, 64, generator=g)
        lit = O.apply_rotary(qk, O.rotary_sinusoids(coords))
        tab = O.rotary_scale_table(coords)                       # [3, 11, 32]
        ref = qk.clone()
        ref[..., :32] = qk[..., :32] * tab[:, :, None, :]
        assert torch.allclose(lit, ref, atol=1e-6)
    # zero coordinates (CLS rows): scale is (-1, +1, -1, +1, ...)
    z = O.rotary_scale_table(torch.zeros(1, 4))
    assert torch.equal(z[0], torch.tensor([-1.0, 1.0] * 16))


def test_rotary_coordinates():
    assert O.rotary_coordinates(5).tolist() == [-2, -1, 1, 2, 3]
    assert O.rotary_coordinates(4).tolist() == [-2, -1, 1, 2]
    assert O.rotary_coordinates(3, center_origin=False).tolist() == [1, 2, 3]
    c = O.rotary_coordinates_2d(2, 4)
    assert c.shape == (8, 2) and math.isclose(c[0, 0].item(), -1 / 5, rel_tol=1e-6) and math.isclose(c[0, 1].item(), -2 / 5, rel_tol=1e-6)
    assert torch.allclose(c[5], torch.tensor([1 / 5, -1 / 5]))
    from merlot_reserve_b200.modeling import get_rotary_coordinates_2d
    assert np.allclose(get_rotary_coordinates_2d(12, 20), O.rotary_coordinates_2d(12, 20).numpy())


def test_one_hot_pool_equals_gather():
    g = torch.Generator().manual_seed(1)
    v = torch.randn(4, 20, 8, generator=g)
    do_pool = torch.rand(4, 20, generator=g) > 0.5
    idx = torch.randint(0, 6, (4, 20), generator=g)
    out = O.one_hot_pool(do_pool, idx, v, 6)
    ref = torch.zeros(4, 6, 8)
    for b in range(4):
        for l in range(20):
            if do_pool[b, l]:
                ref[b, idx[b, l]] += v[b, l]
    assert torch.allclose(out["x"], ref, atol=1e-6)
    folded = O.one_hot_pool(do_pool, idx, v, 6, real_bsize=2)
    assert folded["x"].shape == (2, 6, 8)
    assert torch.allclose(folded["x"][0], ref[0] + ref[1], atol=1e-6)


def test_mask_bias_equals_validity_semantics():
    """-1e10 additive bias == (masked keys get weight 0; a fully masked query attends uniformly to ALL keys)."""
    g = torch.Generator().manual_seed(2)
    B, T, nh = 2, 9, 2
    qkv = torch.randn(B, T, 3 * nh, 64, generator=g)
    valid = torch.tensor([[1, 1, 0, 1, 0, 0, 1, 1, 0], [1, 0, 0, 0, 0, 0, 0, 0, 0]]).bool()
    mask = valid[:, :, None] & valid[:, None, :]
    bias = torch.where(mask[:, None], 0.0, -1e10)
    ref = O.attention(qkv, None, bias)
    q, k, v = qkv[..., :nh, :], qkv[..., nh:2 * nh, :], qkv[..., 2 * nh:, :]
    s = torch.einsum("bqhd,bkhd->bhqk", q / 8, k)
    s = torch.where(valid[:, None, None, :], s, -float("inf"))
    p = torch.softmax(s, -1)
    p = torch.where(valid[:, None, :, None], p, torch.full_like(p, 1.0 / T))
    alt = torch.einsum("bhqk,bkhd->bqhd", p, v)
    assert torch.allclose(ref, alt, atol=1e-5)


def test_audio_conv_equals_reshape_gemm():
    cfg = tiny_config()
    p = synthetic_params(cfg)["audio_encoder"]["embedding"]
    x = torch.randn(3, 60, 65)
    w = torch.from_numpy(p["kernel"])                                 # [2, 65, H]
    conv = torch.nn.functional.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), stride=2).transpose(1, 2)
    gemm = x.reshape(3, 30, 130) @ w.reshape(130, -1)
    assert torch.allclose(conv, gemm, atol=1e-5)


def test_sharded_contrastive_loss_equals_global():
    """8 replicas with rank-major all-gathered negatives == one replica on the global batch, when the
    per-replica losses are averaged (pmean), pretrain_model.py:290,329."""
    g = torch.Generator().manual_seed(3)
    G, n, H = 4, 6, 16
    x = torch.nn.functional.normalize(torch.randn(G * n, H, generator=g), dim=-1) * 3
    y = torch.nn.functional.normalize(torch.randn(G * n, H, generator=g), dim=-1) * 3
    glob, _ = O.contrastive_loss({"a": {"x": x, "y": y}})
    per = []
    for r in range(G):
        sl = slice(r * n, (r + 1) * n)
        # gather(y_local) returns every replica's rows, rank-major
        loss, _ = O.contrastive_loss({"a": {"x": x[sl], "y": y[sl]}},
                                     gather=lambda t: (y if t.shape == y[sl].shape and torch.equal(t, y[sl]) else x))
        per.append(loss)
    assert torch.allclose(torch.stack(per).mean(), glob, atol=1e-5)


def test_prepare_multimodal_inputs_layout():
    cfg = tiny_config()
    orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg))
    H = cfg["model"]["hidden_size"]
    clip = synthetic_clip(cfg, nseg=2)
    tok = torch.from_numpy(clip["tokens"])[None]
    sub = torch.from_numpy(clip["subseg_idxs"])[None]
    aud = torch.arange(6 * 6 * H, dtype=torch.float32).reshape(1, 6, 6, H)
    mm = orc.prepare_multimodal_inputs(tok, token_segment_idx=torch.div(sub, 3, rounding_mode="floor"),
                                       vision_input=torch.ones(1, 12, H), audio_spans=aud, audio_pointers=sub)
    assert mm["x"].shape == (1, 172, H)
    first_audio = int((tok[0] == 5).nonzero()[0])
    # k-th AUDIOSPAN token takes sub-position k % 6 of the span its pointer names
    for k in range(18):
        assert torch.equal(mm["x"][0, first_audio + k], aud[0, 3 + k // 6, k % 6])
    assert mm["is_valid"][0, :30].all() and not mm["is_valid"][0, 40:160].any() and mm["is_valid"][0, 160:].all()
    assert mm["rotary_coords"][0, 159].tolist() == [0.0, 0.0, -1 / 16, 160 / 1024]   # padding: subseg -1 // 3 = -1
    assert mm["rotary_coords"][0, 160, 3] == 0 and mm["rotary_coords"][0, 166, 2] == 1 / 16


def test_embed_video_properties():
    cfg = tiny_config()
    params = synthetic_params(cfg, weight_scale=4.0)
    orc = O.OracleMerlotReserve(cfg, params)
    clip = synthetic_clip(cfg, nseg=2, seed=4)
    out = orc.embed_video(**clip)
    assert out.shape == (160, 128)
    assert torch.allclose(out.norm(dim=-1), torch.ones(160), atol=1e-3)
    # tokens after the text depend on the audio; changing an unused (padding-pointed) audio clip must not matter
    clip2 = dict(clip)
    clip2["audio_clips"] = clip["audio_clips"].copy()
    clip2["audio_clips"][1] += 1.0     # segment 0 is text: its audio slots 0..2 are only reachable via ptr clamp(−1→0)
    out2 = orc.embed_video(**clip2)
    assert torch.allclose(out, out2, atol=1e-6)
    # batch == stack of singles
    b = orc.batch_embed_video(*[np.stack([clip[k], clip[k]]) for k in ("images", "audio_clips", "tokens", "subseg_idxs")])
    assert torch.allclose(b[0], out, atol=1e-6) and torch.allclose(b[1], out, atol=1e-6)


@pytest.mark.parametrize("name", ["tiny_init", "tiny_scale4"])
def test_oracle_reproduces_golden(name):
    import sys
    sys.path.insert(0, str(GOLDEN.parent.parent / "tools"))
    from make_golden import golden_case
    gold = np.load(GOLDEN / f"{name}.npz")
    now = golden_case(float(gold["weight_scale"]))
    for k in gold.files:
        a, b = gold[k], now[k]
        if np.issubdtype(a.dtype, np.integer):
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, atol=2e-5, rtol=1e-4), (k, np.abs(a - b).max())


def test_synthetic_generators_and_param_counts():
    base = load_config("base")
    assert base["model"]["hidden_size"] == 768 and load_config("large")["model"]["vit_num_layers"] == 24
    cfg = tiny_config()
    p1, p2 = synthetic_params(cfg, seed=0), synthetic_params(cfg, seed=0)
    assert np.array_equal(p1["head"]["kernel"], p2["head"]["kernel"])
    k = p1["joint_transformer"]["layer_00"]["attention_layer"]["qkv"]["kernel"]
    assert k.shape == (128, 6, 64) and np.abs(k).max() <= 2 * 0.02 / math.sqrt(2) + 1e-7
    assert p1["vision_encoder"]["transformer"]["layer_00"]["mlp_layer"]["out"]["kernel"].shape == (512, 128)
    clip = synthetic_clip(cfg, nseg=16)
    assert clip["images"].shape == (16, 24, 768) and clip["audio_clips"].shape == (48, 60, 65)
    assert (clip["audio_clips"][..., 64] == 1).all() and clip["tokens"].shape == (160,)
    assert clip["tokens"][11] == 3 and (clip["subseg_idxs"][clip["tokens"] == 0] == -1).all()
    assert synthetic_spans(5).shape == (5, 15)


def test_base_param_count_matches_survey():
    """Base has 316.4 M parameters."""
    n = count_params(synthetic_params(load_config("base")))
    assert abs(n - 316.4e6) / 316.4e6 < 0.01, n


@pytest.mark.parametrize("name,scale", [("tiny_init", 1.0), ("tiny_scale4", 4.0)])
def test_oracle_pretrain_forward_reproduces_golden(name, scale):
    """The committed pretraining vectors (tools/make_golden.py) are what the oracle computes today."""
    from merlot_reserve_b200.synthetic import synthetic_pretrain_batch
    g = np.load(GOLDEN / f"{name}.npz")
    cfg = tiny_config()
    orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=scale))
    preds = orc.pretrain_forward(synthetic_pretrain_batch(cfg, 2, seed=3))
    for head, d in preds.items():
        for k, v in d.items():
            assert np.allclose(v.numpy(), g[f"pretrain/{head}/{k}"], atol=2e-5), (head, k)
    loss, info = O.contrastive_loss(preds)
    assert abs(float(loss) - float(g["pretrain/loss"])) < 1e-4
    # structure of the forward (pretrain_model.py:170-233): 12 b targets + 36 b extras, 48 b spans, sources in {0,1,2}
    assert preds["text_to_audio"]["y"].shape[0] == 24 and preds["text_to_audio"]["y_extra"].shape[0] == 72
    assert preds["stuff_to_span"]["x"].shape[0] == 96 and set(preds["stuff_to_span"]["_sources"].tolist()) <= {0, 1, 2}


def test_pretrain_batch_layout():
    """synthetic_pretrain_batch follows the reference dataloader's pointer conventions (dataloader.py:545-700)."""
    from merlot_reserve_b200.synthetic import synthetic_pretrain_batch
    cfg = load_config("base")
    b = synthetic_pretrain_batch(cfg, 2, seed=9)
    assert b["images"].shape == (2, 16 * 240, 768) and b["audio_clips"].shape == (2, 48 * 60, 65)
    assert b["text2audio"].shape == (2, 2, 160) and b["random_text"].shape == (2, 1, 640) and b["text_spans"].shape == (2, 62, 15)
    for i in range(2):
        t2a, a2t, rt = b["text2audio"][i], b["audio2text"][i], b["random_text"][i]
        assert (t2a == 4).sum() == 12 and (t2a == 3).sum() == 12 and (a2t == 3).sum() == 12 and (rt == 3).sum() == 38
        assert sorted(b["text2audio/text_ptr"][i][t2a == 3].tolist()) == list(range(12))
        assert sorted(b["audio2text/text_ptr"][i][a2t == 3].tolist()) == list(range(12, 24))
        assert sorted(b["random_text/text_ptr"][i][rt == 3].tolist()) == list(range(24, 62))
        atm = b["audio_text_matching"][i, 0]
        assert (atm == 6).sum() == 16 and (atm == 5).sum() == 16 * 18
        # masked audio pointers of group g stay inside that group's 24 sub-segments
        for g in range(2):
            ptr = b["text2audio/audio_ptr"][i, g][t2a[g] == 4]
            assert ((ptr >= 24 * g) & (ptr < 24 * (g + 1))).all()


@pytest.mark.parametrize("scale,min_cos,mean_cos", [(1.0, 0.9999, 0.99995), (4.0, 0.998, 0.9999)])
def test_bf16_graph_noise_floor(scale, min_cos, mean_cos):
    """The oracle's bf16-graph mode (rounding at the points where the reference's dtype=bfloat16 modules round,
    SURVEY.md Appendix A) against its own fp32 mode: the noise floor of the reference's bf16 execution, which is what
    the north_star tolerance (cosine >= 0.999 vs fp32) has to be read against.  With the weight_scale = 4 parameters
    single rows of the bf16 graph already fall to ~0.9987."""
    from oracle.reserve_oracle import bf16_graph
    cfg = tiny_config()
    orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=scale))
    clip = synthetic_clip(cfg, nseg=4, seed=5)
    ref = orc.embed_video(**clip)
    with bf16_graph():
        low = orc.embed_video(**clip)
    again = orc.embed_video(**clip)
    assert torch.equal(ref, again), "the bf16 mode must not leak out of its context"
    cos = torch.nn.functional.cosine_similarity(ref, low, dim=-1)
    assert not torch.equal(ref, low)
    assert cos.min() >= min_cos and cos.mean() >= mean_cos, (float(cos.min()), float(cos.mean()))


@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_bf16_graph_noise_floor_of_contrastive_logits(scale):
    """Same yardstick for the pretraining head: logits x . y^T (temperature 20, |logit| <= 20) of the bf16-graph
    forward against the fp32 forward.  The stated tolerance for the CUDA path, 2e-2 * scale + 2e-2 * |ref| (0.4 at the
    extremes), is ~8x this floor; the loss moves by less than 5e-3."""
    from merlot_reserve_b200.synthetic import synthetic_pretrain_batch
    from oracle.reserve_oracle import bf16_graph, contrastive_loss
    cfg = tiny_config()
    orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=scale))
    batch = synthetic_pretrain_batch(cfg, 2, seed=3)
    p32 = orc.pretrain_forward(batch)
    with bf16_graph():
        p16 = orc.pretrain_forward(batch)
    l32, _ = contrastive_loss(p32)
    l16, _ = contrastive_loss(p16)
    assert abs(float(l32) - float(l16)) < 5e-3
    for head in p32:
        lg32 = p32[head]["x"].float() @ p32[head]["y"].float().T
        lg16 = p16[head]["x"].float() @ p16[head]["y"].float().T
        d = (lg32 - lg16).abs()
        assert float(d.max()) < 0.1, (head, float(d.max()))
        assert bool((d <= 2e-2 * 20.0 + 2e-2 * lg32.abs()).all())


@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_fp64_spot_check_of_the_fp32_oracle(scale):
    """The same restatement in float64 (`fp64()`): on every real token row the fp32 oracle is at 1 - cosine < 1e-10
    of it, so none of the oracle-vs-kernel distance is the oracle's own rounding.  Padding rows are the exception
    that proves a reference behaviour: their queries are fully masked, and only because fp32 (and bf16) absorb the
    scores into the -1e10 bias do they attend UNIFORMLY (modeling.py:343-358) -- in float64 the scores survive next to
    -1e10 and those rows come out differently.  The CUDA path reproduces the fp32 / bf16 behaviour."""
    from oracle.reserve_oracle import fp64
    cfg = tiny_config()
    orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=scale))
    clip = synthetic_clip(cfg, nseg=4, seed=5)
    ref = orc.embed_video(**clip)
    with fp64():
        hi = orc.embed_video(**clip)
    assert ref.dtype == torch.float32 and hi.dtype == torch.float64
    d = 1.0 - torch.nn.functional.cosine_similarity(ref.double(), hi, dim=-1)
    real = torch.from_numpy(clip["tokens"] != 0)
    assert float(d[real].max()) < 1e-10
    assert torch.equal(orc.embed_video(**clip), ref), "fp64 mode must not leak out of its context"


def test_frames_to_patches_is_space_to_depth():
    """oracle.frames_to_patches (preprocess.py:464-481 without the resize) against an independent formulation of
    tf.nn.space_to_depth(NHWC, block 16) + reshape: [B, h1, P, w1, P, 3] -> transpose -> [B, h1*w1, P*P*3]."""
    rng = np.random.default_rng(0)
    gh, gw, P = 3, 5, 16
    f = rng.integers(0, 256, size=(2, gh * P, gw * P, 3), dtype=np.uint8)
    got = O.frames_to_patches(f, gh, gw)
    x = f.astype(np.float32) * np.float32(1.0 / 255.0)
    ref = x.reshape(2, gh, P, gw, P, 3).transpose(0, 1, 3, 2, 4, 5).reshape(2, gh * gw, P * P * 3)
    assert got.dtype == np.float32 and np.array_equal(got, ref)
    assert got.min() >= 0.0 and got.max() <= 1.0 and got[0, 1, 0] == x[0, 0, P, 0]      # patch 1 = block (0, 1)
