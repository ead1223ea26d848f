"""Edge cases of the reference-facing API against the CPU oracle: empty batches, ragged token inputs (all padding,
no audio spans, a single segment), one-token and empty spans, and the longest sequences of the BASELINE configs
(joint T = 2464 = 160 + 16 x 144 at Large 18 x 32) through the masked attention kernel."""
import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReserve
from merlot_reserve_b200.synthetic import synthetic_clip, synthetic_params, synthetic_spans, tiny_config
from oracle.reserve_oracle import OracleMerlotReserve

pytestmark = pytest.mark.gpu
COS_MIN = 0.999


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a)).float().reshape(-1, np.asarray(a).shape[-1])
    b = torch.as_tensor(np.asarray(b)).float().reshape(-1, np.asarray(b).shape[-1])
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


@pytest.fixture(scope="module")
def setup(cuda):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    model = MerlotReserve.from_config(cfg)
    model.bind(params)
    return cfg, params, model, OracleMerlotReserve(cfg, params)


def test_empty_batches(setup):
    cfg, params, model, orc = setup
    H = cfg["model"]["hidden_size"]
    gh, gw = cfg["model"]["output_grid"]
    hw = gh * gw
    assert model.embed_text_spans_only(np.zeros((0, 15), np.int32)).shape == (0, H)
    assert model.get_imgseq_only(np.zeros((0, hw, 768), np.float32)).shape == (0, hw // 4, H)
    assert model.get_audioseq_only(np.zeros((0, 60, 65), np.float32)).shape == (0, 6, H)
    assert model.embed_audio_only(np.zeros((0, 60, 65), np.float32)).shape == (0, H)
    out = model.batch_embed_video(np.zeros((0, 2, hw, 768), np.float32), np.zeros((0, 6, 60, 65), np.float32),
                                  np.zeros((0, 160), np.int32), np.zeros((0, 160), np.int32))
    assert out.shape == (0, 160, H)


def test_ragged_token_inputs(setup):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=2, seed=17)
    L = clip["tokens"].shape[0]
    cases = {}
    # every position padding (token 0, subseg -1): all query rows attend uniformly, vision keys stay valid
    cases["all_padding"] = (np.zeros(L, np.int32), np.full(L, -1, np.int32))
    # text only: no AUDIOSPAN anywhere, everything in segment 0
    tok = np.random.default_rng(0).integers(10, 30000, L).astype(np.int32)
    cases["text_only_full_length"] = (tok, np.zeros(L, np.int32))
    # a single real token
    one = np.zeros(L, np.int32)
    one[0] = 77
    sub = np.full(L, -1, np.int32)
    sub[0] = 0
    cases["one_token"] = (one, sub)
    for name, (t, s) in cases.items():
        got = model.embed_video(clip["images"], clip["audio_clips"], t, s)
        ref = orc.embed_video(clip["images"], clip["audio_clips"], t, s)
        assert np.isfinite(got).all(), name
        cos = _cos(got, ref)
        assert cos.min() >= COS_MIN, f"{name}: min cosine {cos.min():.5f} at row {int(cos.argmin())}"


def test_single_segment_clip(setup):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=1, seed=23)
    got = model.embed_video(**clip)
    assert _cos(got, orc.embed_video(**clip)).min() >= COS_MIN


def test_span_lengths_zero_one_and_full(setup):
    cfg, params, model, orc = setup
    spans = synthetic_spans(6, seed=2).copy()
    spans[0] = 0                      # empty span: only the CLS position is a valid key
    spans[1] = 0
    spans[1, 0] = 1234                # one token
    spans[2] = np.arange(100, 115)    # all 15 positions used
    got = model.embed_text_spans_only(spans)
    ref = orc.embed_text_spans_only(spans)
    assert np.isfinite(got).all()
    assert _cos(got, ref).min() >= COS_MIN
    one = model.embed_text_spans_only(spans[2:3])
    assert _cos(one, got[2:3]).min() >= 0.99999


def test_longest_joint_sequence_masked_attention(cuda, lib):
    """T = 2464 (Large, 16 segments of 18 x 32): 20 key blocks, text padding inside the first 160 positions."""
    B, T, H = 1, 2464, 128
    g = torch.Generator().manual_seed(5)
    qkv = (torch.randn(B * T, 3 * H, generator=g)).to(torch.bfloat16).to(cuda)
    qkv.view(B * T, 3, H)[:, 0] *= 0.35
    valid = torch.ones(B, T, dtype=torch.uint8)
    valid[:, 14:160] = 0
    valid = valid.to(cuda)
    nh = H // 64
    q, k, v = qkv.float().view(B, T, 3, nh, 64).unbind(2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k)
    vb = valid.bool()
    s = s + torch.where(vb[:, :, None] & vb[:, None, :], 0.0, -1e10)[:, None]
    ref = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B * T, H)
    got = lib.attention(qkv, B, T, H, valid=valid.view(-1)).float()
    err = (got - ref).abs()
    assert float(err.max()) <= 2e-2 + 2e-2 * float(ref.abs().max()), float(err.max())
