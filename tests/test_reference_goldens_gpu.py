"""CUDA path vs the REFERENCE-EXECUTED fixtures (tests/golden_ref/, written by tools/run_reference_shim.py from the
reference's own mreserve/modeling.py / pretrain/pretrain_model.py / finetune heads; see tests/test_reference_goldens.py
and oracle/refshim/README.md for what that pins).  No oracle in the loop here: the kernels are compared with what
the reference code itself produced in fp32 on the same seeded weights and inputs.

Tolerances (BASELINE.json north_star): embedding cosine >= 0.999 per row; contrastive logits within
2e-2 * scale + 2e-2 * |ref| (scale = exp(log_scale) = 20); identical argmax on multiple-choice scoring where the
reference's top-2 margin exceeds the logit tolerance.  The reference's own bf16 module dtype moves embed_video rows
by up to 1 - cosine = 3e-4 on this configuration (tests/test_reference_goldens.py::test_bf16_module_dtype_noise_floor).
"""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReserve, MerlotReservePretrainer, loss_fn_given_preds
from merlot_reserve_b200.synthetic import (load_config, synthetic_clip, synthetic_params, synthetic_pretrain_batch,
                                           synthetic_spans, synthetic_tvqa_batch, tiny_config)

pytestmark = pytest.mark.gpu
REF = Path(__file__).resolve().parent / "golden_ref"
COS_MIN = 0.999
MASK, AUDIOSPAN = 3, 5


def _np(x):
    return x.float().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _cos(a, b):
    a, b = _np(a), _np(b)
    a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).reshape(-1, a.shape[-1])
    b = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).reshape(-1, b.shape[-1])
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


@pytest.fixture(scope="module", params=["tiny_init", "tiny_scale4"])
def case(request, cuda):
    ref = np.load(REF / f"{request.param}.npz")
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=float(ref["weight_scale"]))
    return cfg, params, MerlotReserve.from_config(cfg), ref


def test_zero_shot_api_vs_reference(case):
    cfg, params, model, ref = case
    ap = lambda m, *a: _np(model.apply({"params": params}, *a, method=m))
    clip, clip2 = synthetic_clip(cfg, nseg=4, seed=21), synthetic_clip(cfg, nseg=4, seed=22)
    K = ("images", "audio_clips", "tokens", "subseg_idxs")
    for name, got in (
            ("embed_video", ap(model.embed_video, *[clip[k] for k in K])),
            ("imgseq", ap(model.get_imgseq_only, clip["images"])),
            ("audioseq", ap(model.get_audioseq_only, clip["audio_clips"])),
            ("audio_cls", ap(model.embed_audio_only, clip["audio_clips"])),
            ("spans", ap(model.embed_text_spans_only, synthetic_spans(12, seed=3))),
            ("batch_embed_video", ap(model.batch_embed_video, *[np.stack([clip[k], clip2[k]], 0) for k in K]))):
        assert got.shape == ref[name].shape, name
        c = _cos(got, ref[name])
        assert c.min() >= COS_MIN, f"{name}: min cosine {c.min():.5f} at row {int(c.argmin())}"
    text_tokens = clip["tokens"].copy()
    text_tokens[text_tokens == AUDIOSPAN] = 77
    for name, got in (
            ("preencoded_noaudio", ap(model.embed_preencoded_noaudio, ref["imgseq"], text_tokens, clip["subseg_idxs"])),
            ("preencoded_audio", ap(model.embed_preencoded_audio, ref["imgseq"], ref["audioseq"], clip["tokens"],
                                    clip["subseg_idxs"], clip["subseg_idxs"])),
            ("multiimg_prompt", ap(model.embed_singleimg_with_multiimg_prompt, ref["imgseq"][:2], clip["images"][2:],
                                   text_tokens, clip["subseg_idxs"]))):
        c = _cos(got, ref[name])
        assert c.min() >= COS_MIN, f"{name}: min cosine {c.min():.5f}"


def test_prepare_multimodal_inputs_vs_reference(case):
    """The reference's return schema (modeling.py:758): x, rotary_coords, attention_mask."""
    cfg, params, model, ref = case
    mm = model.apply({"params": params}, tokens=ref["mm/tokens"], token_segment_idx=ref["mm/seg"],
                     vision_input=ref["mm/vis"], audio_spans=ref["mm/aspans"], audio_pointers=ref["mm/aptr"],
                     padding_len=56, video_src_idx=ref["mm/vsrc"], method=model.prepare_multimodal_inputs)
    x = _np(mm["x"])
    assert np.abs(x - ref["mm/x"]).max() <= 2e-2 * np.abs(ref["mm/x"]).max()          # bf16 storage of fp32 inputs
    assert np.abs(_np(mm["rotary_coords"]) - ref["mm/rotary_coords"]).max() < 1e-6
    assert np.array_equal(_np(mm["attention_mask"]).astype(bool), ref["mm/attention_mask"])


def test_joint_encoder_reference_schema_and_learned_pe_vs_reference(case):
    """joint_transformer(x, rotary_coords=, attention_mask=) with the reference's own argument schema, and the
    learned-`pe` branch (modeling.py:335-341)."""
    cfg, params, model, ref = case
    jt = lambda **kw: _np(model.joint_transformer(ref["mm/x"], **kw)["seq"])
    model.bind(params)
    got = jt(rotary_coords=ref["mm/rotary_coords"], attention_mask=ref["mm/attention_mask"])
    valid = ref["mm/attention_mask"].diagonal(axis1=1, axis2=2)
    c = _cos(got[valid], ref["jt/rotary"][valid])
    assert c.min() >= COS_MIN, float(c.min())
    c = _cos(got, ref["jt/rotary"])                       # fully masked query rows attend uniformly (finite mask)
    assert c.min() >= COS_MIN, float(c.min())
    p2 = dict(params, joint_transformer=dict(params["joint_transformer"], pe=ref["jt/pe"]))
    model.bind(p2)
    got = jt(attention_mask=ref["mm/attention_mask"])
    c = _cos(got, ref["jt/pe_out"])
    assert c.min() >= COS_MIN, float(c.min())
    with pytest.raises(ValueError):                      # a mask the reference never builds is refused, not mis-read
        bad = ref["mm/attention_mask"].copy()
        bad[0, 0, 1] = not bad[0, 0, 1]
        jt(rotary_coords=ref["mm/rotary_coords"], attention_mask=bad)


def test_multiple_choice_heads_vs_reference(case):
    cfg, params, model, ref = case
    rng = np.random.default_rng(17)
    mc_tokens = rng.integers(10, 32768, size=(3, 4, 24)).astype(np.int32)
    mc_tokens[:, :, -1] = MASK
    mc_images = rng.random((3, 24, 768), dtype=np.float32)
    got = _np(model.apply({"params": params}, mc_images, mc_tokens, ref["mc_proj"], method=model.score_multiple_choice))
    want = ref["mc_logits"]
    tol = 2e-2 * np.abs(want).max() + 2e-2
    assert np.abs(got - want).max() <= tol, (got, want)
    top2 = np.sort(want, -1)
    decided = (top2[:, -1] - top2[:, -2]) > 2 * tol
    assert np.array_equal(got.argmax(-1)[decided], want.argmax(-1)[decided])
    tb = synthetic_tvqa_batch(cfg, 2, num_segments=3, num_answers=3, seq_len=96, seed=4)
    la, lt = model.apply({"params": params}, tb, ref["mc_proj"], method=model.score_tvqa)
    for g, w in ((la, ref["tvqa/logits_from_audio"]), (lt, ref["tvqa/logits_from_text"])):
        assert np.abs(_np(g) - w).max() <= 2e-2 * np.abs(w).max() + 2e-2


def test_pretraining_forward_and_loss_vs_reference(case):
    cfg, params, model, ref = case
    pre = MerlotReservePretrainer.from_config(cfg)
    got = pre.apply({"params": params}, synthetic_pretrain_batch(cfg, 2, seed=3))
    for head in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        for k, v in got[head].items():
            want = ref[f"pretrain/{head}/{k}"]
            if k == "_sources":
                assert np.array_equal(v.cpu().numpy().astype(np.int32), want)
            else:
                c = _cos(v, want)
                assert c.min() >= COS_MIN, (head, k, float(c.min()))
        lg = _np(got[head]["x"]) @ _np(got[head]["y"]).T
        lr = ref[f"pretrain/{head}/x"] @ ref[f"pretrain/{head}/y"].T
        assert (np.abs(lg - lr) <= 2e-2 * 20.0 + 2e-2 * np.abs(lr)).all(), head
    loss, info = loss_fn_given_preds(got)
    assert abs(float(loss) - float(ref["pretrain/loss"])) <= 0.1
    for k, v in info.items():
        want = float(ref[f"pretrain/loss_info/{k}"])
        assert abs(float(v) - want) <= 0.05 + 0.02 * abs(want), k


def test_base_one_clip_vs_reference(cuda):
    """BASELINE.json configs[0] at full size: Reserve-Base, one 16-segment 12x20 clip, literal random init."""
    ref = np.load(REF / "base_cfg1.npz")["embed_video"]
    cfg = load_config("base")
    model = MerlotReserve.from_config(cfg)
    got = _np(model.apply({"params": synthetic_params(cfg, seed=0)}, **synthetic_clip(cfg, nseg=16, seed=1),
                          method=model.embed_video))
    c = _cos(got, ref)
    assert got.shape == ref.shape == (160, 768)
    assert c.min() >= COS_MIN, f"min cosine {c.min():.5f} at row {int(c.argmin())}"


def _fullsize():
    p = REF / "fullsize.npz"
    if not p.exists():
        pytest.skip("tests/golden_ref/fullsize.npz not generated")
    return np.load(p)


def test_cfg3_base_b8_pretrain_forward_and_loss_vs_reference(cuda):
    """SURVEY 8d cfg 3 at size: Reserve-Base, 8 examples per GPU -- every contrastive operand row and the loss."""
    ref = _fullsize()
    cfg = load_config("base")
    pre = MerlotReservePretrainer.from_config(cfg)
    got = pre.apply({"params": synthetic_params(cfg, seed=0, weight_scale=4.0)}, synthetic_pretrain_batch(cfg, 8, seed=5))
    for head in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        for k, v in got[head].items():
            want = ref[f"cfg3/{head}/{k}"]
            assert tuple(v.shape) == want.shape, (head, k)
            if k == "_sources":
                assert np.array_equal(v.cpu().numpy().astype(np.int32), want)
            else:
                c = _cos(v, want)
                assert c.min() >= COS_MIN, (head, k, float(c.min()), int(c.argmin()))
    loss, info = loss_fn_given_preds(got)
    assert abs(float(loss) - float(ref["cfg3/loss"])) <= 0.1, (float(loss), float(ref["cfg3/loss"]))
    for k, v in info.items():
        want = float(ref[f"cfg3/loss_info/{k}"])
        assert abs(float(v) - want) <= 0.05 + 0.02 * abs(want), (k, float(v), want)


def test_cfg4_large_nseg16_vs_reference(cuda):
    """SURVEY 8d cfg 4 at size: Reserve-Large, 18x32 grid, the metric's 16-frame clip (joint T = 2464)."""
    ref = _fullsize()["cfg4/embed_video"]
    cfg = load_config("large", (18, 32))
    model = MerlotReserve.from_config(cfg)
    got = _np(model.apply({"params": synthetic_params(cfg, seed=0, weight_scale=4.0)}, **synthetic_clip(cfg, nseg=16, seed=1),
                          method=model.embed_video))
    assert got.shape == ref.shape == (160, 1024)
    c = _cos(got, ref)
    assert c.min() >= COS_MIN, f"min cosine {c.min():.5f} at row {int(c.argmin())}"


def test_cfg5_large_mc_L144_vs_reference(cuda):
    """SURVEY 8d cfg 5 at size: Reserve-Large VCR-style scoring, 1 frame (18x32) x 4 candidates, L = 144 (ragged
    candidate lengths, MASK at the end of each)."""
    ref = _fullsize()
    cfg = load_config("large", (18, 32))
    model = MerlotReserve.from_config(cfg)
    rng = np.random.default_rng(29)
    mc_tokens = rng.integers(10, 32768, size=(2, 4, 144)).astype(np.int32)
    lens = rng.integers(20, 144, size=(2, 4))
    for b in range(2):
        for c in range(4):
            mc_tokens[b, c, lens[b, c]:] = 0
            mc_tokens[b, c, lens[b, c] - 1] = MASK
    assert np.array_equal(mc_tokens, ref["cfg5/mc_tokens"])
    mc_images = rng.random((2, 576, 768), dtype=np.float32)
    got = _np(model.apply({"params": synthetic_params(cfg, seed=0, weight_scale=4.0)}, mc_images, mc_tokens,
                          ref["cfg5/mc_proj"], method=model.score_multiple_choice))
    want = ref["cfg5/mc_logits"]
    tol = 2e-2 * np.abs(want).max() + 2e-2
    assert np.abs(got - want).max() <= tol, (got, want)
    top2 = np.sort(want, -1)
    decided = (top2[:, -1] - top2[:, -2]) > 2 * tol
    assert np.array_equal(got.argmax(-1)[decided], want.argmax(-1)[decided])
