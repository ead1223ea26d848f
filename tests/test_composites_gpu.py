"""Composite C-ABI entries (include/mreserve_b200.h: mr_vision_encode, mr_audio_encode, mr_span_encode,
mr_joint_encode, mr_embed_video_batched) -- the one-call-per-sub-module form of
mreserve/modeling.py:391-430, 443-476, 488-504, 651-758 + 841-842, 806-846.

Each is compared (a) with the CPU oracle on the same weights / inputs (cosine >= 0.999, BASELINE.json north_star) and
(b) bit for bit with the step-by-step composition of the per-op entries, which enqueues the same kernels one ctypes
call at a time (`_lib.PROFILE` set).  The CUDA-graph form of batch_embed_video must replay to the same bits.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReserve, _lib as L
from merlot_reserve_b200.synthetic import synthetic_batch, synthetic_clip, synthetic_params, synthetic_spans, tiny_config
from oracle.reserve_oracle import MASK, OracleMerlotReserve

pytestmark = pytest.mark.gpu
COS_MIN = 0.999


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a.float().cpu() if isinstance(a, torch.Tensor) else a)).float().reshape(-1, a.shape[-1])
    b = torch.as_tensor(np.asarray(b.float().cpu() if isinstance(b, torch.Tensor) else b)).float().reshape(-1, b.shape[-1])
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


@pytest.fixture(scope="module")
def setup(cuda):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    model = MerlotReserve.from_config(cfg).bind(params)
    return cfg, params, model, OracleMerlotReserve(cfg, params)


class _Stepwise:
    """Runs the façade with per-launch profiling on: every method then composes the per-op entries call by call."""

    def __enter__(self):
        L.PROFILE = []
        return self

    def __exit__(self, *exc):
        L.PROFILE = None


def test_vision_audio_span_composites(setup, cuda):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=3, seed=9)
    img = torch.from_numpy(clip["images"]).to(cuda)
    aud = torch.from_numpy(clip["audio_clips"]).to(cuda)
    n0 = L.launch_count()
    v = model.vision_encoder(img)
    a = model.audio_encoder(aud)
    spans = torch.from_numpy(synthetic_spans(21, seed=2)).to(cuda)
    s = model._span(spans, normalize=True)
    assert L.launch_count() - n0 > 30                       # the composites launched the kernels themselves
    with _Stepwise():
        v2 = model._vision(img, want_cls=True)
        a2 = model._audio(aud, want_cls=True)
        s2 = L.unit_normalize(model._span(spans))
    assert torch.equal(v["seq_attnpool"], v2["seq_attnpool"]) and torch.equal(v["cls"], v2["cls"])
    assert torch.equal(a["seq_attnpool"], a2["seq_attnpool"]) and torch.equal(a["cls"], a2["cls"])
    assert torch.equal(s, s2)
    ref_v = orc.vision_encoder(clip["images"])
    assert _cos(v["seq_attnpool"], ref_v["seq_attnpool"]).min() >= COS_MIN
    assert _cos(v["seq"], ref_v["seq"]).min() >= COS_MIN
    assert _cos(v["cls"], ref_v["cls"]).min() >= COS_MIN
    ref_a = orc.audio_encoder(clip["audio_clips"])
    assert _cos(a["seq_attnpool"], ref_a["seq_attnpool"]).min() >= COS_MIN
    assert _cos(a["cls"], ref_a["cls"]).min() >= COS_MIN
    assert _cos(s, orc.embed_text_spans_only(spans.cpu().numpy())).min() >= COS_MIN
    # token_encoder + span_encoder on EMBEDDED spans (the reference's module signature, modeling.py:488)
    emb = model.token_encoder({"text_spans": spans})["text_spans"]
    cls = model.span_encoder(emb, spans != 0)
    assert _cos(L.unit_normalize(cls), s).min() >= 0.99999


def test_embed_video_batched_composite_and_graph(setup, cuda):
    cfg, params, model, orc = setup
    batch = synthetic_batch(cfg, 3, nseg=4, seed=31)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in batch.items()}
    n0 = L.launch_count()
    got = model.batch_embed_video(**dev)
    launches = L.launch_count() - n0
    with _Stepwise():
        stepwise = model.batch_embed_video(**dev)
    assert torch.equal(got, stepwise)
    for i in range(3):
        ref = orc.embed_video(*[batch[k][i] for k in ("images", "audio_clips", "tokens", "subseg_idxs")])
        c = _cos(got[i], ref)
        assert c.min() >= COS_MIN, (i, float(c.min()))
    # CUDA graph: one launch replays the whole forward; new inputs flow through the static buffers
    run = model.capture_embed_video(3, nseg=4)
    n0 = L.launch_count()
    out = run(batch["images"], batch["audio_clips"], batch["tokens"], batch["subseg_idxs"])
    torch.cuda.synchronize()
    assert L.launch_count() == n0                           # nothing is launched by the library during a replay
    assert torch.equal(out, got)
    other = synthetic_batch(cfg, 3, nseg=4, seed=77)
    out2 = run(other["images"], other["audio_clips"], other["tokens"], other["subseg_idxs"]).clone()
    assert torch.equal(out2, model.batch_embed_video(**{k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in other.items()}))
    assert launches > 50


def test_joint_composite_mc_scoring_row_map(setup, cuda):
    """mr_joint_encode without the head, candidates reading their example's frame through the vision row map."""
    cfg, params, model, orc = setup
    rng = np.random.default_rng(3)
    tokens = rng.integers(10, 32768, size=(3, 4, 24)).astype(np.int32)
    tokens[:, :, 10:] = 0
    tokens[:, :, 9] = MASK
    images = rng.random((3, 24, 768), dtype=np.float32)
    proj = (rng.standard_normal(cfg["model"]["hidden_size"]) * 0.3).astype(np.float32)
    got = model.score_multiple_choice(images, tokens, proj)
    with _Stepwise():
        stepwise = model.score_multiple_choice(images, tokens, proj)
    assert np.array_equal(got, stepwise)
    ref = orc.score_multiple_choice(images, tokens, proj).numpy()
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max() + 2e-2


def test_composite_argument_checks(setup, cuda):
    """Bad descriptors / missing buffers are refused with a status code, nothing is launched."""
    cfg, params, model, orc = setup
    lib = L.load()
    p = model._p
    d = L.VisionDesc(B=2, grid_h=4, grid_w=6, H=128, in_bf16=0, ln_eps=1e-5, enc=p.cw("vis"), pool=p.cw("vis_pool"),
                     emb_w=p.vis_emb[0].data_ptr(), emb_b=p.vis_emb[1].data_ptr(), rot=model._table("vision").data_ptr())
    need = lib.mr_vision_encode_workspace_bytes(C.byref(d))
    assert need > 0
    img = torch.zeros((2, 24, 768), device=cuda)
    out = torch.empty((2, 6, 128), dtype=torch.bfloat16, device=cuda)
    ins = (C.c_void_p * 1)(img.data_ptr())
    outs = (C.c_void_p * 3)(out.data_ptr(), None, None)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = torch.empty((need,), dtype=torch.uint8, device=cuda)
    n0 = L.launch_count()
    assert lib.mr_vision_encode(stream, C.byref(d), ins, outs, C.c_void_p(ws.data_ptr()), need - 256) == -5    # MR_ERR_WORKSPACE
    d.grid_h = 5
    assert lib.mr_vision_encode(stream, C.byref(d), ins, outs, C.c_void_p(ws.data_ptr()), need) == -1          # MR_ERR_BAD_SHAPE
    d.grid_h = 4
    d.emb_w = None
    assert lib.mr_vision_encode(stream, C.byref(d), ins, outs, C.c_void_p(ws.data_ptr()), need) == -4          # MR_ERR_NULL_POINTER
    assert L.launch_count() == n0
    d.B = 0
    assert lib.mr_vision_encode(stream, C.byref(d), ins, outs, None, 0) == 0                                     # empty batch: no-op
