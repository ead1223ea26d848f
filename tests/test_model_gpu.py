"""End-to-end parity of the reference-facing API (merlot_reserve_b200.modeling) against the CPU oracle on
identical random-init weights and synthetic inputs.  Tolerances (BASELINE.json north_star): embedding
cosine >= 0.999 per row; identical argmax on multiple-choice scoring wherever the oracle's top-2 margin
exceeds the stated logit tolerance."""
import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReserve, PretrainedMerlotReserve
from merlot_reserve_b200.synthetic import (synthetic_batch, synthetic_clip, synthetic_params, synthetic_spans,
                                           tiny_config, load_config)
from oracle.reserve_oracle import OracleMerlotReserve, MASK

pytestmark = pytest.mark.gpu
COS_MIN = 0.999


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a)).float().reshape(-1, np.asarray(a).shape[-1])
    b = torch.as_tensor(np.asarray(b)).float().reshape(-1, np.asarray(b).shape[-1])
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


@pytest.fixture(scope="module", params=[1.0, 4.0], ids=["init_scale", "scale4"])
def setup(request, cuda):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=request.param)
    model = MerlotReserve.from_config(cfg)
    return cfg, params, model, OracleMerlotReserve(cfg, params)


def test_vision_audio_span_encoders(setup):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=3)
    got = model.apply({"params": params}, clip["images"], method=model.get_imgseq_only)
    assert _cos(got, orc.get_imgseq_only(clip["images"])).min() >= COS_MIN
    got = model.apply({"params": params}, clip["audio_clips"], method=model.get_audioseq_only)
    assert _cos(got, orc.get_audioseq_only(clip["audio_clips"])).min() >= COS_MIN
    got = model.apply({"params": params}, clip["audio_clips"], method=model.embed_audio_only)
    assert _cos(got, orc.embed_audio_only(clip["audio_clips"])).min() >= COS_MIN
    spans = synthetic_spans(21)
    got = model.apply({"params": params}, spans, method=model.embed_text_spans_only)
    ref = orc.embed_text_spans_only(spans)
    assert _cos(got, ref).min() >= COS_MIN
    assert np.allclose(np.linalg.norm(got, axis=-1), 1.0, atol=1e-3)


@pytest.mark.parametrize("nseg", [2, 8])
def test_embed_video(setup, nseg):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=nseg, seed=5)
    got = model.apply({"params": params}, **clip, method=model.embed_video)
    ref = orc.embed_video(**clip)
    assert got.shape == tuple(ref.shape) == (160, cfg["model"]["hidden_size"])
    cos = _cos(got, ref)
    # every row, including the padding rows whose queries are fully masked (uniform attention)
    assert cos.min() >= COS_MIN, f"min cosine {cos.min():.5f} at row {int(cos.argmin())}"


def test_batch_embed_video_matches_single(setup):
    cfg, params, model, orc = setup
    batch = synthetic_batch(cfg, 3, nseg=4, seed=11)
    got = model.apply({"params": params}, **batch, method=model.batch_embed_video)
    for i in range(3):
        one = model.apply({"params": params}, *[batch[k][i] for k in ("images", "audio_clips", "tokens", "subseg_idxs")],
                          method=model.embed_video)
        assert _cos(got[i], one).min() >= 0.99999
        ref = orc.embed_video(*[batch[k][i] for k in ("images", "audio_clips", "tokens", "subseg_idxs")])
        assert _cos(got[i], ref).min() >= COS_MIN


def test_preencoded_and_prompt_paths(setup):
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=4, seed=3)
    enc_ref = orc.get_imgseq_only(clip["images"])
    aud_ref = orc.get_audioseq_only(clip["audio_clips"])
    a = model.apply({"params": params}, enc_ref.numpy(), clip["tokens"], clip["subseg_idxs"],
                    method=model.embed_preencoded_noaudio)
    assert _cos(a, orc.embed_preencoded_noaudio(enc_ref, clip["tokens"], clip["subseg_idxs"])).min() >= COS_MIN
    b = model.apply({"params": params}, enc_ref.numpy(), aud_ref.numpy(), clip["tokens"], clip["subseg_idxs"],
                    clip["subseg_idxs"], method=model.embed_preencoded_audio)
    assert _cos(b, orc.embed_preencoded_audio(enc_ref, aud_ref, clip["tokens"], clip["subseg_idxs"],
                                              clip["subseg_idxs"])).min() >= COS_MIN
    c = model.apply({"params": params}, enc_ref[:2].numpy(), clip["images"][2:], clip["tokens"], clip["subseg_idxs"],
                    method=model.embed_singleimg_with_multiimg_prompt)
    assert _cos(c, orc.embed_singleimg_with_multiimg_prompt(enc_ref[:2], clip["images"][2:], clip["tokens"],
                                                            clip["subseg_idxs"])).min() >= COS_MIN


def test_multiple_text_and_mc_scoring(setup):
    cfg, params, model, orc = setup
    rng = np.random.default_rng(0)
    clip = synthetic_clip(cfg, nseg=1, seed=2)
    n_cand, Ltok = 4, 32
    toks = rng.integers(10, 32768, size=(n_cand, Ltok)).astype(np.int32)
    toks[:, -1] = MASK
    toks[1, 20:] = 0
    toks[1, 19] = MASK
    sub = np.zeros_like(toks)
    got = model.apply({"params": params}, clip["images"], toks, sub, method=model.embed_singleimg_with_multiple_text)
    ref = orc.embed_singleimg_with_multiple_text(clip["images"], toks, sub)
    assert _cos(got, ref).min() >= COS_MIN
    # VCR-style scoring: logits within tolerance, same argmax where the margin is larger than the tolerance
    B = 6
    imgs = rng.random((B, clip["images"].shape[1], 768), dtype=np.float32)
    tokens = rng.integers(10, 32768, size=(B, n_cand, Ltok)).astype(np.int32)
    tokens[:, :, -1] = MASK
    proj = (rng.standard_normal(cfg["model"]["hidden_size"]) * 0.3).astype(np.float32)
    got = model.apply({"params": params}, imgs, tokens, proj, method=model.score_multiple_choice)
    ref = orc.score_multiple_choice(imgs, tokens, proj).numpy()
    tol = 0.05 + 0.02 * np.abs(ref)
    assert (np.abs(got - ref) <= tol).all(), np.abs(got - ref).max()
    srt = np.sort(ref, -1)
    confident = (srt[:, -1] - srt[:, -2]) > 2 * tol.max()
    assert (got.argmax(-1)[confident] == ref.argmax(-1)[confident]).all()


def test_zero_shot_scoring_argmax(setup):
    """demo/demo_video.py:33-47: MASK rows of embed_video against the span label space, x100."""
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=4, seed=9)
    spans = synthetic_spans(40, seed=1)
    pm = PretrainedMerlotReserve.from_params(cfg, params)
    h = pm.embed_video(**clip)[clip["tokens"] == MASK]
    logits = 100.0 * h @ pm.embed_text_spans_only(spans).T
    h_ref = orc.embed_video(**clip)[torch.from_numpy(clip["tokens"] == MASK)]
    ref = (100.0 * h_ref @ orc.embed_text_spans_only(spans).T).numpy()
    assert np.abs(logits - ref).max() <= 0.5          # logits live in [-100, 100]
    srt = np.sort(ref, -1)
    if (srt[:, -1] - srt[:, -2] > 1.0).all():
        assert (logits.argmax(-1) == ref.argmax(-1)).all()


def test_input_validation(setup):
    cfg, params, model, orc = setup
    model.bind(params)
    clip = synthetic_clip(cfg, nseg=2)
    with pytest.raises(ValueError):
        model.embed_video(clip["images"][..., :700], clip["audio_clips"], clip["tokens"], clip["subseg_idxs"])
    with pytest.raises(ValueError):
        model.embed_video(clip["images"], clip["audio_clips"][:5], clip["tokens"], clip["subseg_idxs"])
    with pytest.raises(ValueError):
        model.embed_video(clip["images"], clip["audio_clips"], clip["tokens"], clip["subseg_idxs"][:10])


def test_layernorm_fold_matches_separate_layernorm(setup):
    """The default path folds every pre-attention / pre-MLP LayerNorm into the consuming GEMM; the unfused path
    (separate LayerNorm kernel) must give the same embeddings, and both must match the oracle."""
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=4, seed=13)
    fused = model.apply({"params": params}, **clip, method=model.embed_video)
    plain = MerlotReserve.from_config(cfg)
    plain.fuse_layernorm = False
    unfused = plain.apply({"params": params}, **clip, method=plain.embed_video)
    assert _cos(fused, unfused).min() >= 0.999
    ref = orc.embed_video(**clip)
    assert _cos(fused, ref).min() >= COS_MIN and _cos(unfused, ref).min() >= COS_MIN


def test_pinned_host_inputs_take_the_overlapped_path_and_match(setup):
    """bench.py's e2e leg: pinned host tensors -> chunked H2D on a side stream overlapped with compute; results
    must equal the resident-input path bit for bit (same kernels, same chunk-independent arithmetic)."""
    cfg, params, model, orc = setup
    model.bind(params)
    batch = synthetic_batch(cfg, 40, nseg=4, seed=17)          # 160 frames -> 2 H2D chunks
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in batch.items()}
    model._H2D_CHUNKS = 2
    a = model.batch_embed_video(**pinned)
    b = model.batch_embed_video(**{k: torch.from_numpy(v).cuda() for k, v in batch.items()})
    model._H2D_CHUNKS = None
    a2 = model.batch_embed_video(**pinned)                     # staging buffer reuse, automatic chunk count
    torch.cuda.synchronize()
    assert a.is_cuda and torch.equal(a, b) and torch.equal(a2, b)


def test_composite_encoder_call_equals_step_by_step(setup, lib):
    """mr_transformer_encode (one C call per encoder) enqueues the same kernels as the Python-composed layer loop
    that bench.py uses for per-launch timing: results must be bit-identical."""
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=3, seed=19)
    a = model.apply({"params": params}, **clip, method=model.embed_video)
    lib.PROFILE = []
    try:
        b = model.apply({"params": params}, **clip, method=model.embed_video)
        names = {n for n, *_ in lib.PROFILE}
    finally:
        lib.PROFILE = None
    assert "mr_transformer_encode" not in names and "mr_gemm_epilogue" in names
    assert np.array_equal(a, b)


def test_tvqa_forward_matches_oracle(setup):
    """MerlotReserveTVQA.__call__ (finetune/tvqa/tvqa_finetune.py:163-258): 2 examples x 3 frames x 4 candidates, audio and
    text-only variants in one joint batch, with the reference's zero "frame 0" trick.  Logits within the bf16
    tolerance of the fp32 oracle and the same argmax wherever the oracle's top-2 margin exceeds it."""
    from merlot_reserve_b200.synthetic import synthetic_tvqa_batch
    cfg, params, model, orc = setup
    H = cfg["model"]["hidden_size"]
    batch = synthetic_tvqa_batch(cfg, 2, num_segments=3, num_answers=4, seq_len=96, seed=4)
    proj = (np.random.default_rng(5).standard_normal((H, 1)) * 0.02).astype(np.float32)
    got_a, got_t = model.apply({"params": params}, batch, proj, method=model.score_tvqa)
    ref_a, ref_t = orc.score_tvqa(batch, proj)
    for got, ref, name in ((got_a, ref_a, "audio"), (got_t, ref_t, "text")):
        ref = ref.numpy()
        assert got.shape == ref.shape == (2, 4)
        tol = 2e-2 * np.abs(ref).max() + 2e-3
        assert np.abs(got - ref).max() <= tol, (name, got, ref)
        top2 = np.sort(ref, -1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 2 * tol
        assert (got.argmax(-1) == ref.argmax(-1))[clear].all()
    assert np.abs(got_a - got_t).max() > 1e-4      # the two variants really are different sequences


def test_product_against_the_reference_bf16_noise_floor(setup):
    """Distance of the B200 path from the fp32 oracle next to the distance of the reference's OWN bf16 graph
    (oracle bf16_graph mode, rounding at every op as the dtype=bfloat16 flax modules do) from the same fp32 result:
    the fused epilogues round once per GEMM, so the product must not be further away than that floor (factor 2 of
    slack on the mean, and the worst row no worse than the floor's worst row by more than 1e-3)."""
    from oracle.reserve_oracle import bf16_graph
    cfg, params, model, orc = setup
    clip = synthetic_clip(cfg, nseg=4, seed=5)
    ref = orc.embed_video(**clip)
    with bf16_graph():
        floor = orc.embed_video(**clip)
    got = model.apply({"params": params}, **clip, method=model.embed_video)
    d_floor, d_got = 1.0 - _cos(floor, ref), 1.0 - _cos(got, ref)
    assert d_got.mean() <= 2.0 * d_floor.mean() + 1e-6, (float(d_got.mean()), float(d_floor.mean()))
    assert d_got.max() <= d_floor.max() + 1e-3, (float(d_got.max()), float(d_floor.max()))
    print(f"1 - cosine vs fp32: product mean {float(d_got.mean()):.2e} max {float(d_got.max()):.2e}; "
          f"reference bf16 graph mean {float(d_floor.mean()):.2e} max {float(d_floor.max()):.2e}")


def test_device_patchify_feeds_the_vision_encoder(setup):
    """uint8 frames -> `patchify` (device) -> get_imgs_enc equals the reference route: fp32 patches made on the host
    by the restatement of preprocess_image_to_patches -> get_imgs_enc.  Bit for bit (same bf16 operand)."""
    from oracle.reserve_oracle import frames_to_patches
    cfg, params, model, orc = setup
    model.bind(params)
    gh, gw = cfg["model"]["output_grid"]
    frames = np.random.default_rng(8).integers(0, 256, size=(3, gh * 16, gw * 16, 3), dtype=np.uint8)
    on_device = model.get_imgs_enc(model.patchify(frames))
    on_host = model.get_imgs_enc(frames_to_patches(frames, gh, gw))
    assert tuple(on_device.shape) == (3, gh * gw // 4, cfg["model"]["hidden_size"])
    assert np.array_equal(np.asarray(on_device.float().cpu()), np.asarray(torch.as_tensor(np.asarray(on_host)).float()))
    assert _cos(on_device.float().cpu(), orc.get_imgseq_only(frames_to_patches(frames, gh, gw))).min() >= COS_MIN
