"""Pins the oracle to the REFERENCE ITSELF.

tests/golden_ref/*.npz were written by tools/run_reference_shim.py: the reference's own
mreserve/modeling.py, pretrain/pretrain_model.py, finetune/vcr/qa_qar_joint_finetune.py (class MerlotReserveVCR) and
finetune/tvqa/tvqa_finetune.py (class MerlotReserveTVQA), imported UNMODIFIED from /root/reference and executed over
numpy stand-ins of jax / flax (oracle/refshim; what that pins vs. what stays restated library semantics is stated in
oracle/refshim/README.md).  Here oracle/reserve_oracle.py -- the checker every -m gpu parity test uses -- must
reproduce every one of those vectors to fp32 rounding; tests/test_reference_goldens_gpu.py compares the CUDA path
with the same fixtures directly.  CPU only, no reference tree needed (the fixtures are committed).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from merlot_reserve_b200.synthetic import (load_config, synthetic_clip, synthetic_params, synthetic_pretrain_batch,
                                           synthetic_spans, synthetic_tvqa_batch, tiny_config)
from oracle import reserve_oracle as O

REF = Path(__file__).resolve().parent / "golden_ref"
ROOT = Path(__file__).resolve().parent.parent


def _close(got, ref, rtol=2e-5, what=""):
    """|got - ref| <= rtol * max|ref| (fp32 rounding through a few layers; the two sides differ only in summation
    order: numpy / torch-CPU kernels vs. torch-CPU kernels with other shapes)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / scale
    assert err <= rtol, f"{what}: max |diff| / max |ref| = {err:.3e}"


@pytest.fixture(scope="module", params=["tiny_init", "tiny_scale4"])
def case(request):
    ref = np.load(REF / f"{request.param}.npz")
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=float(ref["weight_scale"]))
    return cfg, params, O.OracleMerlotReserve(cfg, params), ref


def test_module_functions_match_the_reference():
    """get_rotary_coordinates(_2d), construct_rotary_sinusoids, apply_rotary, apply_attention, my_gelu,
    unit_normalize, one_hot_pool (mreserve/modeling.py:21-200,240-241,541-578) on fixed inputs."""
    h = np.load(REF / "helpers.npz")
    for L in (1, 2, 15, 30, 31):
        assert np.array_equal(O.rotary_coordinates(L).numpy(), h[f"coords/center/{L}"])
        assert np.array_equal(O.rotary_coordinates(L, center_origin=False).numpy(), h[f"coords/plain/{L}"])
    for hw in ("12x20", "18x32", "6x10", "9x16", "4x6", "2x3"):
        hh, ww = map(int, hw.split("x"))
        _close(O.rotary_coordinates_2d(hh, ww).numpy(), h[f"coords2d/{hw}"], 1e-6, hw)
    for n in (1, 2, 4):
        c = torch.from_numpy(h[f"rotary/coords/{n}"])
        s = O.rotary_sinusoids(c)
        _close(s.numpy(), h[f"rotary/sinusoids/{n}"], 2e-6, "sinusoids")
        qk = torch.from_numpy(h[f"rotary/qk/{n}"])
        _close(O.apply_rotary(qk, s).numpy(), h[f"rotary/applied/{n}"], 2e-6, "apply_rotary")
        # the closed form the kernels use (scale table) against the reference's literal output
        tab = O.rotary_scale_table(c)
        closed = qk.clone()
        closed[..., :32] = qk[..., :32] * tab[:, :, None, :]
        _close(closed.numpy(), h[f"rotary/applied/{n}"], 2e-6, "scale table")
    sin = O.rotary_sinusoids(torch.from_numpy(h["attn/coords"]))
    _close(O.attention(torch.from_numpy(h["attn/qkv"]), sin, torch.from_numpy(h["attn/bias"])).numpy(), h["attn/out"], 2e-6, "attention")
    _close(O.my_gelu(torch.from_numpy(h["gelu/x"])).numpy(), h["gelu/y"], 1e-6, "gelu")
    _close(O.unit_normalize(torch.from_numpy(h["unit/x"])).numpy(), h["unit/y"], 1e-6, "unit_normalize")
    v, dp, idx = (torch.from_numpy(h[k]) for k in ("pool/v", "pool/do_pool", "pool/idx"))
    for tag, rb in (("plain", None), ("folded", 2)):
        r = O.one_hot_pool(dp, idx.long(), v, 5, real_bsize=rb)
        _close(r["x"].numpy(), h[f"pool/{tag}/x"], 1e-6, "one_hot_pool")
        assert np.array_equal(r["idx_oh"].numpy(), h[f"pool/{tag}/idx_oh"])


def test_parameter_tree_is_the_reference_tree():
    """Names and shapes of every leaf the reference's modules create (Module.init under the shim) == the synthetic
    generator's tree == what the packer accepts; Base and Large (18x32) included."""
    trees = json.loads((REF / "param_trees.json").read_text())

    def flat(t, p=""):
        out = {}
        for k, v in t.items():
            if isinstance(v, dict):
                out.update(flat(v, p + k + "/"))
            else:
                out[p + k] = list(np.shape(v))
        return out

    for name, cfg in (("tiny", tiny_config()), ("base", load_config("base")), ("large", load_config("large", (18, 32)))):
        if name == "tiny":
            ours = flat(synthetic_params(cfg))
        else:   # shapes only (no 316 M / 853 M parameter draw): the generator's shapes are config-determined
            from merlot_reserve_b200 import synthetic
            ours = synthetic.param_shapes(cfg)
        assert ours == trees[name], (name, sorted(set(ours) ^ set(trees[name]))[:8])


def test_zero_shot_api_matches_the_reference(case):
    """embed_video / batch_embed_video / get_imgseq_only / get_audioseq_only / embed_audio_only /
    embed_text_spans_only / embed_preencoded_* / embed_singleimg_with_multiimg_prompt (modeling.py:767-931)."""
    cfg, params, orc, ref = case
    clip, clip2 = synthetic_clip(cfg, nseg=4, seed=21), synthetic_clip(cfg, nseg=4, seed=22)
    _close(orc.embed_video(**clip).numpy(), ref["embed_video"], what="embed_video")
    _close(orc.get_imgseq_only(clip["images"]).numpy(), ref["imgseq"], what="imgseq")
    _close(orc.get_audioseq_only(clip["audio_clips"]).numpy(), ref["audioseq"], what="audioseq")
    _close(orc.embed_audio_only(clip["audio_clips"]).numpy(), ref["audio_cls"], what="audio_cls")
    _close(orc.embed_text_spans_only(synthetic_spans(12, seed=3)).numpy(), ref["spans"], what="spans")
    stack = {k: np.stack([clip[k], clip2[k]], 0) for k in clip}
    _close(orc.batch_embed_video(**stack).numpy(), ref["batch_embed_video"], what="batch_embed_video")
    text_tokens = clip["tokens"].copy()
    text_tokens[text_tokens == O.AUDIOSPAN] = 77
    _close(orc.embed_preencoded_noaudio(ref["imgseq"], text_tokens, clip["subseg_idxs"]).numpy(), ref["preencoded_noaudio"],
           what="preencoded_noaudio")
    _close(orc.embed_preencoded_audio(ref["imgseq"], ref["audioseq"], clip["tokens"], clip["subseg_idxs"],
                                      clip["subseg_idxs"]).numpy(), ref["preencoded_audio"], what="preencoded_audio")
    _close(orc.embed_singleimg_with_multiimg_prompt(ref["imgseq"][:2], clip["images"][2:], text_tokens,
                                                    clip["subseg_idxs"]).numpy(), ref["multiimg_prompt"], what="multiimg_prompt")


def test_prepare_multimodal_inputs_matches_the_reference(case):
    """x, rotary_coords and the [B, T, T] attention_mask (modeling.py:651-758) with audio splice, vision rows,
    right-padding and video_src_idx packing."""
    cfg, params, orc, ref = case
    T = lambda k: torch.from_numpy(ref[k])
    mm = orc.prepare_multimodal_inputs(tokens=T("mm/tokens").long(), token_segment_idx=T("mm/seg").long(),
                                       vision_input=T("mm/vis"), audio_spans=T("mm/aspans"),
                                       audio_pointers=T("mm/aptr").long(), padding_len=56, video_src_idx=T("mm/vsrc").long())
    _close(mm["x"].numpy(), ref["mm/x"], 1e-6, "x")
    _close(mm["rotary_coords"].numpy(), ref["mm/rotary_coords"], 1e-6, "rotary_coords")
    assert np.array_equal(mm["attention_mask"].numpy().astype(bool), ref["mm/attention_mask"])


def test_joint_encoder_reference_schema_and_learned_pe_branch(case):
    """TransformerEncoder.__call__ (modeling.py:297-376) on (x, rotary_coords, attention_mask), and the learned-`pe`
    branch (:335-341) taken when no rotary_coords are given."""
    cfg, params, orc, ref = case
    x, mask = torch.from_numpy(ref["mm/x"]), torch.from_numpy(ref["mm/attention_mask"])
    got = orc.joint_transformer(x, torch.from_numpy(ref["mm/rotary_coords"]), mask)["seq"]
    _close(got.numpy(), ref["jt/rotary"], what="joint encoder")
    p2 = dict(params, joint_transformer=dict(params["joint_transformer"], pe=ref["jt/pe"]))
    got = O.OracleMerlotReserve(cfg, p2).joint_transformer(x, None, mask)["seq"]
    _close(got.numpy(), ref["jt/pe_out"], what="learned pe")


def test_multiple_choice_heads_match_the_reference(case):
    """MerlotReserveVCR.__call__ (qa_qar_joint_finetune.py:150-170) and MerlotReserveTVQA.__call__
    (tvqa_finetune.py:163-258)."""
    cfg, params, orc, ref = case
    rng = np.random.default_rng(17)
    mc_tokens = rng.integers(10, 32768, size=(3, 4, 24)).astype(np.int32)
    mc_tokens[:, :, -1] = O.MASK
    mc_images = rng.random((3, 24, 768), dtype=np.float32)
    assert np.array_equal(mc_tokens, ref["mc_tokens"])
    logits = orc.score_multiple_choice(mc_images, mc_tokens, ref["mc_proj"]).numpy()
    _close(logits, ref["mc_logits"], what="vcr logits")
    assert np.array_equal(logits.argmax(-1), ref["mc_logits"].argmax(-1))
    tb = synthetic_tvqa_batch(cfg, 2, num_segments=3, num_answers=3, seq_len=96, seed=4)
    la, lt = orc.score_tvqa(tb, ref["mc_proj"])
    _close(la.numpy(), ref["tvqa/logits_from_audio"], what="tvqa audio")
    _close(lt.numpy(), ref["tvqa/logits_from_text"], what="tvqa text")


def test_pretraining_forward_and_loss_match_the_reference(case):
    """MerlotReservePretrainer.__call__ and loss_fn_given_preds (pretrain_model.py:38-303): every x / y / y_extra row,
    the Gumbel-top-k selection order (_sources), the loss and its per-head / per-source breakdown; then two replicas
    (jax.pmap + lax.all_gather in the reference) against the oracle's rank-major gather."""
    cfg, params, orc, ref = case
    b0, b1 = synthetic_pretrain_batch(cfg, 2, seed=3), synthetic_pretrain_batch(cfg, 2, seed=4)
    preds = orc.pretrain_forward(b0)
    for head, d in preds.items():
        for k, v in d.items():
            if k == "_sources":
                assert np.array_equal(v.numpy(), ref[f"pretrain/{head}/{k}"])
            else:
                _close(v.numpy(), ref[f"pretrain/{head}/{k}"], what=f"{head}/{k}")
    loss, info = O.contrastive_loss(preds)
    assert abs(float(loss) - float(ref["pretrain/loss"])) < 2e-5
    for k, v in info.items():
        assert abs(float(v) - float(ref[f"pretrain/loss_info/{k}"])) < 2e-5, k
    # world 2: every replica's loss with the other replica's targets as extra negatives
    all_preds = [preds, orc.pretrain_forward(b1)]

    def gather_for(rank):
        def gather(t):
            for head in all_preds[rank]:
                for name in ("x", "y"):
                    mine = all_preds[rank][head][name]
                    extra = all_preds[rank][head].get(f"{name}_extra")
                    cat = lambda p: torch.cat([p[head][name]] + ([p[head][f"{name}_extra"]] if extra is not None else []), 0)
                    if t.shape == cat(all_preds[rank]).shape and torch.equal(t, cat(all_preds[rank])):
                        return torch.cat([cat(p) for p in all_preds], 0)
            raise AssertionError("unknown gather operand")
        return gather

    for rank in range(2):
        loss_r, info_r = O.contrastive_loss(all_preds[rank], gather=gather_for(rank))
        assert abs(float(loss_r) - float(ref["pretrain_world2/loss"][rank])) < 2e-5
        for k, v in info_r.items():
            assert abs(float(v) - float(ref[f"pretrain_world2/loss_info/{k}"][rank])) < 2e-5, k


def test_bf16_module_dtype_noise_floor():
    """The reference run with use_bfloat16=True (every module output a bf16 array, modeling.py:594) against its own
    fp32 run, and the oracle's bf16_graph() emulation against both: the tolerances of the GPU parity tests
    (cosine >= 0.999 on embeddings) sit above the reference's own bf16 noise."""
    ref16, ref32 = np.load(REF / "tiny_bf16.npz"), np.load(REF / "tiny_scale4.npz")
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    clip = synthetic_clip(cfg, nseg=4, seed=21)
    cos = lambda a, b: ((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))).min()
    floor = cos(ref16["embed_video"], ref32["embed_video"])
    assert 0.99 < floor < 0.99999, floor                     # bf16 is visibly noisier than fp32, and still close
    with O.bf16_graph():
        emu = O.OracleMerlotReserve(cfg, params).embed_video(**clip).numpy()
    assert cos(emu, ref16["embed_video"]) > 0.999            # the emulation tracks the reference's bf16 graph
    assert cos(emu, ref32["embed_video"]) > 0.99


def test_base_one_clip_embed_video_matches_the_reference():
    """BASELINE.json configs[0]: Reserve-Base random-init embed_video on one synthetic 16-frame 192x320 clip, fp32
    (about 4 s of torch-CPU)."""
    ref = np.load(REF / "base_cfg1.npz")
    cfg = load_config("base")
    got = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0)).embed_video(**synthetic_clip(cfg, nseg=16, seed=1)).numpy()
    _close(got, ref["embed_video"], 5e-5, "Base embed_video")


@pytest.mark.skipif(not Path("/root/reference/mreserve/modeling.py").exists(), reason="reference tree not present")
def test_fixtures_regenerate_from_the_reference_tree():
    """Where the reference tree is present (the build container): re-execute the reference under the shim and check
    that the committed fixture is what it produces."""
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; import tools.run_reference_shim as t; "
            "mods = t.reference_modules(); out = t.tiny_case(mods, 4.0); ref = np.load(%r); "
            "bad = [k for k in ref.files if not np.allclose(out[k], ref[k], rtol=0, atol=1e-6 * max(1.0, float(np.abs(ref[k]).max())))]; "
            "print('BAD', bad); sys.exit(1 if bad else 0)") % (str(ROOT), str(REF / "tiny_scale4.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT),
                       env={**os.environ, "PYTHONPATH": ""})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
