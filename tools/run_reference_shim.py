#!/usr/bin/env python
"""Executes the REFERENCE's own model files (imported unmodified from /root/reference) over the numpy stand-ins of
jax / flax in oracle/refshim, on this repo's seeded synthetic weights and inputs, and writes the outputs as golden
vectors under tests/golden_ref/.  Runs in the BUILD container only (it reads /root/reference; the GPU box has no
reference tree): the fixtures are committed and every test reads only them.

    python tools/run_reference_shim.py                 # tiny + helper + Base-cfg1 fixtures (about a minute)
    python tools/run_reference_shim.py --fullsize      # also Base b=8 pretrain, Large 18x32 nseg=16, Large MC L=144
    python tools/run_reference_shim.py --check-oracle  # oracle/ vs the fixtures just written -> REPORT.md

What executing the reference this way pins, and what stays restated library semantics, is in
oracle/refshim/README.md.  With a real jax + flax install `tools/run_reference_jax.py` writes the same keys.

Randomness: the reference draws `split_from_here` (pretrain_model.py:28, jax.random.choice) and the Gumbel uniform
(:222, jax.random.uniform) from threefry keys; both are INPUTS of the synthetic batch ('_rng/*'), injected here
through `jax.random.override` so that the reference code itself runs unmodified.
"""
import argparse
import ast
import contextlib
import io
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
OUT = ROOT / "tests" / "golden_ref"


def _setup_paths():
    for p in (str(ROOT), str(REFERENCE), str(ROOT / "oracle" / "refshim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


@contextlib.contextmanager
def quiet():
    """The reference prints shapes from inside every layer."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


def _load_class_from_script(path: Path, class_name: str, namespace: dict):
    """finetune/*.py are scripts (argparse + training loop at module level): take ONLY the class definition, verbatim
    from the reference file, and execute it in `namespace` (= the names the script had imported at that point)."""
    src = path.read_text()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    code = compile(ast.Module(body=[node], type_ignores=[]), str(path), "exec")
    exec(code, namespace)
    return namespace[class_name], (node.lineno, node.end_lineno)


def _flat(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flat(v, f"{prefix}{k}/"))
        else:
            out[prefix + k] = [int(s) for s in np.shape(v)]
    return out


def reference_modules():
    _setup_paths()
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    with quiet():
        import jax                                   # oracle/refshim/jax
        import flax.linen as nn                      # oracle/refshim/flax
        assert "refshim" in jax.__version__
        import mreserve.modeling as ref_modeling     # /root/reference/mreserve/modeling.py
        import pretrain.pretrain_model as ref_pm     # /root/reference/pretrain/pretrain_model.py
    assert Path(ref_modeling.__file__).resolve() == REFERENCE / "mreserve" / "modeling.py"
    assert Path(ref_pm.__file__).resolve() == REFERENCE / "pretrain" / "pretrain_model.py"
    ns = dict(vars(ref_modeling))
    ns.update(nn=nn, jax=jax, jnp=jax.numpy, np=np)
    VCR, vcr_lines = _load_class_from_script(REFERENCE / "finetune/vcr/qa_qar_joint_finetune.py", "MerlotReserveVCR", dict(ns))
    TVQA, tvqa_lines = _load_class_from_script(REFERENCE / "finetune/tvqa/tvqa_finetune.py", "MerlotReserveTVQA", dict(ns))
    return jax, ref_modeling, ref_pm, VCR, TVQA


def run_pretrain(jax, ref_pm, cfg, params, batches):
    """MerlotReservePretrainer.__call__ per simulated replica, then loss_fn_given_preds under a pmap of len(batches)
    replicas (the shim's all_gather is rank-major like jax's)."""
    pre = ref_pm.MerlotReservePretrainer.from_config(cfg)
    all_preds = []
    for batch in batches:
        # the reference builds mm_inputs['audio2text'] first (uses prngs[0]) and 'text2audio' third (prngs[1])
        splits = [np.asarray(batch["_rng/split_audio2text"]), np.asarray(batch["_rng/split_text2audio"])]
        jb = {k: jax.numpy.asarray(np.array(v)) for k, v in batch.items() if not k.startswith("_rng/")}
        with quiet(), jax.random.override(
                choice=lambda key, a, shape, replace, p: splits.pop(0) - 1,       # split_from_here = 1 + choice(...)
                uniform=lambda key, shape, dtype, lo, hi: np.asarray(batch["_rng/gumbel_u"], np.float32).reshape(shape)):
            preds = pre.apply({"params": params}, jb)
        assert not splits
        all_preds.append(preds)
    stacked = jax.tree_util.tree_multimap(lambda *xs: np.stack(xs, 0), *all_preds)
    with quiet():
        loss, info = jax.pmap(lambda p: ref_pm.loss_fn_given_preds(dict(p)), axis_name="batch")(stacked)
    return all_preds, np.asarray(loss), {k: np.asarray(v) for k, v in info.items()}


def tiny_case(mods, weight_scale, use_bf16=False):
    jax, ref_modeling, ref_pm, VCR, TVQA = mods
    from merlot_reserve_b200.synthetic import (synthetic_clip, synthetic_params, synthetic_pretrain_batch,
                                               synthetic_spans, synthetic_tvqa_batch, tiny_config)
    MASK = ref_modeling.MASK
    cfg = tiny_config()
    cfg["model"]["use_bfloat16"] = bool(use_bf16)
    params = synthetic_params(cfg, seed=0, weight_scale=weight_scale)
    model = ref_modeling.MerlotReserve.from_config(cfg)
    clip = synthetic_clip(cfg, nseg=4, seed=21)
    clip2 = synthetic_clip(cfg, nseg=4, seed=22)
    spans = synthetic_spans(12, seed=3)

    J = jax.numpy.asarray

    def ap(method, *a, **k):
        with quiet():
            return model.apply({"params": params}, *[J(x) for x in a], method=method, **{n: J(v) for n, v in k.items()})

    f32 = lambda x: np.asarray(x).astype(np.float32)
    out = {"weight_scale": np.float32(weight_scale)}
    out["embed_video"] = f32(ap(model.embed_video, clip["images"], clip["audio_clips"], clip["tokens"], clip["subseg_idxs"]))
    out["imgseq"] = f32(ap(model.get_imgseq_only, clip["images"]))
    out["audioseq"] = f32(ap(model.get_audioseq_only, clip["audio_clips"]))
    out["audio_cls"] = f32(ap(model.embed_audio_only, clip["audio_clips"]))
    out["spans"] = f32(ap(model.embed_text_spans_only, spans))
    out["span_tokens"] = spans
    stack = {k: np.stack([clip[k], clip2[k]], 0) for k in clip}
    out["batch_embed_video"] = f32(ap(model.batch_embed_video, stack["images"], stack["audio_clips"], stack["tokens"],
                                      stack["subseg_idxs"]))                                  # jax.vmap, modeling.py:845
    # partial pipelines (modeling.py:848-931)
    imgs_enc, audio_enc = out["imgseq"], out["audioseq"]
    text_tokens = clip["tokens"].copy()
    text_tokens[text_tokens == ref_modeling.AUDIOSPAN] = 77                                    # a text-only sequence
    out["preencoded_noaudio"] = f32(ap(model.embed_preencoded_noaudio, imgs_enc, text_tokens, clip["subseg_idxs"]))
    out["preencoded_audio"] = f32(ap(model.embed_preencoded_audio, imgs_enc, audio_enc, clip["tokens"],
                                     clip["subseg_idxs"], clip["subseg_idxs"]))
    out["multiimg_prompt"] = f32(ap(model.embed_singleimg_with_multiimg_prompt, imgs_enc[:2], clip["images"][2:],
                                    text_tokens, clip["subseg_idxs"]))
    # prepare_multimodal_inputs itself (modeling.py:651-758): x / rotary_coords / attention_mask, with video_src_idx
    # packing and right-padding
    B, L = 2, 24
    rng = np.random.default_rng(5)
    toks = rng.integers(10, 32768, size=(B, L)).astype(np.int32)
    toks[:, 5:11] = ref_modeling.AUDIOSPAN
    toks[:, 20:] = 0
    seg = np.minimum(np.arange(L)[None] // 6, 3).repeat(B, 0).astype(np.int32)
    aptr = rng.integers(0, 12, size=(B, L)).astype(np.int32)
    vis = rng.standard_normal((B, 4 * 6, cfg["model"]["hidden_size"])).astype(np.float32)      # 4 segments x (2x3) pooled
    aspans = rng.standard_normal((B, 12, 6, cfg["model"]["hidden_size"])).astype(np.float32)
    vsrc = np.array([[0, 0, 1, 1], [2, 2, 2, 3]], np.int32)
    mm = ap(model.prepare_multimodal_inputs, tokens=toks, token_segment_idx=seg, vision_input=vis, audio_spans=aspans,
            audio_pointers=aptr, padding_len=56, video_src_idx=vsrc)
    out.update({"mm/tokens": toks, "mm/seg": seg, "mm/aptr": aptr, "mm/vis": vis, "mm/aspans": aspans, "mm/vsrc": vsrc,
                "mm/x": f32(mm["x"]), "mm/rotary_coords": f32(mm["rotary_coords"]),
                "mm/attention_mask": np.asarray(mm["attention_mask"]).astype(np.bool_)})
    # the joint encoder called with the reference's own schema (modeling.py:297-376): rotary + mask, and the
    # learned-`pe` branch (:335-341) with no rotary_coords (needs a joint_transformer/pe leaf, N(0, 0.02) as :338)
    jt = lambda self, x, **kw: self.joint_transformer(x, **kw)["seq"]
    xin = np.asarray(mm["x"], np.float32)
    out["jt/rotary"] = f32(ap(jt, xin, rotary_coords=np.asarray(mm["rotary_coords"]), attention_mask=np.asarray(mm["attention_mask"])))
    pe = (np.random.default_rng(23).standard_normal((56, cfg["model"]["hidden_size"])) * 0.02).astype(np.float32)
    with quiet():
        out["jt/pe_out"] = f32(model.apply({"params": dict(params, joint_transformer=dict(params["joint_transformer"], pe=pe))},
                                           J(xin), attention_mask=J(np.asarray(mm["attention_mask"])), method=jt))
    out["jt/pe"] = pe
    # VCR head: MerlotReserveVCR.__call__ (finetune/vcr/qa_qar_joint_finetune.py:144-170)
    rng = np.random.default_rng(17)
    mc_tokens = rng.integers(10, 32768, size=(3, 4, 24)).astype(np.int32)
    mc_tokens[:, :, -1] = MASK
    mc_images = rng.random((3, clip["images"].shape[1], 768), dtype=np.float32)
    proj = (rng.standard_normal(cfg["model"]["hidden_size"]) * 0.3).astype(np.float32)
    vcr = VCR.from_config(cfg)
    pv = dict(params, proj={"kernel": proj.reshape(-1, 1)})
    with quiet():
        # answers [B, 2, num_ans_per, L]: the reference scores Q->A and QA->R rows together; 2 x 2 candidates here
        logits = vcr.apply({"params": pv}, {"image": J(mc_images), "answers": J(mc_tokens.reshape(3, 2, 2, 24))})
    out.update({"mc_tokens": mc_tokens, "mc_proj": proj, "mc_logits": f32(logits).reshape(3, 4)})
    # TVQA head: MerlotReserveTVQA.__call__ (finetune/tvqa/tvqa_finetune.py:157-258)
    tb = synthetic_tvqa_batch(cfg, 2, num_segments=3, num_answers=3, seq_len=96, seed=4)
    tvqa = TVQA.from_config(cfg)
    with quiet():
        la, lt = tvqa.apply({"params": pv}, {k: J(np.array(v)) for k, v in tb.items()})
    out.update({"tvqa/logits_from_audio": f32(la), "tvqa/logits_from_text": f32(lt)})
    # pretraining forward + loss on 1 replica (b=2) and on 2 simulated replicas (pretrain_model.py:38-303)
    b0 = synthetic_pretrain_batch(cfg, 2, seed=3)
    preds, loss, info = run_pretrain(jax, ref_pm, cfg, params, [b0])
    for head, d in preds[0].items():
        for k, v in d.items():
            out[f"pretrain/{head}/{k}"] = np.asarray(v).astype(np.int32 if k == "_sources" else np.float32)
    out["pretrain/loss"] = np.float32(loss[0])
    for k, v in info.items():
        out[f"pretrain/loss_info/{k}"] = np.float32(v[0])
    b1 = synthetic_pretrain_batch(cfg, 2, seed=4)
    _, loss2, info2 = run_pretrain(jax, ref_pm, cfg, params, [b0, b1])
    out["pretrain_world2/loss"] = loss2.astype(np.float32)                       # per replica, [2]
    for k, v in info2.items():
        out[f"pretrain_world2/loss_info/{k}"] = v.astype(np.float32)
    return out


def helper_case(mods):
    """The module-level functions of mreserve/modeling.py on fixed inputs."""
    jax, m, *_ = mods
    rng = np.random.default_rng(11)
    J = jax.numpy.asarray
    out = {}
    for L in (1, 2, 15, 30, 31):
        out[f"coords/center/{L}"] = np.asarray(m.get_rotary_coordinates(L), np.float32)
        out[f"coords/plain/{L}"] = np.asarray(m.get_rotary_coordinates(L, center_origin=False), np.float32)
    for h, w in ((12, 20), (18, 32), (6, 10), (9, 16), (4, 6), (2, 3)):
        out[f"coords2d/{h}x{w}"] = np.asarray(m.get_rotary_coordinates_2d(h, w), np.float32)
    for n in (1, 2, 4):
        c = rng.uniform(-1, 1, size=(3, 7, n)).astype(np.float32)
        s = np.asarray(m.construct_rotary_sinusoids(J(c)), np.float32)
        qk = rng.standard_normal((3, 7, 6, 64)).astype(np.float32)
        out[f"rotary/coords/{n}"], out[f"rotary/sinusoids/{n}"] = c, s
        out[f"rotary/qk/{n}"], out[f"rotary/applied/{n}"] = qk, np.asarray(m.apply_rotary(J(qk), J(s)), np.float32)
    qkv = rng.standard_normal((2, 9, 6, 64)).astype(np.float32)                 # 2 heads: q, k, v
    c = rng.uniform(-1, 1, size=(9, 1)).astype(np.float32)
    valid = rng.random((2, 9)) > 0.3
    bias = np.where(valid[:, None, None, :] & valid[:, None, :, None], 0.0, -1e10).astype(np.float32)
    out["attn/qkv"], out["attn/coords"], out["attn/bias"] = qkv, c, bias
    out["attn/out"] = np.asarray(m.apply_attention(J(qkv), m.construct_rotary_sinusoids(J(c)), J(bias)), np.float32)
    x = rng.standard_normal((4, 10)).astype(np.float32) * 3
    out["gelu/x"], out["gelu/y"] = x, np.asarray(m.my_gelu(J(x)), np.float32)
    out["unit/x"], out["unit/y"] = x, np.asarray(m.unit_normalize(J(x)), np.float32)
    v = rng.standard_normal((4, 6, 8)).astype(np.float32)
    do_pool = rng.random((4, 6)) > 0.4
    idx = rng.integers(0, 5, size=(4, 6)).astype(np.int32)
    for tag, rb in (("plain", None), ("folded", 2)):
        r = m.one_hot_pool(J(do_pool), J(idx), J(v), 5, real_bsize=rb)
        out[f"pool/{tag}/x"], out[f"pool/{tag}/idx_oh"] = np.asarray(r["x"], np.float32), np.asarray(r["idx_oh"], np.float32)
    out["pool/v"], out["pool/do_pool"], out["pool/idx"] = v, do_pool, idx
    return out


def param_trees(mods):
    """Parameter-tree names and shapes as the reference's modules create them (Module.init)."""
    jax, m, ref_pm, VCR, TVQA = mods
    from merlot_reserve_b200.synthetic import load_config, synthetic_clip, synthetic_spans, tiny_config
    trees = {}
    for name, cfg in (("tiny", tiny_config()), ("base", load_config("base")), ("large", load_config("large", (18, 32)))):
        cfg["model"]["use_bfloat16"] = False
        model = m.MerlotReserve.from_config(cfg)
        clip = synthetic_clip(cfg, nseg=1, seed=1)
        with quiet():
            a = model.init(jax.random.PRNGKey(0), *[jax.numpy.asarray(clip[k]) for k in ("images", "audio_clips", "tokens", "subseg_idxs")],
                           method=model.embed_video)["params"]
            b = model.init(jax.random.PRNGKey(0), jax.numpy.asarray(synthetic_spans(2, seed=1)), method=model.embed_text_spans_only)["params"]
        flat = _flat(a)
        flat.update(_flat(b))
        trees[name] = dict(sorted(flat.items()))
    return trees


def base_cfg1(mods):
    """BASELINE.json configs[0]: Reserve-Base random-init embed_video on one synthetic 16-frame 192x320 clip, fp32."""
    jax, m, *_ = mods
    from merlot_reserve_b200.synthetic import load_config, synthetic_clip, synthetic_params
    cfg = load_config("base")
    cfg["model"]["use_bfloat16"] = False
    params = synthetic_params(cfg, seed=0)
    clip = synthetic_clip(cfg, nseg=16, seed=1)
    model = m.MerlotReserve.from_config(cfg)
    t0 = time.time()
    with quiet():
        out = model.apply({"params": params}, *[jax.numpy.asarray(clip[k]) for k in ("images", "audio_clips", "tokens", "subseg_idxs")],
                          method=model.embed_video)
    dt = time.time() - t0
    return {"embed_video": np.asarray(out, np.float32), "seconds": np.float64(dt), "params_seed": np.int32(0),
            "clip_seed": np.int32(1)}


def fullsize(mods):
    """SURVEY 8d configs at size: cfg 3 (Base b=8 pretrain forward + loss), cfg 4 (Large 18x32, nseg=16, T=2464),
    cfg 5 (Large MC, L=144, 2 examples x 4 candidates).  Weights: synthetic_params(seed 0, weight_scale 4) as in
    tests/test_fullsize_gpu.py (attention far from uniform, rows differ)."""
    jax, m, ref_pm, VCR, TVQA = mods
    from merlot_reserve_b200.synthetic import load_config, synthetic_clip, synthetic_params, synthetic_pretrain_batch
    out = {}
    t0 = time.time()
    cfg = load_config("base")
    cfg["model"]["use_bfloat16"] = False
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    batch = synthetic_pretrain_batch(cfg, 8, seed=5)
    preds, loss, info = run_pretrain(jax, ref_pm, cfg, params, [batch])
    for head, d in preds[0].items():
        for k, v in d.items():
            # rows are unit vectors x exp(log_scale / 2): fp16 storage keeps 11 bits (cosine error ~1e-7), halves the file
            out[f"cfg3/{head}/{k}"] = np.asarray(v).astype(np.int32 if k == "_sources" else np.float16)
    out["cfg3/loss"] = np.float32(loss[0])
    for k, v in info.items():
        out[f"cfg3/loss_info/{k}"] = np.float32(v[0])
    print(f"cfg3 Base b=8 pretrain forward: {time.time() - t0:.1f}s loss {float(loss[0]):.5f}", flush=True)
    t0 = time.time()
    cfg = load_config("large", (18, 32))
    cfg["model"]["use_bfloat16"] = False
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    clip = synthetic_clip(cfg, nseg=16, seed=1)
    model = m.MerlotReserve.from_config(cfg)
    with quiet():
        ev = model.apply({"params": params}, *[jax.numpy.asarray(clip[k]) for k in ("images", "audio_clips", "tokens", "subseg_idxs")],
                         method=model.embed_video)
    out["cfg4/embed_video"] = np.asarray(ev, np.float32)
    print(f"cfg4 Large 18x32 nseg=16: {time.time() - t0:.1f}s", flush=True)
    t0 = time.time()
    rng = np.random.default_rng(29)
    mc_tokens = rng.integers(10, 32768, size=(2, 4, 144)).astype(np.int32)
    lens = rng.integers(20, 144, size=(2, 4))
    for b in range(2):
        for c in range(4):
            mc_tokens[b, c, lens[b, c]:] = 0
            mc_tokens[b, c, lens[b, c] - 1] = m.MASK
    mc_images = rng.random((2, 576, 768), dtype=np.float32)
    proj = (rng.standard_normal(1024) * 0.3).astype(np.float32)
    vcr = VCR.from_config(cfg)
    with quiet():
        logits = vcr.apply({"params": dict(params, proj={"kernel": proj.reshape(-1, 1)})},
                           {"image": jax.numpy.asarray(mc_images), "answers": jax.numpy.asarray(mc_tokens.reshape(2, 2, 2, 144))})
    out.update({"cfg5/mc_tokens": mc_tokens, "cfg5/mc_proj": proj, "cfg5/mc_logits": np.asarray(logits, np.float32).reshape(2, 4)})
    print(f"cfg5 Large MC L=144: {time.time() - t0:.1f}s", flush=True)
    return out


def check_oracle():
    """oracle/reserve_oracle.py against the full-size fixtures (too slow for the CPU test suite: minutes of
    torch-CPU); writes tests/golden_ref/REPORT.md.  The tiny / Base-cfg1 fixtures are checked by
    tests/test_reference_goldens.py on every run."""
    import torch
    from merlot_reserve_b200.synthetic import load_config, synthetic_clip, synthetic_params, synthetic_pretrain_batch
    from oracle import reserve_oracle as O
    ref = np.load(OUT / "fullsize.npz")
    lines = ["# oracle/ vs the reference-executed full-size fixtures (tools/run_reference_shim.py --check-oracle)", "",
             "| fixture | quantity | max abs diff / max abs ref | min row cosine |", "|---|---|---:|---:|"]

    def row(name, what, got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        rel = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
        cos = ""
        if got.ndim >= 2:
            cos = f"{((got * want).sum(-1) / (np.linalg.norm(got, axis=-1) * np.linalg.norm(want, axis=-1) + 1e-30)).min():.7f}"
        lines.append(f"| {name} | {what} | {rel:.2e} | {cos} |")
        print(lines[-1], flush=True)

    with torch.no_grad():
        cfg = load_config("base")
        orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=4.0))
        preds = orc.pretrain_forward(synthetic_pretrain_batch(cfg, 8, seed=5))
        for head, d in preds.items():
            for k, v in d.items():
                if k == "_sources":
                    same = np.array_equal(v.numpy(), ref[f"cfg3/{head}/{k}"])
                    lines.append(f"| cfg3 Base b=8 | {head}/_sources identical | {same} | |")
                else:
                    row("cfg3 Base b=8", f"{head}/{k} (fixture stored fp16)", v.numpy(), ref[f"cfg3/{head}/{k}"])
        loss, info = O.contrastive_loss(preds)
        lines.append(f"| cfg3 Base b=8 | loss | oracle {float(loss):.6f} vs reference {float(ref['cfg3/loss']):.6f} | |")
        cfg = load_config("large", (18, 32))
        orc = O.OracleMerlotReserve(cfg, synthetic_params(cfg, seed=0, weight_scale=4.0))
        row("cfg4 Large 18x32 nseg=16", "embed_video [160, 1024]", orc.embed_video(**synthetic_clip(cfg, nseg=16, seed=1)).numpy(),
            ref["cfg4/embed_video"])
        rng = np.random.default_rng(29)
        rng.integers(10, 32768, size=(2, 4, 144)); rng.integers(20, 144, size=(2, 4))
        mc_images = rng.random((2, 576, 768), dtype=np.float32)
        got = orc.score_multiple_choice(mc_images, ref["cfg5/mc_tokens"], ref["cfg5/mc_proj"]).numpy()
        row("cfg5 Large MC L=144", "logits [2, 4]", got.reshape(-1), ref["cfg5/mc_logits"].reshape(-1))
        lines.append(f"| cfg5 Large MC L=144 | argmax identical | {np.array_equal(got.argmax(-1), ref['cfg5/mc_logits'].argmax(-1))} | |")
    (OUT / "REPORT.md").write_text("\n".join(lines) + "\n")
    print("wrote", OUT / "REPORT.md")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fullsize", action="store_true")
    ap.add_argument("--only-fullsize", action="store_true")
    ap.add_argument("--check-oracle", action="store_true")
    a = ap.parse_args()
    if a.check_oracle:
        _setup_paths()
        return check_oracle()
    if not REFERENCE.exists():
        sys.exit("/root/reference is not present: this tool runs in the build container only")
    mods = reference_modules()
    OUT.mkdir(parents=True, exist_ok=True)
    if not a.only_fullsize:
        for name, scale in (("tiny_init", 1.0), ("tiny_scale4", 4.0)):
            np.savez_compressed(OUT / f"{name}.npz", **tiny_case(mods, scale))
            print("wrote", OUT / f"{name}.npz", flush=True)
        np.savez_compressed(OUT / "tiny_bf16.npz", **tiny_case(mods, 4.0, use_bf16=True))
        print("wrote", OUT / "tiny_bf16.npz", flush=True)
        np.savez_compressed(OUT / "helpers.npz", **helper_case(mods))
        (OUT / "param_trees.json").write_text(json.dumps(param_trees(mods), indent=0))
        c1 = base_cfg1(mods)
        np.savez_compressed(OUT / "base_cfg1.npz", **c1)
        print(f"wrote base_cfg1.npz (reference embed_video on numpy/torch-CPU: {float(c1['seconds']):.1f}s)", flush=True)
    if a.fullsize or a.only_fullsize:
        np.savez_compressed(OUT / "fullsize.npz", **fullsize(mods))
        print("wrote", OUT / "fullsize.npz", flush=True)


if __name__ == "__main__":
    main()
