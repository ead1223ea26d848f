#!/usr/bin/env python
"""Regenerates tests/golden/*.npz with the REAL reference (rowanz/merlot_reserve) instead of the CPU oracle.

CANNOT RUN IN THE BUILD IMAGE (no jax / flax / clu there): it is the audit path for the oracle, to be run on any
machine that has `jax`, `flax` and a checkout of the reference:

    python tools/run_reference_jax.py --reference /path/to/merlot_reserve --out /tmp/golden_ref
    python tools/run_reference_jax.py --compare /tmp/golden_ref        # against tests/golden (cosine / abs diff)

It builds the same synthetic parameters and inputs as tools/make_golden.py (numpy, seeded), feeds them through
`mreserve.modeling.MerlotReserve` / `pretrain.pretrain_model.MerlotReservePretrainer` in fp32
(`use_bfloat16: False`) and writes the same keys.  The two PRNG draws inside the reference's pretraining forward
(pretrain_model.py:28 `jax.random.choice`, :219 `jax.random.uniform`) are replaced by the '_rng/*' arrays of the
synthetic batch, exactly as the oracle and the CUDA path take them.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def generate(reference: Path, out_dir: Path):
    sys.path.insert(0, str(reference))
    import jax
    import jax.numpy as jnp
    from mreserve.modeling import MerlotReserve  # the reference
    from pretrain import pretrain_model as ref_pm
    from merlot_reserve_b200.synthetic import (synthetic_clip, synthetic_params, synthetic_pretrain_batch,
                                               synthetic_spans, tiny_config)

    class Pretrainer(ref_pm.MerlotReservePretrainer):
        _splits = None

        def _augment_video_src_idx(self, video_src_idx, prng_key):          # pretrain_model.py:9-36 without the draw
            B, L = video_src_idx.shape
            split_from_here = jnp.asarray(type(self)._splits.pop(0))
            split_mask = split_from_here[:, None] <= jnp.arange(L)[None]
            return jnp.where(split_mask, video_src_idx + 4 * L, video_src_idx)

    out_dir.mkdir(parents=True, exist_ok=True)
    for name, scale in (("tiny_init", 1.0), ("tiny_scale4", 4.0)):
        cfg = tiny_config()
        cfg["model"]["use_bfloat16"] = False
        params = jax.tree_util.tree_map(jnp.asarray, synthetic_params(cfg, seed=0, weight_scale=scale))
        model = MerlotReserve.from_config(cfg)
        clip = synthetic_clip(cfg, nseg=4, seed=21)
        spans = synthetic_spans(12, seed=3)
        ap = lambda m, *a: np.asarray(model.apply({"params": params}, *a, method=m))
        out = {
            "embed_video": ap(model.embed_video, clip["images"], clip["audio_clips"], clip["tokens"], clip["subseg_idxs"]),
            "imgseq": ap(model.get_imgseq_only, clip["images"]),
            "audioseq": ap(model.get_audioseq_only, clip["audio_clips"]),
            "audio_cls": ap(model.embed_audio_only, clip["audio_clips"]),
            "spans": ap(model.embed_text_spans_only, spans),
        }
        batch = synthetic_pretrain_batch(cfg, 2, seed=3)
        pre = Pretrainer.from_config(cfg)
        # sorted-key order of mm_inputs construction in the reference: audio2text is built first, then text2audio
        Pretrainer._splits = [batch["_rng/split_audio2text"], batch["_rng/split_text2audio"]]
        real_uniform = jax.random.uniform
        jax.random.uniform = lambda key, shape, dtype=jnp.float32, minval=0.0, maxval=1.0: jnp.asarray(batch["_rng/gumbel_u"])
        try:
            jb = {k: jnp.asarray(v) for k, v in batch.items() if not k.startswith("_rng/")}
            preds = pre.apply({"params": params}, jb)
        finally:
            jax.random.uniform = real_uniform
        for head, d in preds.items():
            for k, v in d.items():
                out[f"pretrain/{head}/{k}"] = np.asarray(v)
        # single-replica loss: all_gather over a 1-sized axis via pmap
        loss, info = jax.pmap(lambda p: ref_pm.loss_fn_given_preds(p), axis_name="batch")(
            jax.tree_util.tree_map(lambda x: x[None], preds))
        out["pretrain/loss"] = np.float32(loss[0])
        for k, v in info.items():
            out[f"pretrain/loss_info/{k}"] = np.float32(np.asarray(v)[0])
        np.savez_compressed(out_dir / f"{name}.npz", **out)
        print("wrote", out_dir / f"{name}.npz")


def compare(ref_dir: Path):
    worst = 1.0
    for name in ("tiny_init", "tiny_scale4"):
        ours, ref = np.load(ROOT / "tests" / "golden" / f"{name}.npz"), np.load(ref_dir / f"{name}.npz")
        for k in ref.files:
            a, b = np.asarray(ours[k], np.float64), np.asarray(ref[k], np.float64)
            if a.ndim >= 2:
                cos = (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)
                worst = min(worst, float(cos.min()))
                print(f"{name}/{k}: min cosine {cos.min():.7f}  max |diff| {np.abs(a - b).max():.3e}")
            else:
                print(f"{name}/{k}: max |diff| {np.abs(a - b).max():.3e}")
    print("worst cosine", worst)
    return 0 if worst >= 0.9999 else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", type=Path)
    ap.add_argument("--out", type=Path, default=Path("golden_ref"))
    ap.add_argument("--compare", type=Path)
    a = ap.parse_args()
    if a.compare:
        sys.exit(compare(a.compare))
    generate(a.reference, a.out)
