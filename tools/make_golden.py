"""Writes tests/golden/*.npz from the CPU oracle on the synthetic tiny configuration.

JAX / Flax cannot be installed in the build environment, so these vectors come from the oracle itself
(parity unpinned, see oracle/reserve_oracle.py); `tools/run_reference_jax.py` regenerates the same file
layout from the REAL reference on a machine that has jax + flax.  The -m gpu tests compare the CUDA path
against these committed vectors; the CPU tests check that the oracle still reproduces them.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from merlot_reserve_b200.synthetic import (synthetic_clip, synthetic_params, synthetic_pretrain_batch,  # noqa: E402
                                           synthetic_spans, tiny_config)
from oracle.reserve_oracle import MASK, OracleMerlotReserve, contrastive_loss  # noqa: E402


def golden_case(weight_scale: float):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=weight_scale)
    orc = OracleMerlotReserve(cfg, params)
    clip = synthetic_clip(cfg, nseg=4, seed=21)
    spans = synthetic_spans(12, seed=3)
    rng = np.random.default_rng(17)
    mc_tokens = rng.integers(10, 32768, size=(3, 4, 24)).astype(np.int32)
    mc_tokens[:, :, -1] = MASK
    mc_images = rng.random((3, clip["images"].shape[1], 768), dtype=np.float32)
    proj = (rng.standard_normal(cfg["model"]["hidden_size"]) * 0.3).astype(np.float32)
    out = {
        "embed_video": orc.embed_video(**clip).numpy(),
        "imgseq": orc.get_imgseq_only(clip["images"]).numpy(),
        "audioseq": orc.get_audioseq_only(clip["audio_clips"]).numpy(),
        "audio_cls": orc.embed_audio_only(clip["audio_clips"]).numpy(),
        "spans": orc.embed_text_spans_only(spans).numpy(),
        "mc_logits": orc.score_multiple_choice(mc_images, mc_tokens, proj).numpy(),
        "mc_tokens": mc_tokens, "mc_proj": proj, "span_tokens": spans,
        "tokens": clip["tokens"], "subseg_idxs": clip["subseg_idxs"],
        "images_checksum": np.float64(clip["images"].astype(np.float64).sum()),
        "audio_checksum": np.float64(clip["audio_clips"].astype(np.float64).sum()),
        "weight_scale": np.float32(weight_scale),
    }
    # pretraining forward + loss (pretrain/pretrain_model.py:38-303) on synthetic_pretrain_batch(cfg, 2, seed=3)
    preds = orc.pretrain_forward(synthetic_pretrain_batch(cfg, 2, seed=3))
    loss, info = contrastive_loss(preds)
    for head, d in preds.items():
        for k, v in d.items():
            out[f"pretrain/{head}/{k}"] = v.numpy().astype(np.int32 if k == "_sources" else np.float32)
    out["pretrain/loss"] = np.float32(loss)
    for k, v in info.items():
        out[f"pretrain/loss_info/{k}"] = np.float32(v)
    return out


if __name__ == "__main__":
    gd = ROOT / "tests" / "golden"
    gd.mkdir(parents=True, exist_ok=True)
    for name, s in (("tiny_init", 1.0), ("tiny_scale4", 4.0)):
        np.savez_compressed(gd / f"{name}.npz", **golden_case(s))
        print("wrote", gd / f"{name}.npz")
