"""Parity of the pretraining forward + contrastive loss (merlot_reserve_b200.pretrain_model) against the CPU
oracle's restatement of pretrain/pretrain_model.py:38-303 on identical weights, batch and PRNG draws.
Tolerances: embedding cosine >= 0.999 per row (x / y / y_extra are unit vectors times exp(scale/2));
contrastive logits |d| <= 2e-2 * scale + 2e-2 * |ref| with scale = exp(log_scale) (BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReservePretrainer, loss_fn_given_preds
from merlot_reserve_b200.synthetic import synthetic_params, synthetic_pretrain_batch, tiny_config
from oracle.reserve_oracle import OracleMerlotReserve, contrastive_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1.0, 4.0], ids=["init_scale", "scale4"])
def setup(request, cuda):
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=request.param)
    batch = synthetic_pretrain_batch(cfg, 2, seed=3)
    model = MerlotReservePretrainer.from_config(cfg)
    got = model.apply({"params": params}, batch)
    ref = OracleMerlotReserve(cfg, params).pretrain_forward(batch)
    return cfg, got, ref


def test_pretrain_forward_outputs(setup):
    cfg, got, ref = setup
    scale = math.exp(math.log(20.0) / 2)
    for head in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        for k, r in ref[head].items():
            g = got[head][k]
            assert tuple(g.shape) == tuple(r.shape), (head, k)
            if k == "_sources":
                assert (g.cpu().long() == r.long()).all()
                continue
            assert g.dtype == torch.bfloat16
            g = g.float().cpu()
            cos = torch.nn.functional.cosine_similarity(g, r, dim=-1)
            assert cos.min() >= 0.999, (head, k, float(cos.min()))
            assert torch.allclose(g.norm(dim=-1), torch.full((g.shape[0],), scale), atol=0.03)


def test_contrastive_logits_and_loss(setup):
    cfg, got, ref = setup
    scale = 20.0
    for head in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        lg = got[head]["x"].float().cpu() @ got[head]["y"].float().cpu().T
        lr = ref[head]["x"] @ ref[head]["y"].T
        assert (torch.abs(lg - lr) <= 2e-2 * scale + 2e-2 * lr.abs()).all(), (head, float((lg - lr).abs().max()))
    loss, info = loss_fn_given_preds(got)
    ref_loss, ref_info = contrastive_loss(ref)
    assert set(info) == set(ref_info)
    for k in ref_info:
        assert abs(float(info[k]) - float(ref_info[k])) <= 0.05 + 0.02 * abs(float(ref_info[k])), (k, float(info[k]), float(ref_info[k]))
    assert abs(float(loss) - float(ref_loss)) <= 0.1


def test_loss_kernels_match_oracle_on_same_preds(cuda):
    """loss_fn_given_preds on the ORACLE's (bf16-rounded) predictions: isolates the loss kernels."""
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    ref = OracleMerlotReserve(cfg, params).pretrain_forward(synthetic_pretrain_batch(cfg, 1, seed=5))
    rounded = {h: {k: (v.to(torch.bfloat16) if k != "_sources" else v.to(torch.int32)) for k, v in d.items()}
               for h, d in ref.items()}
    dev_preds = {h: {k: v.cuda().contiguous() for k, v in d.items()} for h, d in rounded.items()}
    loss, info = loss_fn_given_preds(dev_preds)
    ref_loss, ref_info = contrastive_loss({h: {k: (v.float() if k != "_sources" else v) for k, v in d.items()}
                                           for h, d in rounded.items()})
    for k in ref_info:
        assert abs(float(info[k]) - float(ref_info[k])) <= 2e-3, k
    assert abs(float(loss) - float(ref_loss)) <= 5e-3


@pytest.mark.parametrize("name,scale", [("tiny_init", 1.0), ("tiny_scale4", 4.0)])
def test_pretrain_forward_against_committed_golden(cuda, name, scale):
    """CUDA path vs the committed golden vectors (tests/golden/*.npz, tools/make_golden.py)."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / f"{name}.npz")
    cfg = tiny_config()
    model = MerlotReservePretrainer.from_config(cfg)
    got = model.apply({"params": synthetic_params(cfg, seed=0, weight_scale=scale)}, synthetic_pretrain_batch(cfg, 2, seed=3))
    for head in ("imgs_to_audio", "text_to_audio", "stuff_to_span"):
        for k, v in got[head].items():
            ref = torch.from_numpy(g[f"pretrain/{head}/{k}"])
            if k == "_sources":
                assert torch.equal(v.cpu().long(), ref.long())
            else:
                assert torch.nn.functional.cosine_similarity(v.float().cpu(), ref, dim=-1).min() >= 0.999, (head, k)
    loss, _ = loss_fn_given_preds(got)
    assert abs(float(loss) - float(g["pretrain/loss"])) <= 0.1
