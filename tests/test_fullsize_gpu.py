"""Parity at BASELINE.json's FULL sizes (configs[1]: Reserve-Base, 64 clips x 16 segments, bf16).

The CPU oracle needs ~1 s per Base clip, so it checks a sample of the batch; the rest of the batch is covered by
size-independent properties of embed_video (mreserve/modeling.py:806-845): every output row is unit-norm
(unit_normalize, :570-578), a clip's embedding does not depend on which other clips share the batch or on its
position in it (vmap semantics of batch_embed_video), and duplicated clips give duplicated rows."""
import numpy as np
import pytest
import torch

from merlot_reserve_b200 import MerlotReserve
from merlot_reserve_b200.synthetic import load_config, synthetic_batch, synthetic_params
from oracle.reserve_oracle import OracleMerlotReserve

pytestmark = pytest.mark.gpu
COS_MIN = 0.999          # north_star tolerance vs the fp32 oracle
KEYS = ("images", "audio_clips", "tokens", "subseg_idxs")


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a)).float().reshape(-1, np.asarray(a).shape[-1])
    b = torch.as_tensor(np.asarray(b)).float().reshape(-1, np.asarray(b).shape[-1])
    return torch.nn.functional.cosine_similarity(a, b, dim=-1)


@pytest.fixture(scope="module")
def base(cuda):
    cfg = load_config("base")
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)      # scale 4: embeddings differ between rows / clips
    model = MerlotReserve.from_config(cfg)
    batch = synthetic_batch(cfg, 64, nseg=16, seed=21)
    out = np.asarray(model.apply({"params": params}, **batch, method=model.batch_embed_video))
    return cfg, params, model, batch, out


def test_full_batch_shape_and_unit_norm(base):
    cfg, params, model, batch, out = base
    assert out.shape == (64, 160, cfg["model"]["hidden_size"]) and np.isfinite(out).all()
    assert np.abs(np.linalg.norm(out.astype(np.float64), axis=-1) - 1.0).max() < 2e-3


def test_full_batch_sample_matches_cpu_oracle(base):
    cfg, params, model, batch, out = base
    orc = OracleMerlotReserve(cfg, params)
    for i in (0, 37, 63):
        ref = orc.embed_video(*[batch[k][i] for k in KEYS])
        cos = _cos(out[i], ref)
        assert cos.min() >= COS_MIN, f"clip {i}: min cosine {cos.min():.5f} at row {int(cos.argmin())}"


def test_clip_is_independent_of_its_batch(base):
    cfg, params, model, batch, out = base
    # alone
    for i in (5, 63):
        one = np.asarray(model.apply({"params": params}, *[batch[k][i] for k in KEYS], method=model.embed_video))
        assert _cos(out[i], one).min() >= 0.99999
        assert np.abs(out[i] - one).max() < 2e-2
    # permuted, with a duplicate: row order follows the clips, duplicates agree bit for bit
    perm = np.random.default_rng(3).permutation(64)
    perm[10] = perm[50]
    got = np.asarray(model.apply({"params": params}, **{k: np.asarray(batch[k])[perm] for k in KEYS},
                                 method=model.batch_embed_video))
    assert _cos(got.reshape(-1, got.shape[-1]), out[perm].reshape(-1, out.shape[-1])).min() >= 0.99999
    assert np.array_equal(got[10], got[50])


def test_large_resadapt_clip_matches_cpu_oracle(cuda):
    """configs[3]: Reserve-Large, 18 x 32 patch grid (288 x 512 frames), one clip of 8 segments (the reference's own
    maximum, preprocess.py:497) against the CPU oracle."""
    cfg = load_config("large", (18, 32))
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    model = MerlotReserve.from_config(cfg)
    batch = synthetic_batch(cfg, 1, nseg=8, seed=4)
    clip = [batch[k][0] for k in KEYS]
    got = np.asarray(model.apply({"params": params}, *clip, method=model.embed_video))
    ref = OracleMerlotReserve(cfg, params).embed_video(*clip)
    assert got.shape == tuple(ref.shape) == (160, 1024)
    cos = _cos(got, ref)
    assert cos.min() >= COS_MIN, f"min cosine {cos.min():.5f} at row {int(cos.argmin())}"
