"""bf16-moment AdamW (pretrain/optimization.py:36-195): CPU checks of the oracle restatement, then the one-launch
CUDA kernel mr_adamw_multi against it, then PretrainStep.train_step's arenas."""
import math

import ml_dtypes
import numpy as np
import pytest
import torch

from oracle import optimization_oracle as oo

BF16 = ml_dtypes.bfloat16


def test_unsigned_bfloat16_code_known_answers():
    # 1.0^3 is exact in bf16: positive code, decodes to 1
    assert float(oo.unsigned_bfloat16_encode(np.float32(1.0))) == 1.0
    # v^3 = 1 + 2^-9 sits exactly on the "extra mantissa bit": bf16 rounds it to 1.0, the NEGATIVE code restores it
    v = np.cbrt(np.float64(1 + 2 ** -9)).astype(np.float32)
    code = oo.unsigned_bfloat16_encode(v)
    assert float(code) == -1.0
    assert abs(float(oo.unsigned_bfloat16_decode(code)) - float(v)) < 1e-6
    # zero stays zero, sign of +0 decodes through the non-negative branch (optimization.py:38)
    assert float(oo.unsigned_bfloat16_decode(oo.unsigned_bfloat16_encode(np.float32(0.0)))) == 0.0
    assert oo.unsigned_bfloat16_encode(np.ones(3, np.float32)).dtype == BF16


def test_unsigned_bfloat16_code_halves_the_rounding_error():
    rng = np.random.default_rng(0)
    v = np.exp(rng.uniform(np.log(1e-10), np.log(1e2), 20000)).astype(np.float32)
    dec = oo.unsigned_bfloat16_decode(oo.unsigned_bfloat16_encode(v))
    rel = np.abs(dec - v) / v
    plain = np.abs(np.cbrt((v * v * v).astype(BF16).astype(np.float32)) - v) / v
    # bf16 keeps 8 significant bits: half an ulp is 2^-8 of the cube = 2^-8 / 3 of v; the sign bit buys one more
    # level between two codes, which roughly halves the mean error
    assert rel.max() <= 2 ** -8 / 3 * 1.01 and plain.max() <= 2 ** -8 / 3 * 1.01
    assert rel.mean() < 0.7 * plain.mean()


def test_schedule_matches_reference_formula():
    f = oo.lr_scale_linearwarmup_cosinedecay
    assert f(0, 3750, 750000, 0.02) == 0.0 and f(1875, 3750, 750000, 0.02) == 0.5
    assert abs(f(3750, 3750, 750000, 0.02) - 1.0) < 1e-12            # cos(0): top of the cosine
    mid = 3750 + (750000 - 3750 + 1.0) / 2
    assert abs(f(mid, 3750, 750000, 0.02) - (0.02 + 0.98 * 0.5)) < 1e-9
    assert abs(f(10 ** 7, 3750, 750000, 0.02) - 0.02) < 1e-12        # clamped after the last step
    g = oo.lr_scale_linearwarmup_lineardecay
    assert g(5, 10, 110) == 0.5 and abs(g(60, 10, 110) - (1 - 50 / 101.0)) < 1e-12


def test_bf16_adam_first_step_and_chain():
    """count 0, zero moments: mu = (1-b1) g, nu = (1-b2) g^2, no bias correction (reference default) =>
    update = 0.1 g / (sqrt(0.02) |g| + eps); chain adds wd * p on matrices only, times schedule, times -lr."""
    g = np.array([[0.5, -2.0], [1e-3, 4.0]], np.float32)
    upd, m, v = oo.scale_by_bfloat16_adam_update(g, np.zeros_like(g, BF16), np.zeros_like(g, BF16), 0, b1=0.9, b2=0.98,
                                                 eps=1e-6, do_bias_correction=False)
    np.testing.assert_allclose(upd, 0.1 * g / (np.sqrt(0.02) * np.abs(g) + 1e-6), rtol=1e-5)
    assert m.dtype == BF16 and v.dtype == BF16
    np.testing.assert_allclose(m.astype(np.float32), 0.1 * g, rtol=2 ** -8)
    np.testing.assert_allclose(oo.unsigned_bfloat16_decode(v), 0.02 * g * g, rtol=2 ** -8 / 3 * 1.01)
    opt = {"learning_rate": 1e-2, "num_train_steps": 100, "num_warmup_steps": 10, "weight_decay_rate": 0.1, "beta_2": 0.98,
           "eps": 1e-6}
    params = {"w": np.ones((2, 2), np.float32), "b": np.ones(2, np.float32)}
    grads = {"w": g, "b": g[0]}
    st = oo.init_opt_state(params)
    p1, st1 = oo.train_state_update(params, grads, st, opt)
    assert st1["count"] == 1 and np.array_equal(p1["w"], params["w"]) and np.array_equal(p1["b"], params["b"])  # warmup: lr 0
    p2, st2 = oo.train_state_update(p1, grads, st1, opt)
    u_b, _, _ = oo.scale_by_bfloat16_adam_update(grads["b"], st1["mu"]["b"], st1["nu"]["b"], 1, 0.9, 0.98, 1e-6,
                                                 do_bias_correction=False)
    np.testing.assert_allclose(p2["b"], 1 - 1e-2 * 0.1 * u_b, rtol=1e-6)                      # vectors: no decay
    u_w, _, _ = oo.scale_by_bfloat16_adam_update(grads["w"], st1["mu"]["w"], st1["nu"]["w"], 1, 0.9, 0.98, 1e-6,
                                                 do_bias_correction=False)
    np.testing.assert_allclose(p2["w"], 1 - 1e-2 * 0.1 * (u_w + 0.1), rtol=1e-6)              # matrices: + wd * p


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("bf16_moments,bf16_grad", [(True, True), (True, False), (False, True)])
def test_adamw_multi_matches_oracle(bf16_moments, bf16_grad):
    from merlot_reserve_b200 import _lib as L
    L.require_b200(0)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    sizes = [(4100, True, "bf16"), (7, False, "f32"), (2048, True, "bf16"), (333, False, None), (16, True, "f32")]
    offs, n = [], 0
    for k, _, _ in sizes:
        offs.append(n)
        n += (k + 7) // 8 * 8
    p0 = rng.standard_normal(n).astype(np.float32)
    mdt = BF16 if bf16_moments else np.float32
    m0 = (rng.standard_normal(n) * 0.01).astype(mdt)
    v0 = oo.unsigned_bfloat16_encode(np.abs(rng.standard_normal(n)).astype(np.float32) * 1e-3) if bf16_moments \
        else (np.abs(rng.standard_normal(n)) * 1e-3).astype(np.float32)
    to_t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == BF16 else a).to(dev)
    as_dev = lambda a: to_t(a).view(torch.bfloat16) if a.dtype == BF16 else to_t(a)
    p, m, v = as_dev(p0.copy()), as_dev(m0.copy()), as_dev(v0.copy())
    works, slots = [], []
    for (k, decay, wk), off in zip(sizes, offs):
        w = None if wk is None else torch.zeros(k, dtype=torch.bfloat16 if wk == "bf16" else torch.float32, device=dev)
        works.append(w)
        slots.append((off, k, w, decay))
    table = L.adamw_chunk_table(slots, dev)
    assert table.numel() == 24 * (3 + 1 + 1 + 1 + 1)
    hyper = dict(lr=3e-3, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.1, grad_scale=0.5)
    ref_p, ref_m, ref_v = p0.copy(), m0.copy(), v0.copy()
    for step in range(3):
        g0 = (rng.standard_normal(n) * 10.0 ** rng.integers(-4, 1, n)).astype(np.float32)
        g0[3] = np.nan
        gq = g0.astype(BF16) if bf16_grad else g0
        L.adamw_multi(as_dev(gq), table, p, m, v, bias_corr1=1 - 0.9 ** (step + 1), bias_corr2=1 - 0.98 ** (step + 1), **hyper)
        gg = np.nan_to_num(gq.astype(np.float32)) * np.float32(0.5)
        for (k, decay, _), off in zip(sizes, offs):
            sl = slice(off, off + k)
            if bf16_moments:
                upd, ref_m[sl], ref_v[sl] = oo.scale_by_bfloat16_adam_update(gg[sl], ref_m[sl], ref_v[sl], step, 0.9, 0.98, 1e-6)
            else:
                ref_m[sl] = np.float32(0.1) * gg[sl] + np.float32(0.9) * ref_m[sl]
                ref_v[sl] = np.float32(0.02) * gg[sl] ** 2 + np.float32(0.98) * ref_v[sl]
                upd = (ref_m[sl] / np.float32(1 - 0.9 ** (step + 1))) / (np.sqrt(ref_v[sl] / np.float32(1 - 0.98 ** (step + 1))) + 1e-6)
            ref_p[sl] = ref_p[sl] - np.float32(3e-3) * (upd + (0.1 if decay else 0.0) * ref_p[sl])
    got_p = p.cpu().numpy()
    for (k, decay, wk), off, w in zip(sizes, offs, works):
        sl = slice(off, off + k)
        # a nu code one step off (cbrtf vs np.cbrt in the last ulp) moves the update by <= 2^-8 / 6 relative
        np.testing.assert_allclose(got_p[sl], ref_p[sl], rtol=0, atol=3 * 3e-3 * 2e-3 * (10 if bf16_moments else 1) + 1e-6)
        gm = m[sl].float().cpu().numpy()
        np.testing.assert_allclose(gm, ref_m[sl].astype(np.float32), rtol=2 ** -7, atol=1e-12)
        if bf16_moments:
            gv = oo.unsigned_bfloat16_decode(v[sl].view(torch.int16).cpu().numpy().view(BF16))
            np.testing.assert_allclose(gv, oo.unsigned_bfloat16_decode(ref_v[sl]), rtol=3e-3, atol=1e-20)
        else:
            np.testing.assert_allclose(v[sl].cpu().numpy(), ref_v[sl], rtol=1e-5, atol=1e-20)
        if w is not None:
            assert torch.equal(w.float().cpu(), p[sl].to(w.dtype).float().cpu()), "working copy refreshed from the new master"
    pad = np.ones(n, bool)
    for (k, _, _), off in zip(sizes, offs):
        pad[off:off + k] = False
    assert np.array_equal(got_p[pad], p0[pad]), "padding elements between tensor slots are never touched"


@pytest.mark.gpu
def test_train_step_arenas_follow_the_reference_optimizer():
    """PretrainStep.train_step with the reference's default optimizer block: one step from zero moments must equal the
    oracle's `train_state_update` applied to the step's own gradients (bf16-rounded, as the reference's are), for a
    decayed matrix, an undecayed vector and the packed qkv bias ([heads, 64] in flax: decayed)."""
    from merlot_reserve_b200.synthetic import synthetic_params, synthetic_pretrain_batch, tiny_config
    from merlot_reserve_b200.training import PretrainStep
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    batch = synthetic_pretrain_batch(cfg, 2, seed=3)
    step = PretrainStep(cfg, params)
    opt = dict(PretrainStep.DEFAULT_OPT, num_warmup_steps=0, num_train_steps=100, learning_rate=1e-3)
    _, _, grads = step.loss_and_grads(batch)
    step._init_state(opt)
    before = {k: {kk: vv.clone() for kk, vv in d.items()} for k, d in step.optimizer_state().items()}
    assert before["head/kernel"]["m"].dtype == torch.bfloat16
    step.train_step(batch, opt)
    after = step.optimizer_state()
    reg = step._registry()
    for name in ["head/kernel", "joint_transformer/final_ln/scale", "vision_encoder/transformer/layer_00/attention_layer/qkv/bias",
                 "token_encoder/Embed_0/embedding"]:
        g = grads[name].reshape(before[name]["p"].shape).to(torch.bfloat16).float().cpu().numpy()
        decay = reg[name][1]
        p0 = before[name]["p"].cpu().numpy()
        upd, m1, v1 = oo.scale_by_bfloat16_adam_update(g, np.zeros_like(g, BF16), np.zeros_like(g, BF16), 0, 0.9, 0.98, 1e-6,
                                                       do_bias_correction=False)
        want = p0 - 1e-3 * (upd + (0.1 if decay else 0.0) * p0)          # schedule(0) = 1 with no warm-up
        np.testing.assert_allclose(after[name]["p"].cpu().numpy(), want, rtol=0, atol=1e-3 * 4e-3 + 1e-7, err_msg=name)
        np.testing.assert_allclose(after[name]["m"].float().cpu().numpy(), m1.astype(np.float32), rtol=2 ** -7, atol=1e-12)
        live = reg[name][0]
        assert torch.equal(live.float().cpu(), after[name]["p"].to(live.dtype).float().cpu()), f"{name}: live tensor refreshed"
    assert step.step_count == 1


@pytest.mark.gpu
def test_train_state_round_trips_through_the_reference_checkpoint_format(tmp_path):
    """train 2 steps -> state_dict -> save_checkpoint (fp16 params, bf16 moments) -> load -> a fresh PretrainStep:
    parameters equal up to the fp16 storage, moments bit-identical, and the next step runs from the restored state."""
    from merlot_reserve_b200 import checkpoint as ck
    from merlot_reserve_b200.synthetic import synthetic_params, synthetic_pretrain_batch, tiny_config
    from merlot_reserve_b200.training import PretrainStep
    cfg = tiny_config()
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    batch = synthetic_pretrain_batch(cfg, 2, seed=3)
    opt = dict(PretrainStep.DEFAULT_OPT, num_warmup_steps=0, num_train_steps=100, learning_rate=1e-3)
    a = PretrainStep(cfg, params)
    for _ in range(2):
        a.train_step(batch, opt)
    sd = a.state_dict()
    assert sd["step"] == 2 and set(sd["opt_state"]) == {"0", "1", "2", "3"}
    assert sd["params"]["audio_encoder"]["embedding"]["kernel"].shape == (2, 65, cfg["model"]["hidden_size"])
    assert sd["opt_state"]["0"]["mu"]["head"]["kernel"].dtype == BF16
    ck.save_checkpoint(sd, str(tmp_path))
    restored = ck.load_checkpoint(str(tmp_path))
    b = PretrainStep(cfg, synthetic_params(cfg, seed=9, weight_scale=1.0))     # different weights: must be overwritten
    b.load_state_dict(restored, opt)
    sa, sb = a.optimizer_state(), b.optimizer_state()
    for name in sa:
        pa = sa[name]["p"].cpu()
        assert torch.equal(sb[name]["p"].cpu(), pa.to(torch.float16).float()), name      # one fp16 rounding (checkpoint.py:29-36)
        # value-identical codes (the K padding of the audio embedding comes back as +0 instead of the encoder's -0)
        assert torch.equal(sb[name]["m"].float().cpu(), sa[name]["m"].float().cpu()), name
        assert torch.equal(sb[name]["v"].float().cpu(), sa[name]["v"].float().cpu()), name
    assert b.step_count == 2
    assert np.allclose(b.model._p.contrastive_scales, a.model._p.contrastive_scales, rtol=1e-3)
    la, lb = float(a.train_step(batch, opt)[0]), float(b.train_step(batch, opt)[0])
    assert np.isfinite(lb) and abs(la - lb) < 0.05, (la, lb)


def test_gradient_arena_layout_keeps_allreduce_buckets_apart():
    """Host logic of the overlapped gradient all-reduce (training.arena_layout / ALLREDUCE_BUCKETS): every top-level
    module starts on a row of the bf16 cast, the modules of a bucket are adjacent in the arena, and no 1024-element
    row is shared by two buckets -- so a bucket can be cast and reduced while later ones are still being written."""
    from merlot_reserve_b200.training import ALLREDUCE_BUCKETS, arena_layout
    rng = np.random.default_rng(0)
    tops = ["audio_encoder", "head", "joint_transformer", "span_encoder", "token_encoder", "vision_encoder"]
    names = sorted(f"{t}/{sub}/p{i}" for t in tops for sub in ("a", "transformer/layer_00") for i in range(5))
    numel = {k: int(rng.integers(1, 5000)) for k in names}
    offs, total = arena_layout(names, numel)
    assert total % 1024 == 0 and all(o % 8 == 0 for o in offs.values())
    spans = []
    for k in names:                                              # no overlap, sorted order
        spans.append((offs[k], offs[k] + numel[k]))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert {t for b in ALLREDUCE_BUCKETS for t in b} == set(tops)
    rows = []
    for bucket in ALLREDUCE_BUCKETS:
        mine = [k for k in names if k.split("/")[0] in bucket]
        lo = min(offs[k] for k in mine) // 1024
        hi = -(-max(offs[k] + numel[k] for k in mine) // 1024)
        inside = [k for k in names if lo * 1024 <= offs[k] < hi * 1024]
        assert inside == mine, "a bucket's row range covers exactly its own tensors"
        rows.append((lo, hi))
    rows.sort()
    assert all(a[1] <= b[0] for a, b in zip(rows, rows[1:])), "buckets share no row of the cast"
    # transformer layers start on rows of their own too: a bucket can be split between layers (VISION_LAYER_SPLITS)
    names = sorted([f"vision_encoder/transformer/layer_{i:02d}/p{j}" for i in range(4) for j in range(3)] +
                   ["vision_encoder/embedding/kernel", "vision_encoder/transformer/pre_ln/scale"])
    numel = {k: int(rng.integers(1, 3000)) for k in names}
    offs, _ = arena_layout(names, numel)
    for i in range(4):
        first = f"vision_encoder/transformer/layer_{i:02d}/p0"
        assert offs[first] % 1024 == 0
    assert offs["vision_encoder/transformer/pre_ln/scale"] % 1024 == 0
    upper = [k for k in names if "/layer_02/" in k or "/layer_03/" in k]
    lo, hi = min(offs[k] for k in upper), max(offs[k] + numel[k] for k in upper)
    assert [k for k in names if lo <= offs[k] < -(-hi // 1024) * 1024] == upper
