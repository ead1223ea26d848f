"""CPU restatement (numpy) of the reference optimizer, pretrain/optimization.py -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(merlot_reserve_b200/training.py -> mr_adamw_multi) never does.

Parity status: UNPINNED by the reference (it ships no optimizer tests or golden vectors, and optax / jax are not
installable here).  What is restated and from where:
  * `unsigned_bfloat16_decode` / `_encode`  <- optimization.py:36-51 (sign bit of the bf16 code of nu^3 used as one more
    mantissa bit: a negative code means "multiply by 1 + 2^-9");
  * `scale_by_bfloat16_adam_update`         <- optimization.py:54-114 (bf16 mu, coded nu, update from the UN-rounded
    moments, optional bias correction :101-104 -- off by default, :165);
  * `lr_scale_linearwarmup_cosinedecay`     <- optimization.py:117-137;
  * `train_state_update`                    <- optimization.py:158-195, the optax chain
    [adam, add_decayed_weights(mask ndim > 1), scale_by_schedule, scale(-lr)] + optax.apply_updates.
optax semantics restated from library knowledge (optax 0.0.9, requirements.txt:7): `add_decayed_weights` adds
`weight_decay * p` to the update of masked leaves; `scale_by_schedule` multiplies by `schedule(count)` with `count`
the number of updates applied BEFORE this one; `scale(s)` multiplies by s; `apply_updates` is `p + u`.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import ml_dtypes
import numpy as np

BF16 = ml_dtypes.bfloat16
MISSING_PRECISION = np.float32(1 + 1 / 2 ** 9)          # optimization.py:34


def unsigned_bfloat16_decode(v: np.ndarray) -> np.ndarray:
    """optimization.py:36-39."""
    v = np.asarray(v)
    v_abs = np.abs(v.astype(np.float32))
    v_abs = np.where(v.astype(np.float32) >= 0, v_abs, v_abs * MISSING_PRECISION)
    return np.cbrt(v_abs).astype(np.float32)


def unsigned_bfloat16_encode(v: np.ndarray) -> np.ndarray:
    """optimization.py:42-50: bf16 of v^3, negated when `code * (1 + 2^-9)` is the closer of the two."""
    v = np.asarray(v, np.float32)
    v_pow = (v * v * v).astype(np.float32)
    v_bf = v_pow.astype(BF16)
    v_bf32 = v_bf.astype(np.float32)
    err0 = np.abs(v_bf32 - v_pow)
    err1 = np.abs(v_bf32 * MISSING_PRECISION - v_pow)
    return np.where(err0 < err1, v_bf32, -v_bf32).astype(BF16)


def scale_by_bfloat16_adam_update(g: np.ndarray, mu: np.ndarray, nu: np.ndarray, count: int, b1=0.9, b2=0.999, eps=1e-8,
                                  eps_root=0.0, do_bias_correction=True) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One `update_fn` call of optimization.py:84-112 on a single leaf: returns (update, new mu bf16, new nu code)."""
    g = np.asarray(g, np.float32)
    next_m = np.float32(1 - b1) * g + np.float32(b1) * np.asarray(mu).astype(np.float32)           # :86-90
    next_v = np.float32(1 - b2) * np.square(g) + np.float32(b2) * unsigned_bfloat16_decode(nu)    # :92-95
    m_enc = next_m.astype(BF16)                                                                   # :98
    v_enc = unsigned_bfloat16_encode(next_v)                                                      # :100-101
    count_inc = count + 1
    if do_bias_correction:                                                                        # :103-106
        next_m = next_m / np.float32(1 - b1 ** count_inc)
        next_v = next_v / np.float32(1 - b2 ** count_inc)
    update = next_m / (np.sqrt(next_v + np.float32(eps_root)) + np.float32(eps))                  # :109-110
    return update.astype(np.float32), m_enc, v_enc


def lr_scale_linearwarmup_cosinedecay(step: int, num_warmup_steps: int, num_train_steps: int, final_lr_scale=0.1) -> float:
    """optimization.py:117-137."""
    assert num_warmup_steps <= num_train_steps
    if step < num_warmup_steps:
        return step / num_warmup_steps
    s = min((step - num_warmup_steps) / (num_train_steps - num_warmup_steps + 1.0), 1.0)
    s = 1.0 - (1.0 - math.cos(math.pi * s)) / 2.0
    return final_lr_scale + (1.0 - final_lr_scale) * s


def lr_scale_linearwarmup_lineardecay(step: int, num_warmup_steps: int, num_train_steps: int) -> float:
    """optimization.py:140-155."""
    if step < num_warmup_steps:
        return step / num_warmup_steps
    return 1.0 - min((step - num_warmup_steps) / (num_train_steps - num_warmup_steps + 1.0), 1.0)


def init_opt_state(params: Dict[str, np.ndarray], use_bfloat16=True) -> Dict:
    """`init_fn` of optimization.py:79-82 on a flat {path: leaf} dict."""
    dt = BF16 if use_bfloat16 else np.float32
    return {"count": 0, "mu": {k: np.zeros(np.shape(v), dt) for k, v in params.items()},
            "nu": {k: np.zeros(np.shape(v), dt) for k, v in params.items()}}


def train_state_update(params: Dict[str, np.ndarray], grads: Dict[str, np.ndarray], state: Dict, opt_config: Dict):
    """One `apply_gradients` of the TrainState built by optimization.py:158-195 on flat {path: fp32 leaf} dicts.
    Returns (new params, new state).  With `use_bfloat16_adam: False` the reference falls back to
    `optax.scale_by_adam` (:70-72: fp32 moments, bias correction always on)."""
    b1, b2, eps = opt_config.get("beta_1", 0.9), opt_config.get("beta_2", 0.98), opt_config.get("eps", 1e-8)
    use_bf16 = opt_config.get("use_bfloat16_adam", True)
    bias_corr = opt_config.get("do_bias_correction", False) if use_bf16 else True
    lr_scale = lr_scale_linearwarmup_cosinedecay(state["count"], opt_config["num_warmup_steps"],
                                                 opt_config["num_train_steps"], opt_config.get("final_lr_scale", 0.02))
    new_p, new_mu, new_nu = {}, {}, {}
    for k, p in params.items():
        g = np.asarray(grads[k], np.float32)
        if use_bf16:
            upd, m, v = scale_by_bfloat16_adam_update(g, state["mu"][k], state["nu"][k], state["count"], b1, b2, eps,
                                                      do_bias_correction=bias_corr)
        else:
            m = np.float32(1 - b1) * g + np.float32(b1) * state["mu"][k]
            v = np.float32(1 - b2) * np.square(g) + np.float32(b2) * state["nu"][k]
            t = state["count"] + 1
            upd = (m / np.float32(1 - b1 ** t)) / (np.sqrt(v / np.float32(1 - b2 ** t)) + np.float32(eps))
        if np.ndim(p) > 1:                                                        # :181-184
            upd = upd + np.float32(opt_config["weight_decay_rate"]) * np.asarray(p, np.float32)
        upd = upd * np.float32(lr_scale) * np.float32(-opt_config["learning_rate"])   # :185-190
        new_p[k] = (np.asarray(p, np.float32) + upd).astype(np.float32)
        new_mu[k], new_nu[k] = m, v
    return new_p, {"count": state["count"] + 1, "mu": new_mu, "nu": new_nu}
