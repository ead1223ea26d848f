"""flax/linen/attention.py (0.3.4) restated: dot_product_attention_weights / dot_product_attention."""
import numpy as np

import jax
import jax.numpy as jnp


def dot_product_attention_weights(query, key, bias=None, broadcast_dropout=True, dropout_rng=None, dropout_rate=0.,
                                  deterministic=False, dtype=jnp.float32, precision=None):
    """query [..., q, h, d], key [..., k, h, d] -> softmax((q / sqrt(d)) k^T + bias) [..., h, q, k] in `dtype`."""
    assert query.ndim == key.ndim and query.shape[-1] == key.shape[-1]
    assert dropout_rate == 0.
    depth = query.shape[-1]
    query = query / jnp.sqrt(np.float32(depth)).astype(dtype)
    attn_weights = jnp.einsum("...qhd,...khd->...hqk", query, key)
    if bias is not None:
        attn_weights = attn_weights + bias
    return jax.nn.softmax(attn_weights).astype(dtype)


def dot_product_attention(query, key, value, bias=None, broadcast_dropout=True, dropout_rng=None, dropout_rate=0.,
                          deterministic=False, dtype=jnp.float32, precision=None):
    w = dot_product_attention_weights(query, key, bias, dtype=dtype)
    return jnp.einsum("...hqk,...khd->...qhd", w, value)
