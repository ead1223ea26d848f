"""CPU oracle for the MERLOT Reserve encoder + contrastive hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch fp32 on the CPU (plus an emulation of the reference's bf16 graph, `bf16_graph()`,
used to measure that graph's own rounding noise), the arithmetic that the reference
(`/root/reference/mreserve/modeling.py`, `pretrain/pretrain_model.py`) asks XLA to perform through
flax.linen.  It is the checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`merlot_reserve_b200/`) never does.

PARITY UNPINNED: the reference ships no tests or golden vectors, and JAX/Flax cannot be installed in
this environment (no wheels, no network), so this restatement could not be checked against outputs of
the reference itself.  It follows the reference source line by line (citations below) plus the
documented semantics of flax 0.3.4 (`nn.Dense`, `nn.DenseGeneral`, `nn.LayerNorm`, `nn.Conv`,
`nn.Embed`, `nn.MultiHeadDotProductAttention`, `dot_product_attention_weights`), which live outside
the reference tree.  `tools/run_reference_jax.py` regenerates the golden vectors with the real
reference on any machine that has jax + flax, so the oracle can be audited externally.

Parameters use the reference's flax pytree (nested dicts, fp32 leaves) so that a real checkpoint maps
one to one:  vision_encoder/{embedding, transformer/{cls, pre_ln, layer_XX/..., final_ln, cls_proj},
seq_attnpool/{query,key,value,out}}, audio_encoder/..., span_encoder/transformer/...,
joint_transformer/..., token_encoder/Embed_0/embedding, head/{kernel,bias}, contrastive_scales.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

# special token ids -- mreserve/lowercase_encoder.py:9-16
PADDING, START, END, MASK, MASKAUDIO, AUDIOSPAN, LTOVPOOL = 0, 1, 2, 3, 4, 5, 6


_FLOAT = torch.float32      # working precision; `with fp64():` re-runs the same restatement in double precision


class fp64:
    """Spot-check mode: activations and parameters in float64 (coordinates / sinusoid tables stay fp32 constants).
    Shows how much of any oracle-vs-kernel distance is the oracle's own fp32 rounding (it is ~1e-7 in cosine)."""

    def __enter__(self):
        global _FLOAT
        self._prev, _FLOAT = _FLOAT, torch.float64
        return self

    def __exit__(self, *exc):
        global _FLOAT
        _FLOAT = self._prev


def _t(x, dtype=None):
    dtype = dtype or _FLOAT
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x)).to(dtype)


# ------------------------------------------------------------------------------------------------
# bf16-graph emulation (rounding points marked there).  The reference runs its modules with
# dtype=bfloat16 on accelerators (modeling.py:594, pretrain_model.py:323-324): fp32 parameters are cast to bf16 at
# use and every op's output is a bf16 array.  Inside `with bf16_graph():` the layer functions below round to bf16
# at those points (values stay in fp32 tensors), which gives the NOISE FLOOR of the reference's own bf16 execution
# against this file's fp32 mode -- the yardstick for the tolerances stated in the tests.  XLA may fuse elementwise
# chains and keep fp32 inside a fusion, so this is an upper bound on the rounding the reference performs.
# ------------------------------------------------------------------------------------------------
_BF16_GRAPH = False


class bf16_graph:
    def __enter__(self):
        global _BF16_GRAPH
        self._prev, _BF16_GRAPH = _BF16_GRAPH, True
        return self

    def __exit__(self, *exc):
        global _BF16_GRAPH
        _BF16_GRAPH = self._prev


def _r(x: torch.Tensor) -> torch.Tensor:
    """A rounding point of the bf16 graph (identity in fp32 mode)."""
    return x.to(torch.bfloat16).to(x.dtype) if _BF16_GRAPH else x


def _w(x) -> torch.Tensor:
    """A parameter as the module sees it: cast to the module dtype at use."""
    return _r(_t(x))


# ------------------------------------------------------------------------------------------------
# coordinates and the "rotary" scale (modeling.py:21-144)
# ------------------------------------------------------------------------------------------------
def rotary_coordinates(seq_len: int, center_origin: bool = True) -> torch.Tensor:
    """modeling.py:21-35.  Centred: [-L//2 .. -1, 1 .. L - L//2] (zero is skipped); else 1..L."""
    if not center_origin:
        return 1.0 + torch.arange(seq_len, dtype=torch.float32)
    half = seq_len // 2
    neg = torch.arange(half, dtype=torch.float32) - float(half)
    pos = 1.0 + torch.arange(seq_len - half, dtype=torch.float32)
    return torch.cat([neg, pos], 0)


def rotary_coordinates_2d(h: int, w: int) -> torch.Tensor:
    """modeling.py:38-50: row-major [h*w, 2] = (h_coord, w_coord) / (max(h, w) + 1)."""
    scale = 1.0 / (max(h, w) + 1.0)
    hc = scale * rotary_coordinates(h)
    wc = scale * rotary_coordinates(w)
    hh, ww = torch.meshgrid(hc, wc, indexing="ij")
    return torch.stack([hh, ww], -1).reshape(h * w, 2)


def multimodal_coords(B, L, h=None, w=None, segment_idx=None, token_idx=None, max_segment=16.0, max_token=1024):
    """modeling.py:53-78: stack (h, w, segment / 16, token / 1024)."""
    z = torch.zeros(B, L, dtype=torch.float32)
    hv = z if h is None else h
    wv = z if w is None else w
    sv = z if segment_idx is None else segment_idx / max_segment
    tv = z if token_idx is None else token_idx / max_token
    return torch.stack([hv, wv, sv, tv], -1)


def rotary_sinusoids(coords: torch.Tensor, rotary_hsize: int = 32, max_freq: float = 10.0) -> torch.Tensor:
    """modeling.py:81-113.  coords [..., L, n] -> [..., 2 (cos, sin), L, rotary_hsize]."""
    *batch, L, n = coords.shape
    assert rotary_hsize % (2 * n) == 0
    d = rotary_hsize // (2 * n)
    freqs = torch.logspace(0.0, math.log2(max_freq / 2.0), d, base=2.0, dtype=torch.float32)
    rad = coords[..., None] * freqs * math.pi          # [..., L, n, d]
    rad = rad.reshape(*batch, L, n * d)
    sinus = torch.stack([torch.cos(rad), torch.sin(rad)], -3)
    return torch.repeat_interleave(sinus, 2, dim=-1)


def apply_rotary(qk: torch.Tensor, sinusoids: torch.Tensor) -> torch.Tensor:
    """modeling.py:116-144, INCLUDING its quirk: plane 0 (cos) is used where a sine is expected and
    plane 1 (sin) where a cosine is, and the "rotated" vector is [-x0, x1, -x2, x3, ...].
    qk [..., L, heads, dh]; sinusoids [..., 2, L, r]."""
    r = sinusoids.shape[-1]
    first = sinusoids[..., 0, :, None, :]    # the reference calls this `sin`; it holds cos(theta)
    second = sinusoids[..., 1, :, None, :]   # the reference calls this `cos`; it holds sin(theta)
    while first.dim() < qk.dim():
        first, second = first[None], second[None]
    head = qk[..., :r]
    flipped = torch.stack([-head[..., ::2], head[..., 1::2]], -1).reshape(head.shape)
    head = head * second + flipped * first
    return torch.cat([head, qk[..., r:]], -1)


def rotary_scale_table(coords: torch.Tensor) -> torch.Tensor:
    """Closed form of apply_rotary∘rotary_sinusoids: out[2i] = x[2i](sin-cos), out[2i+1] = x[2i+1](sin+cos).
    Used by tests to show the CUDA table is the same function."""
    s = rotary_sinusoids(coords)                      # [..., 2, L, 32]
    cos, sin = s[..., 0, :, :], s[..., 1, :, :]
    tab = torch.empty_like(cos)
    tab[..., 0::2] = (sin - cos)[..., 0::2]
    tab[..., 1::2] = (sin + cos)[..., 1::2]
    return tab


# ------------------------------------------------------------------------------------------------
# flax.linen layers restated (flax 0.3.4 semantics; library not in the reference tree)
# ------------------------------------------------------------------------------------------------
def dense(x, p):
    """nn.Dense / nn.DenseGeneral with the kernel flattened to [in, out]."""
    k = _w(p["kernel"])
    y = _r(x @ k)
    if "bias" in p:
        y = _r(y + _w(p["bias"]).reshape(-1))
    return y


def layer_norm(x, p, eps=1e-5):
    """flax LayerNorm: fp32 stats, var = E[x^2] - E[x]^2, (x - mean) * (rsqrt(var + eps) * scale) + bias."""
    x = x.to(_FLOAT)
    mean = x.mean(-1, keepdim=True)
    mean2 = (x * x).mean(-1, keepdim=True)
    var = mean2 - mean * mean
    mul = torch.rsqrt(var + eps) * _w(p["scale"])
    return _r((x - mean) * mul + _w(p["bias"]))


def my_gelu(x):
    """modeling.py:240."""
    return _r(x * torch.sigmoid(1.702 * x))


def attention(qkv, sinusoids, bias):
    """modeling.py:188-200 + flax dot_product_attention_weights.  qkv [B, T, 3nh, dh]."""
    nh = qkv.shape[-2] // 3
    qk, v = qkv[..., : 2 * nh, :], qkv[..., 2 * nh:, :]
    if sinusoids is not None:
        qk = _r(apply_rotary(qk, _r(sinusoids)))
    q, k = qk[..., :nh, :], qk[..., nh:, :]
    q = q / math.sqrt(q.shape[-1])                    # 1/8: exact in bf16
    w = _r(torch.einsum("bqhd,bkhd->bhqk", q, k))
    if bias is not None:
        w = w + bias
    w = _r(torch.softmax(w, -1))                      # flax computes the softmax of the bf16 logits, returns dtype
    return _r(torch.einsum("bhqk,bkhd->bqhd", w, v))


def transformer_layer(x, p, sinusoids, bias):
    """modeling.py:259-280 (pre-LN residual block), AttentionLayer :205-237, MLPBlock :244-256."""
    B, T, H = x.shape
    nh = H // 64
    h = layer_norm(x, p["pre_attn_ln"])
    ap = p["attention_layer"]
    qkv_k = _w(ap["qkv"]["kernel"]).reshape(H, 3 * H)
    qkv = _r(_r(h @ qkv_k) + _w(ap["qkv"]["bias"]).reshape(3 * H))
    ctx = attention(qkv.reshape(B, T, 3 * nh, 64), sinusoids, bias).reshape(B, T, H)
    x = _r(x + _r(ctx @ _w(ap["attn_proj"]["kernel"]).reshape(H, H)))
    h2 = layer_norm(x, p["pre_mlp_ln"])
    mp = p["mlp_layer"]
    u = my_gelu(_r(_r(h2 @ _w(mp["intermediate"]["kernel"])) + _w(mp["intermediate"]["bias"])))
    x = _r(x + _r(u @ _w(mp["out"]["kernel"])))
    return x


def transformer_encoder(x, p, num_layers, *, rotary_coords=None, attention_mask=None, is_valid=None,
                        add_cls_token=False, return_layers=False):
    """modeling.py:283-376."""
    B, T, H = x.shape
    if add_cls_token:
        assert attention_mask is None
        cls = _w(p["cls"]).reshape(1, 1, H).expand(B, 1, H)
        x = torch.cat([cls, x], 1)
        T += 1
        if is_valid is not None:
            is_valid = torch.cat([torch.ones(B, 1, dtype=torch.bool), is_valid], 1)
        if rotary_coords is not None:
            rotary_coords = torch.cat([torch.zeros_like(rotary_coords[..., :1, :]), rotary_coords], -2)
    sinusoids = None
    if rotary_coords is not None:
        assert rotary_coords.shape[-2] == T
        sinusoids = rotary_sinusoids(rotary_coords)
    else:
        x = _r(x + _w(p["pe"])[None])
    if is_valid is not None:
        assert attention_mask is None
        attention_mask = is_valid[:, :, None] & is_valid[:, None, :]
    bias = None
    if attention_mask is not None:
        m = attention_mask[:, None, :, :]
        bias = torch.where(m, torch.zeros((), dtype=torch.float32), torch.full((), -1e10, dtype=torch.float32))
    x = layer_norm(x, p["pre_ln"])
    layers = []
    for i in range(num_layers):
        x = transformer_layer(x, p[f"layer_{i:02d}"], sinusoids, bias)
        if return_layers:
            layers.append(x)
    xf = layer_norm(x, p["final_ln"])
    out = {}
    if add_cls_token:
        out["cls"] = dense(xf[:, 0], p["cls_proj"])
        out["seq"] = xf[:, 1:]
    else:
        out["seq"] = xf
    if return_layers:
        out["layers"] = layers
    return out


def mha_pool(q_in, kv_in, p):
    """flax MultiHeadDotProductAttention with inputs_q [N, 1, H], inputs_kv [N, W, H]
    (modeling.py:424-426, :468-472): q/k/v DenseGeneral(+bias) -> attention -> out DenseGeneral(+bias)."""
    N, W, H = kv_in.shape
    nh = H // 64

    def proj(x, pp):
        return _r(_r(x @ _w(pp["kernel"]).reshape(H, H)) + _w(pp["bias"]).reshape(H)).reshape(*x.shape[:-1], nh, 64)

    q = proj(q_in, p["query"]) / math.sqrt(64)
    k = proj(kv_in, p["key"])
    v = proj(kv_in, p["value"])
    w = _r(torch.softmax(_r(torch.einsum("nqhd,nkhd->nhqk", q, k)), -1))
    ctx = _r(torch.einsum("nhqk,nkhd->nqhd", w, v)).reshape(N, q_in.shape[1], H)
    return _r(_r(ctx @ _w(p["out"]["kernel"]).reshape(H, H)) + _w(p["out"]["bias"]))


# ------------------------------------------------------------------------------------------------
# the model (modeling.py:379-931)
# ------------------------------------------------------------------------------------------------
class OracleMerlotReserve:
    """fp32 CPU restatement of `MerlotReserve` (modeling.py:581-931).  `config` is the reference's
    YAML dict (`pretrain/configs/*.yaml`); `params` the flax param tree."""

    def __init__(self, config: Dict, params: Dict):
        m = dict(config["model"])
        self.data = dict(config.get("data", {}))
        self.cfg = m
        self.params = params
        self.hidden_size = m["hidden_size"]
        self.output_grid_h, self.output_grid_w = m["output_grid"]
        self.vit_pooling_ratio = m["vit_pooling_ratio"]
        self.audio_seq_length = m["audio_seq_length"]
        self.audio_token_length = m["audio_token_length"]
        self.audio_patch_size = m["audio_patch_size"]
        # modeling.py:611-612
        self.audio_pooling_ratio = self.audio_seq_length // (self.audio_token_length * self.audio_patch_size)

    # --- encoders ---
    def vision_encoder(self, x):
        """VisionTransformer.__call__ modeling.py:391-430.  x [B, hw, 768]."""
        p = self.params["vision_encoder"]
        H, gh, gw, r = self.hidden_size, self.output_grid_h, self.output_grid_w, self.vit_pooling_ratio
        B, hw, pp3 = x.shape
        assert hw == gh * gw and pp3 == (self.cfg["vit_patch_size"] ** 2) * 3
        x = dense(_r(_t(x)), p["embedding"])
        coords = rotary_coordinates_2d(gh, gw)
        t_out = transformer_encoder(x, p["transformer"], self.cfg["vit_num_layers"], rotary_coords=coords,
                                    add_cls_token=True)
        h2, w2 = gh // r, gw // r
        seq = t_out["seq"].reshape(B * h2, r, w2, r, H).transpose(1, 2).reshape(B * h2 * w2, r * r, H)
        pooled = mha_pool(_r(seq.mean(1, keepdim=True)), seq, p["seq_attnpool"])
        t_out["seq_attnpool"] = pooled.reshape(B, h2 * w2, H)
        return t_out

    def audio_encoder(self, x):
        """AudioTransformer.__call__ modeling.py:443-476.  x [B, 60, 65]."""
        p = self.params["audio_encoder"]
        H, ps = self.hidden_size, self.audio_patch_size
        x = _t(x)
        B, raw_len, nm = x.shape
        assert nm == 65 and raw_len % ps == 0
        L = raw_len // ps
        # nn.Conv(kernel [ps], stride [ps], SAME): no padding when raw_len % ps == 0 -> reshape-GEMM
        k = _w(p["embedding"]["kernel"]).reshape(ps * nm, H)
        x = _r(_r(_r(x).reshape(B, L, ps * nm) @ k) + _w(p["embedding"]["bias"]))
        coords = rotary_coordinates(L)[:, None] / L
        t_out = transformer_encoder(x, p["transformer"], self.cfg["audio_num_layers"], rotary_coords=coords,
                                    add_cls_token=True)
        r = self.audio_pooling_ratio
        assert L % r == 0
        seq = t_out["seq"].reshape(-1, r, H)
        pooled = mha_pool(_r(seq.mean(1, keepdim=True)), seq, p["seq_attnpool"])
        t_out["seq_attnpool"] = pooled.reshape(B, L // r, H)
        return t_out

    def span_encoder(self, x, x_isvalid):
        """SpanTransformer.__call__ modeling.py:488-504."""
        L = x.shape[1]
        assert L < 16
        coords = rotary_coordinates(L, center_origin=False)[:, None] / 16
        t_out = transformer_encoder(x, self.params["span_encoder"]["transformer"], self.cfg["span_num_layers"],
                                    rotary_coords=coords, is_valid=x_isvalid, add_cls_token=True)
        return t_out["cls"]

    def token_encoder(self, tokens):
        """TokenEmbedder modeling.py:507-538 (one nn.Embed lookup)."""
        emb = _w(self.params["token_encoder"]["Embed_0"]["embedding"])
        return emb[torch.as_tensor(np.asarray(tokens)).long()]

    def joint_transformer(self, x, rotary_coords, attention_mask, return_layers=False):
        return transformer_encoder(x, self.params["joint_transformer"], self.cfg["joint_num_layers"],
                                   rotary_coords=rotary_coords, attention_mask=attention_mask,
                                   return_layers=return_layers)

    def joint_proj(self, x):
        return dense(x, self.params["head"])

    # --- joint input assembly (modeling.py:651-758) ---
    def prepare_multimodal_inputs(self, tokens, token_segment_idx=None, token_embs=None, vision_input=None,
                                  audio_spans=None, audio_pointers=None, padding_len=None, video_src_idx=None):
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        B, L = tokens.shape
        H = self.hidden_size
        if token_embs is None:
            token_embs = self.token_encoder(tokens)
        if audio_spans is not None and audio_pointers is not None:
            audio_pointers = torch.as_tensor(np.asarray(audio_pointers)).long()
            assert audio_spans.shape[0] == B and audio_spans.shape[2] == self.audio_token_length
            is_audio = tokens == AUDIOSPAN
            ptr = audio_pointers.clamp(min=0)
            subpos = (torch.cumsum(is_audio.int(), -1) - 1).clamp(min=0) % self.audio_token_length
            audio_embs = audio_spans[torch.arange(B)[:, None], ptr, subpos]
            token_embs = torch.where(is_audio[..., None], audio_embs, token_embs)
        token_idx = (1.0 + torch.arange(L, dtype=torch.float32))[None].expand(B, L)
        seg = None
        if token_segment_idx is not None:
            token_segment_idx = torch.as_tensor(np.asarray(token_segment_idx)).long()
            seg = token_segment_idx.float()
        coords = multimodal_coords(B, L, segment_idx=seg, token_idx=token_idx)
        vis_len, vis_seg = 0, None
        if vision_input is not None:
            hp = self.output_grid_h // self.vit_pooling_ratio
            wp = self.output_grid_w // self.vit_pooling_ratio
            pool_coords = rotary_coordinates_2d(hp, wp)
            vis_len = vision_input.shape[1]
            nseg = vis_len // (hp * wp)
            img_coords = pool_coords.repeat(nseg, 1)[None].expand(B, -1, -1)
            vis_seg = torch.arange(nseg).repeat_interleave(hp * wp)[None].expand(B, -1)
            img_mm = multimodal_coords(B, vis_len, h=img_coords[..., 0], w=img_coords[..., 1],
                                       segment_idx=vis_seg.float())
            coords = torch.cat([coords, img_mm], 1)
            token_embs = torch.cat([token_embs, vision_input], 1)
        is_valid = tokens != PADDING
        if vis_len > 0:
            is_valid = torch.cat([is_valid, torch.ones(B, vis_len, dtype=torch.bool)], 1)
        extra = 0
        if padding_len is not None:
            extra = padding_len - is_valid.shape[1]
            assert extra >= 0
            if extra > 0:
                is_valid = torch.cat([is_valid, torch.zeros(B, extra, dtype=torch.bool)], 1)
                coords = torch.cat([coords, torch.zeros(B, extra, 4)], 1)
                token_embs = torch.cat([token_embs, torch.zeros(B, extra, H)], 1)
        mask = is_valid[:, None, :] & is_valid[:, :, None]
        video_src = None
        if video_src_idx is not None and token_segment_idx is not None:
            video_src_idx = torch.as_tensor(np.asarray(video_src_idx)).long()
            rows = torch.arange(B)[:, None]
            parts = [video_src_idx[rows, token_segment_idx]]
            if vis_seg is not None:
                parts.append(video_src_idx[rows, vis_seg])
            if extra > 0:
                parts.append(torch.full((B, extra), -1, dtype=torch.long))
            video_src = torch.cat(parts, -1)
            mask = mask & (video_src[:, None, :] == video_src[:, :, None])
        return {"x": token_embs, "rotary_coords": coords, "attention_mask": mask,
                "is_valid": is_valid, "video_src": video_src}

    # --- zero-shot API (modeling.py:767-931) ---
    def embed_text_spans_only(self, text_spans):
        text_spans = torch.as_tensor(np.asarray(text_spans)).long()
        return unit_normalize(self.span_encoder(self.token_encoder(text_spans), text_spans != PADDING))

    def embed_audio_only(self, audio_clips):
        a = _t(audio_clips)
        lead = a.shape[:-2]
        cls = self.audio_encoder(a.reshape(-1, self.audio_seq_length, 65))["cls"]
        return unit_normalize(cls).reshape(*lead, self.hidden_size)

    def get_imgseq_only(self, imgs):
        x = _t(imgs)
        lead, n, pp3 = x.shape[:-2], x.shape[-2], x.shape[-1]
        enc = self.vision_encoder(x.reshape(-1, n, pp3))["seq_attnpool"]
        return enc.reshape(*lead, n // 4, self.hidden_size)

    def get_audioseq_only(self, audio_clips):
        return self.audio_encoder(_t(audio_clips).reshape(-1, self.audio_seq_length, 65))["seq_attnpool"]

    def _joint_tail(self, mm, token_length):
        enc = self.joint_transformer(mm["x"], mm["rotary_coords"], mm["attention_mask"])["seq"]
        return unit_normalize(self.joint_proj(enc[0, :token_length]))

    def embed_video(self, images, audio_clips, tokens, subseg_idxs):
        """modeling.py:806-843."""
        images, audio_clips = _t(images), _t(audio_clips)
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        subseg = torch.as_tensor(np.asarray(subseg_idxs)).long()
        nseg, npatch, pp3 = images.shape
        assert pp3 == 768
        assert audio_clips.shape == (3 * nseg, self.audio_seq_length, 65)
        (L,) = tokens.shape
        assert subseg.shape == (L,)
        imgs_enc = self.vision_encoder(images)["seq_attnpool"].reshape(nseg * npatch // 4, self.hidden_size)
        audio_enc = self.audio_encoder(audio_clips)["seq_attnpool"]
        mm = self.prepare_multimodal_inputs(
            tokens=tokens[None], token_segment_idx=torch.div(subseg[None], 3, rounding_mode="floor"),
            vision_input=imgs_enc[None], audio_pointers=subseg[None], audio_spans=audio_enc[None])
        return self._joint_tail(mm, L)

    def batch_embed_video(self, images, audio_clips, tokens, subseg_idxs):
        return torch.stack([self.embed_video(i, a, t, s)
                            for i, a, t, s in zip(images, audio_clips, tokens, subseg_idxs)], 0)

    def embed_singleimg_with_multiimg_prompt(self, images_prompt, images, tokens, subseg_idxs):
        """modeling.py:848-878."""
        images_prompt, images = _t(images_prompt), _t(images)
        ns0, (ns1, npatch, _) = images_prompt.shape[0], images.shape
        assert ns0 + ns1 <= 8
        enc = torch.cat([images_prompt, self.vision_encoder(images)["seq_attnpool"]], 0)
        return self.embed_preencoded_noaudio(enc, tokens, subseg_idxs)

    def embed_preencoded_noaudio(self, images_enc, tokens, subseg_idxs):
        """modeling.py:880-904."""
        images_enc = _t(images_enc).reshape(-1, self.hidden_size)
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        subseg = torch.as_tensor(np.asarray(subseg_idxs)).long()
        mm = self.prepare_multimodal_inputs(
            tokens=tokens[None], token_segment_idx=torch.div(subseg[None], 3, rounding_mode="floor"),
            vision_input=images_enc[None])
        return self._joint_tail(mm, tokens.shape[0])

    def embed_preencoded_audio(self, images_enc, audio_enc, tokens, subseg_idxs, audio_pointers):
        """modeling.py:906-931."""
        images_enc = _t(images_enc).reshape(-1, self.hidden_size)
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        subseg = torch.as_tensor(np.asarray(subseg_idxs)).long()
        ptrs = torch.as_tensor(np.asarray(audio_pointers)).long()
        mm = self.prepare_multimodal_inputs(
            tokens=tokens[None], token_segment_idx=torch.div(subseg[None], 3, rounding_mode="floor"),
            vision_input=images_enc[None], audio_pointers=ptrs[None], audio_spans=_t(audio_enc)[None])
        return self._joint_tail(mm, tokens.shape[0])

    # --- names required by BASELINE.json north_star ---
    def get_imgs_enc(self, images):
        x = _t(images)
        lead, n, pp3 = x.shape[:-2], x.shape[-2], x.shape[-1]
        enc = self.vision_encoder(x.reshape(-1, n, pp3))["seq_attnpool"]
        return enc.reshape(*lead, n // (self.vit_pooling_ratio ** 2), self.hidden_size)

    def get_audio_enc(self, audio_clips):
        a = _t(audio_clips)
        lead = a.shape[:-2]
        enc = self.audio_encoder(a.reshape(-1, self.audio_seq_length, 65))["seq_attnpool"]
        return enc.reshape(*lead, self.audio_token_length, self.hidden_size)

    def embed_singleimg_with_multiple_text(self, images, tokens, subseg_idxs):
        """Vision encoded once, broadcast to every candidate, then the joint tail per candidate
        (pattern of finetune/vcr/qa_qar_joint_finetune.py:155-163 + modeling.py:869-878)."""
        images = _t(images)
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        subseg = torch.as_tensor(np.asarray(subseg_idxs)).long()
        n_cand, L = tokens.shape
        enc = self.vision_encoder(images)["seq_attnpool"].reshape(1, -1, self.hidden_size)
        mm = self.prepare_multimodal_inputs(
            tokens=tokens, token_segment_idx=torch.div(subseg, 3, rounding_mode="floor"),
            vision_input=enc.expand(n_cand, -1, -1))
        seq = self.joint_transformer(mm["x"], mm["rotary_coords"], mm["attention_mask"])["seq"]
        return unit_normalize(self.joint_proj(seq[:, :L]))

    def score_multiple_choice(self, images, tokens, proj_kernel):
        """MerlotReserveVCR.__call__ (finetune/vcr/qa_qar_joint_finetune.py:150-170):
        images [B, hw, 768], tokens [B, n_cand, L] -> logits [B, n_cand]."""
        images = _t(images)
        tokens = torch.as_tensor(np.asarray(tokens)).long()
        B, n_cand, L = tokens.shape
        enc = self.vision_encoder(images)["seq_attnpool"]
        enc = enc.repeat_interleave(n_cand, dim=0)
        flat = tokens.reshape(B * n_cand, L)
        mm = self.prepare_multimodal_inputs(tokens=flat, token_segment_idx=torch.zeros_like(flat), vision_input=enc)
        seq = self.joint_transformer(mm["x"], mm["rotary_coords"], mm["attention_mask"])["seq"][:, :L]
        pool_idx = (flat == MASK).int().argmax(1)
        pooled = seq[torch.arange(B * n_cand), pool_idx]
        return (pooled @ _t(proj_kernel).reshape(-1, 1)).reshape(B, n_cand)


    def score_tvqa(self, batch, proj_kernel):
        """MerlotReserveTVQA.__call__ (finetune/tvqa/tvqa_finetune.py:163-258), literally: a zero "frame 0" is put in
        front of the encoded frames so the real frames get segment coordinates 1..n, both joint inputs are assembled,
        and the rows (and mask columns) of that first frame are cut out again (:212-224).
        batch: images [B, n_img, hw, 768]; audio_clips [B, n_seg, n_sub, 60, 65]; textonly_seqs / audio_seqs
        [B, n_ans, L, 2] (token, segment).  Returns (logits_from_audio, logits_from_text), each [B, n_ans]."""
        images = _t(batch["images"])
        B, n_img, hw, pp3 = images.shape
        H = self.hidden_size
        imgs_enc = self.vision_encoder(images.reshape(B * n_img, hw, pp3))["seq_attnpool"].reshape(B, n_img, hw // 4, H)
        imgs_enc = torch.cat([torch.zeros(B, 1, hw // 4, H), imgs_enc], 1)                       # :170-174
        n_img += 1
        text = torch.as_tensor(np.asarray(batch["textonly_seqs"])).long()
        _, n_ans, L, _ = text.shape
        imgs_enc = imgs_enc.reshape(B, n_img * hw // 4, H).repeat_interleave(n_ans, dim=0)       # :178-179
        text_toks = text[..., 0].reshape(B * n_ans, L)
        text_in = self.prepare_multimodal_inputs(tokens=text_toks, token_segment_idx=text[..., 1].reshape(B * n_ans, L),
                                                 vision_input=imgs_enc)
        audio_clips = _t(batch["audio_clips"])
        _, n_seg, n_sub, alen, mels = audio_clips.shape
        audio_enc = self.audio_encoder(audio_clips.reshape(-1, alen, mels))["seq_attnpool"]
        atl = audio_enc.shape[1]
        n_spans = n_seg * n_sub
        audio_enc = audio_enc.reshape(B, n_spans, atl, H).repeat_interleave(n_ans, dim=0)       # :196-197
        aseq = torch.as_tensor(np.asarray(batch["audio_seqs"])).long()
        audio_toks = aseq[..., 0].reshape(B * n_ans, L)
        ptr = torch.div(torch.cumsum((audio_toks == AUDIOSPAN).int(), -1) - 1, atl, rounding_mode="floor") % n_spans  # :200-201
        audio_in = self.prepare_multimodal_inputs(tokens=audio_toks, token_segment_idx=aseq[..., 1].reshape(B * n_ans, L),
                                                  vision_input=imgs_enc, audio_spans=audio_enc, audio_pointers=ptr)
        start = L + hw // 4                                                                      # :212
        for d in (text_in, audio_in):
            for k in ("x", "rotary_coords", "attention_mask"):
                d[k] = torch.cat([d[k][:, :L], d[k][:, start:]], 1)
            d["attention_mask"] = torch.cat([d["attention_mask"][:, :, :L], d["attention_mask"][:, :, start:]], 2)
        x = torch.cat([audio_in["x"], text_in["x"]], 0)                                          # :243-245
        coords = torch.cat([audio_in["rotary_coords"], text_in["rotary_coords"]], 0)
        mask = torch.cat([audio_in["attention_mask"], text_in["attention_mask"]], 0)
        joint = self.joint_transformer(x, coords, mask)["seq"][:, :L]
        toks = torch.cat([audio_toks, text_toks], 0)
        pool_idx = (toks == MASK).float().argmax(1)                                              # :251
        pooled = joint[torch.arange(2 * B * n_ans), pool_idx]
        logits = (pooled @ _t(proj_kernel).reshape(-1, 1)).squeeze(-1)
        la, lt = logits[: B * n_ans], logits[B * n_ans:]
        return la.reshape(B, n_ans), lt.reshape(B, n_ans)

    # --- pretraining forward (pretrain/pretrain_model.py:38-259) ---
    @staticmethod
    def _augment_video_src_idx(video_src_idx, split_from_here):
        """pretrain_model.py:9-36 with the PRNG draw `split_from_here` [B] (values 1..L; L = no split) supplied."""
        B, L = video_src_idx.shape
        if L == 1:
            return video_src_idx
        split_mask = split_from_here[:, None] <= torch.arange(L)[None]
        return torch.where(split_mask, video_src_idx + 4 * L, video_src_idx)

    def pretrain_forward(self, batch):
        """MerlotReservePretrainer.__call__ (pretrain_model.py:38-259), fp32.  The two PRNG draws of the reference
        are inputs: batch['_rng/split_audio2text'], ['_rng/split_text2audio'] (`split_from_here`, :28) and
        ['_rng/gumbel_u'] (the uniform of :219).  jnp.argsort / lax.top_k tie order: stable, lower index first."""
        d = self.data
        G, nsub = d["num_segment_groups"], d["num_audio_subsegments"]
        Ltxt, Lseq = d["lang_seq_len"], d["seq_len"]
        H = self.hidden_size
        lng = lambda x: torch.as_tensor(np.asarray(x)).long()
        images = _t(batch["images"])
        b, nseg_hw, pp3 = images.shape
        hw = self.output_grid_h * self.output_grid_w
        nseg = nseg_hw // hw
        per_group = nseg // G
        imgs_enc = self.vision_encoder(images.reshape(b * nseg, hw, pp3))                                   # :53
        hw1 = hw // (self.vit_pooling_ratio ** 2)
        imgs_seq = imgs_enc["seq_attnpool"].reshape(b, G, per_group * hw1, H)                               # :56-57
        vis_len = imgs_seq.shape[-2]
        audio_enc = self.audio_encoder(_t(batch["audio_clips"]).reshape(b * nseg * nsub, self.audio_seq_length, -1))
        n_aud = nseg * nsub
        audio_seq = audio_enc["seq_attnpool"].reshape(b, n_aud, self.audio_token_length, H)                 # :72-73
        audio_cls = audio_enc["cls"].reshape(b, n_aud, H)                                                   # :74
        bt = {}
        for k1 in ("text2audio", "audio2text"):                                                             # :77-80
            for k2 in ("", "/audio_ptr", "/text_ptr"):
                bt[k1 + k2] = lng(batch[k1 + k2]).reshape(-1, Ltxt)
        for k in ("random_text", "random_text/text_ptr", "audio_text_matching", "audio_text_matching/audio_ptr"):
            bt[k] = lng(batch[k]).reshape(-1, Lseq)                                                         # :82-83
        text_spans = lng(batch["text_spans"]).reshape(-1, self.cfg["text_span_length"])                     # :85
        vsi = lng(batch["video_src_index"]).reshape(-1, per_group)                                          # :92
        n_a2t, n_t2a = d["num_audio2text_seqs"], d["num_text2audio_seqs"]

        def tiled_vsi(n):
            return vsi.reshape(b, G, per_group).repeat(1, n, 1).reshape(-1, per_group)

        fdiv = lambda a, q: torch.div(a, q, rounding_mode="floor")
        mm = {}
        mm["audio2text"] = self.prepare_multimodal_inputs(                                                  # :99-113
            tokens=bt["audio2text"], token_segment_idx=torch.remainder(fdiv(bt["audio2text/audio_ptr"], nsub), per_group),
            vision_input=imgs_seq.repeat(1, n_a2t, 1, 1).reshape(-1, vis_len, H),
            audio_spans=audio_seq.repeat_interleave(G * n_a2t, dim=0), audio_pointers=bt["audio2text/audio_ptr"],
            padding_len=Lseq,
            video_src_idx=self._augment_video_src_idx(tiled_vsi(n_a2t), lng(batch["_rng/split_audio2text"])))
        atm = bt["audio_text_matching"]
        mm["audio_text_matching"] = self.prepare_multimodal_inputs(                                         # :115-122
            tokens=atm, token_segment_idx=torch.cumsum((atm == LTOVPOOL).long(), -1), audio_spans=audio_seq,
            audio_pointers=bt["audio_text_matching/audio_ptr"], padding_len=Lseq)
        mm["text2audio"] = self.prepare_multimodal_inputs(                                                  # :125-137
            tokens=bt["text2audio"], token_segment_idx=torch.remainder(fdiv(bt["text2audio/audio_ptr"], nsub), per_group),
            vision_input=imgs_seq.repeat(1, n_t2a, 1, 1).reshape(-1, vis_len, H),
            audio_pointers=bt["text2audio/audio_ptr"], padding_len=Lseq,
            video_src_idx=self._augment_video_src_idx(tiled_vsi(n_t2a), lng(batch["_rng/split_text2audio"])))
        mm["random_text"] = self.prepare_multimodal_inputs(tokens=bt["random_text"], padding_len=Lseq)      # :138
        keys = sorted(mm)                                                                                   # :140
        x = torch.cat([mm[k]["x"] for k in keys], 0)
        coords = torch.cat([mm[k]["rotary_coords"] for k in keys], 0)
        mask = torch.cat([mm[k]["attention_mask"] for k in keys], 0)
        sizes = [mm[k]["x"].shape[0] for k in keys]
        joint = self.joint_proj(self.joint_transformer(x, coords, mask)["seq"])                             # :150-151
        mo = dict(zip(keys, torch.split(joint, sizes, 0)))
        mo["text2audio"] = mo["text2audio"][:, :Ltxt]                                                       # :154-155
        mo["audio2text"] = mo["audio2text"][:, :Ltxt]

        is_pool = atm == LTOVPOOL                                                                           # :160-165
        a2v = one_hot_pool(is_pool, torch.cumsum(is_pool.long(), -1) - 1, mo["audio_text_matching"],
                           nseg)["x"].reshape(b * nseg, H)
        t2a = one_hot_pool(bt["text2audio"] == MASKAUDIO, bt["text2audio/audio_ptr"], mo["text2audio"], n_aud,
                           real_bsize=b)                                                                    # :170-176
        n_trg = int(n_aud * d["mask_rate"]) * n_t2a                                                         # :178
        is_selected = t2a["idx_oh"].sum(1)
        idx_sort = torch.sort(-is_selected, dim=-1, stable=True)[1]                                         # :181
        best = idx_sort[:, :n_trg].reshape(b * n_trg)
        bi = torch.arange(b).repeat_interleave(n_trg)
        t2a_sel = t2a["x"][bi, best]
        a2t_sel = audio_cls[bi, best]
        extra = idx_sort[:, n_trg:].reshape(b * (n_aud - n_trg))
        a2t_extra = audio_cls[torch.arange(b).repeat_interleave(n_aud - n_trg), extra]                      # :188-190

        n_spans = text_spans.shape[0] // b                                                                  # :195
        t2sp = {}
        for k in ("audio2text", "text2audio", "random_text"):
            r = one_hot_pool(bt[k] == MASK, bt[f"{k}/text_ptr"], mo[k], n_spans, real_bsize=b)
            t2sp[k] = {"x": r["x"], "count": r["idx_oh"].sum(1)}
        sel = t2sp["text2audio"]["x"] + t2sp["audio2text"]["x"] + t2sp["random_text"]["x"]                  # :206
        ct = t2sp["text2audio"]["count"] + t2sp["audio2text"]["count"] + t2sp["random_text"]["count"]
        src = torch.stack([torch.zeros_like(ct), t2sp["text2audio"]["count"], t2sp["audio2text"]["count"],
                           t2sp["random_text"]["count"]], -1).argmax(-1) - 1                                # :208-209
        is_valid = (text_spans != PADDING).any(-1).reshape(b, n_spans) & (ct > 0.0)                         # :212-214
        logits_for_pred = is_valid.float() * 1e6 + math.log(4.0) * (t2sp["text2audio"]["count"] + t2sp["audio2text"]["count"])
        z = -torch.log(-torch.log(_t(batch["_rng/gumbel_u"])))                                              # :219-220
        n_inc = d["num_text_spans_to_include"]
        score = (logits_for_pred + z).reshape(-1)
        best = torch.sort(-score, stable=True)[1][:n_inc * b]                                               # :227 (top_k)
        t2sp_sel = sel.reshape(b * n_spans, H)[best]
        t2sp_src = src.reshape(b * n_spans)[best]
        sp2t_sel = self.span_encoder(self.token_encoder(text_spans[best]), text_spans[best] != PADDING)     # :232-233
        scales = _t(self.params["contrastive_scales"]).clamp(max=math.log(100.0))                           # :236
        out = {
            "imgs_to_audio": {"x": a2v, "y": imgs_enc["cls"], "log_scale": scales[0]},
            "text_to_audio": {"x": t2a_sel, "y": a2t_sel, "y_extra": a2t_extra, "log_scale": scales[1]},
            "stuff_to_span": {"x": t2sp_sel, "y": sp2t_sel, "log_scale": scales[2], "_sources": t2sp_src},
        }
        for k in out:                                                                                       # :243-257
            temp = torch.exp(out[k].pop("log_scale") / 2.0)
            for k2 in ("x", "y", "y_extra"):
                if k2 in out[k]:
                    out[k][k2] = unit_normalize(out[k][k2]) * temp
        return out


def frames_to_patches(frames_u8: np.ndarray, grid_h: int, grid_w: int, patch: int = 16) -> np.ndarray:
    """The tail of preprocess_image_to_patches (mreserve/preprocess.py:464-481) for frames that already have the
    model's resolution: tf.image.convert_image_dtype(uint8 -> float32) = float32(x) * float32(1/255) (:471),
    tf.nn.space_to_depth(img[None], P, 'NHWC') -- output channel (py * P + px) * 3 + c of block (by, bx) -- and the
    reshape to [h1 * w1, P * P * 3] (:478-479).  frames [B, grid_h*P, grid_w*P, 3] uint8 -> [B, h1*w1, P*P*3] fp32."""
    f = np.asarray(frames_u8)
    assert f.dtype == np.uint8 and f.shape[1:] == (grid_h * patch, grid_w * patch, 3)
    x = f.astype(np.float32) * np.float32(1.0 / 255.0)
    B = x.shape[0]
    out = np.empty((B, grid_h * grid_w, patch * patch * 3), np.float32)
    for by in range(grid_h):
        for bx in range(grid_w):
            blk = x[:, by * patch:(by + 1) * patch, bx * patch:(bx + 1) * patch, :]      # [B, py, px, c]
            out[:, by * grid_w + bx] = blk.reshape(B, -1)
    return out


def unit_normalize(x):
    """modeling.py:570-578."""
    x = x.to(_FLOAT)
    return x / torch.sqrt((x * x).sum(-1, keepdim=True) + 1e-5)


def one_hot_pool(do_pool, idx, v, num_segments, real_bsize=None):
    """modeling.py:541-567 (einsum against a one-hot; index -1 -> all-zero row)."""
    B, L, H = v.shape
    if real_bsize is not None:
        l2 = (L * B) // real_bsize
        do_pool, idx, v = do_pool.reshape(real_bsize, l2), idx.reshape(real_bsize, l2), v.reshape(real_bsize, l2, H)
    ptr = torch.where(do_pool, idx, torch.full_like(idx, -1))
    oh = (ptr[..., None] == torch.arange(num_segments)).float()
    return {"x": torch.einsum("bls,blh->bsh", oh, v.float()), "idx_oh": oh}


def contrastive_loss(preds: Dict[str, Dict[str, torch.Tensor]], gather=None):
    """loss_fn_given_preds (pretrain/pretrain_model.py:262-303).  `gather(y)` stands for
    jax.lax.all_gather(y, 'batch').reshape(-1, H); identity when there is one replica."""
    gather = gather or (lambda y: y)
    info = {}
    for name, d in preds.items():
        x, y = d["x"].float(), d["y"].float()
        numer = (x * y).sum(-1)
        info[name] = torch.zeros(())
        if "_sources" in d:
            for k in ("text2audio", "audio2text", "random_text"):
                info[f"_{name}_from_{k}"] = torch.zeros(())
        for a, b in ("xy", "yx"):
            xa, yb = d[a].float(), d[b].float()
            if f"{b}_extra" in d:
                yb = torch.cat([yb, d[f"{b}_extra"].float()], 0)
            logits = xa @ gather(yb).T
            lse = torch.logsumexp(logits, -1)
            info[name] = info[name] + (lse - numer).mean() / 2.0
            if "_sources" in d:
                for i, k in enumerate(("text2audio", "audio2text", "random_text")):
                    m = (d["_sources"] == i).float()
                    info[f"_{name}_from_{k}"] = info[f"_{name}_from_{k}"] + ((lse - numer) * m).sum() / (m.sum() + 1e-5) / 2.0
    return sum(v for k, v in info.items() if not k.startswith("_")), info
