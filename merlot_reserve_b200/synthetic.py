"""Deterministic synthetic weights and inputs (numpy only) -- SURVEY.md §8(d).

No dataset or checkpoint can be fetched in the build environment, so tests and benchmarks run on
random-init weights of the named architecture and synthetic clips of the reference's shapes.

* weights (seed 0): Dense kernels follow the reference's `kernel_init` (mreserve/modeling.py:147-186:
  truncated normal(+-2) * min(18 / fan_in, 0.02) / sqrt(2), with its fan_in rule for 3-D kernels);
  attention-pool projections use lecun-normal (flax default); biases / LayerNorm parameters / cls are
  deliberately NON-trivial (N(0, 0.02), scale 1 + N(0, 0.02)) so that no bias or scale path can be
  dropped silently; embedding N(0, 0.02); contrastive_scales = log(20).
* clip (seeds 1..3): images U[0,1) [nseg, h*w, 768]; audio |N(0,1)|*1.5 with column 64 == 1.0
  (playback speed, mreserve/preprocess.py:141-144); tokens / subseg_idxs laid out like
  `preprocess_video` (mreserve/preprocess.py:497-540): segment 0 is text ending with MASK, the other
  segments are 18 x AUDIOSPAN with subseg = 3 i + k // 6, truncated / padded to 160 with (0, -1).
"""
from __future__ import annotations

import copy
import math
from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import yaml

MASK, AUDIOSPAN = 3, 5
VOCAB = 32768
_CFG_DIR = Path(__file__).resolve().parent / "configs"


def load_config(name: str = "base", output_grid: Tuple[int, int] | None = None, **model_overrides) -> Dict:
    with open(_CFG_DIR / f"{name}.yaml") as f:
        cfg = yaml.safe_load(f)
    if output_grid is not None:
        cfg["model"]["output_grid"] = list(output_grid)
    cfg["model"].update(model_overrides)
    return cfg


def tiny_config(hidden_size=128, layers=2, output_grid=(4, 6), span_layers=2) -> Dict:
    """A small geometry with the same structure, for tests the CPU oracle finishes in seconds."""
    cfg = load_config("base", output_grid)
    cfg["model"].update(hidden_size=hidden_size, joint_num_layers=layers, audio_num_layers=layers,
                        vit_num_layers=layers, span_num_layers=span_layers)
    return cfg


def _trunc_normal(rng, shape):
    x = rng.standard_normal(shape, dtype=np.float32)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()), dtype=np.float32)
        bad = np.abs(x) > 2.0
    return x


def _kernel_init(rng, shape, scale=1.0):
    if len(shape) == 2:
        fan_in = shape[0]
    else:  # (H, heads, dh) or (heads, dh, H): modeling.py:173-180
        fan_in = shape[0]
        if shape[0] < shape[2]:
            fan_in *= shape[1]
    std = min(18.0 / fan_in, 0.02) / math.sqrt(2.0)
    return (_trunc_normal(rng, shape) * np.float32(std * scale)).astype(np.float32)


def _lecun(rng, shape, fan_in, scale=1.0):
    return (_trunc_normal(rng, shape) * np.float32(scale * math.sqrt(1.0 / fan_in) / 0.87962566)).astype(np.float32)


def _small(rng, shape):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)).astype(np.float32)


def _ln(rng, H):
    return {"scale": (1.0 + _small(rng, (H,))).astype(np.float32), "bias": _small(rng, (H,))}


def _encoder_params(rng, H, num_layers, add_cls, scale):
    nh = H // 64
    p = {"pre_ln": _ln(rng, H), "final_ln": _ln(rng, H)}
    if add_cls:
        p["cls"] = _small(rng, (H,))
        p["cls_proj"] = {"kernel": _kernel_init(rng, (H, H), scale), "bias": _small(rng, (H,))}
    for i in range(num_layers):
        p[f"layer_{i:02d}"] = {
            "pre_attn_ln": _ln(rng, H),
            "attention_layer": {
                "qkv": {"kernel": _kernel_init(rng, (H, 3 * nh, 64), scale), "bias": _small(rng, (3 * nh, 64))},
                "attn_proj": {"kernel": _kernel_init(rng, (nh, 64, H), scale)},
            },
            "pre_mlp_ln": _ln(rng, H),
            "mlp_layer": {
                "intermediate": {"kernel": _kernel_init(rng, (H, 4 * H), scale), "bias": _small(rng, (4 * H,))},
                "out": {"kernel": _kernel_init(rng, (4 * H, H), scale)},
            },
        }
    return p


def _pool_params(rng, H, scale):
    nh = H // 64
    p = {k: {"kernel": _lecun(rng, (H, nh, 64), H, scale), "bias": _small(rng, (nh, 64))}
         for k in ("query", "key", "value")}
    p["out"] = {"kernel": _lecun(rng, (nh, 64, H), H, scale), "bias": _small(rng, (H,))}
    return p


def synthetic_params(config: Dict, seed: int = 0, weight_scale: float = 1.0) -> Dict:
    """Parameter tree with the reference's flax names.  `weight_scale` multiplies every
    projection kernel; tests use > 1 so that attention is far from uniform."""
    m = config["model"]
    H = m["hidden_size"]
    rng = np.random.default_rng(seed)
    pp3 = (m["vit_patch_size"] ** 2) * 3
    s = weight_scale
    return {
        "vision_encoder": {
            "embedding": {"kernel": _kernel_init(rng, (pp3, H), s), "bias": _small(rng, (H,))},
            "transformer": _encoder_params(rng, H, m["vit_num_layers"], True, s),
            "seq_attnpool": _pool_params(rng, H, s),
        },
        "audio_encoder": {
            "embedding": {"kernel": _kernel_init(rng, (m["audio_patch_size"], 65, H), s), "bias": _small(rng, (H,))},
            "transformer": _encoder_params(rng, H, m["audio_num_layers"], True, s),
            "seq_attnpool": _pool_params(rng, H, s),
        },
        "span_encoder": {"transformer": _encoder_params(rng, H, m["span_num_layers"], True, s)},
        "joint_transformer": _encoder_params(rng, H, m["joint_num_layers"], False, s),
        "token_encoder": {"Embed_0": {"embedding": _small(rng, (VOCAB, H))}},
        "head": {"kernel": _kernel_init(rng, (H, H), s), "bias": _small(rng, (H,))},
        "contrastive_scales": np.full((3,), math.log(20.0), dtype=np.float32),
    }


class _ShapeOnly:
    """Stands in for an array in `param_shapes`."""

    def __init__(self, shape):
        self.shape = tuple(int(x) for x in shape)

    def astype(self, _):
        return self

    def __radd__(self, _):
        return self


def param_shapes(config: Dict) -> Dict[str, list]:
    """'a/b/c' -> shape of every leaf `synthetic_params(config)` would create, without drawing the values
    (the tree layout of SURVEY.md 8b; compared with the reference's own Module.init in tests/test_reference_goldens.py)."""
    g = globals()
    saved = {k: g[k] for k in ("_kernel_init", "_lecun", "_small")}
    g["_kernel_init"] = lambda rng, shape, scale=1.0: _ShapeOnly(shape)
    g["_lecun"] = lambda rng, shape, fan_in, scale=1.0: _ShapeOnly(shape)
    g["_small"] = lambda rng, shape: _ShapeOnly(shape)
    try:
        tree = synthetic_params(config)
    finally:
        g.update(saved)

    def flat(t, prefix=""):
        out = {}
        for k, v in t.items():
            if isinstance(v, dict):
                out.update(flat(v, f"{prefix}{k}/"))
            else:
                out[prefix + k] = [int(x) for x in v.shape]
        return out
    return flat(tree)


def synthetic_clip(config: Dict, nseg: int = 16, seed: int = 1, token_length: int = 160, text_len: int = 12) -> Dict:
    """One clip in the layout `MerlotReserve.embed_video` takes (modeling.py:806-826)."""
    m = config["model"]
    gh, gw = m["output_grid"]
    images = np.random.default_rng(1000003 * seed + 1).random((nseg, gh * gw, 768), dtype=np.float32)
    audio = np.abs(np.random.default_rng(1000003 * seed + 2).standard_normal(
        (3 * nseg, m["audio_seq_length"], 65), dtype=np.float32)) * np.float32(1.5)
    audio[..., 64] = 1.0
    rng = np.random.default_rng(1000003 * seed + 3)
    toks = list(rng.integers(10, VOCAB, size=text_len - 1)) + [MASK]
    sub = [0] * text_len
    for i in range(1, nseg):
        toks += [AUDIOSPAN] * 18
        sub += (3 * i + np.arange(18) // 6).tolist()
    toks, sub = toks[:token_length], sub[:token_length]
    while len(toks) < token_length:
        toks.append(0)
        sub.append(-1)
    return {"images": images, "audio_clips": audio.astype(np.float32),
            "tokens": np.asarray(toks, dtype=np.int32), "subseg_idxs": np.asarray(sub, dtype=np.int32)}


def synthetic_batch(config: Dict, batch: int, nseg: int = 16, seed: int = 1, **kw) -> Dict:
    clips = [synthetic_clip(config, nseg, seed + i, **kw) for i in range(batch)]
    return {k: np.stack([c[k] for c in clips], 0) for k in clips[0]}


def synthetic_spans(n: int, seed: int = 7, length: int = 15) -> np.ndarray:
    """Tokenised label-space options padded to 15 (PretrainedMerlotReserve.get_label_space, modeling.py:1024)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, length), dtype=np.int32)
    for i in range(n):
        k = int(rng.integers(1, length + 1))
        out[i, :k] = rng.integers(10, VOCAB, size=k)
    return out


def count_params(tree) -> int:
    if isinstance(tree, dict):
        return sum(count_params(v) for v in tree.values())
    return int(np.asarray(tree).size)


def round_tree_bf16(tree):
    """Round every leaf to the nearest bf16 value (kept as fp32) -- lets a test isolate kernel error
    from weight-quantisation error."""
    import torch

    if isinstance(tree, dict):
        return {k: round_tree_bf16(v) for k, v in tree.items()}
    return torch.from_numpy(np.asarray(tree)).to(torch.bfloat16).to(torch.float32).numpy()


def clone_tree(tree):
    return copy.deepcopy(tree)


# ------------------------------------------------------------------------------------------------
# pretraining batch (BASELINE.json configs[2]; SURVEY.md §8d cfg 3)
# ------------------------------------------------------------------------------------------------
MASKAUDIO, LTOVPOOL = 4, 6


def synthetic_pretrain_batch(config: Dict, batch: int, seed: int = 1) -> Dict:
    """One per-device pretraining batch in the layout `MerlotReservePretrainer.__call__` takes
    (pretrain/pretrain_model.py:38-93; shapes from pretrain/dataloader.py:545-700):

      images [b, nseg*hw, 768], audio_clips [b, nseg*3*60, 65], video_src_index [b, G, nseg/G],
      text2audio / audio2text (+ '/audio_ptr', '/text_ptr') [b, G, lang_seq_len],
      audio_text_matching (+ '/audio_ptr'), random_text (+ '/text_ptr') [b, 1, seq_len], text_spans [b, n_spans, 15].

    Per example the two text2audio rows carry 12 [MASK, MASKAUDIO] pairs (text spans 0..11), the two audio2text
    rows 12 MASKs (12..23), random_text 38 MASKs (24..61) -- the reference's counters (dataloader.py:557,589,647).
    The reference draws two things from a JAX PRNG inside the forward (pretrain_model.py:28,219); they are supplied
    as inputs instead: '_rng/split_audio2text', '_rng/split_text2audio' [b*G] (= `split_from_here`) and
    '_rng/gumbel_u' [b, n_spans] (the uniform draw of the Gumbel top-k)."""
    m, d = config["model"], config["data"]
    gh, gw = m["output_grid"]
    hw = gh * gw
    nseg, G, nsub = d["num_segments"], d["num_segment_groups"], d["num_audio_subsegments"]
    per_group = nseg // G
    n_aud = nseg * nsub
    Ltxt, Lseq, atl = d["lang_seq_len"], d["seq_len"], m["audio_token_length"]
    n_trg = int(n_aud * d["mask_rate"])                       # 12 masked audio spans per example and task
    budget = d["text_span_budget"]
    n_spans = 2 * n_trg + budget                               # 62
    rng = np.random.default_rng(7000003 * seed + 11)
    out = {k: [] for k in ("images", "audio_clips", "video_src_index", "text2audio", "text2audio/audio_ptr",
                           "text2audio/text_ptr", "audio2text", "audio2text/audio_ptr", "audio2text/text_ptr",
                           "audio_text_matching", "audio_text_matching/audio_ptr", "random_text",
                           "random_text/text_ptr", "text_spans")}

    def pad_rows(rows, length):
        toks = np.zeros((length,), np.int32)
        ap = np.full((length,), -1, np.int32)
        tp = np.full((length,), -1, np.int32)
        rows = rows[:length]
        if rows:
            a = np.asarray(rows, dtype=np.int32)
            toks[:len(rows)], ap[:len(rows)], tp[:len(rows)] = a[:, 0], a[:, 1], a[:, 2]
        return toks, ap, tp

    for _ in range(batch):
        out["images"].append(rng.random((nseg * hw, 768), dtype=np.float32))
        a = np.abs(rng.standard_normal((n_aud * m["audio_seq_length"], 65), dtype=np.float32)) * np.float32(1.5)
        a[:, 64] = 1.0
        out["audio_clips"].append(a)
        vsrc = np.ones((G, per_group), np.int32)
        vsrc[-1, per_group // 2 + 1:] = 2                       # two videos packed in the last group
        out["video_src_index"].append(vsrc)

        t2a = {k: [] for k in ("t", "a", "p")}
        a2t = {k: [] for k in ("t", "a", "p")}
        per_g = n_trg // G
        for g in range(G):
            subs = np.arange(g * per_group * nsub, (g + 1) * per_group * nsub)
            # --- text2audio: text everywhere, masked spans become [MASK, MASKAUDIO] ---
            masked = set(rng.choice(subs, size=per_g, replace=False).tolist())
            rows, k = [], g * per_g
            for s in subs:
                if s in masked:
                    rows += [(MASK, s, k), (MASKAUDIO, s, k)]
                    k += 1
                else:
                    rows += [(int(t), s, -1) for t in rng.integers(10, VOCAB, size=int(rng.integers(3, 7)))]
            t, ap, tp = pad_rows(rows, Ltxt)
            t2a["t"].append(t), t2a["a"].append(ap), t2a["p"].append(tp)
            # --- audio2text: audio spans, a few textified neighbours, masked spans become MASK ---
            masked = set(rng.choice(subs, size=per_g, replace=False).tolist())
            rows, k = [], n_trg + g * per_g
            for s in subs:
                if s in masked:
                    rows.append((MASK, s, k))
                    k += 1
                elif rng.random() < 0.25:
                    rows += [(int(t), s, -1) for t in rng.integers(10, VOCAB, size=int(rng.integers(3, 7)))]
                else:
                    rows += [(AUDIOSPAN, s, -1)] * atl
            t, ap, tp = pad_rows(rows, Ltxt)
            a2t["t"].append(t), a2t["a"].append(ap), a2t["p"].append(tp)
        for name, dct in (("text2audio", t2a), ("audio2text", a2t)):
            out[name].append(np.stack(dct["t"])), out[name + "/audio_ptr"].append(np.stack(dct["a"]))
            out[name + "/text_ptr"].append(np.stack(dct["p"]))

        # --- audio_text_matching: aux text, then per segment LTOVPOOL + 3 x 6 AUDIOSPAN (dataloader.py:605-620) ---
        rows = [(int(t), -1, -1) for t in rng.integers(10, VOCAB, size=int(rng.integers(20, 60)))]
        for i in range(nseg):
            rows.append((LTOVPOOL, i * nsub, -1))
            for j in range(nsub):
                rows += [(AUDIOSPAN, i * nsub + j, -1)] * atl
        t, ap, _ = pad_rows(rows, Lseq)
        out["audio_text_matching"].append(t[None]), out["audio_text_matching/audio_ptr"].append(ap[None])

        # --- random_text: `budget` MASKs spread through plain text ---
        rows, k = [], 2 * n_trg
        n_txt = min(Lseq - 8, 10 * budget)
        mask_at = set(rng.choice(np.arange(2, n_txt), size=budget, replace=False).tolist())
        for i in range(n_txt):
            if i in mask_at:
                rows.append((MASK, -1, k))
                k += 1
            else:
                rows.append((int(rng.integers(10, VOCAB)), -1, -1))
        t, _, tp = pad_rows(rows, Lseq)
        out["random_text"].append(t[None]), out["random_text/text_ptr"].append(tp[None])

        spans = np.zeros((n_spans, m["text_span_length"]), np.int32)
        for i in range(n_spans):
            n = int(rng.integers(1, m["text_span_length"] + 1))
            spans[i, :n] = rng.integers(10, VOCAB, size=n)
        spans[int(rng.integers(2 * n_trg, n_spans))] = 0        # one empty span: exercises the is_valid filter
        out["text_spans"].append(spans)

    res = {k: np.stack(v, 0) for k, v in out.items()}
    split = np.full((2, batch * G), per_group, np.int32)        # per_group == "no split" (pretrain_model.py:31-33)
    split[0, 0] = per_group // 2
    split[1, -1] = 2
    res["_rng/split_audio2text"], res["_rng/split_text2audio"] = split[0], split[1]
    res["_rng/gumbel_u"] = rng.random((batch, n_spans), dtype=np.float32) * np.float32(0.999) + np.float32(5e-4)
    return res


def synthetic_tvqa_batch(config: Dict, batch: int, num_segments: int = 7, num_answers: int = 5, seq_len: int = 160,
                         seed: int = 1) -> Dict:
    """A batch in the layout `MerlotReserveTVQA.__call__` takes (finetune/tvqa/tvqa_finetune.py:163-258;
    finetune/common_dataloader.py:140-230): per example `num_segments` frames, num_segments x 3 audio sub-segments,
    and for each of `num_answers` candidates two token sequences of (token, segment) pairs -- subtitles as text
    ('textonly_seqs') or as AUDIOSPAN placeholders ('audio_seqs') -- each = question + candidate ending in MASK,
    followed by the per-segment material (segment indices 1..num_segments), padded with (0, -1)."""
    m = config["model"]
    gh, gw = m["output_grid"]
    rng = np.random.default_rng(7000003 * seed + 11)
    images = rng.random((batch, num_segments, gh * gw, 768), dtype=np.float32)
    audio = np.abs(rng.standard_normal((batch, num_segments, 3, m["audio_seq_length"], 65), dtype=np.float32)) * np.float32(1.5)
    audio[..., 64] = 1.0
    text = np.zeros((batch, num_answers, seq_len, 2), np.int32)
    aud = np.zeros((batch, num_answers, seq_len, 2), np.int32)
    text[..., 1] = -1
    aud[..., 1] = -1
    for b in range(batch):
        question = rng.integers(10, VOCAB, size=int(rng.integers(4, 9)))
        subs = [rng.integers(10, VOCAB, size=int(rng.integers(2, 7))) for _ in range(num_segments)]
        for a in range(num_answers):
            head = list(question) + list(rng.integers(10, VOCAB, size=int(rng.integers(2, 6)))) + [MASK]
            t_tok, t_seg = list(head), [0] * len(head)
            a_tok, a_seg = list(head), [0] * len(head)
            for i in range(num_segments):
                t_tok += list(subs[i])
                t_seg += [i + 1] * len(subs[i])
                a_tok += [AUDIOSPAN] * 18          # 3 sub-segments x audio_token_length 6
                a_seg += [i + 1] * 18
            for arr, tok, seg in ((text, t_tok, t_seg), (aud, a_tok, a_seg)):
                n = min(seq_len, len(tok))
                arr[b, a, :n, 0] = tok[:n]
                arr[b, a, :n, 1] = seg[:n]
    return {"images": images, "audio_clips": audio, "textonly_seqs": text, "audio_seqs": aud}
