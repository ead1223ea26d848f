"""B200-native drop-in for the forward + loss of `pretrain/pretrain_model.py`.

`MerlotReservePretrainer.__call__(batch)` mirrors the reference class of the same name (pretrain_model.py:8-259):
same batch keys and shapes, same three output dicts (`imgs_to_audio`, `text_to_audio`, `stuff_to_span`, each with
unit-norm, temperature-scaled bf16 `x` / `y` [/ `y_extra`, `_sources`]).  `loss_fn_given_preds(preds)` mirrors
pretrain_model.py:262-303; with `torch.distributed` initialised it all-gathers the contrastive targets over the
data-parallel group exactly where the reference calls `jax.lax.all_gather(y, 'batch')` (one packed NCCL all-gather
for all six target sets).

The reference draws two random quantities from a JAX PRNG inside the forward (`split_from_here`,
pretrain_model.py:28, and the Gumbel noise of :219).  They can be supplied in the batch ('_rng/split_audio2text',
'_rng/split_text2audio', '_rng/gumbel_u' -- what the parity tests do); otherwise they are drawn on the host from a
generator seeded like the reference seeds its key (the sum of `audio2text/text_ptr`, :95).

Everything that touches activations runs in the sm_100a kernels of libmreserve_b200; torch is used for memory,
integer index preparation (cumsum / sort / top-k of pointers and counts) and the collective.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from .modeling import LTOVPOOL, MASK, MASKAUDIO, PADDING, MerlotReserve, _dev


class MerlotReservePretrainer(MerlotReserve):
    """pretrain_model.py:8-259."""

    @staticmethod
    def _augment_video_src_idx(video_src_idx: torch.Tensor, split_from_here: torch.Tensor) -> torch.Tensor:
        """pretrain_model.py:9-36 (the categorical draw `split_from_here` [B] in 1..L is an input)."""
        B, Lv = video_src_idx.shape
        if Lv == 1:
            return video_src_idx
        split_mask = split_from_here[:, None] <= torch.arange(Lv, device=video_src_idx.device)[None]
        return torch.where(split_mask, video_src_idx + 4 * Lv, video_src_idx)

    def _draw_rng(self, batch: Dict, bG: int, per_group: int, b: int, n_spans: int, dev) -> Dict[str, torch.Tensor]:
        need = [k for k in ("_rng/split_audio2text", "_rng/split_text2audio", "_rng/gumbel_u") if k not in batch]
        if not need:
            return {k: _dev(batch[k]) for k in ("_rng/split_audio2text", "_rng/split_text2audio", "_rng/gumbel_u")}
        seed = int(np.asarray(torch.as_tensor(batch["audio2text/text_ptr"]).cpu()).astype(np.uint32).sum())  # :95
        rng = np.random.default_rng(seed)
        p_split = self.config.get("_augment_video_src_idx_prob", 0.1)
        probs = [p_split / max(per_group - 1, 1)] * (per_group - 1) + [1 - p_split]
        out = {}
        for k in ("_rng/split_audio2text", "_rng/split_text2audio"):
            out[k] = _dev(batch[k]) if k in batch else torch.from_numpy(
                (1 + rng.choice(per_group, size=bG, p=np.asarray(probs))).astype(np.int32)).to(dev)
        k = "_rng/gumbel_u"
        out[k] = _dev(batch[k]) if k in batch else torch.from_numpy(
            rng.uniform(1e-9, 1.0, size=(b, n_spans)).astype(np.float32)).to(dev)
        return out

    def __call__(self, batch: Dict) -> Dict[str, Dict[str, torch.Tensor]]:
        p, H = self._p, self.hidden_size
        data = self.config["data"]
        G, nsub = data["num_segment_groups"], data["num_audio_subsegments"]
        Ltxt, Lseq = data["lang_seq_len"], data["seq_len"]
        n_a2t, n_t2a = data["num_audio2text_seqs"], data["num_text2audio_seqs"]
        gh, gw = self.output_grid_h, self.output_grid_w
        hw = gh * gw
        atl = self.config["audio_token_length"]

        images = _dev(batch["images"], torch.float32)
        dev = images.device
        b, nseg_hw, pp3 = images.shape
        nseg = nseg_hw // hw
        per_group = nseg // G
        n_aud = nseg * nsub
        i32 = lambda k, cols: _dev(batch[k], torch.int32).reshape(-1, cols).contiguous()

        # --- encoders (pretrain_model.py:53-74) ---
        vis = self._vision(images.reshape(b * nseg, hw, pp3), want_cls=True)
        hw1 = hw // 4
        vis_len = per_group * hw1
        imgs_seq = vis["seq_attnpool"].reshape(b * G, vis_len, H)           # row = (example, group)
        aud = self._audio(_dev(batch["audio_clips"], torch.float32).reshape(b * n_aud, self.config["audio_seq_length"], -1),
                          want_cls=True)
        audio_seq = aud["seq_attnpool"].reshape(b, n_aud, atl, H)
        audio_cls = aud["cls"]                                              # [b * n_aud, H]

        tok = {k: i32(k, Ltxt) for k in ("text2audio", "text2audio/audio_ptr", "text2audio/text_ptr", "audio2text",
                                         "audio2text/audio_ptr", "audio2text/text_ptr")}
        tok.update({k: i32(k, Lseq) for k in ("random_text", "random_text/text_ptr", "audio_text_matching",
                                              "audio_text_matching/audio_ptr")})
        text_spans = i32("text_spans", self.config["text_span_length"])
        n_spans = text_spans.shape[0] // b
        vsi = _dev(batch["video_src_index"], torch.int64).reshape(b, G, per_group)
        rng = self._draw_rng(batch, b * G * max(n_a2t, n_t2a), per_group, b, n_spans, dev)

        # --- joint inputs, assembled straight into one [sum(rows), seq_len, H] batch in sorted-key order
        #     (pretrain_model.py:99-144): audio2text, audio_text_matching, random_text, text2audio ---
        rows = {"audio2text": b * G * n_a2t, "audio_text_matching": b, "random_text": b, "text2audio": b * G * n_t2a}
        keys = sorted(rows)
        R = sum(rows.values())
        x = torch.empty((R * Lseq, H), dtype=torch.bfloat16, device=dev)
        rot = torch.empty((R * Lseq, 32), dtype=torch.float32, device=dev)
        valid = torch.empty((R * Lseq,), dtype=torch.uint8, device=dev)
        group = torch.zeros((R, Lseq), dtype=torch.int32, device=dev)
        start, r0 = {}, 0
        for k in keys:
            start[k] = r0
            r0 += rows[k]

        def views(k):
            a, n = start[k] * Lseq, rows[k] * Lseq
            return x[a:a + n], rot[a:a + n], valid[a:a + n]

        hp, wp = gh // 2, gw // 2
        vis_seg = torch.arange(per_group, device=dev).repeat_interleave(hw1)

        def video_src(seg_idx, n_seqs, split):
            """prepare_multimodal_inputs' video_src (modeling.py:743-756): text rows, vision rows, padding = -1."""
            v = vsi.repeat(1, n_seqs, 1).reshape(-1, per_group)
            v = self._augment_video_src_idx(v, split.to(torch.int64))
            Bk = v.shape[0]
            parts = [torch.gather(v, 1, seg_idx.long()), torch.gather(v, 1, vis_seg[None].expand(Bk, -1))]
            pad = Lseq - Ltxt - vis_len
            if pad > 0:
                parts.append(torch.full((Bk, pad), -1, dtype=torch.int64, device=dev))
            return torch.cat(parts, 1).to(torch.int32)

        for k, n_seqs, use_audio in (("audio2text", n_a2t, True), ("text2audio", n_t2a, False)):
            ptr = tok[k + "/audio_ptr"]
            seg = torch.remainder(torch.div(ptr, nsub, rounding_mode="floor"), per_group).to(torch.int32).contiguous()
            L.embed_assemble(tok[k], p.tok_emb, H=H, seg_idx=seg, aud_ptr=ptr if use_audio else None,
                             audio=audio_seq if use_audio else None, vision=imgs_seq,
                             n_aud=n_aud if use_audio else 0, n_vis=vis_len, pool_h=hp, pool_w=wp, T=Lseq,
                             audio_token_length=atl, aud_div=G * n_seqs, vis_group=G, vis_tile=n_seqs, out=views(k))
            group[start[k]:start[k] + rows[k]] = video_src(seg, n_seqs, rng["_rng/split_" + k])
        atm = tok["audio_text_matching"]
        is_pool = atm == LTOVPOOL
        pool_cum = torch.cumsum(is_pool.to(torch.int32), -1).to(torch.int32)
        L.embed_assemble(atm, p.tok_emb, H=H, seg_idx=pool_cum.contiguous(), aud_ptr=tok["audio_text_matching/audio_ptr"],
                         audio=audio_seq, n_aud=n_aud, T=Lseq, audio_token_length=atl, out=views("audio_text_matching"))
        L.embed_assemble(tok["random_text"], p.tok_emb, H=H, T=Lseq, out=views("random_text"))

        # --- joint transformer, final LN + joint_proj only on the rows that are pooled from (:150-155) ---
        xe = self._encoder(x, p.joint, R, Lseq, rot, valid=valid, group=group.reshape(-1))

        def projected(k, keep):
            a = start[k] * Lseq
            h = L.layernorm(xe[a:a + rows[k] * Lseq], *p.joint.final_ln, rows=rows[k] * keep, seq_in=Lseq, seq_out=keep)
            return L.gemm(h, p.head[0], bias=p.head[1], out_f32=True).view(rows[k], keep, H)

        mo = {"audio2text": projected("audio2text", Ltxt), "text2audio": projected("text2audio", Ltxt),
              "audio_text_matching": projected("audio_text_matching", Lseq),
              "random_text": projected("random_text", Lseq)}

        # --- pools (one_hot_pool as segment gathers; :160-204) ---
        def pool(do_pool, idx, v, S, acc=None):
            l2 = (v.shape[0] * v.shape[1]) // b
            return L.segment_pool(do_pool.to(torch.uint8).reshape(b, l2).contiguous(), idx.reshape(b, l2).contiguous(),
                                  v.reshape(b, l2, H), S, accumulate_into=acc)

        a2v, _ = pool(is_pool, pool_cum - 1, mo["audio_text_matching"], nseg)
        t2a_x, is_selected = pool(tok["text2audio"] == MASKAUDIO, tok["text2audio/audio_ptr"], mo["text2audio"], n_aud)
        n_trg = int(n_aud * data["mask_rate"]) * n_t2a
        idx_sort = torch.sort(-is_selected, dim=-1, stable=True)[1]                                  # :181
        base = (torch.arange(b, device=dev) * n_aud)[:, None]
        best_rows = (idx_sort[:, :n_trg] + base).reshape(-1).to(torch.int32).contiguous()
        extra_rows = (idx_sort[:, n_trg:] + base).reshape(-1).to(torch.int32).contiguous()

        counts = {}
        sel = None
        for k in ("text2audio", "audio2text", "random_text"):
            if sel is None:
                sel = pool(tok[k] == MASK, tok[k + "/text_ptr"], mo[k], n_spans)
                counts[k] = sel[1].clone()
            else:
                before = sel[1].clone()
                pool(tok[k] == MASK, tok[k + "/text_ptr"], mo[k], n_spans, acc=sel)
                counts[k] = sel[1] - before
        t2sp_x, t2sp_ct = sel
        src = torch.stack([torch.zeros_like(t2sp_ct), counts["text2audio"], counts["audio2text"],
                           counts["random_text"]], -1).argmax(-1) - 1                                 # :208-209
        is_valid = (text_spans != PADDING).any(-1).reshape(b, n_spans) & (t2sp_ct > 0)                # :212-214
        logits_for_pred = is_valid.float() * 1e6 + math.log(4.0) * (counts["text2audio"] + counts["audio2text"])
        z = -torch.log(-torch.log(rng["_rng/gumbel_u"].float()))
        n_inc = data["num_text_spans_to_include"]
        if n_inc > n_spans:
            raise ValueError("num_text_spans_to_include exceeds the number of text spans (pretrain_model.py:224)")
        best = torch.sort(-(logits_for_pred + z).reshape(-1), stable=True)[1][:n_inc * b]             # :227
        best32 = best.to(torch.int32).contiguous()
        sp_cls = self._span(text_spans[best].contiguous())                                           # :232-233

        # --- unit_normalize * exp(log_scale / 2), bf16 (:236-257); selections are fused as row gathers ---
        ls = np.minimum(p.contrastive_scales.astype(np.float32), np.float32(math.log(100.0)))
        t = [float(np.exp(v / 2.0)) for v in ls]
        un = lambda v, s, idx=None: L.unit_normalize(v.reshape(-1, H), scale=s, out_f32=False, row_idx=idx)
        return {
            "imgs_to_audio": {"x": un(a2v, t[0]), "y": un(vis["cls"], t[0])},
            "text_to_audio": {"x": un(t2a_x, t[1], best_rows), "y": un(audio_cls, t[1], best_rows),
                              "y_extra": un(audio_cls, t[1], extra_rows)},
            "stuff_to_span": {"x": un(t2sp_x, t[2], best32), "y": un(sp_cls, t[2]),
                              "_sources": src.reshape(-1)[best].to(torch.int32).contiguous()},
        }


_PAIRS = (("x", "y"), ("y", "x"))


def gather_contrastive_targets(preds: Dict[str, Dict[str, torch.Tensor]], group=None, comm=None):
    """The exchange step of the path (pretrain_model.py:284-290): for every (head, direction) pair the targets
    `y` (+ `y_extra`) of ALL data-parallel ranks.  The six target sets of a rank are packed into one buffer and
    moved by ONE all-gather; each returned `y_all` is rank-major, i.e. `all_gather(y, 'batch').reshape(-1, H)`.
    Pure tensor plumbing (works on any device / backend); with `comm` (an _lib.Communicator) the all-gather is the
    library's own mr_all_gather on the current stream instead of torch.distributed's.  Returns [(head, x, pos_off, y_all)]: row i of x has its
    positive at row pos_off + i of y_all (this rank's block of the rank-major gather)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    jobs = []
    for name, d in preds.items():
        for k1, k2 in _PAIRS:
            ys = [d[k2]] + ([d[k2 + "_extra"]] if (k2 + "_extra") in d else [])
            jobs.append((name, d[k1], d[k2], ys))
    H = jobs[0][1].shape[-1]
    sizes = [sum(y.shape[0] for y in ys) for _, _, _, ys in jobs]
    packed = torch.cat([y for _, _, _, ys in jobs for y in ys], 0).contiguous()
    if world > 1:
        gathered = torch.empty((world, packed.shape[0], H), dtype=packed.dtype, device=packed.device)
        if comm is not None:
            comm.all_gather(packed, gathered)
        else:
            dist.all_gather_into_tensor(gathered.view(-1, H), packed, group=group)
    else:
        gathered = packed[None]
    out, off = [], 0
    for (name, x, _, _), n in zip(jobs, sizes):
        out.append((name, x, rank * n, gathered[:, off:off + n].reshape(world * n, H)))
        off += n
    return out


def loss_fn_given_preds(preds: Dict[str, Dict[str, torch.Tensor]], group=None, comm=None):
    """pretrain_model.py:262-303 for the contrastive heads: per head and direction
    `mean(logsumexp(x . y_all^T) - sum(x * y)) / 2` plus the per-source breakdown of `stuff_to_span`.
    Returns (loss, loss_info) as 0-d CUDA tensors (this rank's loss; the reference pmean()s metrics later)."""
    info = {}
    for name, x, pos_off, y_all in gather_contrastive_targets(preds, group, comm=comm):
        lse, numer = L.contrastive_lse(x, y_all, pos_off)
        red = L.contrastive_reduce(lse, numer, preds[name].get("_sources"))
        info[name] = info.get(name, 0.0) + red[0] / 2.0
        if "_sources" in preds[name]:
            for i, k in enumerate(("text2audio", "audio2text", "random_text")):
                key = f"_{name}_from_{k}"
                info[key] = info.get(key, 0.0) + red[1 + i] / 2.0
    loss = sum(v for k, v in info.items() if not k.startswith("_"))
    return loss, info
