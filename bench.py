#!/usr/bin/env python
"""Benchmark of the MERLOT Reserve encoder hot path on B200 (BASELINE.json metric: segments/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config base|large] [--clips B]

One "step" = one pass of `MerlotReserve.batch_embed_video` (vision + audio + joint encoders + head +
unit_normalize) over B synthetic clips of 16 segments; a segment = one 192x320 frame (12x20 patches) plus
its 3 audio sub-segments.  N=1 runs BASELINE.json configs[1] (Reserve-Base, 64 clips, bf16).  N>1 (launched
by torchrun, one rank per GPU) runs independent replicas on disjoint clips -- the path shards by clip with
no data-path collective ("weak" scaling); the only communication is the timing barrier / max.

JSON line keys: see the task contract.  `value` is timed with inputs resident in HBM, `e2e` through the
same public call with pinned HOST buffers (H2D of the inputs and D2H of the [B,160,H] result inside the
timed region).  `roofline` describes the dominant kernel class (the tcgen05 GEMM), timed live with CUDA
events around every launch of a separate instrumented step sequence.  `cpu_baseline` / `--impl reference`
time the CPU oracle (torch fp32 restatement of the reference -- the reference's own JAX path cannot be
installed in this environment) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

NSEG = 16
METRIC, UNIT = "segments/sec", "segments/s"


def _enc_flops(T, layers, H):
    return layers * (24.0 * T * H * H + 4.0 * T * T * H)


def flops_vision_frame(cfg):
    m = cfg["model"]
    H = m["hidden_size"]
    hw = m["output_grid"][0] * m["output_grid"][1]
    return _enc_flops(hw + 1, m["vit_num_layers"], H) + 2 * hw * 768 * H + 2 * hw * H * H * 2 + 2 * (hw // 4) * H * H * 2


def flops_audio_subseg(cfg):
    m = cfg["model"]
    H = m["hidden_size"]
    return _enc_flops(31, m["audio_num_layers"], H) + 2 * 30 * 130 * H + 2 * 30 * H * H * 2 + 2 * 6 * H * H * 2


def flops_mc_example(cfg, n_cand=4, L=144):
    """cfg 5 (finetune/vcr/qa_qar_joint_finetune.py:150-170): one frame, n_cand joint sequences of L + hw/4 rows."""
    m = cfg["model"]
    H = m["hidden_size"]
    hw = m["output_grid"][0] * m["output_grid"][1]
    return flops_vision_frame(cfg) + n_cand * _enc_flops(L + hw // 4, m["joint_num_layers"], H)


def flops_pretrain_example(cfg):
    """cfg 3 forward (pretrain_model.py:38-259): 16 frames, 48 audio sub-segments, 6 joint sequences of seq_len
    rows, joint_proj on the pooled rows, 48 spans through the span encoder."""
    m, d = cfg["model"], cfg["data"]
    H = m["hidden_size"]
    G = d["num_segment_groups"]
    n_joint = G * (d["num_audio2text_seqs"] + d["num_text2audio_seqs"]) + 2
    joint = n_joint * _enc_flops(d["seq_len"], m["joint_num_layers"], H)
    span = d["num_text_spans_to_include"] * (_enc_flops(16, m["span_num_layers"], H) + 2 * H * H)
    return (d["num_segments"] * flops_vision_frame(cfg) + 3 * d["num_segments"] * flops_audio_subseg(cfg) + joint + span
            + 2.0 * (2 * G * d["lang_seq_len"] + 2 * d["seq_len"]) * H * H)


def flops_per_clip(cfg, nseg=NSEG, L=160):
    """Algorithmic FLOPs of embed_video per clip: 24 T H^2 + 4 T^2 H per layer, plus
    patch / audio embeds, the two attention pools and the head."""
    m = cfg["model"]
    H = m["hidden_size"]
    gh, gw = m["output_grid"]
    hw = gh * gw

    def enc(T, layers):
        return layers * (24 * T * H * H + 4 * T * T * H)

    vis = nseg * (enc(hw + 1, m["vit_num_layers"]) + 2 * hw * 768 * H + 2 * hw * H * H * 2 + 2 * (hw // 4) * H * H * 2)
    aud = 3 * nseg * (enc(31, m["audio_num_layers"]) + 2 * 30 * 130 * H + 2 * 30 * H * H * 2 + 2 * 6 * H * H * 2)
    joint = enc(L + nseg * hw // 4, m["joint_num_layers"]) + 2 * L * H * H
    return float(vis + aud + joint)


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU oracle arm
# ------------------------------------------------------------------------------------------------
def cpu_oracle_segments_per_s(cfg, params, steps: int, warmup: int, nseg_sample: int):
    """Times the CPU oracle's embed_video on ONE clip of `nseg_sample` segments per step."""
    from merlot_reserve_b200.synthetic import synthetic_clip
    from oracle.reserve_oracle import OracleMerlotReserve
    torch.set_num_threads(os.cpu_count() or 1)
    orc = OracleMerlotReserve(cfg, params)
    clip = synthetic_clip(cfg, nseg=nseg_sample, seed=1)
    with torch.no_grad():
        for _ in range(warmup):
            orc.embed_video(**clip)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.embed_video(**clip)
        dt = (time.perf_counter() - t0) / steps
    return nseg_sample / dt, dt, torch.get_num_threads()


def workload_config(args, world: int):
    """The `config` object of the JSON line (shared by both arms so that the driver compares like with like)."""
    from merlot_reserve_b200.synthetic import load_config
    cfg = load_config(args.config, tuple(args.grid) if args.grid else None)
    gh, gw = cfg["model"]["output_grid"]
    B = args.clips
    return {"workload": f"reserve-{args.config} zero-shot embed_video (vision+audio+joint encoders, head, "
                        f"unit_normalize), {B} clips x 16 segments ({gh}x{gw} patches) + 48 audio sub-segments, "
                        f"L=160, per GPU",
            "clips_per_gpu": B, "segments_per_step_per_gpu": B * NSEG, "parallelism": f"replicas x{world}",
            "l2": "inputs and activations of a step exceed the 126 MB L2 many times over (fp32 patches "
                  "alone are 755 MB at 64 Base clips)"}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path cannot be installed here (jax / flax absent, no
    setup.py), so this times the CPU oracle port of it on all host cores, on a bounded sample (1 clip per step) of
    the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "embed_video":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm covers the embed_video workload"}))
        return
    from merlot_reserve_b200.synthetic import load_config, synthetic_params
    cfg = load_config(args.config, tuple(args.grid) if args.grid else None)
    params = synthetic_params(cfg, seed=0)
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    val, dt, cores = cpu_oracle_segments_per_s(cfg, params, steps, warmup, NSEG)
    sample = (f"CPU oracle (torch fp32 restatement of the reference; jax/flax not installable here): "
              f"embed_video on 1 clip x {NSEG} segments per step (of the {args.clips} clips of the workload), {steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_workload(args, rank, dev):
    """Returns the pieces of one benchmark workload: a resident-input step, an end-to-end (pinned host input) step,
    units and algorithmic FLOPs per step, and the description that goes into `config`."""
    from merlot_reserve_b200 import MerlotReserve, MerlotReservePretrainer, loss_fn_given_preds
    from merlot_reserve_b200.synthetic import (load_config, synthetic_batch, synthetic_params,
                                               synthetic_pretrain_batch)
    grid = tuple(args.grid) if args.grid else None
    cfg = load_config(args.config, grid)
    H = cfg["model"]["hidden_size"]
    gh, gw = cfg["model"]["output_grid"]
    params = synthetic_params(cfg, seed=0)
    pin = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in d.items()}
    w = {"cfg": cfg, "params": params, "metric": METRIC, "unit": UNIT}
    if args.workload == "embed_video":
        B = args.clips
        model = MerlotReserve.from_config(cfg).bind(params)
        host = pin(synthetic_batch(cfg, B, nseg=NSEG, seed=1 + 1000 * rank))      # disjoint clips per rank
        resident = {k: v.to(dev) for k, v in host.items()}
        out_host = torch.empty((B, 160, H), dtype=torch.float32).pin_memory()

        def step_e2e():
            out = model.batch_embed_video(**host)             # H2D of every input happens inside the call
            out_host.copy_(out, non_blocking=True)            # D2H of the step's result
            torch.cuda.current_stream().synchronize()

        w.update(step=lambda: model.batch_embed_video(**resident), step_e2e=step_e2e, units=B * NSEG,
                 flops=flops_per_clip(cfg) * B, h2d=sum(v.numel() * v.element_size() for v in host.values()),
                 d2h=out_host.numel() * 4,
                 workload=f"reserve-{args.config} zero-shot embed_video (vision+audio+joint encoders, head, "
                          f"unit_normalize), {B} clips x 16 segments ({gh}x{gw} patches) + 48 audio sub-segments, "
                          f"L=160, per GPU",
                 extra={"clips_per_gpu": B, "segments_per_step_per_gpu": B * NSEG})
    elif args.workload == "mc_scoring":
        B, n_cand, Ltok = args.clips, 4, 144
        model = MerlotReserve.from_config(cfg).bind(params)
        rng = np.random.default_rng(5 + rank)
        toks = rng.integers(10, 32768, size=(B, n_cand, Ltok)).astype(np.int32)
        toks[:, :, -1] = 3
        host = pin({"images": rng.random((B, gh * gw, 768), dtype=np.float32), "tokens": toks})
        proj = (rng.standard_normal(H) * 0.05).astype(np.float32)
        resident = {k: v.to(dev) for k, v in host.items()}
        out_host = torch.empty((B, n_cand), dtype=torch.float32).pin_memory()

        def step_e2e():
            out = model.score_multiple_choice(host["images"].to(dev, non_blocking=True),
                                              host["tokens"].to(dev, non_blocking=True), proj)
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        w.update(step=lambda: model.score_multiple_choice(resident["images"], resident["tokens"], proj),
                 step_e2e=step_e2e, units=B, flops=flops_mc_example(cfg, n_cand, Ltok) * B,
                 h2d=sum(v.numel() * v.element_size() for v in host.values()), d2h=out_host.numel() * 4,
                 metric="examples/sec", unit="examples/s",
                 workload=f"reserve-{args.config} VCR-style multiple-choice scoring (MerlotReserveVCR forward): batch {B}, "
                          f"1 frame ({gh}x{gw} patches) x {n_cand} candidates, L={Ltok}, proj head, per GPU",
                 extra={"examples_per_gpu": B})
    elif args.workload == "pretrain_fwd":
        b = args.clips
        model = MerlotReservePretrainer.from_config(cfg).bind(params)
        host = pin(synthetic_pretrain_batch(cfg, b, seed=1 + 1000 * rank))
        resident = {k: v.to(dev) for k, v in host.items()}
        out_host = torch.empty((1,), dtype=torch.float32).pin_memory()

        def fwd(batch):
            loss, _ = loss_fn_given_preds(model(dict(batch)))    # all-gathers the targets when world > 1
            return loss

        def step_e2e():
            out_host.copy_(fwd({k: v.to(dev, non_blocking=True) for k, v in host.items()}).reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        w.update(step=lambda: fwd(resident), step_e2e=step_e2e, units=b, flops=flops_pretrain_example(cfg) * b,
                 h2d=sum(v.numel() * v.element_size() for v in host.values()), d2h=4,
                 metric="examples/sec (forward + contrastive loss)", unit="examples/s",
                 workload=f"reserve-{args.config} contrastive pretraining FORWARD + loss (no backward yet): {b} examples "
                          f"per GPU (16 frames {gh}x{gw}, 48 audio sub-segments, 6 joint sequences x 640), targets "
                          f"all-gathered over the ranks",
                 extra={"examples_per_gpu": b})
    elif args.workload == "pretrain_step":
        from merlot_reserve_b200.training import PretrainStep
        b = args.clips
        stepper = PretrainStep(cfg, params)
        host = pin(synthetic_pretrain_batch(cfg, b, seed=1 + 1000 * rank))
        resident = {k: v.to(dev) for k, v in host.items()}
        out_host = torch.empty((1,), dtype=torch.float32).pin_memory()

        def step_e2e():
            loss, _ = stepper.train_step({k: v.to(dev, non_blocking=True) for k, v in host.items()})
            out_host.copy_(loss.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        w.update(step=lambda: stepper.train_step(resident), step_e2e=step_e2e, units=b, flops=3.0 * flops_pretrain_example(cfg) * b,
                 h2d=sum(v.numel() * v.element_size() for v in host.values()), d2h=4,
                 metric="examples/sec (train step: forward + backward + all-reduce + AdamW)", unit="examples/s",
                 workload=f"reserve-{args.config} contrastive pretraining train_step (pretrain_model.py:306-340): {b} examples "
                          f"per GPU (16 frames {gh}x{gw}, 48 audio sub-segments, 6 joint sequences x 640), targets "
                          f"all-gathered and gradients all-reduced (bf16) over the ranks, AdamW with bf16-coded moments",
                 extra={"examples_per_gpu": b})
    else:
        raise ValueError(args.workload)
    return w


def train_step_leg(cfg, params, rank, world, dev, timed, steps=8, warmup=3, b=8):
    """BASELINE configs[2] inside the default line: Reserve-Base contrastive train_step (pretrain_model.py:306-340),
    b examples per GPU, the packed all-gather of the contrastive targets, the packed reduce-scatter of their
    gradients and the bucketed bf16 gradient all-reduce over the `world` ranks.  This is the leg of the N = 1/2/4/8
    curve that moves data between GPUs.  Collective costs are reported three ways: each collective timed alone on its
    real buffers, the step with the gradient all-reduce skipped (timing probe; the difference to the full step is the
    EXPOSED all-reduce), and the step with the single blocking all-reduce after the backward pass."""
    import torch.distributed as dist
    from merlot_reserve_b200 import training
    from merlot_reserve_b200.synthetic import synthetic_pretrain_batch
    H = cfg["model"]["hidden_size"]
    stepper = training.PretrainStep(cfg, params)
    batch = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
             synthetic_pretrain_batch(cfg, b, seed=1 + 1000 * rank).items()}
    for _ in range(warmup):
        stepper.train_step(batch)
    ms_eager = timed(lambda: stepper.train_step(batch), steps) / steps
    ms, graph_note = ms_eager, "eager (one launch per kernel)"
    use_graph = os.environ.get("MR_TRAIN_GRAPH", "1") == "1"
    if use_graph:
        # the whole step (about 1 650 launches, the index glue and -- with more than one rank -- its NCCL collectives) as
        # ONE CUDA-graph launch (capture_train_step)
        try:
            run = stepper.capture_train_step(batch)
            for _ in range(warmup):
                run()
            ms = timed(lambda: run(), steps) / steps
            graph_note = "CUDA graph (capture_train_step): one graph launch per step"
        except Exception as e:      # keep the eager number rather than losing the leg
            use_graph = False
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    out = {"workload": f"reserve-base train_step, {b} examples per GPU (16 frames 12x20, 48 audio sub-segments, 6 joint "
                       f"sequences x 640, 48 spans), bf16, bf16-coded AdamW moments",
           "n_gpus": world, "steps": steps, "warmup": warmup, "examples_per_s": world * b / (ms * 1e-3), "ms_per_step": ms,
           "host_path": graph_note, "ms_per_step_eager": ms_eager,
           "tflops_per_gpu": 3.0 * flops_pretrain_example(cfg) * b / (ms * 1e-3) / 1e12,
           "bucket_bytes": list(getattr(stepper, "last_bucket_bytes", [])),
           "grad_allreduce_bytes": int(stepper._state["g16"].numel() * 2)}
    if world > 1:
        ms = ms_eager if not use_graph else ms
        training._SKIP_GRAD_ALLREDUCE = True                 # timing probe: same kernels, no gradient all-reduce
        for _ in range(2):
            stepper.train_step(batch)
        ms_skip = timed(lambda: stepper.train_step(batch), steps) / steps
        training._SKIP_GRAD_ALLREDUCE = False
        training._OVERLAP_ALLREDUCE = False                  # the single blocking all-reduce after the backward pass
        for _ in range(2):
            stepper.train_step(batch)
        ms_block = timed(lambda: stepper.train_step(batch), steps) / steps
        training._OVERLAP_ALLREDUCE = True
        g16 = stepper._state["g16"]
        rows = 188 * b                                       # the six packed target sets: 188 b rows of H bf16
        src = torch.zeros((rows, H), dtype=torch.bfloat16, device=dev)
        dst = torch.empty((world * rows, H), dtype=torch.bfloat16, device=dev)
        rs_in = torch.zeros((world, rows, H), dtype=torch.float32, device=dev)
        rs_out = torch.empty((rows, H), dtype=torch.float32, device=dev)
        alone = {}
        comm = stepper.communicator()        # the library's collective entries on its own ncclComm_t (None: torch.distributed)
        out["collectives"] = ("mr_all_gather / mr_reduce_scatter / mr_all_reduce (C ABI, ncclComm_t from mr_nccl_comm_init)"
                              if comm is not None else "torch.distributed (NCCL)")
        for name, fn in (("allgather_ms", (lambda: comm.all_gather(src, dst)) if comm else (lambda: dist.all_gather_into_tensor(dst, src))),
                         ("reduce_scatter_ms", (lambda: comm.reduce_scatter(rs_in, rs_out)) if comm else (lambda: dist.reduce_scatter_tensor(rs_out, rs_in))),
                         ("allreduce_standalone_ms", (lambda: comm.all_reduce(g16)) if comm else (lambda: dist.all_reduce(g16)))):
            for _ in range(3):
                fn()
            alone[name] = timed(fn, 10) / 10
        g16.zero_()
        out.update(alone)
        out.update({"ms_per_step_no_grad_allreduce": ms_skip, "allreduce_exposed_ms": max(ms_eager - ms_skip, 0.0),
                    "ms_per_step_blocking_allreduce": ms_block,
                    "allreduce_busbw_gbs": 2.0 * (world - 1) / world * g16.numel() * 2 / (alone["allreduce_standalone_ms"] * 1e-3) / 1e9,
                    "limiting_collective": "bf16 gradient all-reduce (633 MB) against the tail of the backward pass: the "
                                           "vision-encoder bucket is final only when the backward pass ends"})
    del stepper, batch
    return out


def latency_leg(cfg, params, rank, world, dev, timed, iters=20):
    """BASELINE configs[0] on the B200 path: ONE Reserve-Base clip (16 segments), latency in ms.  Three host paths over
    the same kernels: per-op ctypes calls composed in Python (with the profiling events of that mode), the single
    mr_embed_video_batched C call, and its CUDA-graph replay (one graph launch per forward)."""
    from merlot_reserve_b200 import MerlotReserve, _lib
    from merlot_reserve_b200.synthetic import synthetic_batch
    model = MerlotReserve.from_config(cfg).bind(params)
    batch = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in synthetic_batch(cfg, 1, nseg=NSEG, seed=7 + rank).items()}
    out = {"workload": "reserve-base embed_video, 1 clip x 16 segments (12x20 patches) + 48 audio sub-segments, L=160",
           "n_gpus": world, "iters": iters}
    for _ in range(3):
        model.batch_embed_video(**batch)
    out["composite_call_ms"] = timed(lambda: model.batch_embed_video(**batch), iters) / iters
    run = model.capture_embed_video(1, nseg=NSEG)
    args = (batch["images"], batch["audio_clips"], batch["tokens"], batch["subseg_idxs"])
    for _ in range(3):
        run(*args)
    out["cuda_graph_ms"] = timed(lambda: run(*args), iters) / iters
    _lib.PROFILE = []
    for _ in range(3):
        model.batch_embed_video(**batch)
    out["per_op_calls_ms"] = timed(lambda: model.batch_embed_video(**batch), iters) / iters
    out["kernel_launches_per_forward"] = len(_lib.PROFILE) // (iters + 3)
    _lib.PROFILE = None
    out["flops_per_clip"] = flops_per_clip(cfg)
    out["tflops_graph"] = flops_per_clip(cfg) / (out["cuda_graph_ms"] * 1e-3) / 1e12
    return out


def large_legs(rank, world, dev, timed, steps=5, warmup=3):
    """BASELINE configs[3] and [4] inside the default line (short legs, replicas only -- no data-path collective):
    Reserve-Large 18x32 embed_video on the metric's 16-frame clip, and VCR-style multiple-choice scoring, batch 128."""
    from merlot_reserve_b200 import MerlotReserve
    from merlot_reserve_b200.synthetic import load_config, synthetic_batch, synthetic_params
    cfg = load_config("large", (18, 32))
    H = cfg["model"]["hidden_size"]
    model = MerlotReserve.from_config(cfg).bind(synthetic_params(cfg, seed=0))
    out = {}
    clips = 8
    batch = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
             synthetic_batch(cfg, clips, nseg=NSEG, seed=1 + 1000 * rank).items()}
    for _ in range(warmup):
        model.batch_embed_video(**batch)
    ms = timed(lambda: model.batch_embed_video(**batch), steps) / steps
    out["large_embed_video"] = {
        "workload": f"reserve-large resadapt embed_video, {clips} clips x 16 segments (18x32 patches, joint T = 2464) per GPU",
        "n_gpus": world, "steps": steps, "warmup": warmup, "segments_per_s": world * clips * NSEG / (ms * 1e-3),
        "ms_per_step": ms, "tflops_per_gpu": flops_per_clip(cfg) * clips / (ms * 1e-3) / 1e12}
    del batch
    B, n_cand, Ltok = 128, 4, 144
    rng = np.random.default_rng(5 + rank)
    toks = rng.integers(10, 32768, size=(B, n_cand, Ltok)).astype(np.int32)
    toks[:, :, -1] = 3
    images = torch.from_numpy(rng.random((B, 18 * 32, 768), dtype=np.float32)).to(dev)
    tokens = torch.from_numpy(toks).to(dev)
    proj = (rng.standard_normal(H) * 0.05).astype(np.float32)
    for _ in range(warmup):
        model.score_multiple_choice(images, tokens, proj)
    ms = timed(lambda: model.score_multiple_choice(images, tokens, proj), steps) / steps
    out["mc_scoring"] = {
        "workload": f"reserve-large VCR-style multiple-choice scoring, batch {B}, 1 frame (18x32) x {n_cand} candidates, L = {Ltok}, per GPU",
        "n_gpus": world, "steps": steps, "warmup": warmup, "examples_per_s": world * B / (ms * 1e-3), "ms_per_step": ms,
        "tflops_per_gpu": flops_mc_example(cfg, n_cand, Ltok) * B / (ms * 1e-3) / 1e12}
    return out


def run_b200(args):
    import torch.distributed as dist
    from merlot_reserve_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.require_b200(local)
    numa_cpus = None
    if world > 1 and os.environ.get("MR_NUMA_BIND", "1") == "1":
        # before any pinned allocation: staging buffers then live on the GPU's own NUMA node
        from merlot_reserve_b200.numa import bind_to_gpu
        numa_cpus = bind_to_gpu(local)

    w = build_workload(args, rank, dev)
    cfg, params = w["cfg"], w["params"]
    step_resident, step_e2e = w["step"], w["step_e2e"]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        total_ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    ms_step = total_ms / args.steps
    value = world * w["units"] / (ms_step * 1e-3)

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps
    e2e_value = world * w["units"] / (e2e_ms * 1e-3)

    # per-kernel-class timing: CUDA events around every launch of instrumented steps (the model then composes the
    # layer loop call by call instead of through mr_transformer_encode: same kernels, same order)
    peaks = load_peaks()
    _lib.PROFILE = []
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    agg = {}
    for name, (tag, fl, by), a, b in _lib.PROFILE:
        d = agg.setdefault(tag, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += a.elapsed_time(b)
        d["flops"] += fl
        d["bytes"] += by
        d["launches"] += 1
    _lib.PROFILE = None
    prof_ms = sum(d["ms"] for d in agg.values())
    kernels = {t: {"share": d["ms"] / prof_ms, "ms_per_step": d["ms"] / 2, "launches_per_step": d["launches"] // 2,
                   "tflops": (d["flops"] / d["ms"] / 1e9) if d["flops"] else None,
                   "gbs": (d["bytes"] / d["ms"] / 1e6) if (d["bytes"] and not d["flops"]) else None}
               for t, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    gemm = {"ms": 0.0, "flops": 0.0, "launches": 0}
    for t, d in agg.items():
        if t.startswith("gemm"):
            for k in gemm:
                gemm[k] += d[k]
    achieved = gemm["flops"] / gemm["ms"] / 1e9
    peak = peaks["tflops_sustained"]
    traffic, traffic_note = None, None
    tf = ROOT / "profiles" / "gemm_traffic.json"      # dram bytes per launch from the committed ncu --set full capture
    if tf.exists() and args.workload == "embed_video" and args.config == "base":
        tj = json.loads(tf.read_text())
        traffic = tj.get("dram_bytes_per_launch_avg")
        traffic_note = ("dram__bytes_read+write per launch, mean of the four vision-layer GEMM launches "
                        "(M = 246784: algorithmic 1.90 / 1.52 / 0.76 / 1.90 GB for down / qkv / out / up); " + tj.get("source", "")
                        + (f", commit {tj['commit']}" if tj.get("commit") else "") + " (tools/gemm_traffic.py)")
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all epilogue variants)", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
                "avg_launch_ms": gemm["ms"] / gemm["launches"], "flops_per_launch": gemm["flops"] / gemm["launches"],
                "share_of_step": gemm["ms"] / prof_ms, "traffic": traffic, "traffic_note": traffic_note}
    whole = w["flops"] / (ms_step * 1e-3) / 1e12
    w_metric, w_unit, w_h2d, w_d2h = w["metric"], w["unit"], w["h2d"], w["d2h"]
    w_workload, w_extra = w.get("workload"), w.get("extra", {})

    # the other BASELINE configs as short legs of the same line (headline metric / config above are unchanged)
    legs = {}
    if args.workload == "embed_video" and args.config == "base" and not args.grid and not args.no_extra_legs:
        import gc
        w.clear()
        del step_resident, step_e2e
        gc.collect()
        torch.cuda.empty_cache()
        legs["latency_1clip"] = latency_leg(cfg, params, rank, world, dev, timed)
        gc.collect()
        torch.cuda.empty_cache()
        legs["train_step"] = train_step_leg(cfg, params, rank, world, dev, timed)
        gc.collect()
        torch.cuda.empty_cache()
        legs.update(large_legs(rank, world, dev, timed))
        gc.collect()
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "embed_video":
        val, dt, cores = cpu_oracle_segments_per_s(cfg, params, steps=2, warmup=1, nseg_sample=NSEG)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"CPU oracle (torch fp32 port of the reference arithmetic; jax not installable): "
                         f"1 clip x {NSEG} segments, 2 timed passes, {dt:.2f} s each"}
    if rank == 0:
        line = {
            "metric": w_metric, "value": value, "unit": w_unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world) if args.workload == "embed_video" else
                      {"workload": w_workload, **w_extra, "parallelism": f"replicas x{world}",
                       "l2": "inputs and activations of a step exceed the 126 MB L2 many times over"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": w_unit, "ms_per_step": e2e_ms, "h2d_bytes_per_step": w_h2d,
                    "d2h_bytes_per_step": w_d2h},
            "gpu_launches": int(launches),
            "host_placement": (f"rank bound to the {len(numa_cpus)} CPUs local to its GPU (pinned buffers first-touch there)"
                               if numa_cpus else "default (no NUMA binding)"),
            "roofline": roofline,
            "whole_step_tflops": whole, "whole_step_frac_of_sustained_peak": whole / peak,
            "kernels": kernels,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(legs)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="base", choices=["base", "large"])
    ap.add_argument("--clips", type=int, default=None, help="clips / examples per GPU and step")
    ap.add_argument("--workload", default="embed_video", choices=["embed_video", "mc_scoring", "pretrain_fwd", "pretrain_step"],
                    help="embed_video = BASELINE configs[1] (default) / [3] with --config large --grid 18 32; "
                         "mc_scoring = configs[4]; pretrain_fwd = the forward half of configs[2]")
    ap.add_argument("--grid", type=int, nargs=2, default=None, help="output_grid override, e.g. 18 32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip the train_step / large_embed_video / mc_scoring legs of the default line")
    args = ap.parse_args()
    if args.clips is None:
        args.clips = {"embed_video": 64 if args.config == "base" else 16, "mc_scoring": 128, "pretrain_fwd": 8, "pretrain_step": 8}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        with torch.no_grad():
            run_b200(args)


if __name__ == "__main__":
    main()
