"""Device versions of the two host preprocessing functions of `mreserve/preprocess.py` that produce the model's input
tensors, under the reference's names and argument meaning.  No CPU fallback: both need the B200
library.

* `make_spectrogram(waveform, playback_speed=1, sr=22050, pad_size=2)`  <- mreserve/preprocess.py:121-156
* `preprocess_image_to_patches(img, output_grid_size)`                   <- mreserve/preprocess.py:464-481
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib as L


def _dev(x, dtype):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.cuda().contiguous() if not t.is_cuda else t.contiguous()


def make_spectrogram(waveform, playback_speed: int = 1, sr: int = 22050, pad_size: int = 2):
    """waveform [n] (or [B, n]) float -> [3, 60, 65] (or [B, 3, 60, 65]) log-mel sub-clips, as the reference: 64 mels of a
    1536-point STFT at hop 588, log(mel + 0.1) - log(0.1), playback speed in column 64.  Raises ValueError when the
    waveform does not give 180 + 4 * pad_size frames (the reference's check, :148-149)."""
    if sr != 22050 or playback_speed != 1:
        raise ValueError("the device path implements the pretraining parameters: sr = 22050, playback_speed = 1")
    as_np = not isinstance(waveform, torch.Tensor)
    w = _dev(waveform, torch.float32)
    single = w.dim() == 1
    if single:
        w = w[None]
    n_frames = 1 + w.shape[1] // 588
    if n_frames != 180 + 4 * pad_size:
        raise ValueError("provided mel spectrogram {}. target size: {}".format((n_frames, 65), (180 + 4 * pad_size, 65)))
    out = L.logmel(w, pad_size=pad_size, split=True)
    out = out[0] if single else out
    return out.cpu().numpy() if as_np else out


def preprocess_image_to_patches(img, output_grid_size: Tuple[int, int]):
    """uint8 image [h, w, 3] (or a batch [B, h, w, 3]) of any size -> patches [h1 * w1, 768] (or [B, h1 * w1, 768]) in
    the module dtype (bf16), ready for the vision encoder."""
    h1, w1 = output_grid_size
    assert h1 <= 24, "we didn't pretrain on anything bigger than 24x24 or 18x32"
    assert w1 <= 32, "we didn't pretrain on anything bigger than 24x24 or 18x32"
    as_np = not isinstance(img, torch.Tensor)
    f = _dev(img, torch.uint8)
    single = f.dim() == 3
    if single:
        f = f[None]
    out = L.resize_patchify(f, h1, w1)
    out = out[0] if single else out
    return out.float().cpu().numpy() if as_np else out
