"""CPU restatement (numpy) of the preprocessing step in front of the hot path -- TEST INFRASTRUCTURE ONLY (never
imported by the product; see oracle/reserve_oracle.py for the rules).  SURVEY.md 8f rank 4.

* `make_spectrogram`  <- mreserve/preprocess.py:121-156: librosa.feature.melspectrogram(sr 22050, n_fft 1536, hop 588,
  64 mels, fmin 20, fmax 11025, window = scipy.signal.windows.hann passed as a CALLABLE), log(mel + 0.1) - log(0.1), the
  playback-speed column, the three 60-frame sub-clips with `pad_size` gaps.
* `resize_and_pad`    <- pretrain/data_utils.py:42-147 as preprocess_image_to_patches calls it (mreserve/preprocess.py:
  464-481): do_random_scale=False, do_flip_if_vertical=False, bilinear with antialias=True, clip to [0, 1], crop,
  zero pad_to_bounding_box.
* `image_to_patches`  <- mreserve/preprocess.py:464-481: convert_image_dtype, resize_and_pad, space_to_depth(16).

PARITY UNPINNED for these two functions: the arithmetic lives in librosa 0.8.1 / scipy 1.7.0 / numpy 1.19.5
(requirements.txt:12,13,19) and TensorFlow >= 2.6 (requirements.txt:2), none of which is in /root/reference or
installable here, and the reference holds no test vector for them.  What is restated from library knowledge:
  librosa 0.8.1  `stft` (center=True, pad_mode='reflect', window callable -> window(n_fft) i.e. the SYMMETRIC Hann of
                 scipy's default sym=True, frames of n_fft at hop_length, numpy rfft, complex64 result),
                 `melspectrogram` (power 2.0, mel_basis . |S|^2), `filters.mel` (Slaney scale htk=False, norm='slaney',
                 float32 basis);
  TensorFlow     `tf.image.resize(..., method=BILINEAR, antialias=True)` = ScaleAndTranslate with the triangle kernel:
                 per output index a span of source taps, weights = triangle((tap + 0.5 - centre) / kernel_scale) with
                 kernel_scale = max(in / out, 1), normalised over the in-range taps; columns first, then rows, float32;
                 `convert_image_dtype(uint8 -> float32)` = x * (1 / 255); `pad_to_bounding_box` = zeros bottom / right.
"""
from __future__ import annotations

import numpy as np

N_FFT, HOP, N_MELS, SR = 1536, 588, 64, 22050
FMIN, FMAX = 20.0, 11025.0
EPS = 1e-1


# ------------------------------------------------------------------------------------------------
# log-mel spectrogram
# ------------------------------------------------------------------------------------------------
def hann_sym(n: int) -> np.ndarray:
    """scipy.signal.windows.hann(n) with its default sym=True: 0.5 - 0.5 cos(2 pi k / (n - 1))."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))


def hz_to_mel_slaney(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def mel_to_hz_slaney(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> float32 [n_mels, 1 + n_fft // 2]."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None]
    w = np.zeros((n_mels, 1 + n_fft // 2), np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def melspectrogram(waveform: np.ndarray, playback_speed: int = 1) -> np.ndarray:
    """power spectrogram -> mel, [n_mels, 1 + len // hop] float32."""
    y = np.asarray(waveform, np.float32)
    n_fft, hop = N_FFT * playback_speed, HOP * playback_speed
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (ypad.size - n_fft) // hop
    idx = np.arange(n_fft)[None] + hop * np.arange(n_frames)[:, None]
    frames = ypad[idx] * hann_sym(n_fft).astype(np.float32)[None]
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1).astype(np.complex64)          # [frames, 1 + n_fft/2]
    power = (np.abs(spec) ** 2).astype(np.float32)
    return mel_filterbank(n_fft=n_fft) @ power.T


def log_mel(waveform: np.ndarray, playback_speed: int = 1) -> np.ndarray:
    """[frames, 65]: log(mel + 0.1) - log(0.1) and the playback-speed column (preprocess.py:140-145)."""
    mel = melspectrogram(waveform, playback_speed)
    lm = np.log(mel + np.float32(EPS)) - np.log(np.float32(EPS))
    lm = np.concatenate([lm, playback_speed * np.ones((1, lm.shape[1]), lm.dtype)], 0)
    return lm.T.astype(np.float32)


def make_spectrogram(waveform: np.ndarray, playback_speed: int = 1, pad_size: int = 2) -> np.ndarray:
    """mreserve/preprocess.py:121-156 -> [3, 60, 65]."""
    lm = log_mel(waveform, playback_speed)
    seq = 60
    if lm.shape != (seq * 3 + pad_size * 4, 65):
        raise ValueError("provided mel spectrogram {}. target size: {}".format(lm.shape, (seq * 3 + pad_size * 4, 65)))
    return np.stack([lm[pad_size:pad_size + seq], lm[2 * pad_size + seq:2 * pad_size + 2 * seq],
                     lm[3 * pad_size + 2 * seq:3 * pad_size + 3 * seq]])


# ------------------------------------------------------------------------------------------------
# antialiased bilinear resize + pad
# ------------------------------------------------------------------------------------------------
def resize_spans(in_size: int, out_size: int):
    """TensorFlow ScaleAndTranslate spans for the triangle kernel with antialias: (start [out], weights [out, taps])
    in float32 (taps beyond a span carry weight 0)."""
    scale = np.float32(out_size) / np.float32(in_size)
    inv_scale = np.float32(1.0) / scale
    kscale = np.maximum(inv_scale, np.float32(1.0))
    radius = np.float32(1.0)
    taps = min(2 * int(np.ceil(radius * kscale)) + 1, in_size)
    starts = np.zeros(out_size, np.int32)
    weights = np.zeros((out_size, taps), np.float32)
    for x in range(out_size):
        sample_f = (np.float32(x) + np.float32(0.5)) * inv_scale
        lo = int(np.ceil(sample_f - radius * kscale - np.float32(0.5)))
        hi = int(np.floor(sample_f + radius * kscale - np.float32(0.5)))
        lo = min(max(lo, 0), in_size - 1)
        hi = min(max(hi, 0), in_size - 1) + 1
        starts[x] = lo
        w = np.zeros(taps, np.float32)
        total = np.float32(0.0)
        for j, src in enumerate(range(lo, hi)):
            pos = (np.float32(src) + np.float32(0.5) - sample_f) / kscale
            w[j] = np.maximum(np.float32(0.0), np.float32(1.0) - np.abs(pos))
            total = np.float32(total + w[j])
        if total > np.float32(0.0):
            w = (w / total).astype(np.float32)
        weights[x] = w
    return starts, weights


def _resize_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    starts, weights = resize_spans(img.shape[axis], out_size)
    img = np.moveaxis(img, axis, 0)
    out = np.zeros((out_size,) + img.shape[1:], np.float32)
    for j in range(weights.shape[1]):                       # taps in increasing source order, float32 mul then add
        src = np.minimum(starts + j, img.shape[0] - 1)
        out = (out + weights[:, j].reshape((-1,) + (1,) * (img.ndim - 1)) * img[src]).astype(np.float32)
    return np.moveaxis(out, 0, axis)


def resize_and_pad(image: np.ndarray, desired_output_size):
    """pretrain/data_utils.py:42-147 with do_random_scale=False, do_flip_if_vertical=False, bilinear + antialias.
    image float32 [H, W, 3] in [0, 1] -> ([dh, dw, 3] float32, image_info [7])."""
    dh, dw = desired_output_size
    h, w = np.float32(image.shape[0]), np.float32(image.shape[1])
    image_scale = np.minimum(np.float32(dw) / w, np.float32(dh) / h)
    sh, sw = int(np.float32(h * image_scale)), int(np.float32(w * image_scale))
    x = _resize_axis(np.asarray(image, np.float32), sw, 1)              # columns first,
    x = _resize_axis(x, sh, 0)                                          # then rows
    x = np.clip(x, 0.0, 1.0)[:dh, :dw]
    out = np.zeros((dh, dw, 3), np.float32)
    out[:x.shape[0], :x.shape[1]] = x
    info = np.array([min(sh, dh) / np.float32(dh), min(sw, dw) / np.float32(dw), np.float32(1.0) / image_scale, h, w, 0.0, 0.0],
                    np.float32)
    return out, info


def image_to_patches(img_u8: np.ndarray, output_grid_size, patch: int = 16) -> np.ndarray:
    """mreserve/preprocess.py:464-481: uint8 [H, W, 3] (any size) -> float32 [h1 * w1, 768]."""
    h1, w1 = output_grid_size
    assert h1 <= 24 and w1 <= 32
    x = np.asarray(img_u8).astype(np.float32) * np.float32(1.0 / 255.0)
    x, _ = resize_and_pad(x, (h1 * patch, w1 * patch))
    x = x.reshape(h1, patch, w1, patch, 3).transpose(0, 2, 1, 3, 4)     # space_to_depth: block (by, bx), features (py, px, c)
    return x.reshape(h1 * w1, patch * patch * 3)
