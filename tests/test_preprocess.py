"""Device preprocessing: mr_logmel against the numpy restatement of make_spectrogram
(mreserve/preprocess.py:121-156) and mr_resize_patchify against the restatement of preprocess_image_to_patches
(mreserve/preprocess.py:464-481; pretrain/data_utils.py:42-147).  PARITY UNPINNED for these two (librosa / TensorFlow
are not available and the reference holds no vectors for them): see oracle/preprocess_oracle.py for what is restated.

Tolerances: log-mel |diff| <= 2e-3 (the model reads these features in bf16, whose spacing at a typical value of 2 is
1.6e-2); resized patches bit-exact in fp32 (every operation of the kernel is an explicitly rounded fp32 op in the
oracle's order) and therefore bit-exact in bf16."""
import numpy as np
import pytest
import scipy.signal
import torch

from oracle import preprocess_oracle as P


def test_oracle_window_filterbank_and_stft():
    """The restatement against scipy where scipy has the same function: the symmetric Hann window, and the STFT power
    through scipy.signal.stft on the reflect-padded signal."""
    assert np.allclose(P.hann_sym(1536), scipy.signal.windows.hann(1536), atol=1e-15)
    fb = P.mel_filterbank()
    assert fb.shape == (64, 769) and fb.dtype == np.float32 and (fb >= 0).all()
    assert ((fb > 0).sum(0) <= 2).all()                          # triangles: a bin feeds at most two filters
    assert (np.diff(fb.argmax(1)) > 0).all()                     # centre frequencies increase
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(110250) * 0.1).astype(np.float32)
    ypad = np.pad(y, 768, mode="reflect")
    _, _, Z = scipy.signal.stft(ypad, window=scipy.signal.windows.hann(1536), nperseg=1536, noverlap=1536 - 588, boundary=None,
                                padded=False, scaling="spectrum")
    Z = Z * scipy.signal.windows.hann(1536).sum()
    ref = P.melspectrogram(y)
    assert ref.shape == (64, 188)
    assert np.abs(fb @ (np.abs(Z) ** 2) - ref).max() <= 1e-5 * np.abs(ref).max()
    spec = P.make_spectrogram(y)
    assert spec.shape == (3, 60, 65) and (spec[..., 64] == 1.0).all() and (spec[..., :64] >= 0).all()
    with pytest.raises(ValueError):
        P.make_spectrogram(y[:100000])


def test_oracle_resize_spans_and_identity():
    starts, w = P.resize_spans(640, 320)                         # 2x down: triangle of half-width 2
    assert w.shape == (320, 5) and np.allclose(w.sum(1), 1.0, atol=1e-6)
    assert np.allclose(w[1, :4], [0.125, 0.375, 0.375, 0.125]) and starts[1] == 1
    starts, w = P.resize_spans(100, 100)                         # identity: a single tap of weight 1
    assert (starts + w.argmax(1) == np.arange(100)).all() and np.allclose(w.max(1), 1.0) and np.allclose(w.sum(1), 1.0)
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(192, 320, 3), dtype=np.uint8)
    from oracle.reserve_oracle import frames_to_patches
    assert np.array_equal(P.image_to_patches(img, (12, 20)), frames_to_patches(img[None], 12, 20)[0])
    tall = rng.integers(0, 256, size=(400, 300, 3), dtype=np.uint8)      # scale 0.48: 192 x 144, zero padded to 320 wide
    p = P.image_to_patches(tall, (12, 20)).reshape(12, 20, 16, 16, 3)
    assert (p[:, 9:] == 0).all() and (p[:, :9].reshape(12, -1).max(1) > 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["noise", "tone", "quiet", "speechlike"])
def test_logmel_matches_make_spectrogram(cuda, kind):
    from merlot_reserve_b200 import preprocess
    rng = np.random.default_rng(5)
    n = 110250
    t = np.arange(n) / 22050.0
    if kind == "noise":
        y = rng.standard_normal((3, n)) * 0.2
    elif kind == "tone":                                         # full-scale tone: the leakage floor must stay below 0.1
        y = np.stack([0.99 * np.sin(2 * np.pi * f * t) for f in (440.0, 3000.3, 9000.7)])
    elif kind == "quiet":
        y = rng.standard_normal((3, n)) * 1e-3
    else:
        env = np.abs(np.sin(2 * np.pi * 3 * t)) ** 4
        y = np.stack([env * (np.sin(2 * np.pi * 180 * k * t) + 0.3 * rng.standard_normal(n)) * 0.3 for k in (1, 2, 3)])
    y = y.astype(np.float32)
    got = preprocess.make_spectrogram(y)
    ref = np.stack([P.make_spectrogram(w) for w in y])
    assert got.shape == ref.shape == (3, 3, 60, 65)
    assert np.abs(got - ref).max() <= 2e-3, float(np.abs(got - ref).max())
    single = preprocess.make_spectrogram(y[1])
    assert np.array_equal(single, got[1])
    with pytest.raises(ValueError):
        preprocess.make_spectrogram(y[0, :100000])


@pytest.mark.gpu
def test_logmel_whole_spectrogram_and_batch_chunks(cuda, lib):
    """split = 0 (all frames, any length) and a batch larger than one GEMM chunk (64 segments)."""
    rng = np.random.default_rng(6)
    y = (rng.standard_normal((70, 30000)) * 0.1).astype(np.float32)
    got = lib.logmel(torch.from_numpy(y).to(cuda), split=False).cpu().numpy()
    assert got.shape == (70, 1 + 30000 // 588, 65)
    for b in (0, 63, 64, 69):
        assert np.abs(got[b] - P.log_mel(y[b])).max() <= 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,grid", [(360, 640, (12, 20)), (192, 320, (12, 20)), (720, 1280, (18, 32)), (400, 300, (12, 20)),
                                      (97, 131, (24, 24)), (1080, 1920, (12, 20))])
def test_resize_patchify_is_bit_exact(cuda, lib, h, w, grid):
    """Down-scaling, identity, up-scaling, portrait (zero padded columns), odd sizes; fp32 and bf16 outputs."""
    rng = np.random.default_rng(h + w)
    frames = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    frames[1, : h // 2] = 255                                    # saturated region: the clip to [0, 1] is exercised
    ref = np.stack([P.image_to_patches(f, grid) for f in frames])
    dev = torch.from_numpy(frames).to(cuda)
    got32 = lib.resize_patchify(dev, *grid, out_f32=True).cpu().numpy()
    assert got32.shape == ref.shape == (2, grid[0] * grid[1], 768)
    assert np.array_equal(got32, ref), float(np.abs(got32 - ref).max())
    got16 = lib.resize_patchify(dev, *grid).float().cpu().numpy()
    assert np.array_equal(got16, torch.from_numpy(ref).to(torch.bfloat16).float().numpy())


@pytest.mark.gpu
def test_preprocessing_feeds_the_encoders(cuda):
    """uint8 frames of another size and raw waveforms -> device preprocessing -> vision / audio encoders equal the
    route through the oracle's host preprocessing (cosine >= 0.999)."""
    from merlot_reserve_b200 import MerlotReserve, preprocess
    from merlot_reserve_b200.synthetic import synthetic_params, tiny_config
    cfg = tiny_config(output_grid=(4, 6))
    params = synthetic_params(cfg, seed=0, weight_scale=4.0)
    model = MerlotReserve.from_config(cfg).bind(params)
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, size=(3, 150, 260, 3), dtype=np.uint8)
    on_device = model.get_imgs_enc(preprocess.preprocess_image_to_patches(torch.from_numpy(frames).to(cuda), (4, 6)))
    host = model.get_imgs_enc(np.stack([P.image_to_patches(f, (4, 6)) for f in frames]))
    cos = torch.nn.functional.cosine_similarity(torch.as_tensor(np.asarray(on_device.float().cpu())).reshape(-1, 128),
                                                torch.as_tensor(np.asarray(host)).float().reshape(-1, 128), dim=-1)
    assert cos.min() >= 0.999
    y = (rng.standard_normal((2, 110250)) * 0.2).astype(np.float32)
    a_dev = model.get_audio_enc(preprocess.make_spectrogram(torch.from_numpy(y).to(cuda)).reshape(6, 60, 65))
    a_host = model.get_audio_enc(np.concatenate([P.make_spectrogram(w) for w in y]))
    cos = torch.nn.functional.cosine_similarity(torch.as_tensor(np.asarray(a_dev.float().cpu() if isinstance(a_dev, torch.Tensor) else a_dev)).reshape(-1, 128),
                                                torch.as_tensor(np.asarray(a_host)).float().reshape(-1, 128), dim=-1)
    assert cos.min() >= 0.999
