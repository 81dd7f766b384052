"""GPU parity of the control-side loudness extractors (csrc/nws_loudness.cu, through the C ABI and the host mirror
of data/utils/loudness_extraction.py) against the numpy oracle (oracle/loudness_oracle.py: librosa 0.8.0 restated).
Stated tolerances: dB spectrogram 1e-3 dB, per-frame loudness 1e-3 dB = 1.25e-5 normalised, rms 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import loudness_oracle as lo
from tests.test_loudness_cpu import _signal

pytestmark = pytest.mark.gpu


def _mirror():
    from neural_waveshaping_synthesis.data.utils import loudness_extraction as le
    return le


@pytest.mark.parametrize("n,n_fft,hop", [(64000, 1024, 128),     # gin/data/urmp_4second_crepe.gin:11-14, a 4 s segment
                                         (64000, 2048, 512),     # the function's own defaults (colab cell 14)
                                         (5000, 256, 100),       # length not a multiple of the hop
                                         (2100, 4096, 64),       # barely longer than the reflect padding
                                         (1000, 64, 7)])
def test_loudness_and_spectrogram_match_oracle(n, n_fft, hop):
    le = _mirror()
    x = _signal(n, seed=n_fft)
    x[n // 3: n // 3 + n // 10] = 0.0   # a stretch of digital silence: bins clamped at amin / at the -80 dB floor
    ref_db = lo.compute_power_spectrogram(x, n_fft, hop, "hann", 1e-5)
    db = le.compute_power_spectrogram(x, n_fft, hop, "hann", 1e-5)
    assert db.shape == ref_db.shape and db.dtype == np.float32
    assert np.abs(db - ref_db).max() < 1e-3, np.abs(db - ref_db).max()
    for normalise in (True, False):
        ref = lo.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None, normalise=normalise)
        got = le.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None, normalise=normalise)
        assert got.shape == ref.shape == (1 + n // hop,)
        tol = 1.25e-5 if normalise else 1e-3
        assert np.abs(got - ref).max() < tol, (normalise, np.abs(got - ref).max())


@pytest.mark.parametrize("tag", ["gin", "default", "ragged"])
def test_against_fixtures_of_the_real_reference_functions(tag):
    """tests/golden/loudness_*.npz: outputs of the reference's own loudness_extraction.py functions
    (oracle/gen_golden_loudness.py), every variant the file offers."""
    import os
    le = _mirror()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loudness_%s.npz" % tag))
    x, n_fft, hop = z["audio"], int(z["n_fft"]), int(z["hop_length"])
    assert np.abs(le.compute_power_spectrogram(x, n_fft, hop, "hann", 1e-5) - z["db"]).max() < 1e-3
    assert np.abs(le.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None) - z["loudness_frames"]).max() < 1.25e-5
    assert np.abs(le.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None, normalise=False)
                  - z["loudness_frames_db"]).max() < 1e-3
    got = le.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop)
    assert got.shape == z["loudness_samples"].shape and np.abs(got - z["loudness_samples"]).max() < 1.25e-5
    assert np.abs(le.extract_rms(x, n_fft, hop, interpolate_fn=None) - z["rms_frames"]).max() < 1e-6
    got = le.extract_rms(x, n_fft, hop)
    assert got.shape == z["rms_samples"].shape and np.abs(got - z["rms_samples"]).max() < 1e-6


def test_batch_rows_are_independent_and_interpolation_matches():
    le = _mirror()
    xs = np.stack([_signal(32000, seed=s) * g for s, g in ((1, 1.0), (2, 0.01), (3, 0.0))])   # loud, quiet, silent
    out = le.perceptual_loudness_batch(torch.from_numpy(xs).cuda(), 1024, 128)
    assert out.shape == (3, 251) and out.is_cuda
    for i in range(3):
        ref = lo.extract_perceptual_loudness(xs[i], n_fft=1024, hop_length=128, interpolate_fn=None)
        assert np.abs(out[i].cpu().numpy() - ref).max() < 1.25e-5
    assert torch.all(out[2] == 1.0)     # silence: every bin at the maximum
    # default arguments: interpolated to one value per sample (upsampling.py:20-36)
    ref = lo.extract_perceptual_loudness(xs[0])
    got = le.extract_perceptual_loudness(xs[0])
    assert got.shape == ref.shape == (32000,) and np.abs(got - ref).max() < 1.25e-5


@pytest.mark.parametrize("n,window,hop", [(64000, 2048, 512), (5000, 300, 77), (4097, 1025, 128)])
def test_rms_matches_oracle(n, window, hop):
    le = _mirror()
    x = _signal(n, seed=5)
    ref = lo.extract_rms(x, window, hop, interpolate_fn=None)
    got = le.extract_rms(x, window, hop, interpolate_fn=None)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-6 * max(1.0, float(ref.max()))


def test_loudness_errors():
    from neural_waveshaping_synthesis_b200._lib import NwsError
    le = _mirror()
    x = torch.zeros(1, 4096, device="cuda")
    with pytest.raises(NwsError):
        le.perceptual_loudness_batch(x, n_fft=1000, hop_length=128)      # not a power of two
    with pytest.raises(NwsError):
        le.perceptual_loudness_batch(x[:, :500], n_fft=1024, hop_length=128)   # shorter than the reflect padding
    with pytest.raises(ValueError):
        le.perceptual_loudness_batch(torch.zeros(1, 4096), 1024, 128)    # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        le.extract_perceptual_loudness(np.zeros(4096, np.float32), window="hamming")


def test_linear_interpolation_on_device_matches_reference_fixture():
    """upsampling.linear_interpolation (nws_interp_frames: np.interp's float64 arithmetic, one thread per sample)
    against vectors made by the reference's own upsampling.py — including F = 1 and the un-cropped form."""
    from neural_waveshaping_synthesis.data.utils import upsampling as up
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "upsampling.npz"))
    for i, (F, window, hop, orig) in enumerate(z["cases"]):
        frames = z["frames_%d" % i]
        ref = z["linear_%d" % i]
        got = up.linear_interpolation(frames, int(window), int(hop), original_length=int(orig) or None)
        assert got.dtype == np.float64 and got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, float(np.abs(ref).max())), i
    # batched, device in / device out
    fr = torch.from_numpy(np.stack([z["frames_0"], z["frames_0"][::-1].copy()])).cuda()
    out = up.interp_frames_batch(fr, 2048, 512, 24000)
    assert out.shape == (2, 24000) and out.is_cuda
    assert np.abs(out[0].cpu().numpy() - z["linear_0"]).max() <= 1e-11
