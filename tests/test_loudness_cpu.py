"""CPU tests of the loudness-extraction oracle (oracle/loudness_oracle.py, a restatement of librosa 0.8.0's
stft / amplitude_to_db as used by the reference's data/utils/loudness_extraction.py) and of the float64 FFT
header the CUDA extractor uses.  librosa is absent, so the oracle is cross-checked against an independent STFT
(torch.stft, float64) rather than pinned to reference outputs — see the oracle's header."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import loudness_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))


def _signal(n, seed=0, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    f0 = 220.0 * 2 ** (0.3 * np.sin(2 * np.pi * 1.3 * t))
    phase = 2 * np.pi * np.cumsum(f0) / sr
    env = 0.5 * (1 - np.cos(2 * np.pi * np.minimum(t / t[-1], 1.0))) * 0.4
    x = env * sum(np.sin(k * phase) / k for k in range(1, 9)) + 1e-3 * rng.standard_normal(n)
    return x.astype(np.float32)


@pytest.mark.parametrize("n,n_fft,hop", [(64000, 1024, 128), (16000, 2048, 512), (5000, 256, 100)])
def test_stft_restatement_matches_torch_stft(n, n_fft, hop):
    x = _signal(n)
    S = lo.stft(x, n_fft, hop)
    assert S.dtype == np.complex64 and S.shape == (n_fft // 2 + 1, 1 + n // hop)
    ref = torch.stft(torch.from_numpy(x).double(), n_fft, hop_length=hop,
                     window=torch.hann_window(n_fft, periodic=True, dtype=torch.float64),
                     center=True, pad_mode="reflect", return_complex=True).numpy()
    assert np.abs(S - ref).max() <= 2e-7 * np.abs(ref).max()


def test_db_spectrogram_properties_and_edge_cases():
    x = _signal(16000, seed=1)
    db = lo.compute_power_spectrogram(x, 1024, 128, "hann", 1e-5)
    assert db.dtype == np.float32 and db.max() == 0.0 and db.min() >= -80.0
    loud = lo.extract_perceptual_loudness(x, n_fft=1024, hop_length=128, interpolate_fn=None)
    assert loud.shape == (126,) and np.all(loud >= 0.0) and np.all(loud <= 1.0)
    # digital silence: every bin sits at amin, i.e. at the maximum -> 0 dB everywhere -> normalised loudness 1
    silent = lo.extract_perceptual_loudness(np.zeros(4096, np.float32), n_fft=1024, hop_length=128, interpolate_fn=None)
    assert np.all(silent == 1.0)
    # the interpolated variant has one value per audio sample
    up = lo.extract_perceptual_loudness(x, n_fft=1024, hop_length=128)
    assert up.shape == (16000,)
    r = lo.extract_rms(x, 1024, 256, interpolate_fn=None)
    assert r.shape == (1 + 16000 // 256,) and r.max() < 1.0


def test_f64_fft_header_matches_numpy():
    out_so = os.path.join(tempfile.mkdtemp(prefix="nws_fft_"), "libfft_harness.so")
    src = os.path.join(HERE, "cpu_harness", "fft_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-shared", "-fPIC", "-o", out_so, src])
    L = ctypes.CDLL(out_so)
    P = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(3)
    for log_n in range(1, 13):
        n = 1 << log_n
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        tw = np.exp(-2j * np.pi * np.arange(max(n // 2, 1)) / n)
        out = np.zeros_like(x)
        L.h_fft_f64(log_n, x.ctypes.data_as(P), tw.ctypes.data_as(P), out.ctypes.data_as(P))
        ref = np.fft.fft(x)
        assert np.abs(out - ref).max() <= 2e-15 * np.abs(ref).max() * log_n


def test_mirror_has_no_cpu_fallback():
    from neural_waveshaping_synthesis.data.utils.loudness_extraction import extract_perceptual_loudness, perceptual_loudness_batch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        extract_perceptual_loudness(np.zeros(4096, np.float32))
    with pytest.raises(ValueError):
        perceptual_loudness_batch(torch.zeros(1, 4096))


def test_adjust_controls_matches_the_notebook_cell():
    """utils/control_adjust.py against a literal numpy/torch transcription of colab cell 15."""
    from neural_waveshaping_synthesis_b200.utils.control_adjust import adjust_controls
    rng = np.random.default_rng(0)
    T = 300
    f0 = (200 + 100 * rng.random(T)).astype(np.float32)
    loudness = rng.random(T).astype(np.float32)
    confidence = rng.random(T).astype(np.float32)
    data_mean = np.array([[300.0], [0.4], [0.5]])
    data_std = np.array([[80.0], [0.2], [0.1]])
    for kw in (dict(), dict(octave_shift=-1, loudness_scale=1.3, loudness_floor=0.2, loudness_conf_filter=0.3,
                            pitch_conf_filter=0.25, pitch_smoothing=4, loudness_smoothing=7)):
        p = dict(octave_shift=1, loudness_scale=0.5, loudness_floor=0, loudness_conf_filter=0, pitch_conf_filter=0,
                 pitch_smoothing=0, loudness_smoothing=0)
        p.update(kw)
        # --- the cell, verbatim in structure
        f0_filtered = f0 * (confidence > p["pitch_conf_filter"])
        loudness_filtered = loudness * (confidence > p["loudness_conf_filter"])
        f0_shifted = f0_filtered * (2 ** p["octave_shift"])
        loudness_floored = loudness_filtered * (loudness_filtered > p["loudness_floor"]) - p["loudness_floor"]
        loudness_scaled = loudness_floored * p["loudness_scale"]
        loud_norm = (loudness_scaled - data_mean[1]) / data_std[1]
        f0_t = torch.tensor(f0_shifted).float()
        loud_norm_t = torch.tensor(loud_norm).float()
        if p["pitch_smoothing"] != 0:
            f0_t = torch.nn.functional.conv1d(f0_t.expand(1, 1, -1), torch.ones(1, 1, p["pitch_smoothing"] * 2 + 1) /
                                              (p["pitch_smoothing"] * 2 + 1), padding=p["pitch_smoothing"]).squeeze()
        if p["loudness_smoothing"] != 0:
            loud_norm_t = torch.nn.functional.conv1d(loud_norm_t.expand(1, 1, -1),
                                                     torch.ones(1, 1, p["loudness_smoothing"] * 2 + 1) /
                                                     (p["loudness_smoothing"] * 2 + 1), padding=p["loudness_smoothing"]).squeeze()
        f0_norm_t = torch.tensor((f0_t.cpu() - data_mean[0]) / data_std[0]).float()
        control = torch.stack((f0_norm_t, loud_norm_t), dim=0)
        # ---
        got_f0, got_control = adjust_controls(torch.from_numpy(f0), torch.from_numpy(loudness), torch.from_numpy(confidence),
                                              data_mean, data_std, **kw)
        assert got_control.shape == (2, T) and got_control.dtype == torch.float32
        assert torch.allclose(got_f0, f0_t, rtol=1e-6, atol=1e-4)
        assert torch.allclose(got_control, control, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["gin", "default", "ragged"])
def test_oracle_glue_matches_the_real_reference_functions(tag):
    """tests/golden/loudness_*.npz were produced by the reference's own loudness_extraction.py / upsampling.py
    (oracle/gen_golden_loudness.py; librosa replaced by the restatement): the oracle's versions of those functions
    must reproduce them exactly — frame counts, normalisation, interpolation, rms framing."""
    z = np.load(os.path.join(HERE, "golden", "loudness_%s.npz" % tag))
    x, n_fft, hop = z["audio"], int(z["n_fft"]), int(z["hop_length"])
    assert np.array_equal(lo.compute_power_spectrogram(x, n_fft, hop, "hann", 1e-5), z["db"])
    assert np.array_equal(lo.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None), z["loudness_frames"])
    assert np.array_equal(lo.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop, interpolate_fn=None, normalise=False),
                          z["loudness_frames_db"])
    assert np.array_equal(lo.extract_perceptual_loudness(x, n_fft=n_fft, hop_length=hop), z["loudness_samples"])
    assert np.array_equal(lo.extract_rms(x, n_fft, hop, interpolate_fn=None), z["rms_frames"])
    assert np.array_equal(lo.extract_rms(x, n_fft, hop), z["rms_samples"])


def test_upsampling_host_interpolators_match_reference_fixture():
    """cubic_spline_interpolation / overlap_add_upsample (host numpy/scipy by the interface's definition) and the
    oracle's linear_interpolation against vectors made by the reference's own upsampling.py
    (oracle/gen_golden_upsampling.py)."""
    from neural_waveshaping_synthesis.data.utils import upsampling as up
    z = np.load(os.path.join(HERE, "golden", "upsampling.npz"))
    for i, (F, window, hop, orig) in enumerate(z["cases"]):
        frames = z["frames_%d" % i]
        kw = dict(window_length=int(window), hop_length=int(hop), original_length=int(orig) or None)
        assert np.array_equal(lo.linear_interpolation(frames, **kw), z["linear_%d" % i])
        if "cubic_%d" % i in z.files:
            got = up.cubic_spline_interpolation(frames, **kw)
            assert got.shape == z["cubic_%d" % i].shape and np.abs(got - z["cubic_%d" % i]).max() < 1e-10
        if "ola_%d" % i in z.files:
            got = up.overlap_add_upsample(frames, **kw)
            assert got.shape == z["ola_%d" % i].shape and np.abs(got - z["ola_%d" % i]).max() < 1e-12
