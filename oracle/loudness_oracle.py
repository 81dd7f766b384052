"""ORACLE — test infrastructure, NOT product code.

CPU restatement (numpy) of the control-side loudness extractors of
ben-hayes/neural-waveshaping-synthesis — SURVEY.md §8(f) rank 4, the step before the hot path in the
timbre-transfer use case (colab cell 14, scripts/create_dataset.py):

    extract_perceptual_loudness   neural_waveshaping_synthesis/data/utils/loudness_extraction.py:43-68
    compute_power_spectrogram     .../loudness_extraction.py:11-23
    perform_perceptual_weighting  .../loudness_extraction.py:26-40
    extract_rms                   .../loudness_extraction.py:71-90
    linear_interpolation          neural_waveshaping_synthesis/data/utils/upsampling.py:20-36

Parity status: the reference's own code (the glue: which librosa calls, with which arguments, frame counts, the mean
over bins, normalisation, interpolation, rms framing) is PINNED — ``oracle/gen_golden_loudness.py`` runs the real
loudness_extraction.py / upsampling.py and ``tests/test_loudness_cpu.py`` requires this restatement to reproduce
their outputs exactly (tests/golden/loudness_*.npz).  The layer below it is UNPINNED: all arithmetic of these functions
lives in a third-party dependency that is absent from /root/reference and from this image: librosa (pinned
``librosa==0.8.0``, requirements.txt:5), which the fixture generator has to replace by the restatement below.
Its published algorithm is restated here from the 0.8.0 sources —

    librosa.stft              core/spectrum.py   window = scipy get_window(name, n_fft, fftbins=True);
                                                 y padded by n_fft//2 with np.pad(mode="reflect"); frames of
                                                 n_fft every hop_length; numpy rfft (float64) of window*frame
                                                 stored as complex64; n_frames = 1 + len(y) // hop_length
    librosa.amplitude_to_db   core/spectrum.py   magnitude = |S|; ref_value = ref(magnitude); power = magnitude**2;
                                                 power_to_db(power, ref=ref_value**2, amin=amin**2, top_db)
    librosa.power_to_db       core/spectrum.py   10*log10(max(amin, S)) - 10*log10(max(amin, ref)), then
                                                 max(., max(.) - top_db)   (top_db = 80)
    librosa.util.frame        util/utils.py      [frame_length, n_frames] strided view

— and anchored on the reference's own call sites (the argument values it passes, gin/data/urmp_4second_crepe.gin:
11-14).  ``tests/test_loudness_cpu.py`` cross-checks the STFT against an independent implementation
(torch.stft, float64) so at least the transform definition is not self-referential.

Only ``tests/`` may import this module.  The product path never does and has no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import scipy.signal


# ------------------------------------------------------------------ upsampling.py:9-36
def get_padded_length(frames: int, window_length: int, hop_length: int) -> int:
    return frames * hop_length + window_length - hop_length


def linear_interpolation(signal: np.ndarray, window_length: int, hop_length: int,
                         original_length: Optional[int] = None) -> np.ndarray:
    padded_length = get_padded_length(signal.size, window_length, hop_length)
    source_x = np.linspace(0, signal.size - 1, signal.size)
    target_x = np.linspace(0, signal.size - 1, padded_length)
    interpolated = np.interp(target_x, source_x, signal)
    if original_length:
        interpolated = interpolated[window_length // 2:]
        interpolated = interpolated[:original_length]
    return interpolated


# ------------------------------------------------------------------ librosa 0.8.0 restated
def stft(y: np.ndarray, n_fft: int, hop_length: int, window: str = "hann") -> np.ndarray:
    """librosa.stft(y, n_fft, hop_length, window=window) with its defaults (win_length = n_fft, center=True,
    pad_mode="reflect", dtype complex64 for float32 input).  Returns [1 + n_fft//2, 1 + len(y)//hop_length]."""
    y = np.asarray(y)
    fft_window = scipy.signal.get_window(window, n_fft, fftbins=True).reshape((-1, 1))   # float64
    y_pad = np.pad(y, int(n_fft // 2), mode="reflect")
    n_frames = 1 + (y_pad.shape[-1] - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    y_frames = y_pad[idx]                                                                 # [n_fft, n_frames]
    dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    return np.fft.rfft(fft_window * y_frames, axis=0).astype(dtype)


def power_to_db(S: np.ndarray, ref: float, amin: float, top_db: Optional[float] = 80.0) -> np.ndarray:
    magnitude = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S: np.ndarray, ref: Callable = np.max, amin: float = 1e-5, top_db: Optional[float] = 80.0) -> np.ndarray:
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude)
    power = np.square(magnitude, out=magnitude)
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


# ------------------------------------------------------------------ loudness_extraction.py
def compute_power_spectrogram(audio: np.ndarray, n_fft: int, hop_length: int, window: str, epsilon: float) -> np.ndarray:
    """loudness_extraction.py:11-23."""
    spectrogram = stft(audio, n_fft=n_fft, hop_length=hop_length, window=window)
    magnitude_spectrogram = np.abs(spectrogram)
    return amplitude_to_db(magnitude_spectrogram, ref=np.max, amin=epsilon)


def perform_perceptual_weighting(power_spectrogram_in_db: np.ndarray, sample_rate: float, n_fft: int) -> np.ndarray:
    """loudness_extraction.py:26-40.  The reference computes librosa.A_weighting(fft_frequencies) and then does NOT
    add it (`weighted_spectrogram = power_spectrogram_in_db  # + weights`, :39): the spectrogram is returned as is."""
    return power_spectrogram_in_db


def extract_perceptual_loudness(audio: np.ndarray, sample_rate: float = 16000, n_fft: int = 2048, hop_length: int = 512,
                                window: str = "hann", epsilon: float = 1e-5,
                                interpolate_fn: Optional[Callable] = linear_interpolation,
                                normalise: bool = True) -> np.ndarray:
    """loudness_extraction.py:43-68."""
    power_spectrogram = compute_power_spectrogram(audio, n_fft=n_fft, hop_length=hop_length, window=window, epsilon=epsilon)
    weighted = perform_perceptual_weighting(power_spectrogram, sample_rate=sample_rate, n_fft=n_fft)
    loudness = np.mean(weighted, axis=0)
    if interpolate_fn:
        loudness = interpolate_fn(loudness, n_fft, hop_length, original_length=audio.size)
    if normalise:
        loudness = (loudness + 80) / 80
    return loudness


def extract_rms(audio: np.ndarray, window_size: int = 2048, hop_length: int = 512, sample_rate: Optional[float] = 16000.0,
                interpolate_fn: Optional[Callable] = linear_interpolation) -> np.ndarray:
    """loudness_extraction.py:71-90 (librosa.util.frame restated as an index gather)."""
    padded_audio = np.pad(audio, (window_size // 2, window_size // 2))
    n_frames = 1 + (padded_audio.shape[-1] - window_size) // hop_length
    idx = np.arange(window_size)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = padded_audio[idx]
    squared = frames ** 2
    mean = np.mean(squared, axis=0)
    root = np.sqrt(mean)
    if interpolate_fn:
        assert sample_rate is not None, "Must provide sample rate if upsampling"
        root = interpolate_fn(root, window_size, hop_length, original_length=audio.size)
    return root
