"""Control-side loudness extraction on the GPU — mirror of the reference's data/utils/loudness_extraction.py:11-90
(same function names, arguments, defaults and return shapes), computed by csrc/nws_loudness.cu through the C ABI
(`nws_extract_loudness`, `nws_extract_rms`).  The reference's arithmetic here is librosa 0.8.0's (requirements.txt:5);
see the kernel file for the recipe.  No CPU fallback: a CUDA device and the built library are required.

numpy in / numpy out like the reference; the `*_batch` functions take and return CUDA tensors ([B, N] -> [B, frames])
and never leave the device — that is what a pipeline feeding `NeuralWaveshaping.forward` should call."""
from typing import Callable, Optional

import gin
import numpy as np
import torch

from ... import _lib
from .upsampling import linear_interpolation


def _as_cuda_batch(audio) -> torch.Tensor:
    if isinstance(audio, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("loudness extraction needs a CUDA device (there is no CPU fallback)")
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda()
    if not audio.is_cuda:
        raise ValueError("audio tensor must live on a CUDA device (there is no CPU fallback)")
    if audio.dtype != torch.float32:
        raise ValueError("audio must be float32")
    return audio.contiguous()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def perceptual_loudness_batch(audio: torch.Tensor, n_fft: int = 2048, hop_length: int = 512, epsilon: float = 1e-5,
                              normalise: bool = True, return_spectrogram: bool = False):
    """audio [B, N] float32 CUDA -> loudness [B, 1 + N // hop_length] (and, optionally, the dB spectrogram
    [B, n_fft // 2 + 1, frames] that compute_power_spectrogram returns)."""
    audio = _as_cuda_batch(audio)
    if audio.dim() != 2:
        raise ValueError("expected audio of shape [B, N]")
    lib = _lib.load_library()
    B, N = audio.shape
    frames, bins = 1 + N // hop_length, n_fft // 2 + 1
    with torch.cuda.device(audio.device):
        nbytes = lib.nws_loudness_workspace_bytes(B, N, n_fft, hop_length)
        ws = torch.empty(max(int(nbytes), 4), dtype=torch.uint8, device=audio.device)
        out = torch.empty(B, frames, dtype=torch.float32, device=audio.device)
        db = torch.empty(B, frames, bins, dtype=torch.float32, device=audio.device) if return_spectrogram else None
        _lib.check(lib.nws_extract_loudness(audio.data_ptr(), B, N, n_fft, hop_length, float(epsilon), 1 if normalise else 0,
                                            out.data_ptr(), db.data_ptr() if db is not None else None,
                                            ws.data_ptr(), ws.numel(), _stream(audio)))
    if return_spectrogram:
        return out, db.transpose(1, 2)
    return out


def rms_batch(audio: torch.Tensor, window_size: int = 2048, hop_length: int = 512) -> torch.Tensor:
    """audio [B, N] float32 CUDA -> rms [B, frames]."""
    audio = _as_cuda_batch(audio)
    if audio.dim() != 2:
        raise ValueError("expected audio of shape [B, N]")
    lib = _lib.load_library()
    B, N = audio.shape
    frames = 1 + (N + 2 * (window_size // 2) - window_size) // hop_length
    with torch.cuda.device(audio.device):
        out = torch.empty(B, max(frames, 0), dtype=torch.float32, device=audio.device)
        _lib.check(lib.nws_extract_rms(audio.data_ptr(), B, N, window_size, hop_length, out.data_ptr(), _stream(audio)))
    return out


def _check_window(window: str):
    if window != "hann":
        raise NotImplementedError("only the reference's window (\"hann\", loudness_extraction.py:48) is built; got %r" % (window,))


def compute_power_spectrogram(audio: np.ndarray, n_fft: int, hop_length: int, window: str, epsilon: float):
    """loudness_extraction.py:11-23: dB spectrogram [n_fft // 2 + 1, frames] relative to its own maximum."""
    _check_window(window)
    _, db = perceptual_loudness_batch(_as_cuda_batch(audio).view(1, -1), n_fft, hop_length, epsilon, False, True)
    return db[0].contiguous().cpu().numpy()


def perform_perceptual_weighting(power_spectrogram_in_db: np.ndarray, sample_rate: float, n_fft: int):
    """loudness_extraction.py:26-40.  The reference computes the A-weighting curve and then leaves it out
    (`weighted_spectrogram = power_spectrogram_in_db  # + weights`, :39); kept as is for drop-in results."""
    return power_spectrogram_in_db


@gin.configurable
def extract_perceptual_loudness(audio: np.ndarray, sample_rate: float = 16000, n_fft: int = 2048, hop_length: int = 512,
                                window: str = "hann", epsilon: float = 1e-5,
                                interpolate_fn: Optional[Callable] = linear_interpolation, normalise: bool = True):
    """loudness_extraction.py:43-68 for one 1-D signal: per-frame loudness (frames = 1 + len // hop_length), optionally
    interpolated to the sample rate by `interpolate_fn` (host callable, as in the reference) and mapped to (x + 80) / 80."""
    _check_window(window)
    a = _as_cuda_batch(audio)
    if a.dim() != 1:
        raise ValueError("expected a 1-D signal")
    loudness = perceptual_loudness_batch(a.view(1, -1), n_fft, hop_length, epsilon, False)[0].cpu().numpy()
    if interpolate_fn:
        loudness = interpolate_fn(loudness, n_fft, hop_length, original_length=a.numel())
    if normalise:
        loudness = (loudness + 80) / 80
    return loudness


@gin.configurable
def extract_rms(audio: np.ndarray, window_size: int = 2048, hop_length: int = 512, sample_rate: Optional[float] = 16000.0,
                interpolate_fn: Optional[Callable] = linear_interpolation):
    """loudness_extraction.py:71-90."""
    a = _as_cuda_batch(audio)
    if a.dim() != 1:
        raise ValueError("expected a 1-D signal")
    root = rms_batch(a.view(1, -1), window_size, hop_length)[0].cpu().numpy()
    if interpolate_fn:
        assert sample_rate is not None, "Must provide sample rate if upsampling"
        root = interpolate_fn(root, window_size, hop_length, original_length=a.numel())
    return root
