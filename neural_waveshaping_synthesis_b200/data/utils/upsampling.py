"""Frame-rate -> sample-rate interpolation helpers — mirror of the reference's data/utils/upsampling.py:9-36
(the `interpolate_fn` callables that extract_perceptual_loudness / extract_rms accept).  They run on the host in
numpy exactly as in the reference: the extractors take an arbitrary Python callable here, so this step is host
code by the interface's own definition (the gin data config sets it to None, urmp_4second_crepe.gin:2,12)."""
from typing import Optional

import gin
import numpy as np


def get_padded_length(frames: int, window_length: int, hop_length: int):
    return frames * hop_length + window_length - hop_length


def get_source_target_axes(frames: int, window_length: int, hop_length: int):
    padded_length = get_padded_length(frames, window_length, hop_length)
    return np.linspace(0, frames - 1, frames), np.linspace(0, frames - 1, padded_length)


@gin.configurable
def linear_interpolation(signal: np.ndarray, window_length: int, hop_length: int, original_length: Optional[int] = None):
    source_x, target_x = get_source_target_axes(signal.size, window_length, hop_length)
    interpolated = np.interp(target_x, source_x, signal)
    if original_length:
        interpolated = interpolated[window_length // 2:]
        interpolated = interpolated[:original_length]
    return interpolated
