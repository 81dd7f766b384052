"""Signal generators — host-side mirror of the reference's modules/generators.py
(FIRNoiseSynth :11-35, HarmonicOscillator :38-66).  Parameters/buffers and constructor signatures
match; the arithmetic runs in csrc/nws_noise.cu and, for the oscillator bank, inside the fused
audio-rate kernel (csrc/nws_audio.cu)."""
import math
from typing import Callable

import gin
import torch
import torch.nn as nn

from ._bound import BoundToRoot


@gin.configurable
class FIRNoiseSynth(nn.Module, BoundToRoot):
    def __init__(self, ir_length: int, hop_length: int, window_fn: Callable = torch.hann_window):
        super().__init__()
        self.ir_length = ir_length
        self.hop_length = hop_length
        self.register_buffer("window", window_fn(ir_length))

    def forward(self, H_re, noise=None):
        """H_re [B,129,T] -> [B,1,128*T].  `noise` (optional, [128*T-1]) replaces the uniform draw the
        reference makes at generators.py:30."""
        root = self._root()
        eng = root._engine_for(H_re)
        if noise is None:
            noise = torch.rand(self.hop_length * H_re.shape[-1] - 1, device=H_re.device)
        return eng.noise(H_re, noise).unsqueeze(1)


@gin.configurable
class HarmonicOscillator(nn.Module):
    def __init__(self, n_harmonics, sample_rate):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_harmonics = n_harmonics
        self.register_buffer("harmonic_axis", torch.arange(1, n_harmonics + 1).view(1, -1, 1))
        self.register_buffer("rand_phase", torch.ones(1, n_harmonics, 1) * math.tau)

    def forward(self, f0):
        raise NotImplementedError(
            "the oscillator bank is never materialised: it is generated inside the fused audio-rate kernel "
            "(csrc/nws_audio.cu); call NeuralWaveshaping.forward")
