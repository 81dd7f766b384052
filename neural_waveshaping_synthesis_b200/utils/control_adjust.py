"""Control-signal adjustment of the timbre-transfer notebook (colab/NEWT_Timbre_Transfer.ipynb, cell 15): the host-side
step between feature extraction (f0 + confidence from CREPE, loudness from `extract_perceptual_loudness`) and
`NeuralWaveshaping.forward`.  In the reference this is notebook code made of elementwise torch/numpy ops and a box
filter; it is restated here as one function on tensors so a pipeline can keep the features on the device:

    loudness = perceptual_loudness_batch(audio)                      # csrc/nws_loudness.cu
    f0_t, control = adjust_controls(f0, loudness, confidence, data_mean, data_std, octave_shift=1)
    audio_out = model(f0_t.view(1, 1, -1), control.unsqueeze(0))     # the fused forward

Plain torch (plumbing, any device); no kernel of ours is involved."""
from typing import Tuple

import torch


def _box_filter(x: torch.Tensor, radius: int) -> torch.Tensor:
    """conv1d with a ones kernel of 2*radius+1 taps / (2*radius+1) and zero padding `radius` (cell 15)."""
    k = 2 * radius + 1
    w = torch.ones(1, 1, k, device=x.device, dtype=x.dtype) / k
    return torch.nn.functional.conv1d(x.expand(1, 1, -1), w, padding=radius).squeeze()


def adjust_controls(f0: torch.Tensor, loudness: torch.Tensor, confidence: torch.Tensor, data_mean, data_std,
                    octave_shift: int = 1, loudness_scale: float = 0.5, loudness_floor: float = 0.0,
                    loudness_conf_filter: float = 0.0, pitch_conf_filter: float = 0.0, pitch_smoothing: int = 0,
                    loudness_smoothing: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """f0 [T] Hz, loudness [T] (normalised, as extract_perceptual_loudness returns it), confidence [T];
    data_mean / data_std: the checkpoint's feature statistics (index 0 = f0, 1 = loudness).
    Returns (f0_t [T] in Hz — what the notebook passes as `f0` — and control [2, T] = (normalised f0, normalised
    loudness)), float32, on f0's device.  Defaults are the notebook's slider defaults."""
    f0 = torch.as_tensor(f0)
    dev = f0.device
    loudness = torch.as_tensor(loudness, device=dev)
    confidence = torch.as_tensor(confidence, device=dev)
    stat = lambda a, i: float(torch.as_tensor(a).reshape(-1)[i])   # the checkpoints store [n_features, 1] arrays
    mean0, mean1, std0, std1 = stat(data_mean, 0), stat(data_mean, 1), stat(data_std, 0), stat(data_std, 1)
    f0_filtered = f0 * (confidence > pitch_conf_filter)
    loudness_filtered = loudness * (confidence > loudness_conf_filter)
    f0_shifted = f0_filtered * (2 ** octave_shift)
    loudness_floored = loudness_filtered * (loudness_filtered > loudness_floor) - loudness_floor
    loudness_scaled = loudness_floored * loudness_scale
    loud_norm_t = ((loudness_scaled - mean1) / std1).float()
    f0_t = f0_shifted.float()
    if pitch_smoothing != 0:
        f0_t = _box_filter(f0_t, int(pitch_smoothing))
    if loudness_smoothing != 0:
        loud_norm_t = _box_filter(loud_norm_t, int(loudness_smoothing))
    f0_norm_t = ((f0_t - mean0) / std0).float()
    return f0_t, torch.stack((f0_norm_t, loud_norm_t), dim=0)
