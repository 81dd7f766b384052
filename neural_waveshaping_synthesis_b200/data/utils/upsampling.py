"""Frame rate -> sample rate: the `interpolate_fn` callables the feature extractors accept
(reference: data/utils/upsampling.py — linear_interpolation :20-36, cubic_spline_interpolation :38-55,
overlap_add_upsample :57-79; same names, arguments and results).

`linear_interpolation` — the extractors' default — runs on the GPU (`nws_interp_frames`, csrc/nws_loudness.cu:
np.interp's float64 arithmetic, one thread per sample); `interp_frames_batch` is its tensor-in / tensor-out form
for pipelines that stay on the device.  The other two are rarely-used alternatives that no gin file selects; they
stay host numpy/scipy (the interface hands them a host array), written over whole arrays instead of the
reference's per-frame Python loop."""
from typing import Optional

import gin
import numpy as np
import torch

from ... import _lib


def _frame_grid(n_frames: int, window_length: int, hop_length: int):
    """(number of output points, positions of those points on the frame axis): the padded signal the frames were
    cut from has n_frames * hop + (window - hop) samples and spans frame 0 .. frame n_frames - 1 end to end."""
    n_points = n_frames * hop_length + (window_length - hop_length)
    return n_points, np.linspace(0.0, n_frames - 1.0, n_points)


def _centre_crop(x: np.ndarray, lead: int, length: Optional[int]):
    return x if not length else x[lead:lead + length]


def interp_frames_batch(frames: torch.Tensor, window_length: int, hop_length: int,
                        original_length: Optional[int] = None) -> torch.Tensor:
    """frames [B, F] float32 CUDA -> [B, samples] float64 CUDA (np.interp semantics, see the module docstring)."""
    if not frames.is_cuda:
        raise ValueError("frames must live on a CUDA device (there is no CPU fallback)")
    frames = frames.to(torch.float32).contiguous()
    if frames.dim() != 2:
        raise ValueError("expected frames of shape [B, F]")
    lib = _lib.load_library()
    B, F = frames.shape
    n = lib.nws_interp_frames_len(F, int(window_length), int(hop_length), int(original_length or 0))
    with torch.cuda.device(frames.device):
        out = torch.empty(B, max(n, 0), dtype=torch.float64, device=frames.device)
        _lib.check(lib.nws_interp_frames(frames.data_ptr(), B, F, int(window_length), int(hop_length),
                                         int(original_length or 0), out.data_ptr(),
                                         torch.cuda.current_stream(frames.device).cuda_stream))
    return out


@gin.configurable
def linear_interpolation(signal: np.ndarray, window_length: int, hop_length: int, original_length: Optional[int] = None):
    if not torch.cuda.is_available():
        raise RuntimeError("linear_interpolation runs on the GPU (nws_interp_frames); there is no CPU fallback")
    frames = torch.from_numpy(np.ascontiguousarray(signal, dtype=np.float32).reshape(1, -1)).cuda()
    return interp_frames_batch(frames, window_length, hop_length, original_length)[0].cpu().numpy()


@gin.configurable
def cubic_spline_interpolation(signal: np.ndarray, window_length: int, hop_length: int,
                               original_length: Optional[int] = None):
    import scipy.interpolate
    _, at = _frame_grid(signal.size, window_length, hop_length)
    # interp1d(kind="cubic") is the not-a-knot interpolating cubic B-spline through the frame values
    spline = scipy.interpolate.make_interp_spline(np.arange(signal.size, dtype=np.float64), signal, k=3)
    return _centre_crop(spline(at), window_length // 2, original_length)


@gin.configurable
def overlap_add_upsample(signal: np.ndarray, window_length: int, hop_length: int, window_fn: str = "hann",
                         window_scale: int = 2, original_length: Optional[int] = None):
    import scipy.signal.windows
    span = hop_length * window_scale
    bump = scipy.signal.windows.get_window(window_fn, span)
    n_points, _ = _frame_grid(signal.size, window_length, hop_length)
    # every frame value scales one window placed at its hop; frames are added in order, like the reference's loop,
    # clipped where a window would run past the padded length
    starts = np.arange(signal.size) * hop_length
    acc = np.zeros(max(n_points, int(starts[-1]) + span) if signal.size else n_points)
    for lane in range(window_scale):       # windows `window_scale` frames apart never overlap: add them as one strided block
        sel = np.arange(lane, signal.size, window_scale)
        if sel.size:
            block = (signal[sel, None] * bump[None, :]).reshape(-1)
            view = acc[starts[sel[0]]: starts[sel[0]] + block.size]
            view += block[: view.size]
    acc = acc[:n_points]
    return _centre_crop(acc, (n_points - original_length) // 2 if original_length else 0, original_length)
