"""Timing helpers shared by this repo's scripts/: CUDA-event timing on the current stream (the
reference's scripts use wall clock without a device synchronise, which on a GPU measures launch
time only — SURVEY.md App. C.2)."""
import time

import torch


def time_forward(fn, iters: int, device, warmup: int = 3):
    """Seconds per call of `fn` (list of `iters` floats): CUDA events on CUDA devices, perf_counter on CPU."""
    device = torch.device(device)
    for _ in range(warmup):
        fn()
    out = []
    if device.type == "cuda":
        torch.cuda.synchronize(device)
        pairs = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            pairs.append((a, b))
        torch.cuda.synchronize(device)
        out = [a.elapsed_time(b) * 1e-3 for a, b in pairs]
    else:
        for _ in range(iters):
            t0 = time.perf_counter()
            fn()
            out.append(time.perf_counter() - t0)
    return out


def build_model(gin_file: str, use_fast_newt: bool, device, checkpoint: str = None):
    import gin
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    from neural_waveshaping_synthesis.models.neural_waveshaping import NeuralWaveshaping
    gin.parse_config_file(gin_file)
    model = NeuralWaveshaping.load_from_checkpoint(checkpoint) if checkpoint else NeuralWaveshaping()
    model.eval()
    if use_fast_newt:
        model.newt = FastNEWT(model.newt)
    return model.to(device)
