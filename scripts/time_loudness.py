"""Throughput of the control-side loudness extractor: a batch of 4 s segments through
`perceptual_loudness_batch` (csrc/nws_loudness.cu), CUDA-event timed with the audio resident on the device.
gin/data/urmp_4second_crepe.gin settings (n_fft 1024, hop 128).

    PYTHONPATH=. python scripts/time_loudness.py [--batch 64] [--iters 50] [--json out.json]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_waveshaping_synthesis.data.utils.loudness_extraction import perceptual_loudness_batch  # noqa: E402


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    B, iters, N, n_fft, hop = arg("--batch", 64), arg("--iters", 50), 64000, 1024, 128
    torch.manual_seed(0)
    audio = (torch.rand(B, N, device="cuda") * 2 - 1) * 0.3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        perceptual_loudness_batch(audio, n_fft, hop)
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        perceptual_loudness_batch(audio, n_fft, hop)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    row = {"batch": B, "samples_per_segment": N, "n_fft": n_fft, "hop_length": hop, "ms_median": float(np.median(ms)),
           "ms_p90": float(np.percentile(ms, 90)), "audio_samples_per_s": float(B * N / (np.median(ms) * 1e-3)),
           "l2": "flushed between iterations (256 MiB write)"}
    print(json.dumps(row))
    if "--json" in sys.argv:
        json.dump(row, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
