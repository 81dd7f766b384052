"""Development aid: where the two-lane HostPipeline's per-batch time goes beyond the device-resident throughput —
the same pipeline without the download, with the download in pieces, and with a pinned result buffer per lane only."""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from neural_waveshaping_synthesis_b200 import streaming  # noqa: E402


def run(pipe, f0h, ch, n=200):
    for _ in pipe.run((f0h, ch) for _ in range(8)):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in pipe.run((f0h, ch) for _ in range(n)):
        pass
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    m = bench.build_weights()
    m.newt = FastNEWT(m.newt)
    m = m.to("cuda:0")
    f0h, ch = torch.rand(64, 1, 500).pin_memory(), torch.rand(64, 2, 500).pin_memory()
    base = run(streaming.HostPipeline(m, "cuda:0", lanes=2), f0h, ch)
    print("two lanes, as shipped:            %.4f ms per batch" % base, flush=True)

    orig_copy = torch.Tensor.copy_

    def no_d2h(self, src, non_blocking=False):
        if not self.is_cuda and src.is_cuda and src.numel() > 1 << 20:
            return self          # skip the download
        return orig_copy(self, src, non_blocking=non_blocking)

    def chunked_d2h(self, src, non_blocking=False):
        if not self.is_cuda and src.is_cuda and src.numel() > 1 << 20:
            k = int(os.environ.get("NWS_D2H_CHUNKS", "8"))
            rows = src.shape[0] // k
            for i in range(k):
                orig_copy(self[i * rows:(i + 1) * rows], src[i * rows:(i + 1) * rows], non_blocking=non_blocking)
            return self
        return orig_copy(self, src, non_blocking=non_blocking)

    torch.Tensor.copy_ = no_d2h
    try:
        print("two lanes, download skipped:      %.4f ms per batch" % run(streaming.HostPipeline(m, "cuda:0", lanes=2), f0h, ch), flush=True)
    finally:
        torch.Tensor.copy_ = orig_copy
    torch.Tensor.copy_ = chunked_d2h
    try:
        print("two lanes, download in 8 pieces:  %.4f ms per batch" % run(streaming.HostPipeline(m, "cuda:0", lanes=2), f0h, ch), flush=True)
    finally:
        torch.Tensor.copy_ = orig_copy
    print("three lanes, as shipped:          %.4f ms per batch" % run(streaming.HostPipeline(m, "cuda:0", lanes=3), f0h, ch), flush=True)


if __name__ == "__main__":
    main()
