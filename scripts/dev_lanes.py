"""Development aid: two-lane throughput (bench.lane_throughput_ms) and single-forward latency of the headline workload under
the current environment switches (NWS_PIPE_FIRST / NWS_PIPE_BLOCK / NWS_TILE_CHUNK ...)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    dev = torch.device("cuda:0")
    variant = sys.argv[1] if len(sys.argv) > 1 else "fastnewt"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    model = bench.build_weights()
    if variant == "fastnewt":
        model.newt = FastNEWT(model.newt)
    model = model.to(dev)
    T = 500

    def make_inputs(i):
        g = torch.Generator(device=dev).manual_seed(100 + i)
        return torch.rand(B, 1, T, device=dev, generator=g), torch.rand(B, 2, T, device=dev, generator=g)
    n_sets = (160 << 20) // (B * 3 * T * 4) + 1
    thr = min(bench.lane_throughput_ms(model, dev, make_inputs, 100, n_sets) for _ in range(2))
    f0, control = make_inputs(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    with torch.no_grad():
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); model(f0, control); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    ts.sort()
    env = {k: v for k, v in os.environ.items() if k.startswith("NWS_")}
    print("%s B %d %s: throughput %.4f ms / batch, latency %.4f ms" % (variant, B, env, thr, ts[len(ts) // 2]), flush=True)


if __name__ == "__main__":
    main()
