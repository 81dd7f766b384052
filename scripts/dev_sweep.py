"""Buffer-size sweep (BASELINE config C4): latency of independent forwards at B=1 — CUDA-event timed eager
launches, wall clock, and CUDA-graph replay (the forward is capture-safe)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_waveshaping_synthesis_b200.timing import build_model, time_forward  # noqa: E402

out = {}
for fast in (False, True):
    import gin
    gin.clear_config()
    torch.manual_seed(0)
    model = build_model(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gin", "models", "newt.gin"),
                        fast, "cuda:0")
    with torch.no_grad():
        for bs in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 64000):
            T = bs // 128
            f0, c = torch.rand(1, 1, T, device="cuda"), torch.rand(1, 2, T, device="cuda")
            u, nz = torch.rand(101, device="cuda"), torch.rand(128 * T - 1, device="cuda")
            secs = np.array(time_forward(lambda: model(f0, c), 50, "cuda:0", warmup=10))
            t0 = time.perf_counter()
            for _ in range(50):
                model(f0, c)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 50
            model(f0, c, phase_shift=u, noise=nz)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                model(f0, c, phase_shift=u, noise=nz)
            gs = np.array(time_forward(g.replay, 50, "cuda:0", warmup=5))
            key = "%s_%d" % ("fast" if fast else "newt", bs)
            out[key] = {"event_ms_median": float(np.median(secs) * 1e3), "wall_ms_mean": wall * 1e3,
                        "graph_ms_median": float(np.median(gs) * 1e3)}
            print("%s bs=%6d  eager %.3f ms  wall %.3f ms  graph replay %.3f ms  (RTF %.5f)" %
                  ("FastNEWT" if fast else "NEWT", bs, np.median(secs) * 1e3, wall * 1e3, np.median(gs) * 1e3,
                   np.median(gs) / (bs / 16000)), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
