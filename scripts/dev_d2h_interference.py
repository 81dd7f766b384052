"""Does a concurrent D2H copy stretch any stage of the forward?  Stage times (serial order, profiling events)
with and without an independent 16 MB D2H running on another stream from the start of the forward."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT  # noqa: E402

dev = torch.device("cuda", 0)
model = copy.deepcopy(bench.build_weights())
model.newt = FastNEWT(model.newt)
model = model.to(dev)
B, T = 64, 500
f0, control = torch.rand(B, 1, T, device=dev), torch.rand(B, 2, T, device=dev)
src = torch.rand(B, T * 128, device=dev)
dst = torch.empty(B, T * 128).pin_memory()
hsrc = torch.rand(B, 3, T).pin_memory()
hdst = torch.empty(B, 3, T, device=dev)
cs = torch.cuda.Stream(dev)
with torch.no_grad():
    for _ in range(5):
        model(f0, control)
    eng = model._engine_for(f0)
    eng.set_profiling(True)
    for mode in ("alone", "with D2H", "with D2H x3 (0.9 ms)", "alone"):
        acc = {}
        for _ in range(20):
            torch.cuda.synchronize()
            if mode != "alone":
                with torch.cuda.stream(cs):
                    for _ in range(3 if "x3" in mode else 1):
                        dst.copy_(src, non_blocking=True)
            model(f0, control)
            torch.cuda.synchronize()
            for k, v in eng.stage_times_ms().items():
                acc[k] = acc.get(k, 0.0) + v / 20
        print("%-22s" % mode, " ".join("%s %.3f" % (k, v) for k, v in acc.items()), " sum %.3f" % sum(acc.values()), flush=True)
    eng.set_profiling(False)
    # whole pipelined forward, event-timed, with a D2H started at the same time
    for mode in ("alone", "with D2H", "with D2H x3 (0.9 ms)", "alone"):
        ts = []
        for _ in range(20):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if mode != "alone":
                with torch.cuda.stream(cs):
                    for _ in range(3 if "x3" in mode else 1):
                        dst.copy_(src, non_blocking=True)
            a.record()
            model(f0, control)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("pipelined %-22s %.3f ms" % (mode, sum(ts) / len(ts)), flush=True)
