"""Buffer-size sweep (BASELINE config C4): latency of independent forwards, B=1, event-timed."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_waveshaping_synthesis_b200.timing import build_model, time_forward

out = {}
for fast in (False, True):
    import gin
    gin.clear_config()
    torch.manual_seed(0)
    model = build_model("gin/models/newt.gin", fast, "cuda:0")
    with torch.no_grad():
        for bs in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 64000):
            T = bs // 128
            f0, c = torch.rand(1, 1, T, device="cuda"), torch.rand(1, 2, T, device="cuda")
            secs = np.array(time_forward(lambda: model(f0, c), 50, "cuda:0", warmup=10))
            import time
            t0 = time.perf_counter()
            for _ in range(50):
                model(f0, c)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 50
            out["%s_%d" % ("fast" if fast else "newt", bs)] = {"event_ms_median": float(np.median(secs) * 1e3), "wall_ms_mean": wall * 1e3}
            print("%s bs=%6d  event median %.3f ms  wall mean %.3f ms  (RTF %.5f)" % ("FastNEWT" if fast else "NEWT", bs, np.median(secs) * 1e3, wall * 1e3, wall / (bs / 16000)), flush=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
