"""Streaming latency (BASELINE config C4 as a real stream): one utterance synthesised buffer by buffer through
`model.stream()` (nws_stream_push: state carried on the device) at the buffer sizes of the reference's
scripts/time_buffer_sizes.py:13, next to the stateless forward the reference script times.  CUDA-event timed.

    PYTHONPATH=. python scripts/time_streaming.py [--pushes 200] [--json out.json]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_waveshaping_synthesis_b200.timing import build_model, time_forward  # noqa: E402

BUFFER_SIZES = [256, 512, 1024, 2048, 4096]
SR = 16000


def main():
    pushes = int(sys.argv[sys.argv.index("--pushes") + 1]) if "--pushes" in sys.argv else 200
    gin_file = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gin", "models", "newt.gin")
    rows = {}
    for fast in (True, False):
        import gin
        gin.clear_config()
        torch.manual_seed(0)
        model = build_model(gin_file, fast, "cuda:0")
        with torch.no_grad():
            for bs in BUFFER_SIZES:
                n = bs // 128
                f0, c = torch.rand(1, 1, n, device="cuda"), torch.rand(1, 2, n, device="cuda")
                st = model.stream(batch_size=1, max_frames=n)
                st.reset()
                secs = np.array(time_forward(lambda: st.push(f0, c), pushes, "cuda:0", warmup=10))
                lat = []
                import time
                for _ in range(pushes):       # submit -> audio complete on the device, one push at a time
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    st.push(f0, c)
                    torch.cuda.synchronize()
                    lat.append(time.perf_counter() - t0)
                lat = np.array(lat)
                stateless = np.array(time_forward(lambda: model(f0, c), 50, "cuda:0", warmup=10))
                key = "%s_%d" % ("fast" if fast else "newt", bs)
                rows[key] = {"push_ms_median": float(np.median(secs) * 1e3), "push_ms_p90": float(np.percentile(secs, 90) * 1e3),
                             "push_latency_ms_median": float(np.median(lat) * 1e3), "push_latency_ms_p99": float(np.percentile(lat, 99) * 1e3),
                             "stateless_forward_ms_median": float(np.median(stateless) * 1e3),
                             "buffer_ms": 1e3 * bs / SR, "rtf": float(np.median(secs) / (bs / SR))}
                print("%-8s buffer %5d (%6.2f ms of audio): push %.3f ms (p90 %.3f)  RTF %.4f  submit->done %.3f ms (p99 %.3f)   stateless forward %.3f ms" %
                      ("FastNEWT" if fast else "NEWT", bs, 1e3 * bs / SR, rows[key]["push_ms_median"],
                       rows[key]["push_ms_p90"], rows[key]["rtf"], rows[key]["push_latency_ms_median"], rows[key]["push_latency_ms_p99"],
                       rows[key]["stateless_forward_ms_median"]), flush=True)
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
