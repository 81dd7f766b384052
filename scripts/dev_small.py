"""Development aid: one stateless B = 1 forward and one stream push per buffer size (for `ncu --metrics gpu__time_duration.sum`
launch lists) and the library's stage times at those sizes."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    dev = torch.device("cuda:0")
    model = bench.build_weights()
    model.newt = FastNEWT(model.newt)
    model = model.to(dev)
    sizes = [int(a) for a in sys.argv[1:]] or [256, 4096]
    with torch.no_grad():
        for bs in sizes:
            T = bs // 128
            f0, control = torch.rand(1, 1, T, device=dev), torch.rand(1, 2, T, device=dev)
            eng = model._engine_for(f0)
            for _ in range(3):
                model(f0, control)
            eng.set_profiling(True)
            acc = {}
            for _ in range(20):
                model(f0, control)
                for k, v in eng.stage_times_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / 20
            eng.set_profiling(False)
            import time
            for rep in range(2):
                evs = []
                t0 = time.perf_counter()
                for _ in range(100):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); model(f0, control); b.record(); evs.append((a, b))
                host = (time.perf_counter() - t0) / 100
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) for a, b in evs)
                print("bs %d eager median %.1f us, host time per call %.1f us" % (bs, ts[50] * 1e3, host * 1e6), flush=True)
            u_g, nz_g = torch.rand(101, device=dev), torch.rand(128 * T - 1, device=dev)
            model(f0, control, phase_shift=u_g, noise=nz_g)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = model(f0, control, phase_shift=u_g, noise=nz_g)
            g.replay(); torch.cuda.synchronize()
            evs = []
            for _ in range(50):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b.record(); evs.append((a, b))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            print("bs %d graph replay median %.1f us" % (bs, ts[len(ts) // 2] * 1e3), flush=True)
            print("bs %d stages (us):" % bs, {k: round(v * 1e3, 1) for k, v in acc.items() if v}, "sum %.1f" % (sum(acc.values()) * 1e3), flush=True)
            st = model.stream(batch_size=1, max_frames=max(T, 2))
            st.reset()
            for _ in range(3):
                st.push(f0 * 200 + 100, control)
            torch.cuda.synchronize()
            print("bs %d marker: stream pushes done" % bs, flush=True)


if __name__ == "__main__":
    main()
