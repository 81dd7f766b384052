"""Development aid: forward time of the pipelined order per GRU implementation / time-block size (NWS_PIPE_BLOCK), with the
difference between the two implementations on identical draws."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for variant in ("fastnewt", "newt"):
        model = bench.build_weights()
        if variant == "fastnewt":
            model.newt = FastNEWT(model.newt)
        model = model.to(dev)
        for B in (64, 256):
            torch.manual_seed(B)
            f0, control = torch.rand(B, 1, 500, device=dev), torch.rand(B, 2, 500, device=dev)
            u, nz = torch.rand(101, device=dev), torch.rand(128 * 500 - 1, device=dev)
            eng = model._engine_for(f0)
            ys, line = {}, []
            for impl, nzf in ((0, 0), (0, 1), (1, 0), (1, 1)):
                eng.set_gru_impl(impl)
                eng.set_noise_fused(bool(nzf))
                with torch.no_grad():
                    ys[impl + 2 * nzf] = model(f0, control, phase_shift=u, noise=nz).clone()
                    for _ in range(3):
                        model(f0, control)
                    ts = []
                    for _ in range(20):
                        flush.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(); model(f0, control); b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                    eng.set_profiling(True)
                    acc = {}
                    for _ in range(5):
                        model(f0, control)
                        for k, v in eng.stage_times_ms().items():
                            acc[k] = acc.get(k, 0.0) + v / 5
                    eng.set_profiling(False)
                print("   stages gru%d nzfused%d:" % (impl, nzf), {k: round(v, 4) for k, v in acc.items() if v}, flush=True)
                ts.sort()
                line.append("gru%d nzfused%d %.4f ms (min %.4f)" % (impl, nzf, ts[len(ts) // 2], ts[0]))
            print("blk", os.environ.get("NWS_PIPE_BLOCK", "default"), variant, "B", B, " | ".join(line),
                  "audio diff gru %.3g fused %.3g" % (float((ys[0] - ys[1]).abs().max()), float((ys[0] - ys[2]).abs().max())), flush=True)


if __name__ == "__main__":
    main()
