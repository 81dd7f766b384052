"""Development aid: the tensor-core recurrence (nws_set_gru_impl(1)) against the fp32 one — embedding difference on the
trained violin weights and the random-init ones, the GRU's stage time and the whole forward, per batch size.
NWS_GRU_ACT=0|1|2 selects the activation variant (csrc/nws_gru_mma.cu)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    dev = torch.device("cuda:0")
    out = {"act": os.environ.get("NWS_GRU_ACT", "1")}
    kat = np.load(os.path.join(REPO, "tests", "golden", "kat_vn_fast.npz"))
    for tag in ("vn", "randinit"):
        z = np.load(os.path.join(REPO, "tests", "golden", "weights_%s.npz" % tag))
        model = bench.build_weights()
        model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("data_")})
        model.eval()
        model.newt = FastNEWT(model.newt)
        model = model.to(dev)
        for B in (2, 5, 8, 64, 256):
            g = torch.Generator().manual_seed(B)
            if tag == "vn":
                c1 = torch.from_numpy(kat["control"])   # [1,2,500] realistic normalised controls
                f1 = torch.from_numpy(kat["f0"])
                scale = 1.0 + 0.05 * torch.randn(B, 1, 1, generator=g)
                control = (c1 * scale + 0.02 * torch.randn(B, 2, 500, generator=g)).to(dev)
                f0 = (f1 * (1.0 + 0.2 * torch.rand(B, 1, 1, generator=g))).to(dev)
            else:
                control = torch.rand(B, 2, 500, generator=g).to(dev)
                f0 = torch.rand(B, 1, 500, generator=g).to(dev)
            eng = model._engine_for(f0)
            res = {}
            embs = {}
            for impl in (0, 1):
                eng.set_gru_impl(impl)
                embs[impl] = eng.control_embedding(control)
                with torch.no_grad():
                    u, nz = torch.rand(101, device=dev), torch.rand(128 * 500 - 1, device=dev)
                    y = model(f0, control, phase_shift=u, noise=nz)
                    res["audio%d" % impl] = y
                    eng.set_profiling(True)
                    acc = 0.0
                    for _ in range(5):
                        model(f0, control)
                        acc += eng.stage_times_ms()["gru"] / 5
                    eng.set_profiling(False)
                    for _ in range(3):
                        model(f0, control)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(10):
                        model(f0, control)
                    b.record()
                    torch.cuda.synchronize()
                res["gru_ms_impl%d" % impl] = acc
                res["forward_ms_impl%d" % impl] = a.elapsed_time(b) / 10
            r = {k: v for k, v in res.items() if not k.startswith("audio")}
            r["emb_max_abs_diff"] = float((embs[0] - embs[1]).abs().max())
            r["emb_max_abs"] = float(embs[0].abs().max())
            r["audio_max_abs_diff"] = float((res["audio0"] - res["audio1"]).abs().max())
            out["%s_B%d" % (tag, B)] = r
            print(tag, B, r, flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "dev_gru_act%s.json" % out["act"]), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
