"""Development aid for ncu captures: a few whole-batch forwards in the SERIAL order (one launch per kernel: the MLP chain
over all 192 tiles, the noise filter, the three reverb passes), FastNEWT or NEWT (argv[1] = fastnewt | newt)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    variant = sys.argv[1] if len(sys.argv) > 1 else "fastnewt"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda:0")
    model = bench.build_weights()
    if variant == "fastnewt":
        model.newt = FastNEWT(model.newt)
    model = model.to(dev)
    torch.manual_seed(1)
    f0, control = torch.rand(64, 1, 500, device=dev), torch.rand(64, 2, 500, device=dev)
    eng = model._engine_for(f0)
    eng.set_pipeline(False)
    with torch.no_grad():
        for _ in range(n):
            model(f0, control)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
