"""Test tool (run by hand: python tests/stream_stress.py [reps]).  Determinism stress of the streaming path: the same chunked run repeated, every repetition must be bit-identical."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nws_oracle as oracle  # noqa: E402
from tests.test_gpu_parity import _model, _run_stream  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for tag, fast, chunks in (("vn", True, [8, 8]), ("vn", False, [5, 11]), ("randinit", True, [2, 2, 2, 2, 2])):
    m, w = _model(tag, fast)
    T = sum(chunks)
    if tag == "randinit":
        g = torch.Generator().manual_seed(5)
        f0, control = torch.rand(2, 1, T, generator=g), torch.rand(2, 2, T, generator=g)
    else:
        f0, control = oracle.realistic_inputs(T, w["data_mean"].numpy(), w["data_std"].numpy(), B=2)
        f0[1] *= 0.5
    u, noise = oracle.draw_rng(T, 11)
    for reverb in (False, True):
        ref = _run_stream(m, f0, control, u, noise, chunks, reverb)
        bad = 0
        for r in range(reps):
            y = _run_stream(m, f0, control, u, noise, chunks, reverb)
            if not torch.equal(y, ref):
                d = (y - ref).abs()
                idx = torch.nonzero(d > 0)
                bad += 1
                print("  MISMATCH %s fast=%s reverb=%s rep %d: max %.3g, %d samples differ, first at %s last at %s" %
                      (tag, fast, reverb, r, float(d.max()), idx.shape[0], idx[0].tolist(), idx[-1].tolist()), flush=True)
        print("%s fast=%s chunks=%s reverb=%s: %d/%d repetitions differ" % (tag, fast, chunks, reverb, bad, reps), flush=True)
