"""Golden vectors for the frame-rate -> sample-rate interpolators, produced by the REFERENCE's own file
(neural_waveshaping_synthesis/data/utils/upsampling.py, loaded unmodified from /root/reference):

    python oracle/gen_golden_upsampling.py        ->  tests/golden/upsampling.npz

Test infrastructure only (see oracle/nws_oracle.py's header): nothing under the package imports this."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.gen_golden_loudness import OUT, librosa_stand_in, load_reference_modules  # noqa: E402


def main():
    librosa_stand_in()   # loudness_extraction.py is loaded alongside (package-relative import); librosa is absent here
    _, up = load_reference_modules()
    rng = np.random.default_rng(7)
    out = {}
    cases = [(47, 2048, 512, 24000), (63, 1024, 128, 8000), (51, 256, 100, 5003), (1, 64, 16, 0), (6, 64, 32, 0)]
    out["cases"] = np.asarray(cases, dtype=np.int64)
    for i, (F, window, hop, orig) in enumerate(cases):
        frames = (rng.standard_normal(F) * 20 - 40).astype(np.float32)
        out["frames_%d" % i] = frames
        kw = dict(window_length=window, hop_length=hop, original_length=orig or None)
        out["linear_%d" % i] = up.linear_interpolation(frames, **kw)
        if F >= 4:
            out["cubic_%d" % i] = up.cubic_spline_interpolation(frames, **kw)
        if 2 * hop <= window:
            out["ola_%d" % i] = up.overlap_add_upsample(frames, **kw)
    path = os.path.join(OUT, "upsampling.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
