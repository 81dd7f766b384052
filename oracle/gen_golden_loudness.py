"""ORACLE tooling — generates tests/golden/loudness_*.npz by running the REAL reference functions
(neural_waveshaping_synthesis/data/utils/loudness_extraction.py and upsampling.py, loaded unmodified from
/root/reference) on seeded signals.

Runs only in the authoring container.  The reference's code is the glue — which librosa calls are made, with which
arguments, the mean over bins, the (x + 80) / 80 mapping, the interpolation — and all arithmetic below it is
librosa 0.8.0, which is absent here.  So `librosa` is a stand-in module built from the restatement in
oracle/loudness_oracle.py: these fixtures pin the restatement's *glue* to the reference's own code exactly (any
drift in argument handling, frame counts, normalisation or interpolation shows up), while the librosa layer itself
stays restated, not pinned (see the oracle's header).

    python oracle/gen_golden_loudness.py          # writes tests/golden/loudness_*.npz
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NWS_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
UTILS = os.path.join(REF, "neural_waveshaping_synthesis", "data", "utils")


def librosa_stand_in():
    sys.path.insert(0, REPO)
    from oracle import loudness_oracle as lo
    m = types.ModuleType("librosa")
    m.stft = lambda y, n_fft=2048, hop_length=None, window="hann": lo.stft(y, n_fft, hop_length or n_fft // 4, window)
    m.amplitude_to_db = lo.amplitude_to_db
    m.fft_frequencies = lambda sr=22050, n_fft=2048: np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)

    def a_weighting(frequencies, min_db=-80.0):   # librosa 0.8.0 core/convert.py (IEC 61672); unused by the reference (:39)
        f_sq = np.asanyarray(frequencies) ** 2.0
        const = np.array([12200, 20.6, 107.7, 737.9]) ** 2.0
        weights = 2.0 + 20.0 * (np.log10(const[0]) + 2 * np.log10(f_sq) - np.log10(f_sq + const[0]) - np.log10(f_sq + const[1])
                                - 0.5 * np.log10(f_sq + const[2]) - 0.5 * np.log10(f_sq + const[3]))
        return weights if min_db is None else np.maximum(min_db, weights)

    m.A_weighting = a_weighting
    util = types.ModuleType("librosa.util")

    def frame(x, frame_length, hop_length):
        n_frames = 1 + (x.shape[-1] - frame_length) // hop_length
        idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
        return x[idx]

    util.frame = frame
    m.util = util
    sys.modules["librosa"] = m
    sys.modules["librosa.util"] = util


def load_reference_modules():
    """The two reference files as members of a synthetic package (so `from .upsampling import ...` resolves) without
    importing the reference package's __init__ chain (pytorch_lightning, torchcrepe, resampy: absent)."""
    import gin  # noqa: F401  (the repo's shim; the reference decorates with @gin.configurable)
    pkg = types.ModuleType("nws_ref_data_utils")
    pkg.__path__ = [UTILS]
    sys.modules["nws_ref_data_utils"] = pkg
    mods = {}
    for name in ("upsampling", "loudness_extraction"):
        spec = importlib.util.spec_from_file_location("nws_ref_data_utils." + name, os.path.join(UTILS, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["loudness_extraction"], mods["upsampling"]


def signal(n, seed, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    f0 = 196.0 * 2 ** (0.25 * np.sin(2 * np.pi * 0.9 * t))
    phase = 2 * np.pi * np.cumsum(f0) / sr
    env = 0.35 * np.abs(np.sin(2 * np.pi * 0.7 * t)) ** 1.5
    x = env * sum(np.sin(k * phase) / k ** 1.2 for k in range(1, 12)) + 3e-4 * rng.standard_normal(n)
    x[n // 2: n // 2 + n // 16] = 0.0
    return x.astype(np.float32)


def main():
    librosa_stand_in()
    le, up = load_reference_modules()
    os.makedirs(OUT, exist_ok=True)
    cases = {"gin": dict(n=32000, n_fft=1024, hop_length=128),           # gin/data/urmp_4second_crepe.gin:11-14
             "default": dict(n=24000, n_fft=2048, hop_length=512),       # the function's defaults (colab cell 14)
             "ragged": dict(n=5003, n_fft=256, hop_length=100)}
    for tag, c in cases.items():
        x = signal(c["n"], seed=len(tag))
        out = {"audio": x, "n_fft": c["n_fft"], "hop_length": c["hop_length"]}
        out["db"] = le.compute_power_spectrogram(x, n_fft=c["n_fft"], hop_length=c["hop_length"], window="hann", epsilon=1e-5)
        out["loudness_frames"] = le.extract_perceptual_loudness(x, n_fft=c["n_fft"], hop_length=c["hop_length"], interpolate_fn=None)
        out["loudness_frames_db"] = le.extract_perceptual_loudness(x, n_fft=c["n_fft"], hop_length=c["hop_length"],
                                                                   interpolate_fn=None, normalise=False)
        out["loudness_samples"] = le.extract_perceptual_loudness(x, n_fft=c["n_fft"], hop_length=c["hop_length"],
                                                                 interpolate_fn=up.linear_interpolation)
        out["rms_frames"] = le.extract_rms(x, c["n_fft"], c["hop_length"], interpolate_fn=None)
        out["rms_samples"] = le.extract_rms(x, c["n_fft"], c["hop_length"], interpolate_fn=up.linear_interpolation)
        path = os.path.join(OUT, "loudness_%s.npz" % tag)
        np.savez_compressed(path, **out)
        print(path, {k: (v.shape, v.dtype) for k, v in out.items() if isinstance(v, np.ndarray)})


if __name__ == "__main__":
    main()
