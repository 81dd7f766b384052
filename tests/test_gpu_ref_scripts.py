"""The reference's OWN script texts, unmodified (tests/golden/ref_scripts/*.py.txt = byte copies of
/root/reference/scripts/*.py, see the README there), executed against the drop-in package on a GPU:
scripts/time_forward_pass.py:26-58, scripts/time_buffer_sizes.py:34-75, scripts/resynthesise_dataset.py:38-76.
(tests/test_gpu_scripts.py runs this repo's own rewrites of the same CLIs.)"""
import os
import types

import numpy as np
import pytest
import torch
from click.testing import CliRunner

from oracle import nws_oracle as oracle
from tests.helpers import load_weights

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GIN = os.path.join(REPO, "gin", "models", "newt.gin")
FIXTURES = os.path.join(REPO, "tests", "golden", "ref_scripts")


def _reference_script(name):
    """Module object made from the reference's script text: /root/reference (or baseline/_ref) when present, else the
    verbatim fixture.  `__name__` is not "__main__", so only the click command `main` is defined."""
    for root in ("/root/reference/scripts", os.path.join(REPO, "baseline", "_ref", "scripts")):
        path = os.path.join(root, name + ".py")
        if os.path.exists(path):
            break
    else:
        path = os.path.join(FIXTURES, name + ".py.txt")
    mod = types.ModuleType("ref_script_" + name)
    mod.__file__ = path
    exec(compile(open(path).read(), path, "exec"), mod.__dict__)
    return mod


@pytest.fixture(autouse=True)
def _fresh_gin():
    import gin
    gin.clear_config()
    yield
    gin.clear_config()


@pytest.mark.parametrize("fast", [False, True])
def test_reference_time_forward_pass(fast):
    args = ["--gin-file", GIN, "--num-iters", "3", "--batch-size", "2", "--device", "cuda:0"]
    r = CliRunner().invoke(_reference_script("time_forward_pass").main, args + (["--use-fast-newt"] if fast else []))
    assert r.exit_code == 0, (r.output, r.exception)
    assert "Mean RTF" in r.output and "90th percentile RTF" in r.output


def test_reference_time_buffer_sizes(tmp_path):
    import pandas as pd
    out = tmp_path / "sweep.csv"
    r = CliRunner().invoke(_reference_script("time_buffer_sizes").main,
                           ["--gin-file", GIN, "--output-file", str(out), "--num-iters", "2", "--device", "cuda:0",
                            "--use-fast-newt"])
    assert r.exit_code == 0, (r.output, r.exception)
    df = pd.read_csv(out)
    assert len(df) == 8 * 2 and set(df.iloc[:, 3]) == {256, 512, 1024, 2048, 4096, 8192, 16384, 32768}
    assert set(df.iloc[:, 2]) == {"gpu"} and (df.iloc[:, 4] > 0).all()


def test_reference_resynthesise_dataset(tmp_path):
    from scipy.io import wavfile
    w = load_weights("vn")
    mean, std = w["data_mean"].numpy(), w["data_std"].numpy()
    ckpt = tmp_path / "last.ckpt"
    torch.save({"state_dict": {k: v for k, v in w.items() if not k.startswith("data_")},
                "hyper_parameters": {"n_waveshapers": 64, "control_hop": 128, "sample_rate": 16000},
                "pytorch-lightning_version": "1.2.8"}, ckpt)
    root = tmp_path / "ds"
    for kind in ("audio", "control"):
        os.makedirs(root / "test" / kind)
    np.save(root / "data_mean.npy", mean)
    np.save(root / "data_std.npy", std)
    T = 24
    _, control2 = oracle.realistic_inputs(T, mean, std, B=3)
    rng = np.random.default_rng(0)
    for i in range(3):
        control = rng.normal(size=(19, T)).astype(np.float32)
        control[0:2] = control2[i].numpy()
        np.save(root / "test" / "control" / ("control_clip_%d.npy" % i), control)
        np.save(root / "test" / "audio" / ("audio_clip_%d.npy" % i), rng.normal(size=T * 128).astype(np.float32) * 0.1)
    outdir = tmp_path / "out"
    for extra in ([], ["--use-fastnewt"]):
        r = CliRunner().invoke(_reference_script("resynthesise_dataset").main,
                               ["--model-gin", GIN, "--model-checkpoint", str(ckpt), "--dataset-root", str(root),
                                "--output-path", str(outdir), "--batch-size", "2", "--num_workers", "0",
                                "--device", "cuda:0"] + extra)
        assert r.exit_code == 0, (r.output, r.exception)
        names = sorted(os.listdir(outdir))
        assert len(names) == 6 and "clip_0.output.wav" in names and "clip_2.target.wav" in names
        sr, audio = wavfile.read(outdir / "clip_1.output.wav")
        assert sr == 16000 and audio.shape == (T * 128,) and np.isfinite(audio).all() and np.abs(audio).max() > 1e-3
