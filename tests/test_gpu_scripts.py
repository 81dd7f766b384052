"""The three drop-in callers (this repo's scripts/, same CLI as the reference's scripts/) end to end on a GPU:
time_forward_pass.py, time_buffer_sizes.py and resynthesise_dataset.py over a synthetic on-disk dataset."""
import importlib.util
import os

import numpy as np
import pytest
import torch
from click.testing import CliRunner

from oracle import nws_oracle as oracle
from tests.helpers import err, load_weights

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GIN = os.path.join(REPO, "gin", "models", "newt.gin")


def _script(name):
    spec = importlib.util.spec_from_file_location("nws_script_" + name, os.path.join(REPO, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_time_forward_pass_cli():
    import gin
    gin.clear_config()
    r = CliRunner().invoke(_script("time_forward_pass").main,
                           ["--gin-file", GIN, "--num-iters", "3", "--batch-size", "2", "--use-fast-newt"])
    assert r.exit_code == 0, r.output
    assert "Mean RTF" in r.output and "90th percentile RTF" in r.output


def test_time_buffer_sizes_cli(tmp_path):
    import gin
    gin.clear_config()
    out = tmp_path / "sweep.csv"
    r = CliRunner().invoke(_script("time_buffer_sizes").main,
                           ["--gin-file", GIN, "--output-file", str(out), "--num-iters", "2"])
    assert r.exit_code == 0, r.output
    import pandas as pd
    df = pd.read_csv(out)
    assert len(df) == 8 * 2 and set(df.iloc[:, 3]) == {256, 512, 1024, 2048, 4096, 8192, 16384, 32768}


def test_resynthesise_dataset_cli(tmp_path):
    import gin
    from scipy.io import wavfile
    gin.clear_config()
    w = load_weights("vn")
    mean, std = w["data_mean"].numpy(), w["data_std"].numpy()
    state = {k: v for k, v in w.items() if not k.startswith("data_")}
    ckpt = tmp_path / "last.ckpt"
    torch.save({"state_dict": state, "hyper_parameters": {"n_waveshapers": 64, "control_hop": 128, "sample_rate": 16000},
                "pytorch-lightning_version": "1.2.8"}, ckpt)
    root = tmp_path / "ds"
    for kind in ("audio", "control"):
        os.makedirs(root / "test" / kind)
    np.save(root / "data_mean.npy", mean)
    np.save(root / "data_std.npy", std)
    T = 40
    f0, control2 = oracle.realistic_inputs(T, mean, std, B=3)
    rng = np.random.default_rng(0)
    for i in range(3):
        control = rng.normal(size=(19, T)).astype(np.float32)
        control[0:2] = control2[i].numpy() * (1.0 + 0.1 * i)
        np.save(root / "test" / "control" / ("control_clip_%d.npy" % i), control)
        np.save(root / "test" / "audio" / ("audio_clip_%d.npy" % i), rng.normal(size=T * 128).astype(np.float32) * 0.1)
    outdir = tmp_path / "out"
    r = CliRunner().invoke(_script("resynthesise_dataset").main,
                           ["--model-gin", GIN, "--model-checkpoint", str(ckpt), "--dataset-root", str(root),
                            "--output-path", str(outdir), "--batch-size", "2", "--num_workers", "0", "--use-fastnewt"])
    assert r.exit_code == 0, (r.output, r.exception)
    names = sorted(os.listdir(outdir))
    assert len(names) == 6 and "clip_0.output.wav" in names and "clip_0.target.wav" in names
    sr, audio = wavfile.read(outdir / "clip_1.output.wav")
    assert sr == 16000 and audio.shape == (T * 128,) and np.isfinite(audio).all() and np.abs(audio).max() > 1e-3
    # the draws are the forward's own Philox stream, so compare the deterministic part: run the same item through
    # the module with injected draws against the oracle
    from neural_waveshaping_synthesis.data.urmp import URMPDataset
    from tests.test_gpu_parity import _model
    item = URMPDataset(str(root), "test", True)[1]
    m, _ = _model("vn", True)
    u, noise = oracle.draw_rng(T, 5)
    f0_t = torch.from_numpy(item["f0"]).float()[None]
    c_t = torch.from_numpy(item["control"]).float()[None]
    with torch.no_grad():
        y = m(f0_t.cuda(), c_t.cuda(), phase_shift=u.reshape(-1).cuda(), noise=noise.cuda())
    ref = oracle.forward(w, f0_t, c_t, u, noise, lut=oracle.build_lookup_table(w))
    e = err(y, ref)
    assert e[0] < 1e-4 and e[1] < 1e-5, e
