"""ORACLE tooling — generates tests/golden/*.npz by running the REAL reference.

Runs only in the authoring container, where /root/reference exists (it does not on
the GPU box).  The reference's own modules are imported unmodified; the three
third-party packages it needs that are absent from this image are replaced by
throw-away stubs defined below (gin -> the repo's shim, pytorch_lightning -> a
LightningModule that is a plain nn.Module, auraloss -> empty), exactly as
SURVEY.md §8(c) describes.

    python oracle/gen_golden.py            # writes tests/golden/

Each fixture stores the inputs (f0, control), the two RNG draws of the forward
(u_phase = rand_like(rand_phase) [101], noise = rand(128T-1)), and the reference's
outputs (plus intermediates for the small cases), so tests need nothing from
/root/reference at run time.
"""
from __future__ import annotations

import copy
import math
import os
import pickle
import sys
import types
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NWS_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")


def install_stubs():
    sys.path.insert(0, REPO)           # `gin` shim
    import gin  # noqa: F401

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl
    sys.modules["auraloss"] = types.ModuleType("auraloss")
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            sys.modules["wandb"] = types.ModuleType("wandb")
    # the reference package must win over the repo's drop-in package of the same name
    sys.path.insert(0, REF)


class _PermissiveUnpickler(pickle.Unpickler):
    """Lightning 1.2.8 checkpoints reference classes that are not installed
    (ModelCheckpoint as a dict key, AttributeDict): stub anything unknown."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (dict,), {"__module__": module, "__hash__": lambda self: id(self),
                                        "__setstate__": lambda self, s: None,
                                        "__reduce_ex__": None})


class _PermissivePickle:
    __name__ = "permissive_pickle"
    Unpickler = _PermissiveUnpickler

    @staticmethod
    def load(f, **kw):
        return _PermissiveUnpickler(f, **kw).load()


def load_ckpt(path):
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PermissivePickle)


def state_to_npz(sd):
    out = {}
    for k, v in sd.items():
        if k.startswith("newt.shaping_fn.") and False:
            continue
        out[k] = v.detach().cpu().numpy()
    return out


def realistic_inputs(T, mean, std, B=1, f_lo=None, f_hi=None):
    u = torch.linspace(0, 1, T)
    if f_lo is None:
        f0 = 440.0 * torch.pow(2.0, 0.5 * torch.sin(2 * math.pi * 1.5 * u))
    else:  # exponential sweep: exercises the anti-alias mask (k*f0 >= 8000) heavily
        f0 = f_lo * torch.pow(torch.tensor(f_hi / f_lo), u)
    loud = 0.10 + 0.03 * torch.sin(2 * math.pi * 3 * u)
    control = torch.stack(((f0 - float(mean[0])) / float(std[0]), (loud - float(mean[1])) / float(std[1])))
    f0 = f0.view(1, 1, T).expand(B, 1, T).contiguous().float()
    control = control.view(1, 2, T).expand(B, 2, T).contiguous().float()
    return f0, control


def run_reference(model, f0, control, seed, want_parts=False, store_noise=True):
    """One reference forward.  The RNG draws are recorded by replaying the same seed.  With
    store_noise=False only `rng_seed` is kept for the 128T-1 noise vector (tests regenerate it with
    oracle.draw_rng and cross-check the stored u_phase) — keeps the 4 s fixtures small."""
    T = f0.shape[-1]
    torch.manual_seed(seed)
    u_phase = torch.rand(1, 101, 1)
    noise = torch.rand(128 * T - 1)
    torch.manual_seed(seed)
    parts = {}
    with torch.no_grad():
        if want_parts:
            import torch.nn.functional as F
            st = torch.get_rng_state()
            f0_up = F.interpolate(f0, T * 128, mode="linear")
            exciter = model.render_exciter(f0_up)            # consumes RNG draw #1
            emb = model.get_embedding(control)
            film = model.newt.mlp(emb)
            newt_out = model.newt(exciter, emb)
            H = model.h_generator(emb)
            noise_out = model.noise_synth(H)                  # consumes RNG draw #2
            dry = torch.cat((newt_out, noise_out), dim=1).sum(1)
            parts = dict(f0_up=f0_up, exciter=exciter, emb=emb, film=film, newt_out=newt_out, H=H,
                         noise_out=noise_out, dry=dry)
            torch.set_rng_state(st)
        y = model(f0, control)
    rec = dict(f0=f0, control=control, u_phase=u_phase.reshape(-1), out=y,
               rng_seed=torch.tensor(seed, dtype=torch.int64))
    if store_noise:
        rec["noise"] = noise
    rec.update({"part_" + k: v for k, v in parts.items()})
    return {k: v.detach().cpu().numpy() for k, v in rec.items()}


def main():
    warnings.filterwarnings("ignore")
    install_stubs()
    os.chdir(REF)
    import gin
    from neural_waveshaping_synthesis.models.neural_waveshaping import NeuralWaveshaping
    from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
    import neural_waveshaping_synthesis
    assert neural_waveshaping_synthesis.__file__.startswith(REF), neural_waveshaping_synthesis.__file__

    gin.parse_config_file("gin/models/newt.gin")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ------------------------------------------------ random-init model (SURVEY App. B KAT #1/#2)
    torch.manual_seed(0)
    m = NeuralWaveshaping().eval()
    np.savez_compressed(os.path.join(OUT, "weights_randinit.npz"), **state_to_npz(m.state_dict()))
    mf = copy.deepcopy(m)
    mf.newt = FastNEWT(mf.newt)
    mf.eval()
    lut = mf.newt.lookup_table.detach().numpy()
    # the 1 MiB table is pinned by every 16th column plus float64 row sums (the oracle rebuilds it in full)
    np.savez_compressed(os.path.join(OUT, "lut_randinit.npz"), lut_sub=lut[:, ::16],
                        row_sum=lut.astype(np.float64).sum(1))

    torch.manual_seed(1)
    f0 = torch.rand(1, 1, 500)
    control = torch.rand(1, 2, 500)
    np.savez_compressed(os.path.join(OUT, "kat_randinit_newt.npz"), **run_reference(m, f0, control, 2, store_noise=False))
    np.savez_compressed(os.path.join(OUT, "kat_randinit_fast.npz"), **run_reference(mf, f0, control, 2, store_noise=False))

    # small multi-batch cases with intermediates (stage-level parity)
    torch.manual_seed(11)
    f0s = torch.rand(2, 1, 6)
    cs = torch.rand(2, 3, 6)          # 3 control channels: channel 2 must be ignored (neural_waveshaping.py:70)
    np.savez_compressed(os.path.join(OUT, "small_randinit_newt.npz"), **run_reference(m, f0s, cs, 12, True))
    np.savez_compressed(os.path.join(OUT, "small_randinit_fast.npz"), **run_reference(mf, f0s, cs, 12, True))
    torch.manual_seed(13)
    f0m = torch.rand(3, 1, 2)
    cm = torch.rand(3, 2, 2)          # T=2: the smallest size the reference accepts
    np.savez_compressed(os.path.join(OUT, "min_randinit_newt.npz"), **run_reference(m, f0m, cm, 14))
    np.savez_compressed(os.path.join(OUT, "min_randinit_fast.npz"), **run_reference(mf, f0m, cm, 14))
    # buffer-sweep shapes of scripts/time_buffer_sizes.py:13 (B=1): 256 ... 32768 samples, one seed each
    sweep = {}
    for bs in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
        torch.manual_seed(100 + bs)
        f0b, cb = torch.rand(1, 1, bs // 128), torch.rand(1, 2, bs // 128)
        r = run_reference(m, f0b, cb, 200 + bs, store_noise=False)
        rf = run_reference(mf, f0b, cb, 200 + bs, store_noise=False)
        for k, v in r.items():
            sweep["bs%d_%s" % (bs, k)] = v
        sweep["bs%d_out_fast" % bs] = rf["out"]
    np.savez_compressed(os.path.join(OUT, "sweep_randinit.npz"), **sweep)

    # ------------------------------------------------ shipped checkpoints (SURVEY App. B KAT #3)
    for inst in ("vn", "fl", "tpt"):
        ck = load_ckpt(os.path.join(REF, "checkpoints", "nws", inst, "last.ckpt"))
        hp = dict(ck["hyper_parameters"])
        mc = NeuralWaveshaping(**{k: hp[k] for k in ("n_waveshapers", "control_hop", "sample_rate") if k in hp})
        missing = mc.load_state_dict(ck["state_dict"])
        print(inst, "load_state_dict:", missing)
        mc.eval()
        mean = np.load(os.path.join(REF, "checkpoints", "nws", inst, "data_mean.npy"))
        std = np.load(os.path.join(REF, "checkpoints", "nws", inst, "data_std.npy"))
        np.savez_compressed(os.path.join(OUT, "weights_%s.npz" % inst), **state_to_npz(mc.state_dict()),
                            data_mean=mean, data_std=std)
        mcf = copy.deepcopy(mc)
        mcf.newt = FastNEWT(mcf.newt)
        mcf.eval()
        f0r, cr = realistic_inputs(500, mean, std)
        np.savez_compressed(os.path.join(OUT, "kat_%s_newt.npz" % inst), **run_reference(mc, f0r, cr, 2, store_noise=False))
        np.savez_compressed(os.path.join(OUT, "kat_%s_fast.npz" % inst), **run_reference(mcf, f0r, cr, 2, store_noise=False))
        if inst == "vn":
            lv = mcf.newt.lookup_table.detach().numpy()
            np.savez_compressed(os.path.join(OUT, "lut_vn.npz"), lut_sub=lv[:, ::16],
                                row_sum=lv.astype(np.float64).sum(1))
            # sweep 150 Hz -> 3 kHz over 10 frames, B=2 (second item an octave lower): mask edge cases
            f0a, ca = realistic_inputs(10, mean, std, B=1, f_lo=150.0, f_hi=3000.0)
            f0b, cb = realistic_inputs(10, mean, std, B=1, f_lo=75.0, f_hi=9000.0)
            f0w, cw = torch.cat((f0a, f0b)), torch.cat((ca, cb))
            np.savez_compressed(os.path.join(OUT, "small_vn_newt.npz"), **run_reference(mc, f0w, cw, 21, True))
            np.savez_compressed(os.path.join(OUT, "small_vn_fast.npz"), **run_reference(mcf, f0w, cw, 21, True))

    for f in sorted(os.listdir(OUT)):
        print("%-28s %8.1f KB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024))


if __name__ == "__main__":
    main()
