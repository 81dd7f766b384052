"""Shared test helpers: golden-fixture loading (tests/golden/, made by oracle/gen_golden.py)."""
import os

import numpy as np
import torch

from oracle import nws_oracle as oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_path(name):
    return os.path.join(GOLDEN, name)


def load_weights(tag):
    """tag in {'randinit', 'vn', 'fl', 'tpt'} -> dict of torch CPU tensors keyed like the
    reference state-dict (+ data_mean/data_std for checkpoints)."""
    z = np.load(golden_path("weights_%s.npz" % tag))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_case(name, prefix=""):
    """A golden case as a dict of torch tensors.  Cases stored without the noise vector carry
    `rng_seed`; the noise is regenerated with the reference's draw order and the stored u_phase
    guards against RNG drift."""
    z = np.load(golden_path(name + ".npz"))
    keys = [k for k in z.files if k.startswith(prefix)] if prefix else z.files
    case = {k[len(prefix):]: torch.from_numpy(np.asarray(z[k])) for k in keys}
    if "noise" not in case:
        T = case["f0"].shape[-1]
        u, noise = oracle.draw_rng(T, int(case["rng_seed"]))
        assert torch.equal(u.reshape(-1), case["u_phase"]), "torch CPU RNG stream changed"
        case["noise"] = noise
    return case


def err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    d = (a - b).abs()
    return float(d.max()), float(d.pow(2).mean().sqrt())
