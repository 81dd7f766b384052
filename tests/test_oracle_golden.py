"""Pins the oracle (oracle/nws_oracle.py) against vectors produced by the real reference
(tests/golden/, generator oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import nws_oracle as oracle
from tests.helpers import err, golden_path, load_case, load_weights

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


@pytest.mark.parametrize("tag", ["randinit", "vn"])
def test_lut_matches_reference(tag):
    w = load_weights(tag)
    lut = oracle.build_lookup_table(w)
    z = np.load(golden_path("lut_%s.npz" % tag))
    assert np.array_equal(lut.numpy()[:, ::16], z["lut_sub"])
    assert np.allclose(lut.double().sum(1).numpy(), z["row_sum"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("case,tag,fast", [
    ("kat_randinit_newt", "randinit", False), ("kat_randinit_fast", "randinit", True),
    ("kat_vn_newt", "vn", False), ("kat_vn_fast", "vn", True),
    ("kat_fl_newt", "fl", False), ("kat_fl_fast", "fl", True),
    ("kat_tpt_newt", "tpt", False), ("kat_tpt_fast", "tpt", True),
    ("small_randinit_newt", "randinit", False), ("small_randinit_fast", "randinit", True),
    ("small_vn_newt", "vn", False), ("small_vn_fast", "vn", True),
    ("min_randinit_newt", "randinit", False), ("min_randinit_fast", "randinit", True),
])
def test_forward_matches_reference(case, tag, fast):
    w = load_weights(tag)
    c = load_case(case)
    lut = oracle.build_lookup_table(w) if fast else None
    y = oracle.forward(w, c["f0"], c["control"], c["u_phase"], c["noise"], lut=lut)
    assert y.shape == c["out"].shape
    # same torch ops in the same order: the restatement is bit-exact with the reference
    assert torch.equal(y, c["out"]), err(y, c["out"])


def test_known_answers_survey_appendix_b():
    c = load_case("kat_randinit_newt")
    y = c["out"]
    assert y.shape == (1, 64000)
    assert np.allclose(y[0, :4].numpy(), [-0.03329345, -0.03229887, -0.09698731, 0.03409825], atol=2e-8)
    assert abs(float(y[0, 32000]) - -0.06161531) < 2e-8
    assert abs(float(y.abs().mean()) - 0.1058435) < 1e-6
    assert abs(float(y.abs().max()) - 0.342327) < 1e-6


@pytest.mark.parametrize("bs", [256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_buffer_sweep_shapes(bs):
    z = np.load(golden_path("sweep_randinit.npz"))
    w = load_weights("randinit")
    f0 = torch.from_numpy(z["bs%d_f0" % bs])
    control = torch.from_numpy(z["bs%d_control" % bs])
    u, noise = oracle.draw_rng(bs // 128, int(z["bs%d_rng_seed" % bs]))
    assert torch.equal(u.reshape(-1), torch.from_numpy(z["bs%d_u_phase" % bs]))
    y = oracle.forward(w, f0, control, u, noise)
    assert torch.equal(y, torch.from_numpy(z["bs%d_out" % bs]))
    yf = oracle.forward(w, f0, control, u, noise, lut=oracle.build_lookup_table(w))
    assert torch.equal(yf, torch.from_numpy(z["bs%d_out_fast" % bs]))


def test_intermediates_small_vn():
    w = load_weights("vn")
    c = load_case("small_vn_newt")
    y, p = oracle.forward(w, c["f0"], c["control"], c["u_phase"], c["noise"], return_parts=True)
    for k in ("f0_up", "exciter", "emb", "film", "H", "newt_out", "noise_out", "dry"):
        assert torch.equal(p[k], c["part_" + k]), (k, err(p[k], c["part_" + k]))


def test_literal_recipes_agree_with_torch_ops():
    """The scalar recipes the CUDA kernels implement against the torch ops."""
    c = load_case("small_vn_newt")
    w = load_weights("vn")
    # A.1 linear upsample: bit-exact
    up = oracle.upsample_linear(c["f0"], c["f0"].shape[-1] * 128)
    lit = oracle.upsample_linear_literal(c["f0"].numpy(), 128)
    assert np.array_equal(up.numpy(), lit)
    # A.6 GRU gate maths: close (different evaluation order than MKL-DNN)
    h_lit = oracle.gru_literal(w, c["control"])
    emb_lit = torch.nn.functional.conv1d(h_lit.transpose(1, 2), w["embedding.proj.weight"], w["embedding.proj.bias"])
    assert err(emb_lit, c["part_emb"])[0] < 5e-4
    # A.4 noise branch, time-domain circular form
    nz = oracle.fir_noise_literal(c["part_H"].numpy(), c["noise"].numpy())
    assert err(nz, c["part_noise_out"][:, 0])[0] < 5e-5
    # A.5 reverb, fold form
    rv = oracle.reverb_literal(w["reverb.ir"].numpy(), c["part_dry"].numpy())
    assert err(rv, c["out"])[0] < 2e-5
