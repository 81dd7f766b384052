"""CPU unit tests of csrc/nws_math.h — the scalar recipes every CUDA kernel uses — built with g++
from the same header (no GPU needed), against double precision / torch CPU / the golden vectors."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import nws_oracle as oracle
from tests.helpers import load_case, load_weights

HERE = os.path.dirname(os.path.abspath(__file__))
F32P = ctypes.POINTER(ctypes.c_float)
I32P = ctypes.POINTER(ctypes.c_int)


def _p(a, t=F32P):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(tempfile.mkdtemp(prefix="nws_math_"), "libmath_harness.so")
    src = os.path.join(HERE, "cpu_harness", "math_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src])
    L = ctypes.CDLL(out)
    L.h_div_check.restype = ctypes.c_long
    L.h_div_check.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_long]
    return L


def test_sinf_accuracy_large_arguments(lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.uniform(-4, 4, 200000), rng.uniform(-1300, 1300, 200000),      # dummy-input regime
        rng.uniform(-3e5, 3e5, 400000), rng.uniform(-1.2e6, 1.2e6, 400000),  # 4 s of a real f0
        np.linspace(-np.pi, np.pi, 10001), np.arange(-4000, 4000) * (np.pi / 4),
    ]).astype(np.float32)
    y = np.empty_like(x)
    lib.h_sinf(_p(x), _p(y), ctypes.c_long(x.size))
    ref = np.sin(x.astype(np.float64))
    err = np.abs(y.astype(np.float64) - ref)
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert err.max() < 1.3e-7, err.max()
    assert (err / np.maximum(ulp, 2.0 ** -30)).max() < 1e3  # tiny results: absolute bound governs
    # against torch CPU's sin (what the reference calls, generators.py:64, shaping.py:12)
    t = torch.sin(torch.from_numpy(x)).numpy()
    assert np.abs(t.astype(np.float64) - y).max() < 2e-7


def test_markstein_division_by_6_is_exact(lib):
    # every fp32 with 2^-100 <= |a| <= 2^26, and 0.  The shaper index numerator a = 4096*fp32(x+3) is
    # either 0 or >= 4096*ulp(3) ~ 1e-3, far inside; below 2^-100 the quotient nears the subnormal
    # range where the remainder FMA is no longer exact (2.8M mismatches, all |a| < 2.4e-38).
    assert lib.h_div_check(6.0, 2.0 ** -100, float(2 ** 26), 1) == 0


def test_markstein_division_other_spans(lib):
    for d in (2.0, 3.0, 5.0, 7.5, 10.0):
        assert lib.h_div_check(d, 2.0 ** -100, float(2 ** 20), 97) == 0


def test_upsample_bit_exact(lib):
    rng = np.random.default_rng(1)
    for T in (2, 3, 16, 500):
        x = (rng.uniform(50, 2000, T)).astype(np.float32)
        y = np.empty(T * 128, np.float32)
        lib.h_upsample(_p(x), T, 128, _p(y))
        ref = oracle.upsample_linear(torch.from_numpy(x).view(1, 1, T), T * 128)[0, 0].numpy()
        assert np.array_equal(y, ref)
    x = rng.uniform(0, 1, 500).astype(np.float32)
    y = np.empty(500 * 128, np.float32)
    lib.h_upsample(_p(x), 500, 128, _p(y))
    ref = oracle.upsample_linear(torch.from_numpy(x).view(1, 1, 500), 64000)[0, 0].numpy()
    assert np.array_equal(y, ref)


@pytest.mark.parametrize("case", ["kat_vn_newt", "kat_randinit_newt"])
def test_phase_pipeline_bit_exact(lib, case):
    c = load_case(case)
    f0 = c["f0"][0, 0].numpy().copy()
    T = f0.size
    csum = np.empty(T * 128, np.float32)
    phase = np.empty(T * 128, np.float32)
    lib.h_phase(_p(f0), T, 128, ctypes.c_float(16000.0), _p(csum), _p(phase))
    f0_up = oracle.upsample_linear(c["f0"], T * 128)
    _, parts = oracle.harmonic_oscillator(f0_up[:, 0], c["u_phase"], return_parts=True)
    assert np.array_equal(csum, parts["csum"][0].numpy())
    assert np.array_equal(phase, parts["phase"][0].numpy())
    shift = np.empty(101, np.float32)
    rp = np.full(101, np.float32(2 * np.pi), np.float32)
    u = c["u_phase"].numpy().copy()
    lib.h_phase_shift(_p(u), _p(rp), 101, _p(shift))
    assert np.array_equal(shift, oracle.phase_shift_from_uniform(c["u_phase"]).reshape(-1).numpy())
    for k in (1, 2, 17, 64, 101):
        arg = np.empty_like(phase)
        lib.h_harmonic_arg(_p(phase), ctypes.c_long(phase.size), k, ctypes.c_float(shift[k - 1]), _p(arg))
        assert np.array_equal(arg, parts["arg"][0, k - 1].numpy())


def test_lut_index_bit_exact(lib):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-3.5, 3.5, 500000), rng.normal(0, 1, 500000),
                        np.linspace(-3, 3, 4096), np.array([-3.0, 3.0, -10.0, 10.0, 2.9999998, -2.9999998, 0.0])
                        ]).astype(np.float32)
    lower = np.empty(x.size, np.int32)
    upper = np.empty(x.size, np.int32)
    fract = np.empty(x.size, np.float32)
    lib.h_lut_index(_p(x), ctypes.c_long(x.size), 4096, ctypes.c_float(-3.0), ctypes.c_float(3.0),
                    _p(lower, I32P), _p(upper, I32P), _p(fract))
    _, lo, up, fr = oracle.lut_indices(torch.from_numpy(x))
    assert np.array_equal(lower, lo.numpy())
    assert np.array_equal(upper, up.numpy())
    assert np.array_equal(fract, fr.numpy())
    # and the value lerp
    w = load_weights("vn")
    lut = oracle.build_lookup_table(w)[5].numpy()
    out = np.empty(x.size, np.float32)
    lov, upv = lut[lower].copy(), lut[upper].copy()
    lib.h_lut_lerp(_p(lov), _p(upv), _p(fract), ctypes.c_long(x.size), _p(out))
    ref = oracle.lut_shaping_fn(torch.from_numpy(lut).view(1, -1), torch.from_numpy(x).view(1, 1, -1))[0, 0].numpy()
    assert np.array_equal(out, ref)


def test_lut_index_pow2_fold_is_bit_identical(lib):
    """The fused kernel's index for the default 4096-entry table (no multiply by the size) == nws_lut_index's."""
    lib.h_idx_pow2_check.restype = ctypes.c_long
    lib.h_idx_pow2_check.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_long]
    assert lib.h_idx_pow2_check(4096, -3.0, 3.0, 1.0e4, 3) == 0          # every third float up to |x| = 1e4
    assert lib.h_idx_pow2_check(4096, -1.0, 2.5, 100.0, 11) == 0
    assert lib.h_idx_pow2_check(1024, -3.0, 3.0, 100.0, 11) == 0


def test_philox_known_answer_and_uniformity(lib):
    raw = np.empty(4, np.uint32)
    # Random123 known-answer vector: counter = 0, key = 0
    lib.h_philox_raw(ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0), raw.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    assert [hex(v) for v in raw] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = np.empty(4 * 50000, np.float32)
    lib.h_philox(ctypes.c_uint64(1234), ctypes.c_uint64(1), ctypes.c_uint64(0), ctypes.c_long(50000), _p(out))
    assert out.min() >= 0.0 and out.max() < 1.0
    assert abs(out.mean() - 0.5) < 5e-3 and abs(out.var() - 1 / 12) < 2e-3


@pytest.mark.parametrize("n1", [125, 250])
@pytest.mark.parametrize("inverse", [0, 1])
def test_mixed_radix_column_fft_matches_numpy(n1, inverse):
    """csrc/nws_fft_mixed.cuh (the column transforms of the reverb's exact-length plans: 32000 = 125 x 256,
    64000 = 250 x 256) run serially on the CPU against numpy's FFT in float64."""
    out_so = os.path.join(tempfile.mkdtemp(prefix="nws_fft_"), "libfft_harness.so")
    src = os.path.join(HERE, "cpu_harness", "fft_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-shared", "-fPIC", "-o", out_so, src])
    L = ctypes.CDLL(out_so)
    rng = np.random.default_rng(n1 + inverse)
    tw = np.exp(-2j * np.pi * np.arange(n1) / n1).astype(np.complex64)
    for log_w in (0, 3, 4):
        W = 1 << log_w
        x = (rng.standard_normal((n1, W)) + 1j * rng.standard_normal((n1, W))).astype(np.complex64)
        out = np.zeros_like(x)
        assert L.h_fft_mixed(n1, inverse, log_w, _p(x), _p(tw), _p(out)) == 0
        x64 = x.astype(np.complex128)
        ref = np.fft.ifft(x64, axis=0) * n1 if inverse else np.fft.fft(x64, axis=0)
        assert np.abs(out - ref).max() < 4e-7 * np.abs(ref).max()
    assert L.h_fft_mixed(128, 0, 0, _p(x), _p(tw), _p(out)) == -1
