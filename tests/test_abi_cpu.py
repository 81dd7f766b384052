"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol
include/nws_b200.h declares, the host mirror keeps the reference's state-dict layout and error
behaviour, and nothing silently falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from neural_waveshaping_synthesis_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(libpath):
    header = open(os.path.join(REPO, "include", "nws_b200.h")).read()
    declared = set(re.findall(r"\b(nws_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(libpath)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from neural_waveshaping_synthesis_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) <= declared
    lib.nws_api_version.restype = ctypes.c_int
    assert lib.nws_api_version() == 1     # no device needed


def test_tensor_order_matches_header():
    from neural_waveshaping_synthesis_b200 import _lib
    header = open(os.path.join(REPO, "include", "nws_b200.h")).read()
    assert "NWS_T_COUNT" in header
    assert _lib.N_TENSORS == 9 + 14 + 9 + 2 + 14 + 2
    z = np.load(os.path.join(REPO, "tests", "golden", "weights_vn.npz"))
    assert all(k in z.files for k in _lib.TENSOR_KEYS)


def test_tensor_numel_matches_shape_table(libpath):
    """The library's own element counts (nws_tensor_numel) agree with the binding's shape table and with the
    shipped checkpoints, so NwsEngine.load_weights can refuse other configurations before any raw pointer is read."""
    from neural_waveshaping_synthesis_b200 import _lib
    lib = ctypes.CDLL(libpath)
    lib.nws_tensor_numel.restype = ctypes.c_size_t
    lib.nws_tensor_numel.argtypes = [ctypes.c_int]
    z = np.load(os.path.join(REPO, "tests", "golden", "weights_vn.npz"))
    for i, (k, shp) in enumerate(zip(_lib.TENSOR_KEYS, _lib.TENSOR_SHAPES)):
        assert tuple(z[k].shape) == shp, k
        assert lib.nws_tensor_numel(i) == int(np.prod(shp)), k
    assert lib.nws_tensor_numel(-1) == 0 and lib.nws_tensor_numel(_lib.N_TENSORS) == 0


def _model():
    import gin
    from neural_waveshaping_synthesis.models.neural_waveshaping import NeuralWaveshaping
    gin.clear_config()
    gin.parse_config_file(os.path.join(REPO, "gin", "models", "newt.gin"))
    return NeuralWaveshaping


def test_seeded_construction_reproduces_reference_init():
    NW = _model()
    torch.manual_seed(0)
    m = NW().eval()
    z = np.load(os.path.join(REPO, "tests", "golden", "weights_randinit.npz"))
    sd = m.state_dict()
    assert set(sd) == set(z.files)
    for k in z.files:
        assert np.array_equal(sd[k].numpy(), z[k]), k
    assert sum(p.numel() for p in m.parameters()) == 266945
    assert m.sample_rate == 16000 and m.control_hop == 128


def test_checkpoint_state_dicts_load():
    NW = _model()
    m = NW()
    for tag in ("vn", "fl", "tpt"):
        z = np.load(os.path.join(REPO, "tests", "golden", "weights_%s.npz" % tag))
        res = m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("data_")})
        assert not res.missing_keys and not res.unexpected_keys


def test_no_cpu_fallback():
    NW = _model()
    m = NW().eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.rand(1, 1, 8), torch.rand(1, 2, 8))
    if not torch.cuda.is_available():
        from neural_waveshaping_synthesis.models.modules.shaping import FastNEWT
        with pytest.raises(RuntimeError, match="CUDA"):
            FastNEWT(m.newt)
    with pytest.raises(NotImplementedError):
        m.newt(torch.rand(1, 64, 256), torch.rand(1, 128, 2))


def test_gin_shim_scopes_and_macros():
    import gin
    gin.clear_config()
    gin.parse_config_file(os.path.join(REPO, "gin", "models", "newt.gin"))
    assert gin.query_parameter("%sample_rate") == 16000
    assert gin.query_parameter("noise_synth/TimeDistributedMLP.out_size") == 129
    from neural_waveshaping_synthesis.models.modules.dynamic import TimeDistributedMLP
    with gin.config_scope("noise_synth"):
        mlp = TimeDistributedMLP()
    assert mlp.out_size == 129 and mlp.depth == 4
    with pytest.raises(TypeError):
        TimeDistributedMLP()  # unscoped: no bindings -> missing required arguments, as with real gin
    mlp2 = TimeDistributedMLP(8, 8, 4)
    assert mlp2.depth == 3  # explicit arguments win, defaults stay


def test_dataset_mirror(tmp_path):
    from neural_waveshaping_synthesis.data.urmp import URMPDataset
    root = tmp_path / "ds"
    for kind in ("audio", "control"):
        os.makedirs(root / "test" / kind)
    rng = np.random.default_rng(0)
    for name in ("a_0", "b_1"):
        np.save(root / "test" / "audio" / ("audio_%s.npy" % name), rng.normal(size=64000).astype(np.float32))
        np.save(root / "test" / "control" / ("control_%s.npy" % name), rng.normal(size=(19, 500)).astype(np.float32))
    mean, std = rng.normal(size=(19, 1)), np.abs(rng.normal(size=(19, 1))) + 0.1
    np.save(root / "data_mean.npy", mean)
    np.save(root / "data_std.npy", std)
    ds = URMPDataset(str(root), "test", True)
    assert len(ds) == 2
    item = ds[0]
    assert item["f0"].shape == (1, 500) and item["control"].shape == (19, 500) and item["name"] == "a_0"
    assert np.allclose(item["f0"], item["control"][0:1] * std[0] + mean[0])


def test_header_is_plain_c(tmp_path):
    """include/nws_b200.h is the boundary a non-C++ host binds: it must compile as strict C99 on its own, and a C
    translation unit calling every entry point through it must compile without warnings (no torch, no C++ types)."""
    import subprocess
    src = tmp_path / "use_abi.c"
    src.write_text(r"""
#include "nws_b200.h"
int use_abi(void* stream) {
  NwsConfig cfg;
  NwsHandle h = 0;
  NwsStreamHandle st = 0;
  int n_out = 0, frames = 0;
  long long first = 0;
  float ms[NWS_N_STAGES];
  nws_default_config(&cfg);
  if (nws_create(&cfg, &h) != NWS_OK) return nws_api_version() + (nws_last_error() != 0);
  nws_load_weights(h, 0, NWS_T_COUNT, stream);
  nws_build_lut(h, 4096, -3.0f, 3.0f, 0, stream);
  nws_set_lut(h, 0, 4096, -3.0f, 3.0f, stream);
  nws_get_lut(h, 0, stream);
  nws_forward(h, 0, 0, 2, 0, 0, 0u, 0u, 0, 1, 2, 1, 0, nws_workspace_bytes(h, 1, 2), stream);
  nws_forward_host(h, 0, 0, 2, 0, 0, 0u, 0u, 0, 1, 2, 1, 0, 0, stream);
  nws_stage_control_embedding(h, 0, 2, 0, 1, 2, 0, 0, stream);
  nws_stage_td_mlp(h, NWS_MLP_FILM, 0, 0, 1, 2, 0, 0, stream);
  nws_stage_audio(h, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0, stream);
  nws_stage_lut_lookup(h, 0, 0, 0, 1, 256, stream);
  nws_stage_noise(h, 0, 0, 0, 1, 2, 0, 0, stream);
  nws_stage_reverb(h, 0, 0, 1, 256, 0, nws_reverb_workspace_bytes(h, 1, 256), stream);
  nws_stage_control_to_params(h, 0, 2, 0, 0, 1, 2, 0, 0, stream);
  nws_shaper_eval(0, 0, 0, 16, 0, stream);
  nws_set_profiling(h, 1);
  nws_get_stage_times(h, ms, NWS_N_STAGES);
  nws_set_pipeline(h, 0);
  nws_set_mlp_impl(h, 1);
  nws_set_audio_impl(h, 1);
  nws_set_shaper_impl(h, 1);
  nws_set_gru_impl(h, 1);
  nws_set_noise_fused(h, 1);
  nws_set_reverb_direct(h, 1);
  nws_set_small_path(h, 1);
  nws_selftest_umma(0, 0, 0, 8, 0, 0, stream);
  nws_selftest_sin(0, 0, 0, 0, 0, stream);
  nws_selftest_ffma_peak(0, 1, 1, 0, stream);
  nws_stream_create(h, 1, 8, &st);
  nws_stream_reset(st, 0, 0u, 0u, stream);
  nws_stream_window(st, 2, &first, &frames);
  nws_stream_push(st, 0, 0, 2, 2, 0, 1, 0, 1, 0, &n_out, stream);
  nws_stream_destroy(st);
  nws_extract_loudness(0, 1, 4096, 1024, 128, 1e-5, 1, 0, 0, 0, nws_loudness_workspace_bytes(1, 4096, 1024, 128), stream);
  nws_extract_rms(0, 1, 4096, 2048, 512, 0, stream);
  (void)nws_shaper_eval_scratch_bytes();
  (void)nws_launch_count(1);
  (void)nws_tensor_numel(NWS_T_REVERB_IR);
  (void)nws_status(h);
  nws_interp_frames(0, 1, 4, 1024, 128, nws_interp_frames_len(4, 1024, 128, 0), 0, stream);
  return nws_destroy(h);
}
""")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(REPO, "include"),
                           "-c", str(src), "-o", str(tmp_path / "use_abi.o")])
    # every function the header declares is exercised by the C file above
    header = open(os.path.join(REPO, "include", "nws_b200.h")).read()
    declared = set(re.findall(r"\b(nws_[a-z_0-9]+)\s*\(", header))
    used = set(re.findall(r"\b(nws_[a-z_0-9]+)\s*\(", src.read_text()))
    assert declared <= used, sorted(declared - used)


def test_model_copies_and_pickles_without_its_engines():
    """The per-device engines hold ctypes handles: they are caches, not state.  deepcopy / torch.save of a model
    that has already run must work and must not share a handle (both used to raise)."""
    import copy
    import io
    NW = _model()
    m = NW().eval()
    for sub in m.modules():
        if hasattr(sub, "_bind_root"):
            sub._bind_root(m)
    m._engines["cuda:0"] = ctypes.pointer(ctypes.c_int(5))    # what a first forward leaves behind (unpicklable)
    m._loaded["cuda:0"] = ("sig", None)
    m2 = copy.deepcopy(m)
    assert m2._engines == {} and m2._loaded == {} and m2.embedding._root() is m2 and m.embedding._root() is m
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert m3._engines == {} and all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m3.state_dict().values()))
    with torch.inference_mode():
        m4 = NW()
    from neural_waveshaping_synthesis_b200.models.neural_waveshaping import _version_of
    assert all(isinstance(_version_of(t), int) for t in m4.state_dict().values())


def test_ref_script_fixtures_are_verbatim():
    """tests/golden/ref_scripts/*.py.txt are byte copies of the reference's scripts (what tests/test_gpu_ref_scripts.py
    executes on the GPU box, where /root/reference does not exist)."""
    import hashlib
    d = os.path.join(REPO, "tests", "golden", "ref_scripts")
    sums = dict(reversed(line.split()) for line in open(os.path.join(d, "SHA256SUMS")) if line.strip())
    assert len(sums) == 3
    for name, digest in sums.items():
        blob = open(os.path.join(d, name), "rb").read()
        assert hashlib.sha256(blob).hexdigest() == digest, name
        ref = os.path.join("/root/reference/scripts", name[:-len(".txt")])
        if os.path.exists(ref):
            assert open(ref, "rb").read() == blob, name
