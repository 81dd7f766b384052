"""ctypes binding of libnws_b200.so (C ABI: include/nws_b200.h).

Loading fails loudly: if the library is missing it is built with nvcc (build.py); if that is not
possible a RuntimeError is raised.  There is no CPU or PyTorch fallback for any entry point.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p

_LOCK = threading.Lock()
_LIB = None

# NwsTensor order (include/nws_b200.h) as reference state-dict keys
_MLP_KEYS = ["net.0.weight", "net.0.bias", "net.1.layer_norm.weight", "net.1.layer_norm.bias",
             "net.3.weight", "net.3.bias", "net.4.layer_norm.weight", "net.4.layer_norm.bias",
             "net.6.weight", "net.6.bias", "net.7.layer_norm.weight", "net.7.layer_norm.bias",
             "net.9.weight", "net.9.bias"]
SHAPER_KEYS = ["newt.shaping_fn.input_scale", "newt.shaping_fn.net.0.weight", "newt.shaping_fn.net.0.bias",
               "newt.shaping_fn.net.2.weight", "newt.shaping_fn.net.2.bias", "newt.shaping_fn.net.4.weight",
               "newt.shaping_fn.net.4.bias", "newt.shaping_fn.net.6.weight", "newt.shaping_fn.net.6.bias"]
TENSOR_KEYS = (
    ["embedding.gru.weight_ih_l0", "embedding.gru.weight_hh_l0", "embedding.gru.bias_ih_l0",
     "embedding.gru.bias_hh_l0", "embedding.proj.weight", "embedding.proj.bias", "osc.rand_phase",
     "harmonic_mixer.weight", "harmonic_mixer.bias"]
    + ["newt.mlp." + k for k in _MLP_KEYS]
    + SHAPER_KEYS
    + ["newt.mixer.0.weight", "newt.mixer.0.bias"]
    + ["h_generator." + k for k in _MLP_KEYS]
    + ["noise_synth.window", "reverb.ir"]
)
N_TENSORS = len(TENSOR_KEYS)  # == NWS_T_COUNT


class NwsConfig(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "sample_rate", "control_hop", "n_harmonics", "n_waveshapers", "embedding_size", "shaping_fn_size",
        "shaping_fn_depth", "noise_bands", "ir_length", "reverb_length")]


class NwsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libnws_b200 error %d: %s" % (code, message))
        self.code = code


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnws_b200.so")


def load_library():
    """Returns the loaded library, building it first if needed.  Raises if neither is possible."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.exists(path):
            from . import build as _build
            _build.build()
        try:
            import torch  # noqa: F401  (brings libcudart into the process)
        except Exception:
            pass
        try:
            L = ctypes.CDLL(path)
        except OSError as e:
            raise RuntimeError("cannot load %s: %s — the CUDA extension is required, there is no CPU fallback" % (path, e))
        fp, vp = POINTER(c_float), c_void_p
        L.nws_last_error.restype = c_char_p
        L.nws_api_version.restype = c_int
        L.nws_launch_count.restype = c_uint64
        L.nws_launch_count.argtypes = [c_int]
        L.nws_default_config.argtypes = [POINTER(NwsConfig)]
        L.nws_default_config.restype = None
        L.nws_create.argtypes = [POINTER(NwsConfig), POINTER(vp)]
        L.nws_destroy.argtypes = [vp]
        L.nws_load_weights.argtypes = [vp, POINTER(vp), c_int, vp]
        L.nws_build_lut.argtypes = [vp, c_int, c_float, c_float, vp, vp]
        L.nws_set_lut.argtypes = [vp, vp, c_int, c_float, c_float, vp]
        L.nws_get_lut.argtypes = [vp, vp, vp]
        L.nws_workspace_bytes.argtypes = [vp, c_int, c_int]
        L.nws_workspace_bytes.restype = c_size_t
        L.nws_reverb_workspace_bytes.argtypes = [vp, c_int, c_int]
        L.nws_reverb_workspace_bytes.restype = c_size_t
        L.nws_forward.argtypes = [vp, vp, vp, c_int, vp, vp, c_uint64, c_uint64, vp, c_int, c_int, c_int, vp, c_size_t, vp]
        L.nws_forward_host.argtypes = L.nws_forward.argtypes
        L.nws_stage_control_embedding.argtypes = [vp, vp, c_int, vp, c_int, c_int, vp, c_size_t, vp]
        L.nws_stage_td_mlp.argtypes = [vp, c_int, vp, vp, c_int, c_int, vp, c_size_t, vp]
        L.nws_stage_audio.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, vp, c_size_t, vp]
        L.nws_stage_lut_lookup.argtypes = [vp, vp, vp, vp, c_int, c_int, vp]
        L.nws_stage_lut_lookup.restype = c_int
        L.nws_stage_noise.argtypes = [vp, vp, vp, vp, c_int, c_int, vp, c_size_t, vp]
        L.nws_stage_reverb.argtypes = [vp, vp, vp, c_int, c_int, vp, c_size_t, vp]
        L.nws_selftest_sin.argtypes = [vp, vp, vp, vp, ctypes.c_longlong, vp]
        L.nws_selftest_sin.restype = c_int
        L.nws_selftest_ffma_peak.argtypes = [vp, c_int, c_int, POINTER(ctypes.c_double), vp]
        L.nws_selftest_ffma_peak.restype = c_int
        L.nws_set_pipeline.argtypes = [vp, c_int]
        L.nws_set_pipeline.restype = c_int
        L.nws_set_mlp_impl.argtypes = [vp, c_int]
        L.nws_set_mlp_impl.restype = c_int
        L.nws_stage_control_to_params.argtypes = [vp, vp, c_int, vp, vp, c_int, c_int, vp, c_size_t, vp]
        L.nws_stage_control_to_params.restype = c_int
        L.nws_set_audio_impl.argtypes = [vp, c_int]
        L.nws_set_audio_impl.restype = c_int
        L.nws_set_small_path.argtypes = [vp, c_int]
        L.nws_set_small_path.restype = c_int
        L.nws_set_reverb_direct.argtypes = [vp, c_int]
        L.nws_set_reverb_direct.restype = c_int
        L.nws_set_shaper_impl.argtypes = [vp, c_int]
        L.nws_set_shaper_impl.restype = c_int
        L.nws_set_gru_impl.argtypes = [vp, c_int]
        L.nws_set_gru_impl.restype = c_int
        L.nws_set_noise_fused.argtypes = [vp, c_int]
        L.nws_set_noise_fused.restype = c_int
        L.nws_selftest_umma.argtypes = [vp, vp, vp, c_int, c_int, vp, vp]
        L.nws_selftest_umma.restype = c_int
        L.nws_set_profiling.argtypes = [vp, c_int]
        L.nws_set_profiling.restype = c_int
        L.nws_get_stage_times.argtypes = [vp, POINTER(c_float), c_int]
        L.nws_get_stage_times.restype = c_int
        L.nws_stream_create.argtypes = [vp, c_int, c_int, POINTER(vp)]
        L.nws_stream_destroy.argtypes = [vp]
        L.nws_stream_reset.argtypes = [vp, vp, c_uint64, c_uint64, vp]
        L.nws_stream_window.argtypes = [vp, c_int, POINTER(ctypes.c_longlong), POINTER(c_int)]
        L.nws_stream_push.argtypes = [vp, vp, vp, c_int, c_int, vp, c_int, c_int, c_int, vp, POINTER(c_int), vp]
        for name in ("nws_stream_create", "nws_stream_destroy", "nws_stream_reset", "nws_stream_window", "nws_stream_push"):
            getattr(L, name).restype = c_int
        L.nws_loudness_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
        L.nws_loudness_workspace_bytes.restype = c_size_t
        L.nws_extract_loudness.argtypes = [vp, c_int, c_int, c_int, c_int, ctypes.c_double, c_int, vp, vp, vp, c_size_t, vp]
        L.nws_extract_loudness.restype = c_int
        L.nws_extract_rms.argtypes = [vp, c_int, c_int, c_int, c_int, vp, vp]
        L.nws_extract_rms.restype = c_int
        L.nws_interp_frames_len.argtypes = [c_int, c_int, c_int, c_int]
        L.nws_interp_frames_len.restype = c_int
        L.nws_interp_frames.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, vp, vp]
        L.nws_interp_frames.restype = c_int
        L.nws_tensor_numel.argtypes = [c_int]
        L.nws_tensor_numel.restype = c_size_t
        L.nws_status.argtypes = [vp]
        L.nws_status.restype = c_int
        L.nws_shaper_eval_scratch_bytes.restype = c_size_t
        L.nws_shaper_eval.argtypes = [POINTER(vp), vp, vp, c_int, vp, vp]
        for name in ("nws_create", "nws_destroy", "nws_load_weights", "nws_build_lut", "nws_set_lut", "nws_get_lut",
                     "nws_forward", "nws_forward_host", "nws_stage_control_embedding", "nws_stage_td_mlp",
                     "nws_stage_audio", "nws_stage_noise", "nws_stage_reverb", "nws_shaper_eval"):
            getattr(L, name).restype = c_int
        _LIB = L
        return L


def check(rc: int):
    if rc != 0:
        raise NwsError(rc, load_library().nws_last_error().decode("utf-8", "replace"))


# symbols include/nws_b200.h declares (tests check the library exports every one)
EXPORTED_SYMBOLS = [
    "nws_last_error", "nws_api_version", "nws_create", "nws_destroy", "nws_default_config", "nws_load_weights",
    "nws_build_lut", "nws_set_lut", "nws_get_lut", "nws_workspace_bytes", "nws_forward", "nws_forward_host",
    "nws_stage_control_embedding", "nws_stage_td_mlp", "nws_stage_audio", "nws_stage_lut_lookup", "nws_stage_noise", "nws_stage_reverb",
    "nws_reverb_workspace_bytes", "nws_shaper_eval_scratch_bytes", "nws_shaper_eval", "nws_launch_count",
    "nws_set_profiling", "nws_get_stage_times", "nws_selftest_umma", "nws_set_audio_impl", "nws_set_mlp_impl", "nws_stage_control_to_params", "nws_selftest_sin", "nws_set_pipeline",
    "nws_stream_create", "nws_stream_destroy", "nws_stream_reset", "nws_stream_window", "nws_stream_push",
    "nws_loudness_workspace_bytes", "nws_extract_loudness", "nws_extract_rms", "nws_selftest_ffma_peak",
    "nws_tensor_numel", "nws_status", "nws_interp_frames_len", "nws_interp_frames",
    "nws_set_shaper_impl", "nws_set_reverb_direct", "nws_set_small_path", "nws_set_gru_impl", "nws_set_noise_fused",
]

# Shapes the kernels are built for (gin/models/newt.gin) in TENSOR_KEYS order.  The C side reads
# raw pointers, so NwsEngine.load_weights refuses anything else (a model built from other gin bindings, or a
# checkpoint with other hyper-parameters) instead of reading out of bounds.
_MLP_SHAPES = lambda n_out: [(128, 128, 1), (128,), (128,), (128,)] * 3 + [(n_out, 128, 1), (n_out,)]
TENSOR_SHAPES = (
    [(384, 2), (384, 128), (384,), (384,), (128, 128, 1), (128,), (1, 101, 1), (64, 101, 1), (64,)]
    + _MLP_SHAPES(256)
    + [(1, 64, 1), (512, 1, 1), (512,), (512, 8, 1), (512,), (512, 8, 1), (512,), (64, 8, 1), (64,)]
    + [(1, 64, 1), (1,)]
    + _MLP_SHAPES(129)
    + [(256,), (1, 31999)]
)
assert len(TENSOR_SHAPES) == N_TENSORS
STAGE_NAMES = ["rng", "phase_carry", "gru", "proj", "film_mlp", "noise_mlp", "noise_spectrum", "noise_filter",
               "audio_fused", "reverb"]
