"""Thin host-side owner of one libnws_b200 handle: torch is used for device memory, streams and
nothing else.  All compute goes through the C ABI (include/nws_b200.h)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib

HOP = 128
N_HARMONICS = 101
N_SHAPERS = 64
EMB = 128
N_BANDS = 129


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _as_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class NwsEngine:
    """One C handle bound to one CUDA device."""

    def __init__(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("NwsEngine needs a CUDA device (got %s): the B200 path has no CPU fallback" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.lib = _lib.load_library()
        cfg = _lib.NwsConfig()
        self.lib.nws_default_config(ctypes.byref(cfg))
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.nws_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None
        self._ws_bytes = {}          # (B, T) -> workspace bytes (saves a library call per forward)
        self._keep = None          # tensors whose pointers the last load call used
        self.has_lut = False
        self.lut_shape = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.nws_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_for(self, B: int, T: int) -> torch.Tensor:
        n = self._ws_bytes.get((B, T))
        if n is None:
            n = self.lib.nws_workspace_bytes(self.handle, B, T)
            if n == 0:
                raise ValueError("unsupported shape B=%d T=%d" % (B, T))
            self._ws_bytes[(B, T)] = n
        ws = self._ws
        return ws if ws is not None and ws.numel() >= n else self._workspace(n)

    def _next_rng(self, n_noise: int):
        """(seed, offset) of this forward's Philox draws, reserved on torch's CUDA generator exactly as
        torch's own CUDA kernels reserve theirs — so torch.manual_seed(s) makes the forward reproducible
        and consecutive forwards consume consecutive parts of the stream."""
        gen = torch.cuda.default_generators[self.device.index]
        off = gen.get_offset()
        gen.set_offset(off + 4 * ((n_noise + 3) // 4 + 32))
        return gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, off // 4

    # ------------------------------------------------------------------ weights / LUT
    def load_weights(self, state: Dict[str, torch.Tensor]):
        """state: reference state-dict keys -> tensors on any device."""
        missing = [k for k in _lib.TENSOR_KEYS if k not in state]
        if missing:
            raise KeyError("missing weights: %s" % missing)
        bad = ["%s: %s, built for %s" % (k, tuple(state[k].shape), shp)
               for k, shp in zip(_lib.TENSOR_KEYS, _lib.TENSOR_SHAPES) if tuple(state[k].shape) != shp]
        if bad:
            raise NotImplementedError("the CUDA path is built for the gin/models/newt.gin sizes only; these tensors "
                                      "have other shapes (the reference would run them, this library reads raw "
                                      "pointers and refuses): " + "; ".join(bad))
        tensors = [_as_f32(state[k], self.device) for k in _lib.TENSOR_KEYS]
        for i, t in enumerate(tensors):   # belt and braces: the library's own element counts
            if t.numel() != self.lib.nws_tensor_numel(i):
                raise NotImplementedError("tensor %s has %d elements, the library reads %d" %
                                          (_lib.TENSOR_KEYS[i], t.numel(), self.lib.nws_tensor_numel(i)))
        arr = (ctypes.c_void_p * _lib.N_TENSORS)(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_load_weights(self.handle, arr, _lib.N_TENSORS, self._stream()))
        self._keep = tensors
        self.has_lut = False

    def build_lut(self, table_size: int = 4096, table_min: float = -3.0, table_max: float = 3.0,
                  sample_points: Optional[torch.Tensor] = None):
        pts = None if sample_points is None else _as_f32(sample_points, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_build_lut(self.handle, table_size, table_min, table_max, _ptr(pts), self._stream()))
        self.has_lut, self.lut_shape = True, (N_SHAPERS, table_size)

    def set_lut(self, lut: torch.Tensor, table_min: float = -3.0, table_max: float = 3.0):
        lut = _as_f32(lut, self.device)
        if lut.dim() != 2 or lut.shape[0] != N_SHAPERS:
            raise ValueError("lookup table must be [64, table_size]")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_set_lut(self.handle, _ptr(lut), lut.shape[1], table_min, table_max, self._stream()))
        self.has_lut, self.lut_shape = True, tuple(lut.shape)

    def get_lut(self) -> torch.Tensor:
        if not self.has_lut:
            raise RuntimeError("no lookup table loaded")
        out = torch.empty(self.lut_shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_get_lut(self.handle, _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ forward
    def _check_inputs(self, f0, control):
        if f0.dim() != 3 or f0.shape[1] != 1:
            raise ValueError("f0 must be [B, 1, T] (got %s)" % (tuple(f0.shape),))
        if control.dim() != 3 or control.shape[1] < 2:
            raise ValueError("control must be [B, C>=2, T] (got %s)" % (tuple(control.shape),))
        if control.shape[0] != f0.shape[0] or control.shape[2] != f0.shape[2]:
            raise ValueError("f0 %s and control %s disagree on batch size or frame count" %
                             (tuple(f0.shape), tuple(control.shape)))
        if f0.shape[2] < 2:
            raise ValueError("at least 2 control frames are required (got T=%d)" % f0.shape[2])
        for name, t in (("f0", f0), ("control", control)):
            if t.dtype != torch.float32:
                raise ValueError("%s must be float32 (got %s)" % (name, t.dtype))
            if t.device != self.device:
                raise ValueError("%s is on %s but the model is on %s" % (name, t.device, self.device))

    def forward(self, f0: torch.Tensor, control: torch.Tensor, u_phase: Optional[torch.Tensor] = None,
                noise: Optional[torch.Tensor] = None, use_lut: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check_inputs(f0, control)
        B, _, T = f0.shape
        N = T * HOP
        f0c, cc = f0.contiguous(), control.contiguous()
        up = None if u_phase is None else _as_f32(u_phase, self.device).reshape(-1)
        nz = None if noise is None else _as_f32(noise, self.device).reshape(-1)
        if up is not None and up.numel() != N_HARMONICS:
            raise ValueError("u_phase must have 101 elements")
        if nz is not None and nz.numel() != N - 1:
            raise ValueError("noise must have 128*T-1 = %d elements" % (N - 1))
        if out is None:
            out = torch.empty(B, N, dtype=torch.float32, device=self.device)
        ws = self.workspace_for(B, T)
        seed, off = (0, 0) if (up is not None and nz is not None) else self._next_rng(N - 1)
        if torch.cuda.current_device() == self.device.index:      # the common case: no device switch needed
            rc = self.lib.nws_forward(self.handle, _ptr(f0c), _ptr(cc), cc.shape[1], _ptr(up), _ptr(nz), seed, off,
                                      _ptr(out), B, T, 1 if use_lut else 0, _ptr(ws), ws.numel(), self._stream())
        else:
            with torch.cuda.device(self.device):
                rc = self.lib.nws_forward(self.handle, _ptr(f0c), _ptr(cc), cc.shape[1], _ptr(up), _ptr(nz), seed, off,
                                          _ptr(out), B, T, 1 if use_lut else 0, _ptr(ws), ws.numel(), self._stream())
        if rc:
            _lib.check(rc)
        return out

    def forward_host(self, f0: torch.Tensor, control: torch.Tensor, out: torch.Tensor, u_phase=None, noise=None,
                     use_lut: bool = False) -> torch.Tensor:
        """Host buffers in, host buffer out (nws_forward_host): H2D + forward + D2H + stream sync."""
        B, _, T = f0.shape
        assert f0.device.type == "cpu" and control.device.type == "cpu" and out.device.type == "cpu"
        assert f0.is_contiguous() and control.is_contiguous() and out.is_contiguous()
        ws = self.workspace_for(B, T)
        seed, off = (0, 0) if (u_phase is not None and noise is not None) else self._next_rng(T * HOP - 1)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_forward_host(self.handle, _ptr(f0), _ptr(control), control.shape[1], _ptr(u_phase),
                                                 _ptr(noise), seed, off, _ptr(out), B, T, 1 if use_lut else 0,
                                                 _ptr(ws), ws.numel(), self._stream()))
        return out

    def set_audio_impl(self, impl: int):
        """1 = tcgen05 harmonic mixer (default), 0 = fp32 SIMT mixer."""
        _lib.check(self.lib.nws_set_audio_impl(self.handle, impl))

    def set_small_path(self, enable: bool):
        """Few frames per utterance: fp32 small-batch MLP chain (default) or always the tensor-core tile kernel."""
        _lib.check(self.lib.nws_set_small_path(self.handle, 1 if enable else 0))

    def set_reverb_direct(self, enable: bool):
        """Short buffers: direct-form reverb (default) or always the FFT path."""
        _lib.check(self.lib.nws_set_reverb_direct(self.handle, 1 if enable else 0))

    def set_shaper_impl(self, impl: int):
        """NEWT shaper hidden layers: 1 = tensor cores (mma.sync, default), 0 = fp32 FMA (paired lanes)."""
        _lib.check(self.lib.nws_set_shaper_impl(self.handle, impl))

    def set_noise_fused(self, enable: bool):
        """Filtered-noise branch as its own launch ahead of the fused audio kernel (default, faster) or inside it."""
        _lib.check(self.lib.nws_set_noise_fused(self.handle, 1 if enable else 0))

    def set_gru_impl(self, impl: int):
        """GRU recurrence: 1 = tensor cores (8 utterances per CTA) from 64 utterances on, fp32 SIMT below (default);
        0 = fp32 SIMT always; 2 = tensor cores always."""
        _lib.check(self.lib.nws_set_gru_impl(self.handle, impl))

    def set_pipeline(self, enable: bool):
        """Pipelined forward (GRU time blocks on an internal stream overlapped with rendering); default on."""
        _lib.check(self.lib.nws_set_pipeline(self.handle, 1 if enable else 0))

    def set_mlp_impl(self, impl: int):
        """1 = tcgen05 MLP chain (default), 0 = fp32 SIMT layer kernels."""
        _lib.check(self.lib.nws_set_mlp_impl(self.handle, impl))

    def control_to_params(self, control: torch.Tensor):
        """control [B,C,T] -> (film [B,256,T], bands [B,129,T])."""
        control = _as_f32(control, self.device)
        B, C, T = control.shape
        film = torch.empty(B, 256, T, dtype=torch.float32, device=self.device)
        bands = torch.empty(B, N_BANDS, T, dtype=torch.float32, device=self.device)
        ws = self.workspace_for(B, T)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_control_to_params(self.handle, _ptr(control), C, _ptr(film), _ptr(bands), B, T,
                                                            _ptr(ws), ws.numel(), self._stream()))
        return film, bands

    def set_profiling(self, enable: bool):
        _lib.check(self.lib.nws_set_profiling(self.handle, 1 if enable else 0))

    def stage_times_ms(self):
        buf = (ctypes.c_float * len(_lib.STAGE_NAMES))()
        _lib.check(self.lib.nws_get_stage_times(self.handle, buf, len(_lib.STAGE_NAMES)))
        return dict(zip(_lib.STAGE_NAMES, [float(v) for v in buf]))

    # ------------------------------------------------------------------ stages
    def control_embedding(self, control: torch.Tensor) -> torch.Tensor:
        control = _as_f32(control, self.device)
        B, C, T = control.shape
        out = torch.empty(B, EMB, T, dtype=torch.float32, device=self.device)
        ws = self.workspace_for(B, T)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_control_embedding(self.handle, _ptr(control), C, _ptr(out), B, T, _ptr(ws),
                                                            ws.numel(), self._stream()))
        return out

    def td_mlp(self, which: int, emb: torch.Tensor) -> torch.Tensor:
        emb = _as_f32(emb, self.device)
        B, C, T = emb.shape
        if C != EMB:
            raise ValueError("embedding must have 128 channels")
        out = torch.empty(B, 256 if which == 0 else N_BANDS, T, dtype=torch.float32, device=self.device)
        ws = self.workspace_for(B, T)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_td_mlp(self.handle, which, _ptr(emb), _ptr(out), B, T, _ptr(ws), ws.numel(),
                                                 self._stream()))
        return out

    def audio(self, f0: torch.Tensor, film: torch.Tensor, u_phase: torch.Tensor, use_lut: bool = False,
              want_exciter: bool = False):
        f0, film = _as_f32(f0, self.device), _as_f32(film, self.device)
        up = _as_f32(u_phase, self.device).reshape(-1)
        B, _, T = f0.shape
        N = T * HOP
        out = torch.empty(B, N, dtype=torch.float32, device=self.device)
        exc = torch.empty(B, N_SHAPERS, N, dtype=torch.float32, device=self.device) if want_exciter else None
        ws = self.workspace_for(B, T)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_audio(self.handle, _ptr(f0), _ptr(film), _ptr(up), _ptr(out), _ptr(exc), B, T,
                                                1 if use_lut else 0, _ptr(ws), ws.numel(), self._stream()))
        return (out, exc) if want_exciter else out

    def lut_lookup(self, x: torch.Tensor):
        """FastNEWT.shaping_fn on x [B,64,N]: returns (y, lower_index)."""
        x = _as_f32(x, self.device)
        B, C, N = x.shape
        if C != N_SHAPERS:
            raise ValueError("x must be [B,64,N]")
        y = torch.empty_like(x)
        lower = torch.empty(x.shape, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_lut_lookup(self.handle, _ptr(x), _ptr(y), _ptr(lower), B, N, self._stream()))
        return y, lower

    def noise(self, H: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        H, noise = _as_f32(H, self.device), _as_f32(noise, self.device).reshape(-1)
        B, C, T = H.shape
        if C != N_BANDS or noise.numel() != T * HOP - 1:
            raise ValueError("H must be [B,129,T] and noise [128*T-1]")
        out = torch.empty(B, T * HOP, dtype=torch.float32, device=self.device)
        ws = self.workspace_for(B, T)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_noise(self.handle, _ptr(H), _ptr(noise), _ptr(out), B, T, _ptr(ws), ws.numel(),
                                                self._stream()))
        return out

    def reverb(self, x: torch.Tensor) -> torch.Tensor:
        x = _as_f32(x, self.device)
        B, N = x.shape
        out = torch.empty_like(x)
        n = self.lib.nws_reverb_workspace_bytes(self.handle, B, N)
        if n == 0:
            raise ValueError("reverb: unsupported length %d" % N)
        ws = self._workspace(n)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nws_stage_reverb(self.handle, _ptr(x), _ptr(out), B, N, _ptr(ws), ws.numel(), self._stream()))
        return out


def shaper_eval(shaper_tensors, x: torch.Tensor) -> torch.Tensor:
    """out[c, i] = shaper_c(x[i]) on the CUDA device of `x` (nws_shaper_eval)."""
    lib = _lib.load_library()
    dev = x.device
    if dev.type != "cuda":
        raise ValueError("shaper_eval needs CUDA tensors")
    ts = [_as_f32(t, dev) for t in shaper_tensors]
    xs = _as_f32(x, dev).reshape(-1)
    out = torch.empty(N_SHAPERS, xs.numel(), dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.nws_shaper_eval_scratch_bytes(), dtype=torch.uint8, device=dev)
    arr = (ctypes.c_void_p * 9)(*[t.data_ptr() for t in ts])
    with torch.cuda.device(dev):
        _lib.check(lib.nws_shaper_eval(arr, _ptr(xs), _ptr(out), xs.numel(), _ptr(scratch),
                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out
