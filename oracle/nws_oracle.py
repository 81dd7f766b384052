"""ORACLE — test infrastructure, NOT product code.

CPU restatement (torch-CPU fp32 ops, functional style) of the hot path of
ben-hayes/neural-waveshaping-synthesis: ``NeuralWaveshaping.forward(f0, control)``
in eval / no-grad mode (reference ``neural_waveshaping_synthesis/models/
neural_waveshaping.py:74-90``).  Every function cites the reference file:line it
restates.  All arithmetic comes from torch (the reference pins torch==1.7.1,
requirements.txt:10, and has no arithmetic of its own); the de-facto oracle device
is torch CPU.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference from
/root/reference (possible only in the authoring container), runs it on seeded
inputs / the three shipped checkpoints and stores input+output vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against
those vectors (bit-exact where the op sequence is identical).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(neural_waveshaping_synthesis_b200/) never does and has no CPU fallback.

Weights are passed as a flat dict keyed by the reference's state-dict names
(SURVEY.md App. B), values torch CPU tensors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, torch.Tensor]

SAMPLE_RATE = 16000          # gin/models/newt.gin:1
CONTROL_HOP = 128            # gin/models/newt.gin:5
N_HARMONICS = 101            # gin/models/newt.gin:7
N_WAVESHAPERS = 64           # gin/models/newt.gin:4
IR_LENGTH = 256              # gin/models/newt.gin:24


# --------------------------------------------------------------------------- RNG
def draw_rng(T: int, seed: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two RNG draws of one reference forward, in the reference's order:
    first ``rand_like(rand_phase)`` [1,101,1] (generators.py:55, reached from
    neural_waveshaping.py:76), then ``rand(hop*T-1)`` (generators.py:30, reached from
    neural_waveshaping.py:83).  Uses the global CPU generator like the reference."""
    if seed is not None:
        torch.manual_seed(seed)
    u_phase = torch.rand(1, N_HARMONICS, 1)
    noise = torch.rand(CONTROL_HOP * T - 1)
    return u_phase, noise


# ---------------------------------------------------------------- hop-rate blocks
def upsample_linear(x: torch.Tensor, size: int) -> torch.Tensor:
    """F.upsample(x, size, mode='linear') == F.interpolate(..., align_corners=False)
    (neural_waveshaping.py:75, shaping.py:69)."""
    return F.interpolate(x, size, mode="linear", align_corners=False)


def upsample_linear_literal(x: np.ndarray, factor: int) -> np.ndarray:
    """Scalar recipe of the same op, numpy fp32; used to pin the
    formula the CUDA kernels implement.  x: [..., T] float32."""
    x = np.asarray(x, dtype=np.float32)
    T = x.shape[-1]
    N = T * factor
    scale = np.float32(T) / np.float32(N)
    n = np.arange(N, dtype=np.float32)
    src = np.maximum(scale * (n + np.float32(0.5)) - np.float32(0.5), np.float32(0.0)).astype(np.float32)
    i0 = src.astype(np.int64)
    i1 = np.minimum(i0 + 1, T - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    x0 = x[..., i0].astype(np.float64)
    x1 = x[..., i1].astype(np.float64)
    # fmaf(l0, x0, fp32(l1*x1)): the product l0*x0 is exact in double, one rounding at the end
    p1 = (l1.astype(np.float32) * x[..., i1]).astype(np.float32).astype(np.float64)
    return (l0.astype(np.float64) * x0 + p1).astype(np.float32)


def control_module(w: Weights, control: torch.Tensor) -> torch.Tensor:
    """get_embedding + ControlModule (neural_waveshaping.py:69-72, 17-26):
    channels 0 and 1 of `control` -> GRU(2->128, batch_first, h0=0) -> Conv1d(128->128, k=1)."""
    x = torch.cat((control[:, 0:1], control[:, 1:2]), dim=1)
    hidden = w["embedding.gru.weight_hh_l0"].shape[1]
    gru = torch.nn.GRU(x.shape[1], hidden, batch_first=True)
    with torch.no_grad():
        gru.weight_ih_l0.copy_(w["embedding.gru.weight_ih_l0"])
        gru.weight_hh_l0.copy_(w["embedding.gru.weight_hh_l0"])
        gru.bias_ih_l0.copy_(w["embedding.gru.bias_ih_l0"])
        gru.bias_hh_l0.copy_(w["embedding.gru.bias_hh_l0"])
        h, _ = gru(x.transpose(1, 2))
    return F.conv1d(h.transpose(1, 2), w["embedding.proj.weight"], w["embedding.proj.bias"])


def gru_literal(w: Weights, control: torch.Tensor) -> torch.Tensor:
    """Step-loop restatement of nn.GRU's gate maths, fp32, used to
    pin the gate order/bias placement the CUDA GRU kernel follows.  Returns h [B,T,128]."""
    W_ih, W_hh = w["embedding.gru.weight_ih_l0"], w["embedding.gru.weight_hh_l0"]
    b_ih, b_hh = w["embedding.gru.bias_ih_l0"], w["embedding.gru.bias_hh_l0"]
    Hd = W_hh.shape[1]
    x = torch.cat((control[:, 0:1], control[:, 1:2]), dim=1).transpose(1, 2)  # [B,T,2]
    B, T, _ = x.shape
    h = torch.zeros(B, Hd)
    out = []
    for t in range(T):
        gi = x[:, t] @ W_ih.t() + b_ih
        gh = h @ W_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n = torch.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
        h = (1 - z) * n + z * h
        out.append(h)
    return torch.stack(out, dim=1)


def td_mlp(w: Weights, prefix: str, x: torch.Tensor, depth: int = 4) -> torch.Tensor:
    """TimeDistributedMLP (dynamic.py:20-40): [Conv1d k=1 -> LayerNorm(C) over channels
    (dynamic.py:11-17) -> LeakyReLU(0.01)] x (depth-1) -> Conv1d k=1.  `prefix` is
    'newt.mlp' (shaping.py:53-55) or 'h_generator' (neural_waveshaping.py:58)."""
    for i in range(depth):
        j = 3 * i
        x = F.conv1d(x, w["%s.net.%d.weight" % (prefix, j)], w["%s.net.%d.bias" % (prefix, j)])
        if i < depth - 1:
            g = w["%s.net.%d.layer_norm.weight" % (prefix, j + 1)]
            b = w["%s.net.%d.layer_norm.bias" % (prefix, j + 1)]
            x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), g, b, 1e-5).transpose(1, 2)
            x = F.leaky_relu(x, 0.01)
    return x


# ------------------------------------------------------------------ audio-rate
def phase_shift_from_uniform(u_phase: torch.Tensor) -> torch.Tensor:
    """_create_phase_shift (generators.py:54-56): rand * tau - pi with rand_phase == tau."""
    rand_phase = torch.ones(1, N_HARMONICS, 1) * math.tau          # generators.py:45
    return u_phase.reshape(1, N_HARMONICS, 1) * rand_phase - math.pi


def harmonic_oscillator(f0_up: torch.Tensor, u_phase: torch.Tensor, sample_rate: float = SAMPLE_RATE,
                        return_parts: bool = False):
    """HarmonicOscillator.forward (generators.py:58-66).  f0_up: [B,N] Hz."""
    harmonic_axis = torch.arange(1, N_HARMONICS + 1).view(1, -1, 1)  # int64, generators.py:47-48
    csum = f0_up.cumsum(-1)                                          # CPU: double accumulate
    phase = math.tau * csum / sample_rate
    harmonic_phase = harmonic_axis * phase.unsqueeze(1)
    arg = harmonic_phase + phase_shift_from_uniform(u_phase)
    mask = (f0_up.unsqueeze(1) * harmonic_axis) < (sample_rate / 2)  # generators.py:50-52
    out = torch.sin(arg) * mask
    if return_parts:
        return out, dict(csum=csum, phase=phase, arg=arg, mask=mask)
    return out


def render_exciter(w: Weights, f0_up: torch.Tensor, u_phase: torch.Tensor) -> torch.Tensor:
    """render_exciter (neural_waveshaping.py:64-67): oscillator bank -> Conv1d(101->64,k=1)."""
    sig = harmonic_oscillator(f0_up[:, 0], u_phase)
    return F.conv1d(sig, w["harmonic_mixer.weight"], w["harmonic_mixer.bias"])


def trainable_nonlinearity(w: Weights, x: torch.Tensor, prefix: str = "newt.shaping_fn") -> torch.Tensor:
    """TrainableNonlinearity.forward (shaping.py:15-37) with Sine everywhere (shaping.py:59-61,
    gin depth 4): grouped Conv1d(groups=64,k=1) -> sin, four times, on input_scale * x."""
    C = w[prefix + ".input_scale"].shape[1]
    x = w[prefix + ".input_scale"] * x
    for j in (0, 2, 4, 6):
        x = torch.sin(F.conv1d(x, w["%s.net.%d.weight" % (prefix, j)], w["%s.net.%d.bias" % (prefix, j)], groups=C))
    return x


def build_lookup_table(w: Weights, table_size: int = 4096, table_min: float = -3.0,
                       table_max: float = 3.0) -> torch.Tensor:
    """FastNEWT._init_lookup_table (shaping.py:107-119): shaping_fn on linspace -> [64, table_size]."""
    C = w["newt.shaping_fn.input_scale"].shape[1]
    sample_values = torch.linspace(table_min, table_max, table_size).expand(1, C, table_size)
    return trainable_nonlinearity(w, sample_values)[0]


def lut_indices(x: torch.Tensor, table_size: int = 4096, table_min: float = -3.0, table_max: float = 3.0):
    """Index arithmetic of FastNEWT.shaping_fn (shaping.py:137-146): idx, lower, upper, fract."""
    idx = table_size * (x - table_min) / (table_max - table_min)
    lower = torch.floor(idx).long()
    lower[lower < 0] = 0
    lower[lower >= table_size] = table_size - 1
    upper = lower + 1
    upper[upper >= table_size] = table_size - 1
    fract = idx - lower
    return idx, lower, upper, fract


def lut_shaping_fn(lut: torch.Tensor, x: torch.Tensor, table_min: float = -3.0, table_max: float = 3.0,
                   faithful_loop: bool = False) -> torch.Tensor:
    """FastNEWT.shaping_fn (shaping.py:136-151).  `faithful_loop` reproduces the reference's
    per-(batch, shaper) Python indexing loop of _lookup (shaping.py:121-134) — same values as the
    gather form, kept so the CPU baseline pays what the reference pays."""
    table_size = lut.shape[-1]
    _, lower, upper, fract = lut_indices(x, table_size, table_min, table_max)
    if faithful_loop:
        def look(idx):
            return torch.stack([torch.stack([lut[s, idx[b, s]] for s in range(idx.shape[1])], dim=0)
                                for b in range(idx.shape[0])], dim=0)
    else:
        def look(idx):
            return torch.gather(lut.unsqueeze(0).expand(idx.shape[0], -1, -1), 2, idx)
    lower_v, upper_v = look(lower), look(upper)
    return (upper_v - lower_v) * fract + lower_v


def newt(w: Weights, exciter: torch.Tensor, emb: torch.Tensor, lut: Optional[torch.Tensor] = None,
         faithful_loop: bool = False, return_parts: bool = False):
    """NEWT.forward (shaping.py:67-79); with `lut` it is FastNEWT (shaping.py:82-151)."""
    film = td_mlp(w, "newt.mlp", emb)
    film_up = upsample_linear(film, exciter.shape[-1])
    C = exciter.shape[1]
    g_i, b_i, g_n, b_n = torch.split(film_up, C, 1)
    x = g_i * exciter + b_i                                          # FiLM, dynamic.py:6-8
    y = trainable_nonlinearity(w, x) if lut is None else lut_shaping_fn(lut, x, faithful_loop=faithful_loop)
    z = g_n * y + b_n
    out = F.conv1d(z, w["newt.mixer.0.weight"], w["newt.mixer.0.bias"])
    if return_parts:
        return out, dict(film=film, shaper_in=x, shaper_out=y)
    return out


def fir_noise_synth(w: Weights, H_re: torch.Tensor, noise: torch.Tensor, return_ir: bool = False):
    """FIRNoiseSynth.forward (generators.py:21-35): zero-phase IR design (irfft -> roll -> hann)
    then STFT-domain product with the noise STFT and istft(center=False)."""
    ir_length, hop = IR_LENGTH, CONTROL_HOP
    H_z = torch.complex(H_re, torch.zeros_like(H_re))
    h = torch.fft.irfft(H_z.transpose(1, 2))
    h = h.roll(ir_length // 2, -1)
    h = h * w["noise_synth.window"].view(1, 1, -1)
    Hf = torch.fft.rfft(h)
    window = torch.ones(ir_length)  # the reference passes window=None (rectangular)
    X = torch.stft(noise, ir_length, hop, window=window, return_complex=True).unsqueeze(0)
    Y = X * Hf.transpose(1, 2)
    y = torch.istft(Y, ir_length, hop, window=window, center=False)
    out = y.unsqueeze(1)[:, :, : H_re.shape[-1] * hop]
    if return_ir:
        return out, h
    return out


def fir_noise_literal(H_re: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """Time-domain statement of the same branch, float64 numpy: per-frame
    256-point CIRCULAR convolution of reflect-padded rectangular frames with the windowed
    zero-phase IR, overlap-add, divide by the {1,2} envelope.  H_re: [B,129,T]."""
    B, K, T = H_re.shape
    L, hop = IR_LENGTH, CONTROL_HOP
    n = np.arange(L)
    k = np.arange(1, K - 1)
    cosm = np.cos(2 * np.pi * np.outer(n, k) / L)                      # [256,127]
    H = H_re.astype(np.float64).transpose(0, 2, 1)                     # [B,T,129]
    h0 = (H[..., :1] + ((-1.0) ** n) * H[..., -1:] + 2.0 * H[..., 1:-1] @ cosm.T) / L
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * n / L)
    h = np.roll(h0, L // 2, axis=-1) * hann                            # [B,T,256]
    xp = np.pad(noise.astype(np.float64), (L // 2, L // 2), mode="reflect")
    out = np.zeros((B, hop * T + hop))
    for t in range(T):
        xt = xp[hop * t: hop * t + L]
        circ = xt[(n[:, None] - n[None, :]) % L]                       # circ[m,j] = x[(m-j) mod L]
        out[:, hop * t: hop * t + L] += h[:, t] @ circ.T
    env = np.full(hop * T + hop, 2.0)
    env[:hop] = 1.0
    env[-hop:] = 1.0
    return (out / env)[:, : hop * T]


def reverb(w: Weights, x: torch.Tensor) -> torch.Tensor:
    """Reverb.forward (shaping.py:161-173): x + circular convolution, length max(N, 32000),
    with [0, ir]."""
    ir_ = torch.cat((w["reverb.initial_zero"], w["reverb.ir"]), dim=-1)
    if x.shape[-1] > ir_.shape[-1]:
        ir_ = F.pad(ir_, (0, x.shape[-1] - ir_.shape[-1]))
        x_ = x
    else:
        x_ = F.pad(x, (0, ir_.shape[-1] - x.shape[-1]))
    return x + torch.fft.irfft(torch.fft.rfft(x_) * torch.fft.rfft(ir_))[..., : x.shape[-1]]


def reverb_literal(ir: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Fold form of the same op, float64: linear convolution with [0, ir]
    then out[n] = x[n] + ylin[n] + ylin[n+L], L = max(N, 32000)."""
    ir_ = np.concatenate(([0.0], ir.astype(np.float64).ravel()))
    N = x.shape[-1]
    L = max(N, ir_.shape[0])
    nfft = 1 << int(np.ceil(np.log2(N + ir_.shape[0])))
    ylin = np.fft.irfft(np.fft.rfft(x.astype(np.float64), nfft) * np.fft.rfft(ir_, nfft), nfft)
    tail = np.zeros_like(ylin[..., :N])
    m = min(N, nfft - L)
    tail[..., :m] = ylin[..., L: L + m]
    return x + ylin[..., :N] + tail


def reverb_causal(ir: np.ndarray, x: np.ndarray) -> np.ndarray:
    """The reverb as a causal (linear) convolution, float64: out[n] = x[n] + sum_j [0, ir][j] x[n-j].
    This is what a stream can produce — Reverb.forward (shaping.py:161-173) additionally wraps the tail
    ylin[n + L] back onto the start (reverb_literal), which needs the whole utterance.  Oracle of the
    streaming extension (nws_stream_push)."""
    ir_ = np.concatenate(([0.0], ir.astype(np.float64).ravel()))
    N = x.shape[-1]
    nfft = 1 << int(np.ceil(np.log2(N + ir_.shape[0])))
    ylin = np.fft.irfft(np.fft.rfft(x.astype(np.float64), nfft) * np.fft.rfft(ir_, nfft), nfft)
    return x + ylin[..., :N]


# -------------------------------------------------------------------- full path
def forward(w: Weights, f0: torch.Tensor, control: torch.Tensor, u_phase: torch.Tensor,
            noise: torch.Tensor, lut: Optional[torch.Tensor] = None, faithful_loop: bool = False,
            return_parts: bool = False):
    """NeuralWaveshaping.forward (neural_waveshaping.py:74-90) with the two RNG draws made
    explicit (`u_phase` = rand_like(rand_phase), `noise` = rand(128T-1); see draw_rng)."""
    with torch.no_grad():
        f0_up = upsample_linear(f0, f0.shape[-1] * CONTROL_HOP)
        exciter = render_exciter(w, f0_up, u_phase)
        emb = control_module(w, control)
        if return_parts:
            x, parts = newt(w, exciter, emb, lut=lut, faithful_loop=faithful_loop, return_parts=True)
        else:
            x, parts = newt(w, exciter, emb, lut=lut, faithful_loop=faithful_loop), None
        H = td_mlp(w, "h_generator", emb)
        nz = fir_noise_synth(w, H, noise)
        dry = torch.cat((x, nz), dim=1).sum(1)
        out = reverb(w, dry)
    if return_parts:
        parts.update(f0_up=f0_up, exciter=exciter, emb=emb, H=H, newt_out=x, noise_out=nz, dry=dry)
        return out, parts
    return out


# ------------------------------------------------------------------- utilities
def random_weights(seed: int = 0) -> Weights:
    """Random weights with the reference's shapes and roughly its init scales
    (Conv1d/GRU uniform ±1/sqrt(fan_in), input_scale ~ N(0,10²) shaping.py:21, reverb ir ~ N(0,1e-12)
    shaping.py:158).  NOT the reference's RNG stream — for reference-identical random-init weights
    use tests/golden/ (generated by oracle/gen_golden.py)."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*shape, generator=g) * 2 - 1) * b

    w: Weights = {}
    w["embedding.gru.weight_ih_l0"] = u((384, 2), 128)
    w["embedding.gru.weight_hh_l0"] = u((384, 128), 128)
    w["embedding.gru.bias_ih_l0"] = u((384,), 128)
    w["embedding.gru.bias_hh_l0"] = u((384,), 128)
    w["embedding.proj.weight"] = u((128, 128, 1), 128)
    w["embedding.proj.bias"] = u((128,), 128)
    w["harmonic_mixer.weight"] = u((64, 101, 1), 101)
    w["harmonic_mixer.bias"] = u((64,), 101)
    for prefix, out in (("newt.mlp", 256), ("h_generator", 129)):
        for j in (0, 3, 6):
            w["%s.net.%d.weight" % (prefix, j)] = u((128, 128, 1), 128)
            w["%s.net.%d.bias" % (prefix, j)] = u((128,), 128)
            w["%s.net.%d.layer_norm.weight" % (prefix, j + 1)] = 1 + 0.1 * torch.randn(128, generator=g)
            w["%s.net.%d.layer_norm.bias" % (prefix, j + 1)] = 0.1 * torch.randn(128, generator=g)
        w["%s.net.9.weight" % prefix] = u((out, 128, 1), 128)
        w["%s.net.9.bias" % prefix] = u((out,), 128)
    w["newt.shaping_fn.input_scale"] = torch.randn(1, 64, 1, generator=g) * 10
    w["newt.shaping_fn.net.0.weight"] = u((512, 1, 1), 1)
    w["newt.shaping_fn.net.0.bias"] = u((512,), 1)
    for j in (2, 4):
        w["newt.shaping_fn.net.%d.weight" % j] = u((512, 8, 1), 8)
        w["newt.shaping_fn.net.%d.bias" % j] = u((512,), 8)
    w["newt.shaping_fn.net.6.weight"] = u((64, 8, 1), 8)
    w["newt.shaping_fn.net.6.bias"] = u((64,), 8)
    w["newt.mixer.0.weight"] = u((1, 64, 1), 64)
    w["newt.mixer.0.bias"] = u((1,), 64)
    w["noise_synth.window"] = torch.hann_window(256)
    w["reverb.ir"] = torch.randn(1, 31999, generator=g) * 1e-6
    w["reverb.initial_zero"] = torch.zeros(1, 1)
    return w


def realistic_inputs(T: int, mean: np.ndarray, std: np.ndarray, B: int = 1):
    """The 'realistic' input set of SURVEY.md §8(d): vibrato around 440 Hz, slow loudness LFO,
    control normalised with a checkpoint's data_mean/std."""
    u = torch.linspace(0, 1, T)
    f0 = 440.0 * torch.pow(2.0, 0.5 * torch.sin(2 * math.pi * 1.5 * u))
    loud = 0.10 + 0.03 * torch.sin(2 * math.pi * 3 * u)
    control = torch.stack(((f0 - float(mean[0])) / float(std[0]), (loud - float(mean[1])) / float(std[1])))
    f0 = f0.view(1, 1, T).expand(B, 1, T).contiguous().float()
    control = control.view(1, 2, T).expand(B, 2, T).contiguous().float()
    return f0, control


def load_weights_npz(path: str) -> Weights:
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}
