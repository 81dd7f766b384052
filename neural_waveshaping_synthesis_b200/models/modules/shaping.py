"""Waveshapers and reverb — host-side mirror of the reference's modules/shaping.py (Sine :10-12,
TrainableNonlinearity :15-37, NEWT :40-79, FastNEWT :82-151, Reverb :154-173).  Same class names,
constructor signatures, gin bindings and state-dict keys; the arithmetic runs in
csrc/nws_audio.cu (shaper MLP / LUT inside the fused kernel) and csrc/nws_reverb.cu."""
import gin
import torch
import torch.nn as nn

from ... import engine as _engine
from ._bound import BoundToRoot
from .dynamic import FiLM, TimeDistributedMLP


class Sine(nn.Module):
    def forward(self, x: torch.Tensor):
        return torch.sin(x)


@gin.configurable
class TrainableNonlinearity(nn.Module):
    """Per-shaper sine MLP held as grouped 1x1 convolutions (one group per waveshaper)."""

    def __init__(self, channels, width, nonlinearity=nn.ReLU, final_nonlinearity=Sine, depth=3):
        super().__init__()
        self.input_scale = nn.Parameter(torch.randn(1, channels, 1) * 10)
        stack = []
        for i in range(depth):
            last = i == depth - 1
            stack.append(nn.Conv1d(channels if i == 0 else channels * width, channels if last else channels * width,
                                   1, groups=channels))
            stack.append(final_nonlinearity() if last else nonlinearity())
        self.net = nn.Sequential(*stack)
        self.channels, self.width, self.depth = channels, width, depth

    def shaper_tensors(self):
        convs = [m for m in self.net if isinstance(m, nn.Conv1d)]
        out = [self.input_scale]
        for c in convs:
            out += [c.weight, c.bias]
        return out

    def evaluate_on_grid(self, points: torch.Tensor) -> torch.Tensor:
        """out[c, i] = shaper_c(points[i]) — the call pattern of FastNEWT's table initialisation.
        Runs nws_shaper_eval on a CUDA device (parameters are staged there if they live on the CPU)."""
        if self.channels != 64 or self.width != 8 or self.depth != 4 or not all(
                isinstance(m, (nn.Conv1d, Sine)) for m in self.net):
            raise NotImplementedError("the CUDA shaper is built for 64 sine shapers of width 8, depth 4 (newt.gin)")
        dev = self.input_scale.device
        if dev.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("building a FastNEWT table needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        table = _engine.shaper_eval([t.detach().to(dev) for t in self.shaper_tensors()], points.to(dev))
        return table.to(self.input_scale.device)

    def forward(self, x):
        raise NotImplementedError(
            "TrainableNonlinearity runs per sample inside the fused audio-rate kernel (csrc/nws_audio.cu); "
            "use evaluate_on_grid() for table construction or NeuralWaveshaping.forward")


@gin.configurable
class NEWT(nn.Module, BoundToRoot):
    def __init__(self, n_waveshapers: int, control_embedding_size: int, shaping_fn_size: int = 16,
                 out_channels: int = 1):
        super().__init__()
        self.n_waveshapers = n_waveshapers
        self.mlp = TimeDistributedMLP(control_embedding_size, control_embedding_size, n_waveshapers * 4, depth=4)
        self.waveshaping_index = FiLM()
        self.shaping_fn = TrainableNonlinearity(n_waveshapers, shaping_fn_size, nonlinearity=Sine)
        self.normalising_coeff = FiLM()
        self.mixer = nn.Sequential(nn.Conv1d(n_waveshapers, out_channels, 1))

    def forward(self, exciter, control_embedding):
        raise NotImplementedError(
            "NEWT consumes the exciter inside the fused audio-rate kernel and never sees it as a tensor; "
            "call NeuralWaveshaping.forward (csrc/nws_audio.cu)")


class FastNEWT(NEWT):
    """NEWT with each shaper replaced by a lookup table sampled on linspace(table_min, table_max,
    table_size) and read with linear interpolation.  Shares mlp / FiLM / mixer with the source NEWT."""

    def __init__(self, newt: NEWT, table_size: int = 4096, table_min: float = -3.0, table_max: float = 3.0):
        super().__init__()  # gin-configured throw-away NEWT, as in the reference (consumes the same RNG)
        self.table_size = table_size
        self.table_min = table_min
        self.table_max = table_max
        self.n_waveshapers = newt.n_waveshapers
        self.mlp = newt.mlp
        self.waveshaping_index = newt.waveshaping_index
        self.normalising_coeff = newt.normalising_coeff
        self.mixer = newt.mixer
        self.lookup_table = self._init_lookup_table(newt, table_size, self.n_waveshapers, table_min, table_max)
        self.to(next(iter(newt.parameters())).device)

    def _init_lookup_table(self, newt: NEWT, table_size: int, n_waveshapers: int, table_min: float, table_max: float):
        device = next(iter(newt.parameters())).device
        # the grid is computed where the reference computes it (torch.linspace on the module's device)
        points = torch.linspace(table_min, table_max, table_size, device=device)
        with torch.no_grad():
            table = newt.shaping_fn.evaluate_on_grid(points)
        return nn.Parameter(table)

    def shaping_fn(self, x):  # shadows the sub-module, like the reference (shaping.py:136)
        raise NotImplementedError("the table lookup runs inside the fused audio-rate kernel (csrc/nws_audio.cu)")


@gin.configurable
class Reverb(nn.Module, BoundToRoot):
    def __init__(self, length_in_seconds, sr):
        super().__init__()
        self.ir = nn.Parameter(torch.randn(1, sr * length_in_seconds - 1) * 1e-6)
        self.register_buffer("initial_zero", torch.zeros(1, 1))

    def forward(self, x):
        root = self._root()
        return root._engine_for(x).reverb(x)
