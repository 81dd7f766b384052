"""Hop-rate building blocks — host-side mirror of the reference's modules/dynamic.py (FiLM :6-8,
TimeDistributedLayerNorm :11-17, TimeDistributedMLP :20-40).  The modules own the parameters
(same state-dict keys); the arithmetic runs in csrc/nws_encoder.cu (nws_linear128_kernel)."""
import gin
import torch.nn as nn

from ._bound import BoundToRoot


class FiLM(nn.Module):
    """gamma * x + beta.  Inside the model this is fused into the audio-rate kernel."""

    def forward(self, x, gamma, beta):
        return gamma * x + beta


class TimeDistributedLayerNorm(nn.Module):
    def __init__(self, size: int):
        super().__init__()
        self.layer_norm = nn.LayerNorm(size)

    def forward(self, x):
        raise NotImplementedError("TimeDistributedLayerNorm is fused into the TimeDistributedMLP CUDA stage")


@gin.configurable
class TimeDistributedMLP(nn.Module, BoundToRoot):
    """[Conv1d(k=1) -> LayerNorm -> LeakyReLU] x (depth-1) -> Conv1d(k=1) over [B, C, T]."""

    def __init__(self, in_size: int, hidden_size: int, out_size: int, depth: int = 3):
        super().__init__()
        if depth < 3:
            raise AssertionError("Depth must be at least 3")
        stack = []
        for i in range(depth):
            last = i == depth - 1
            stack.append(nn.Conv1d(in_size if i == 0 else hidden_size, out_size if last else hidden_size, 1))
            if not last:
                stack += [TimeDistributedLayerNorm(hidden_size), nn.LeakyReLU()]
        self.net = nn.Sequential(*stack)
        self.in_size, self.hidden_size, self.out_size, self.depth = in_size, hidden_size, out_size, depth

    def forward(self, x):
        root = self._root()
        which = root._mlp_index(self)
        return root._engine_for(x).td_mlp(which, x)
