"""Drop-in `NeuralWaveshaping` — host-side mirror of the reference's
models/neural_waveshaping.py (ControlModule :17-26, NeuralWaveshaping :29-90).

Same gin-configurable names, constructor arguments, sub-module names and state-dict keys, so
`gin.parse_config_file(...); NeuralWaveshaping()`, `load_from_checkpoint`, `model.newt =
FastNEWT(model.newt)`, `.eval()`, `.to(device)` and `model(f0, control)` behave as the reference
scripts expect (scripts/time_forward_pass.py:41-51, time_buffer_sizes.py:35-68,
resynthesise_dataset.py:47-59).  `forward` runs entirely in libnws_b200.so (hand-written sm_100a
CUDA behind the C ABI in include/nws_b200.h); it requires a CUDA device — there is no CPU path.

Training (the reference's LightningModule steps, :92-165) is out of scope.
"""
import pickle

import gin
import torch
import torch.nn as nn

from .. import _lib
from ..engine import NwsEngine
from .modules._bound import BoundToRoot
from .modules.dynamic import TimeDistributedMLP
from .modules.generators import FIRNoiseSynth, HarmonicOscillator
from .modules.shaping import NEWT, FastNEWT, Reverb

gin.external_configurable(nn.GRU, module="torch.nn")
gin.external_configurable(nn.Conv1d, module="torch.nn")


@gin.configurable
class ControlModule(nn.Module, BoundToRoot):
    """GRU(control_size -> hidden) over frames, then a 1x1 convolution to the embedding."""

    def __init__(self, control_size: int, hidden_size: int, embedding_size: int):
        super().__init__()
        self.gru = nn.GRU(control_size, hidden_size, batch_first=True)
        self.proj = nn.Conv1d(hidden_size, embedding_size, 1)

    def forward(self, x):
        return self._root()._engine_for(x).control_embedding(x)


class _PermissiveUnpickler(pickle.Unpickler):
    """Lightning checkpoints pickle classes (ModelCheckpoint, AttributeDict) that need not be
    installed; only `state_dict` and `hyper_parameters` are read, so unknown classes become dicts."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (dict,), {"__module__": module, "__hash__": lambda self: id(self),
                                        "__setstate__": lambda self, state: None})


class _PermissivePickle:
    __name__ = "nws_permissive_pickle"
    Unpickler = _PermissiveUnpickler

    @staticmethod
    def load(f, **kw):
        return _PermissiveUnpickler(f, **kw).load()


_GENERATION = __import__("itertools").count()


def _version_of(t: torch.Tensor) -> int:
    """In-place modification counter of a tensor; tensors created under torch.inference_mode() do not track one
    (reading it raises): fall back to the pointer alone for those."""
    try:
        return t._version
    except Exception:
        return -1


@gin.configurable
class NeuralWaveshaping(nn.Module):
    def __init__(self, n_waveshapers: int, control_hop: int, sample_rate: float = 16000,
                 learning_rate: float = 1e-3, lr_decay: float = 0.9, lr_decay_interval: int = 10000,
                 log_audio: bool = False):
        super().__init__()
        self.hparams = dict(n_waveshapers=n_waveshapers, control_hop=control_hop, sample_rate=sample_rate,
                            learning_rate=learning_rate, lr_decay=lr_decay, lr_decay_interval=lr_decay_interval,
                            log_audio=log_audio)
        self.learning_rate = learning_rate
        self.lr_decay = lr_decay
        self.lr_decay_interval = lr_decay_interval
        self.control_hop = control_hop
        self.log_audio = log_audio
        self.sample_rate = sample_rate

        # construction order == the reference's, so a seeded construction yields the same weights
        self.embedding = ControlModule()
        self.osc = HarmonicOscillator()
        self.harmonic_mixer = nn.Conv1d(self.osc.n_harmonics, n_waveshapers, 1)
        self.newt = NEWT()
        with gin.config_scope("noise_synth"):
            self.h_generator = TimeDistributedMLP()
            self.noise_synth = FIRNoiseSynth()
        self.reverb = Reverb()

        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_loaded", {})
        if control_hop != 128 or n_waveshapers != 64 or int(sample_rate) != 16000:
            raise NotImplementedError("the CUDA path is built for gin/models/newt.gin "
                                      "(control_hop 128, 64 waveshapers, 16 kHz)")

    # ------------------------------------------------------------------ copies / pickles
    # The engines own ctypes handles (one C context per device): they are per-instance caches, never state.  A copy
    # or an unpickled model starts without them and builds its own on its first forward — sharing a handle between
    # two modules would free it twice.
    _TRANSIENT = ("_engines", "_loaded", "_wt_cache")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._TRANSIENT:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_loaded", {})

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__setstate__({k: copy.deepcopy(v, memo) for k, v in self.__getstate__().items()})
        for m in new.modules():
            if isinstance(m, BoundToRoot):
                m._bind_root(new)
        return new

    # ------------------------------------------------------------------ engine management
    def _mlp_index(self, module) -> int:
        if module is self.newt.mlp:
            return 0
        if module is self.h_generator:
            return 1
        raise NotImplementedError("this TimeDistributedMLP is not part of the model")

    def _state_for_engine(self):
        sd = {}
        for k, v in self.state_dict(keep_vars=True).items():
            sd[k] = v
        # a FastNEWT keeps the source shaper weights out of reach (its method shadows the sub-module)
        if isinstance(self.newt, FastNEWT):
            for k, v in self.newt._modules["shaping_fn"].state_dict(keep_vars=True).items():
                sd["newt.shaping_fn." + k] = v
        return sd

    def _weight_tensors(self):
        """The 49 tensors of the C ABI's NwsTensor order, cached per (module-structure) so the per-call
        change check is 49 version reads instead of a state_dict() walk."""
        key = (id(self.newt), id(self.newt._modules.get("shaping_fn")), id(self.embedding), id(self.reverb))
        cache = self.__dict__.get("_wt_cache")
        if cache is None or cache[0] != key:
            sd = self._state_for_engine()
            cache = (key, sd, [sd[k] for k in _lib.TENSOR_KEYS], next(_GENERATION))
            object.__setattr__(self, "_wt_cache", cache)
        return cache[1], cache[2]

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace parameter storage without touching version counters
        object.__setattr__(self, "_wt_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        object.__setattr__(self, "_wt_cache", None)
        return super().load_state_dict(*args, **kwargs)

    def _engine_for(self, like: torch.Tensor, lane: int = 0) -> NwsEngine:
        """The engine of the input's device with this module's current weights (and lookup table) loaded.  `lane` > 0
        selects a further, independent engine of the same device (own C context: streams, workspace, scheduler
        counters) so that two forwards can be in flight on two streams at once (streaming.HostPipeline).

        Called on every forward, so the up-to-date check is kept cheap (a 256-sample forward is ~35 us of GPU work):
        sub-module identity (`model.newt = FastNEWT(model.newt)`), moves and casts (`_apply`), `load_state_dict`, and
        in-place updates of any of the 49 tensors (their version counters) are noticed; re-pointing a single
        parameter's `.data` at other storage is not — call `model.load_state_dict(model.state_dict())` after that."""
        dev = like.device
        if dev.type != "cuda":
            raise RuntimeError("NeuralWaveshaping (B200) runs on CUDA only: inputs are on %s. There is no CPU "
                               "fallback; move the model and inputs with .to('cuda')." % dev)
        slot = dev if lane == 0 else (dev, lane)
        eng = self._engines.get(slot)
        if eng is None:
            eng = NwsEngine(dev)
            self._engines[slot] = eng
            for m in self.modules():
                if isinstance(m, BoundToRoot):
                    m._bind_root(self)
        sd, tensors = self._weight_tensors()
        try:
            vsum = sum([t._version for t in tensors])
        except Exception:     # tensors created under torch.inference_mode() do not track versions
            vsum = -1
        fast = self.newt.lookup_table if isinstance(self.newt, FastNEWT) else None
        tag = self._loaded.get(slot)
        key = (self._wt_cache[3], vsum)      # (generation of the tensor list, sum of the in-place version counters)
        if tag is None or tag[0] != key:
            if tensors[0].device != dev:
                raise RuntimeError("the model's parameters are on %s but the inputs are on %s" % (tensors[0].device, dev))
            for m in self.modules():
                if isinstance(m, BoundToRoot):
                    m._bind_root(self)
            eng.load_weights(sd)
            tag = (key, None)
        if fast is not None:
            lsig = (fast.data_ptr(), _version_of(fast), self.newt.table_min, self.newt.table_max)
            if tag[1] != lsig:
                eng.set_lut(fast, self.newt.table_min, self.newt.table_max)
                tag = (tag[0], lsig)
        self._loaded[slot] = tag
        return eng

    # ------------------------------------------------------------------ reference API
    def render_exciter(self, f0):
        raise NotImplementedError("the exciter exists only inside the fused audio-rate kernel; "
                                  "see NwsEngine.audio(..., want_exciter=True) for a debugging tap")

    def get_embedding(self, control):
        return self.embedding(control)

    def forward(self, f0, control, phase_shift=None, noise=None):
        """f0 [B,1,T] Hz, control [B,C>=2,T] (channels 0,1 used) -> audio [B, T*control_hop].

        `phase_shift` ([101], the uniform draw of the oscillator's random phase, i.e. torch.rand_like
        (rand_phase)) and `noise` ([128*T-1] uniform) are optional test hooks replacing the two RNG
        draws of the reference forward; by default both come from an on-device Philox stream seeded from
        torch.initial_seed()."""
        if not isinstance(f0, torch.Tensor) or not isinstance(control, torch.Tensor):
            raise TypeError("f0 and control must be tensors")
        return self.forward_lane(0, f0, control, phase_shift, noise)

    def forward_lane(self, lane, f0, control, phase_shift=None, noise=None, out=None):
        """forward() through engine `lane` of the input's device (lane 0 = the one every other entry point uses; extension,
        not in the reference).  A lane is an independent C context — streams, workspace, scheduler counters — so forwards
        issued on different lanes from different CUDA streams run concurrently (streaming.HostPipeline keeps two in
        flight); forwards on one lane must be issued on one stream at a time.  `out`: optional result buffer [B, 128*T]."""
        eng = self._engine_for(f0, lane)
        return eng.forward(f0, control, u_phase=phase_shift, noise=noise, use_lut=isinstance(self.newt, FastNEWT), out=out)

    _forward_lane = forward_lane

    def synthesise_from_host(self, f0, control, out=None, phase_shift=None, noise=None):
        """Host tensors in, host tensor out through nws_forward_host (H2D, forward, D2H, sync)."""
        dev = next(self.parameters()).device
        eng = self._engine_for(torch.empty(0, device=dev))
        B, _, T = f0.shape
        if out is None:
            out = torch.empty(B, T * self.control_hop, dtype=torch.float32, pin_memory=True)
        return eng.forward_host(f0.contiguous(), control.contiguous(), out, phase_shift, noise,
                                use_lut=isinstance(self.newt, FastNEWT))

    def stream(self, batch_size: int = 1, max_frames: int = 32):
        """Stateful streaming synthesis (extension): see streaming.SynthStream."""
        from ..streaming import SynthStream
        return SynthStream(self, batch_size, max_frames)

    # ------------------------------------------------------------------ Lightning-compatible loading
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
        """Reads a PyTorch-Lightning checkpoint written by the reference's training script
        (`state_dict` + `hyper_parameters`), without needing pytorch_lightning."""
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False,
                          pickle_module=_PermissivePickle)
        hp = dict(ckpt.get("hyper_parameters", {}) or {})
        hp.update(kwargs)
        names = ("n_waveshapers", "control_hop", "sample_rate", "learning_rate", "lr_decay", "lr_decay_interval",
                 "log_audio")
        model = cls(**{k: hp[k] for k in names if k in hp})
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    def save_hyperparameters(self, *args, **kwargs):
        pass

    def log(self, *args, **kwargs):
        pass

    def configure_optimizers(self):
        raise NotImplementedError("training is outside the scope of the B200 forward path")

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is outside the scope of the B200 forward path")

    validation_step = training_step
    test_step = training_step
