"""Sub-modules of the drop-in model run their CUDA stage through the engine of the
NeuralWaveshaping instance that owns them (weights are loaded into the C handle as one set)."""
import weakref


class BoundToRoot:
    """Mixin: `self._nws_root` is a weak reference to the owning NeuralWaveshaping."""
    _nws_root_ref = None

    def _bind_root(self, root):
        object.__setattr__(self, "_nws_root_ref", weakref.ref(root))

    def _root(self):
        root = self._nws_root_ref() if self._nws_root_ref is not None else None
        if root is None:
            raise NotImplementedError(
                "%s.forward runs as a CUDA stage of a NeuralWaveshaping model; construct it through "
                "NeuralWaveshaping() (stand-alone evaluation is not provided, and there is no CPU fallback)"
                % type(self).__name__)
        return root
