"""Sub-modules of the drop-in model run their CUDA stage through the engine of the
NeuralWaveshaping instance that owns them (weights are loaded into the C handle as one set)."""
import weakref

# sub-module -> weak reference to its root.  Kept outside the modules' __dict__: a weak reference is a per-process
# cache, and inside the module it would make torch.save(model) / pickle fail.
_ROOTS = weakref.WeakKeyDictionary()


class BoundToRoot:
    """Mixin: the owning NeuralWaveshaping is looked up through a weak side table."""

    def _bind_root(self, root):
        _ROOTS[self] = weakref.ref(root)

    def _root(self):
        ref = _ROOTS.get(self)
        root = ref() if ref is not None else None
        if root is None:
            raise NotImplementedError(
                "%s.forward runs as a CUDA stage of a NeuralWaveshaping model; construct it through "
                "NeuralWaveshaping() (stand-alone evaluation is not provided, and there is no CPU fallback)"
                % type(self).__name__)
        return root
