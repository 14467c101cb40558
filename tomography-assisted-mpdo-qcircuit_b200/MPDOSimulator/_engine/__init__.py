"""Device engine of the B200 build: C-ABI binding (lib), primitives (prims) and the update path (steps)."""
from .steps import Engine  # noqa: F401


def default_prims():
    """The product back end. Raises when libmpdo_b200.so or a CUDA device is missing (no CPU fallback)."""
    from .prims import CudaPrims
    return CudaPrims()


_PRIMS = None          # process-wide device back end, created on first use
_ENGINES = {}


def get_prims():
    global _PRIMS
    if _PRIMS is None:
        _PRIMS = default_prims()
    return _PRIMS


def engine_for(dtype):
    """Engine (kernel sequences of the update path) for a circuit dtype."""
    p = get_prims()
    key = (id(p), dtype)
    if key not in _ENGINES:
        import os
        if getattr(p, 'name', '') == 'cuda' and os.environ.get('MPDO_ENGINE', 'native') != 'py':
            from .native import NativeEngine          # one C call per step (csrc/engine.cu)
            _ENGINES[key] = NativeEngine(p, dtype)
        else:
            _ENGINES[key] = Engine(p, dtype)          # the same sequences over the Python primitive wrappers
    return _ENGINES[key]


GROUPING_MAX_BATCH = 4


def grouping_enabled(batch):
    """Whether same-shape strands (the bulk brick pairs of a layer, the per-site inner-index truncations) are stacked
    along the batch axis and run as ONE launch sequence instead of one CUDA stream each. For a single circuit (or a
    few) the layer is bound by launch and synchronisation latency and stacking wins (cfg2: 36.2 -> 34.3 ms per layer,
    a third fewer launches); for large parameter-sweep batches every launch already fills the device and one stream
    only serialises what separate streams overlap (cfg4: 8.4 vs 12.9 circuits/s, profiles/r2_grouping.md).
    MPDO_GROUPING=0 / 1 forces the choice."""
    import os
    knob = os.environ.get('MPDO_GROUPING')
    if knob is not None:
        return knob == '1'
    return int(batch) <= GROUPING_MAX_BATCH
