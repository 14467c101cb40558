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
