"""Device engine of the B200 build: C-ABI binding (lib), primitives (prims) and the update path (steps)."""
from .steps import Engine  # noqa: F401


def default_prims():
    """The product back end. Raises when libmpdo_b200.so or a CUDA device is missing (no CPU fallback)."""
    from .prims import CudaPrims
    return CudaPrims()
