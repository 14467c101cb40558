"""MPDOSimulator - B200-native build of the noisy-gate update path of
WeiguoMa/Tomography-assisted-MPDO-QCircuit (drop-in for the reference package's public API:
TensorCircuit, Tools, dmOperations). The TensorNetwork-pytorch backend and the patched decompositions.py are
replaced by hand-written sm_100a CUDA kernels behind the C ABI in include/mpdo_b200.h; no CPU fallback."""
__latestUpdate__ = '10.17.2026'
__version__ = "1.0.0+b200"

from . import Tools
from . import dmOperations
from .Circuit import TensorCircuit

__all__ = ['TensorCircuit', 'Tools', 'dmOperations']
