"""MPDOSimulator - B200-native build of the noisy-gate update path of
WeiguoMa/Tomography-assisted-MPDO-QCircuit (drop-in for the reference package's public API:
TensorCircuit, Tools, dmOperations). The TensorNetwork-pytorch backend and the patched decompositions.py are
replaced by hand-written sm_100a CUDA kernels behind the C ABI in include/mpdo_b200.h; no CPU fallback."""
import os as _os

# Load every kernel of libmpdo_b200.so when the CUDA context is created instead of at its first launch: which kernels a
# layer needs is data dependent (probes of the top-kappa iteration, eigen route for shapes the Cholesky kernels cannot
# schedule), and a first launch in the middle of a run costs a module load - measured as a single 100-150 ms step in
# the first process of a fresh machine (cold file cache). Only effective if set before CUDA is initialised.
_os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')

__latestUpdate__ = '10.18.2026'
__version__ = "1.0.0+b200"

from . import Tools
from . import dmOperations
from .Circuit import TensorCircuit

__all__ = ['TensorCircuit', 'Tools', 'dmOperations']
