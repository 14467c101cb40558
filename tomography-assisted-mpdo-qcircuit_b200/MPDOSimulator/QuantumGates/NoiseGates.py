"""Tomography (chi-matrix) two-qubit gates (reference QuantumGates/NoiseGates.py): CZEXPGate, CPEXPGate.
`.tensor` is the (2,2,2,2,K) Kraus-like tensor [p0,p1,s0,s1,K] produced by RealNoise."""
import warnings
from typing import Optional, Union

from torch import Tensor, complex64

from .AbstractGate import QuantumGate
from ..RealNoise import czExp_channel, cpExp_channel


def _exp_gate(cls_name, gate_name, fallback, what):
    class _G(QuantumGate):
        def __init__(self, tensor: Optional[Tensor] = None, dtype=complex64, device: Union[str, int] = 'cpu'):
            super(_G, self).__init__(dtype=dtype, device=device)
            self.ideal = False
            self._matrix = tensor

        name = property(lambda self: gate_name)
        rank = property(lambda self: 5)
        dimension = property(lambda self: [[2, 2], [2, 2], [16]])
        single = property(lambda self: False)
        variational = property(lambda self: False)

        @property
        def tensor(self):
            if self._matrix is None:
                warnings.warn(f'No (sufficient) {what} input files, use default tensor.')
                return fallback().to(dtype=self.dtype, device=self.device)
            return self._matrix.to(dtype=self.dtype, device=self.device)

    _G.__name__ = _G.__qualname__ = cls_name
    return _G


CZEXPGate = _exp_gate('CZEXPGate', 'CZEXP', lambda: czExp_channel(filename='MPDOSimulator/chi/czDefault.mat'), 'CZ')
CPEXPGate = _exp_gate('CPEXPGate', 'CPEXP', lambda: cpExp_channel(), 'CP')
