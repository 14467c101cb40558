"""Single-qubit gates (reference MPDOSimulator/QuantumGates/SingleGates.py): IGate, HGate, U1Gate,
U2Gate, U3Gate, ArbSingleGate, MeasureX/Y/Z, Reset0/1. Matrices are M[out, in]."""
from typing import Optional, Union
from warnings import warn

import numpy as np
from torch import Tensor, complex64, cos, exp, sin
from torch import tensor as torch_tensor

from .AbstractGate import QuantumGate, make_gate

_S2 = np.sqrt(2)

IGate = make_gate('IGate', 'I', True, False, lambda: [[1, 0], [0, 1]])
HGate = make_gate('HGate', 'H', True, False, lambda: [[1 / _S2, 1 / _S2], [1 / _S2, -1 / _S2]])
# the reference divides an exact +-1 matrix by sqrt(2) (SingleGates.py:64); 1/sqrt(2) rounds identically in
# complex64 and differs by at most one ulp in complex128, where _h_exact below is used instead.


def _h_tensor(self):
    return torch_tensor([[1, 1], [1, -1]], dtype=self.dtype, device=self.device) / _S2


HGate.tensor = property(_h_tensor)

U1Gate = make_gate('U1Gate', 'U1', True, True, lambda theta: [[1, 0], [0, exp(1j * theta)]], ('theta',))


def _u2(phi, lam):
    # reference SingleGates.py:139 reads an undefined self._lam; this is the evidently intended matrix
    lM, pM = exp(1j * lam), exp(1j * phi)
    return [[1 / _S2, -lM / _S2], [pM / _S2, lM * pM / _S2]]


U2Gate = make_gate('U2Gate', 'U2', True, True, _u2, ('phi', 'lam'))


def _u3(theta, phi, lam):                                     # reference SingleGates.py:179-185
    lM, pM = exp(1j * lam), exp(1j * phi)
    tC, tS = cos(theta / 2), sin(theta / 2)
    return [[tC, -lM * tS], [pM * tS, lM * pM * tC]]


U3Gate = make_gate('U3Gate', 'U3', True, True, _u3, ('theta', 'phi', 'lam'))


class ArbSingleGate(QuantumGate):
    """Arbitrary single-qubit gate; a (2,2,K) tensor is taken as an already-noisy gate."""

    def __init__(self, tensor: Tensor, ideal: Optional[bool] = None, dtype=complex64, device: Union[str, int] = 'cpu'):
        super(ArbSingleGate, self).__init__(ideal=ideal, dtype=dtype, device=device)
        self._matrix = tensor.to(dtype=self.dtype, device=self.device)

    name = property(lambda self: 'ArbSingleGate')
    rank = property(lambda self: 2)
    dimension = property(lambda self: [2, 2])
    single = property(lambda self: True)
    variational = property(lambda self: False)

    @property
    def tensor(self):
        if self._matrix.shape != (2, 2):
            warn('You are probably adding a noisy single qubit gate, current shape is {}'.format(self._matrix.shape))
        return self._matrix.reshape(2, 2, -1).squeeze()


def _fixed_measure(cls_name, gate_name, data, scale=1.0):
    class _G(QuantumGate):
        def __init__(self, dtype=complex64, device: Union[str, int] = 'cpu'):
            super(_G, self).__init__(ideal=True, dtype=dtype, device=device)
            self._matrix = scale * torch_tensor(data, dtype=self.dtype, device=self.device)

        name = property(lambda self: gate_name)
        tensor = property(lambda self: self._matrix)
        rank = property(lambda self: 2)
        dimension = property(lambda self: [2, 2])
        single = property(lambda self: True)
        variational = property(lambda self: False)

    _G.__name__ = _G.__qualname__ = cls_name
    return _G


MeasureX = _fixed_measure('MeasureX', 'MeasureX', [[1, 1], [1, -1]], 1 / _S2)
MeasureY = _fixed_measure('MeasureY', 'MeasureY', [[1, -1j], [1, 1j]], 1 / _S2)
MeasureZ = _fixed_measure('MeasureZ', 'MeasureZ', [[1, 0], [0, 1]])
Reset0 = _fixed_measure('Reset0', 'Reset0', [[1, 0], [0, 0]])
Reset1 = _fixed_measure('Reset1', 'Reset1', [[0, 0], [0, 1]])
