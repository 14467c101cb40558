"""Two-qubit gates (reference QuantumGates/DoubleGates.py): IIGate, CNOTGate, ISWAPGate, SWAPGate,
PSWAPGate, ArbDoubleGate, XXPlusYYGate. 4x4 matrices in the basis |q_oqs[0] q_oqs[1]>, reshaped (2,2,2,2).
PSWAPGate and XXPlusYYGate report the name 'SWAP', as the reference does (DoubleGates.py:167,243)."""
from typing import Optional, Union
from warnings import warn

from torch import Tensor, complex64, cos, exp, sin

from .AbstractGate import QuantumGate, make_gate

IIGate = make_gate('IIGate', 'II', False, False, lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
CNOTGate = make_gate('CNOTGate', 'CNOT', False, False,
                     lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
ISWAPGate = make_gate('ISWAPGate', 'ISWAP', False, False,
                      lambda: [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
SWAPGate = make_gate('SWAPGate', 'SWAP', False, False,
                     lambda: [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])


def _pswap(theta):
    tC, tS = cos(theta), sin(theta)
    return [[1, 0, 0, 0], [0, tC, tS, 0], [0, tS, tC, 0], [0, 0, 0, 1]]


PSWAPGate = make_gate('PSWAPGate', 'SWAP', False, True, _pswap, ('theta',))


def _xx_yy(theta, beta):
    tC, tS = cos(theta / 2), -1j * sin(theta / 2)
    return [[1, 0, 0, 0], [0, tC, tS * exp(1j * beta), 0], [0, tS * exp(-1j * beta), tC, 0], [0, 0, 0, 1]]


XXPlusYYGate = make_gate('XXPlusYYGate', 'SWAP', False, True, _xx_yy, ('theta', 'beta'))


class ArbDoubleGate(QuantumGate):
    """Arbitrary two-qubit gate; a (2,2,2,2,K) tensor is taken as an already-noisy gate."""

    def __init__(self, matrix: Tensor, ideal: Optional[bool] = None, dtype=complex64, device: Union[str, int] = 'cpu'):
        super(ArbDoubleGate, self).__init__(ideal=ideal, dtype=dtype, device=device)
        self._matrix = matrix.to(dtype=self.dtype, device=self.device)

    name = property(lambda self: 'ArbDoubleGate')
    rank = property(lambda self: 4)
    dimension = property(lambda self: [[2, 2], [2, 2]])
    single = property(lambda self: False)
    variational = property(lambda self: True)

    @property
    def tensor(self):
        if self._matrix.shape != (2, 2, 2, 2):
            warn('You are probably adding a noisy double qubit gate, current shape is {}'.format(self._matrix.shape))
        return self._matrix.reshape(2, 2, 2, 2, -1).squeeze()
