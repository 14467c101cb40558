"""Z-family gates (reference QuantumGates/ZGates.py): ZGate, RZGate, CZGate, RZZGate."""
from torch import exp

from .AbstractGate import make_gate

ZGate = make_gate('ZGate', 'Z', True, False, lambda: [[1, 0], [0, -1]])
RZGate = make_gate('RZGate', 'RZ', True, True,
                   lambda theta: [[exp(-1j * theta / 2), 0], [0, exp(1j * theta / 2)]], ('theta',))
CZGate = make_gate('CZGate', 'CZ', False, False,
                   lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]])


def _rzz(theta):
    mi, pl = exp(-1j * theta / 2), exp(1j * theta / 2)
    return [[mi, 0, 0, 0], [0, pl, 0, 0], [0, 0, pl, 0], [0, 0, 0, mi]]


RZZGate = make_gate('RZZGate', 'RZZ', False, True, _rzz, ('theta',))
