"""X-family gates (reference QuantumGates/XGates.py): XGate, RXGate, CXGate, RXXGate."""
from torch import cos, sin

from .AbstractGate import make_gate

XGate = make_gate('XGate', 'X', True, False, lambda: [[0, 1], [1, 0]])


def _rx(theta):
    tC, tS = cos(theta / 2), sin(theta / 2)
    return [[tC, -1j * tS], [-1j * tS, tC]]


RXGate = make_gate('RXGate', 'RX', True, True, _rx, ('theta',))
CXGate = make_gate('CXGate', 'CX', False, False,
                   lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])


def _rxx(theta):
    tC, tS = cos(theta / 2), sin(theta / 2)
    return [[tC, 0, 0, -1j * tS], [0, tC, -1j * tS, 0], [0, -1j * tS, tC, 0], [-1j * tS, 0, 0, tC]]


RXXGate = make_gate('RXXGate', 'RXX', False, True, _rxx, ('theta',))
