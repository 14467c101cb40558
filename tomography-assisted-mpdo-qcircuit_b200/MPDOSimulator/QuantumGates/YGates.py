"""Y-family gates (reference QuantumGates/YGates.py): YGate, RYGate, CYGate, RYYGate."""
from torch import cos, sin

from .AbstractGate import make_gate

YGate = make_gate('YGate', 'Y', True, False, lambda: [[0, -1j], [1j, 0]])


def _ry(theta):
    tC, tS = cos(theta / 2), sin(theta / 2)
    return [[tC, -tS], [tS, tC]]


RYGate = make_gate('RYGate', 'RY', True, True, _ry, ('theta',))
CYGate = make_gate('CYGate', 'CY', False, False,
                   lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]])


def _ryy(theta):
    tC, tS = cos(theta / 2), sin(theta / 2)
    return [[tC, 0, 0, -tS], [0, tC, tS, 0], [0, -tS, tC, 0], [tS, 0, 0, tC]]


RYYGate = make_gate('RYYGate', 'RYY', False, True, _ryy, ('theta',))
