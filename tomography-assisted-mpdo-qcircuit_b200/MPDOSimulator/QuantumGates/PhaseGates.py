"""Phase gates (reference QuantumGates/PhaseGates.py): SGate, SDGGate, TGate, PGate, CPGate."""
import numpy as np
from torch import exp

from .AbstractGate import make_gate

SGate = make_gate('SGate', 'S', True, False, lambda: [[1, 0], [0, 1j]])
SDGGate = make_gate('SDGGate', 'Sdg', True, False, lambda: [[1, 0], [0, -1j]])
TGate = make_gate('TGate', 'T', True, False, lambda: [[1, 0], [0, (1 + 1j) / np.sqrt(2)]])
PGate = make_gate('PGate', 'P', True, True, lambda theta: [[1, 0], [0, exp(theta * 1j)]], ('theta',))
CPGate = make_gate('CPGate', 'CP', False, True,
                   lambda theta: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, exp(1j * theta)]], ('theta',))
