"""Kraus operands of the idealNoise / unified models (reference MPDOSimulator/NoiseChannel.py).

Attributes (same names and axis conventions as the reference, :40-46):
  decayTensor     [out, in, 2]        amplitude damping      K0 = diag(1, sqrt(1-q)), K1 = sqrt(q)|0><1|
  dephasingTensor [out, in, 3]        sqrt(1-q) I, sqrt(q)|0><0|, sqrt(q)|1><1|
  dpCTensor       [out, in, 4]        single-qubit depolarizing
  dpCTensor2      [o0, o1, i0, i1, 16] two-qubit depolarizing, Pauli order I,X,Y,Z, index j = 4*j0 + j1
  apdeCTensor     [out, in, 3]        amplitude-phase damping (built, never consumed by the circuit)
Host side only: these are operands of the absorption kernels."""
import warnings
from typing import Optional, Union

import numpy as np
import torch as tc

from .ChipInfo import ChipInformation
from .Tools import select_device


class NoiseChannel:
    def __init__(self, chip: Optional[str] = None, dtype=tc.complex64, device: Union[str, int] = 'cpu'):
        self.dtype = dtype
        self.device = select_device(device)
        self.chip = getattr(ChipInformation(), chip or 'worst')()
        self.bath_rate = self.chip.bath_rate
        self.decay_rate = self.chip.decay_rate
        self.dephasing_rate = self.chip.dephasing_rate
        self.T1, self.T2 = self.chip.T1, self.chip.T2
        self.GateTime = self.chip.gateTime
        self.dpc_errorRate = self.chip.dpc_errorRate
        mk = lambda rows: tc.tensor(rows, dtype=dtype, device=self.device)
        self._basisPauli = [mk([[1, 0], [0, 1]]), mk([[0, 1], [1, 0]]), mk([[0, -1j], [1j, 0]]), mk([[1, 0], [0, -1]])]
        self.decayTensor = self.decay(self.decay_rate, self.GateTime)
        self.dephasingTensor = self.dephasing(self.dephasing_rate, self.GateTime)
        self.dpCTensor = self.depolarization_noise_channel(p=self.dpc_errorRate)
        self.dpCTensor2 = self.depolarization_noise_channel(p=self.dpc_errorRate, qn=2)
        self.apdeCTensor = self.amp_phase_damping_error(time=self.GateTime, T1=self.T1, T2=self.T2)

    def _kraus(self, mats):
        # stacked [k, out, in] -> [out, in, k]
        return tc.tensor(mats, dtype=self.dtype, device=self.device).permute((1, 2, 0))

    def depolarization_noise_channel(self, p: float, qn: int = 1) -> tc.Tensor:
        """eps(rho) = (1 - (4^n - 1)p/4^n) rho + p/4^n sum_{P != I} P rho P  (NoiseChannel.py:51-91)."""
        if not 0 <= p <= 1:
            raise ValueError('Probability p must be in the range [0, 1].')
        d = 4 ** qn
        weights = [np.sqrt(1 - (d - 1) * p / d)] + [np.sqrt(p / d)] * (d - 1)
        ops = list(self._basisPauli)
        for _ in range(qn - 1):
            ops = [tc.kron(left, right) for left in ops for right in self._basisPauli]
        w = tc.diag(tc.tensor(weights, dtype=self.dtype, device=self.device))
        t = tc.einsum('ij, jfk -> fki', w, tc.stack(ops).to(device=self.device))
        return t.reshape([2] * (2 * qn) + [t.shape[-1]])

    def amp_phase_damping_error(self, time: float, T1: float, T2: float) -> tc.Tensor:
        """NoiseChannel.py:93-142."""
        if time < 0 or T1 <= 0 or T2 <= 0:
            raise ValueError('The time, T1, T2 must be greater than or equal to 0, '
                             'for some special cases time = 0 is allowed.')
        if time == 0:
            warnings.warn('The time is 0, which means the noise is not applied.')
        T2p = 2 * T1 * T2 / (2 * T1 - T2)
        p1 = 1 - np.exp(-time / T1)
        p2 = 1 - np.exp(-time / T2p)
        return self._kraus([[[1, 0], [0, np.sqrt(1 - (p1 + p2))]],
                            [[0, 0], [0, np.sqrt(p2)]],
                            [[0, np.sqrt(p1)], [0, 0]]])

    def decay(self, gamma: float, gate_time: float):
        """NoiseChannel.py:144-155."""
        if gate_time <= 0:
            raise ValueError('The gate time must be greater than 0, or set the gate to IDEAL.')
        q = 1 - np.exp(- self.bath_rate * gamma * gate_time)
        return self._kraus([[[1, 0], [0, np.sqrt(1 - q)]], [[0, np.sqrt(q)], [0, 0]]])

    def dephasing(self, gamma: float, gate_time: float):
        """NoiseChannel.py:157-170."""
        if gate_time <= 0:
            raise ValueError('The gate time must be greater than 0, or set the gate to IDEAL.')
        q = 1 - np.exp(- self.bath_rate * gamma * gate_time)
        sq, s1q = np.sqrt(q), np.sqrt(1 - q)
        return self._kraus([[[s1q, 0], [0, s1q]], [[sq, 0], [0, 0]], [[0, 0], [0, sq]]])
