"""Density-operator readout of the B200 build (reference MPDOSimulator/dmOperations.py): traces, purity,
expectation values and Pauli expectations of the un-contracted MPDO, evaluated as transfer-matrix chains
  E_k(O)[(l,l'),(r,r')] = sum_{s,s',a} O[s',s] T_k[l,s,a,r] conj(T_k[l',s',a,r'])
on the device. rho is never normalised (the reference does not normalise either)."""
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor, tensor

from . import _engine
from ._node import DenseNode

PAULI_DICT = {
    0: tensor([[0, 1], [1, 0]]),
    1: tensor([[0, -1j], [1j, 0]]),
    2: tensor([[1, 0], [0, -1]])
}


class _ConjNode(DenseNode):
    """Bra-side partner of a state node: same storage, lazily conjugated, physical axis named con_physics_k."""

    def __init__(self, partner: DenseNode):
        self._partner = partner
        self.index = partner.index
        self.name = f'con_{partner.name}'

    data = property(lambda self: self._partner.data.conj())
    has_left = property(lambda self: self._partner.has_left)
    has_right = property(lambda self: self._partner.has_right)
    has_inner = property(lambda self: self._partner.has_inner)

    def _axes(self):
        return [(f'con_{n}' if n.startswith('physics') else n, d) for n, d in self._partner._axes()]


class DmNodes(list):
    """What cal_dmNodes() returns: n ket-side nodes followed by n conjugated nodes (like the reference's
    `_state + _qubits_conj`), plus the set of qubits whose physical legs are already traced."""

    def __init__(self, state_nodes: List[DenseNode], reduced=None):
        self.state_nodes = list(state_nodes)
        self.conj_nodes = [_ConjNode(n) for n in self.state_nodes]
        super().__init__(self.state_nodes + self.conj_nodes)
        self.reduced = set(reduced or [])

    @property
    def qnumber(self):
        return len(self.state_nodes)

    def tensors(self):
        nodes = self.state_nodes
        B = max(n.data.shape[0] for n in nodes)
        return [n.data if n.data.shape[0] == B else n.data.expand(B, *n.data.shape[1:]).contiguous() for n in nodes]

    def engine(self):
        return _engine.engine_for(self.state_nodes[0].data.dtype)


def _as_dm(dmNodes) -> DmNodes:
    if isinstance(dmNodes, DmNodes):
        return dmNodes
    n = len(dmNodes) // 2
    return DmNodes(list(dmNodes)[:n])


def _out(v: Tensor) -> Tensor:
    return v[0] if v.shape[0] == 1 else v


def reduce_dmNodes(qubits_nodes, conj_qubits_nodes=None, residual_index=None, reduced_index=None):
    """Mark qubits as traced (reference :18-37 connects physics_k with con_physics_k)."""
    if reduced_index is None:
        return None
    dm = qubits_nodes if isinstance(qubits_nodes, DmNodes) else None
    reduced_index = [reduced_index] if isinstance(reduced_index, int) else reduced_index
    if not isinstance(reduced_index, list):
        raise TypeError('reduced_index should be int or list[int]')
    n = dm.qnumber if dm is not None else len(qubits_nodes)
    if reduced_index and max(reduced_index) >= n:
        raise ValueError(f'Reduced index should not be larger than the qubit number. {max(reduced_index)}-{n}')
    if dm is not None:
        dm.reduced.update(reduced_index)


def trace_rho(dmNodes) -> Tensor:
    """Tr rho (reference :40-48)."""
    dm = _as_dm(dmNodes)
    return _out(dm.engine().chain_value(dm.tensors()).real)


def trace_rho_rho(dmNodes_0, dmNodes_1=None) -> Tensor:
    """Tr(rho_0 rho_1) (reference :51-66)."""
    dm0 = _as_dm(dmNodes_0)
    dm1 = dm0 if dmNodes_1 is None else _as_dm(dmNodes_1)
    if dm1.qnumber != dm0.qnumber:
        raise ValueError('Density matrices must have the same number of nodes.')
    return _out(dm0.engine().chain_overlap(dm0.tensors(), dm1.tensors()).real)


def trace_rho2(dmNodes) -> Tensor:
    return trace_rho_rho(dmNodes)


def trace_composited_rho(*dmNodes) -> Tensor:
    """Reference :73-88 sums trace_rho_rho over the slices (sic)."""
    total = 0
    for d in dmNodes:
        total = total + trace_rho_rho(d)
    return total


def trace_composited_rho2(*dmNodes) -> Tensor:
    """Tr[(sum_i rho_i / slices)^2] (reference :91-132)."""
    k = len(dmNodes)
    diag = sum(trace_rho2(d) for d in dmNodes)
    cross = sum(trace_rho_rho(dmNodes[i], dmNodes[j]) for i in range(k) for j in range(i + 1, k))
    return (diag + 2 * cross) / (k ** 2)


def _site_ops(obs: Tensor, oq: List[int]):
    """Split an operator on len(oq) qubits into single-site factors if it is a product; else None."""
    return None


def expect(dmNodes, observables: Union[Tensor, List[Tensor]], oqs: Union[int, List]) -> Union[List[Tensor], Tensor]:
    """Tr(O rho) for each (observable, qubits) pair (reference :135-171); O is a dense 2^m x 2^m matrix on the
    listed qubits, applied through its operator-Schmidt (Pauli-product) expansion so that every term is a
    single transfer-matrix chain."""
    dm = _as_dm(dmNodes)
    qn = dm.qnumber
    oqs = [oqs] if isinstance(oqs, int) else oqs
    observables = [observables] if isinstance(observables, Tensor) else observables
    eng, Ts = dm.engine(), dm.tensors()
    dev = Ts[0].device
    paulis = [torch.eye(2, dtype=torch.complex128), PAULI_DICT[0].to(torch.complex128),
              PAULI_DICT[1].to(torch.complex128), PAULI_DICT[2].to(torch.complex128)]
    values = []
    for j, (obs, oq) in enumerate(zip(observables, oqs)):
        oq = [oq] if isinstance(oq, int) else list(oq)
        m = len(oq)
        try:
            obs_t = obs.reshape([2] * 2 * m)
        except RuntimeError:
            raise ValueError(f'Shape of the No.{j} obs is not valid, which is: {obs.shape}.')
        if m > qn:
            raise ValueError(f'Dim of No.{j} - oqs: {m} or obs: {m} exceeds the system size.')
        mat = obs_t.reshape(2 ** m, 2 ** m).to(torch.complex128).cpu()
        total = 0
        import itertools
        for combo in itertools.product(range(4), repeat=m):
            P = paulis[combo[0]]
            for c in combo[1:]:
                P = torch.kron(P, paulis[c])
            coeff = torch.trace(P.mH @ mat) / (2 ** m)
            if abs(coeff) < 1e-15:
                continue
            ops = {q: paulis[c].to(dev) for q, c in zip(oq, combo) if c != 0}
            total = total + coeff.to(dev) * eng.chain_value(Ts, ops)
        values.append(_out(total.real if isinstance(total, Tensor) else torch.zeros(1, dtype=torch.float64, device=dev)))
    return values


def pauli_expect(dmNodes, observables: Union[int, List[int], Tuple], oqs: Union[int, List[int], Tuple]):
    """Tr(P_{o_1} x ... x P_{o_m} rho) with 0 = X, 1 = Y, 2 = Z on the listed qubits (reference :174-201)."""
    dm = _as_dm(dmNodes)
    oqs = [oqs] if isinstance(oqs, int) else list(oqs)
    observables = [observables] if isinstance(observables, int) else list(observables)
    eng, Ts = dm.engine(), dm.tensors()
    ops = {q: PAULI_DICT[o].to(dtype=torch.complex128, device=Ts[0].device) for o, q in zip(observables, oqs)}
    return _out(eng.chain_value(Ts, ops).real)
