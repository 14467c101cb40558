"""Helpers of the B200 build (reference MPDOSimulator/Tools.py): initial states, device selection, the
Pauli-product basis of RealNoise, density2prob, fidelity, counting. Plotting stays out of scope (needs
matplotlib, presentation only) and raises if called without it."""
import itertools
import warnings
from collections import Counter
from functools import reduce
from typing import Dict, List, Optional, Union

import numpy as np
import torch
from torch import Tensor, complex64

from ._node import DenseNode, replicate_nodes  # noqa: F401

__all__ = [
    'create_ket0Series', 'create_ket1Series', 'create_ketPlusSeries', 'create_ketMinusSeries',
    'create_ketRandomSeries', 'create_ketHadamardSeries', 'plot_histogram', 'density2prob', 'select_device',
    'tc_expect', 'EdgeName2AxisName', 'count_item', 'cal_fidelity', 'random_measurementScheme', 'replicate_nodes',
]

PUBLIC_SQRT2 = 1 / np.sqrt(2)


def EdgeName2AxisName(_nodes):
    """No-op: dense nodes derive their axis names from the layout (the reference needs this after every
    tensornetwork operation, Tools.py:44-72)."""
    return None


def ket0(dtype, device: Union[str, int] = 'cpu'):
    return torch.tensor([1. + 0.j, 0. + 0.j], dtype=dtype, device=device)


def ket1(dtype, device: Union[str, int] = 'cpu'):
    return torch.tensor([0. + 0.j, 1. + 0.j], dtype=dtype, device=device)


def ket_hadamard(dtype, device: Union[str, int] = 'cpu'):
    return PUBLIC_SQRT2 * torch.tensor([1., 1.], dtype=dtype, device=device)


def ket_plus(dtype, device: Union[str, int] = 'cpu'):
    return PUBLIC_SQRT2 * torch.tensor([1., 1.], dtype=dtype, device=device)


def ket_minus(dtype, device: Union[str, int] = 'cpu'):
    return PUBLIC_SQRT2 * torch.tensor([1., -1.], dtype=dtype, device=device)


def _series(qnumber, vec):
    assert qnumber > 0
    return [DenseNode(vec.clone(), k, name=f'qubit_{k}') for k in range(qnumber)]


def create_ket0Series(qnumber: int, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, ket0(dtype, device))


def create_ket1Series(qnumber: int, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, ket1(dtype, device))


def create_ketHadamardSeries(qnumber: int, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, ket_hadamard(dtype, device))


def create_ketPlusSeries(qnumber: int, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, ket_plus(dtype, device))


def create_ketMinusSeries(qnumber: int, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, ket_minus(dtype, device))


def create_ketRandomSeries(qnumber: int, tensor: Tensor, dtype=complex64, device: Union[str, int] = 'cpu') -> list:
    return _series(qnumber, tensor.to(dtype=dtype, device=device))


def tc_expect(oper, state):
    """<O> for state vectors (<psi|O|psi>) or density matrices (Tr O rho), or lists thereof."""
    def one(o, s):
        if s.dim() == 1:
            return torch.einsum('i, ij, j', s.conj(), o, s)
        if s.dim() == 2:
            return torch.trace(torch.matmul(o, s))
        raise ValueError("State must be a vector or matrix.")

    if isinstance(oper, Tensor) and isinstance(state, Tensor):
        return one(oper, state)
    if isinstance(oper, (list, tuple)) and isinstance(state, Tensor):
        return torch.tensor([one(o, state) for o in oper], dtype=state.dtype)
    if isinstance(state, (list, tuple)):
        return torch.tensor([one(oper, x) for x in state], dtype=oper.dtype)
    raise TypeError('Arguments must be torch.Tensors or lists thereof')


def density2prob(rho_in: Tensor, bases: Optional[Dict] = None, tol: Optional[float] = None,
                 _dict: Optional[bool] = True) -> Union[Dict, np.ndarray]:
    """Normalised probabilities <b|rho|b> over the computational basis or the given bases (Tools.py:242-273)."""
    rho_in = rho_in.detach().cpu()
    qn = int(np.log2(rho_in.shape[0]))
    if bases is None:
        probs = [abs(rho_in[i, i]).item() for i in range(2 ** qn)]
        names = [''.join(b) for b in itertools.product('01', repeat=qn)]
    else:
        try:
            vecs, names = bases['Bases'], bases['BasesName']
        except (ValueError, KeyError, TypeError):
            raise ValueError("The input bases should be a dict with format "
                             "{'Bases': List[torch_tensor], 'BasesName': List[str]}")
        probs = [abs(tc_expect(rho_in, v.view(-1))).item() for v in vecs]
    total = sum(probs)
    if _dict:
        return {n: p / total for n, p in zip(names, probs) if tol is None or p >= tol}
    return np.array(probs) / total


def plot_histogram(*args, **kwargs):
    raise NotImplementedError('plotting is outside the scope of the B200 build (presentation only, needs matplotlib)')


def select_device(device: Optional[Union[str, int]] = None):
    """Reference Tools.py:452-463: a string is returned as is; otherwise cuda:<n> when CUDA is available."""
    if isinstance(device, str):
        return device
    if torch.cuda.is_available():
        return 'cuda:0' if device is None else f'cuda:{device}'
    warnings.warn('CUDA is not available, use CPU instead.')
    return 'cpu'


def gates_list(N: int, basis_gates: Optional[List[str]] = None) -> List[str]:
    basis_gates = basis_gates or ['I', 'X', 'Y', 'Z']
    return [''.join(g) for g in itertools.product(basis_gates, repeat=N)]


def name2matrix(operation_name: str, dtype=complex64, device: Union[str, int] = 'cpu'):
    """Kronecker product of I, X, Y := -i sigma_y, Z named by the string (Tools.py:488-509)."""
    ops = {
        'I': torch.eye(2, dtype=dtype, device=device),
        'X': torch.tensor([[0, 1], [1, 0]], dtype=dtype, device=device),
        'Y': -1j * torch.tensor([[0, -1j], [1j, 0]], dtype=dtype, device=device),
        'Z': torch.tensor([[1, 0], [0, -1]], dtype=dtype, device=device),
    }
    return reduce(torch.kron, [ops[c] for c in operation_name])


def sqrt_matrix(density_matrix: Tensor) -> Tensor:
    evs, vecs = torch.linalg.eigh(density_matrix)
    evs = torch.where(evs > 1e-10, evs, torch.zeros_like(evs))
    return vecs @ torch.diag(torch.sqrt(evs)).to(vecs.dtype) @ vecs.conj().transpose(-2, -1)


def cal_fidelity(rho: Tensor, sigma: Tensor) -> Tensor:
    """F = (Tr sqrt(sqrt(rho) sigma sqrt(rho)))^2 (Tools.py:528-551)."""
    if rho.shape != sigma.shape:
        raise ValueError('The shape of rho and sigma should be equal.')
    if rho.shape[0] != rho.shape[1] or rho.shape[0] != sigma.shape[1]:
        raise ValueError('The shape of rho and sigma should be square.')
    sr = sqrt_matrix(rho)
    evs = torch.linalg.eigvalsh(sr @ sigma @ sr)
    evs = torch.where(evs > 1e-12, evs, torch.zeros_like(evs))
    return torch.sum(torch.sqrt(evs)) ** 2


def count_item(data):
    def key(x):
        if isinstance(x[0], tuple):
            return int(''.join(str(int(v)) for v in x[0]))
        if isinstance(x[0], str):
            return x[0]
        return int(x[0])

    counted = dict(Counter([tuple(item) if isinstance(item, list) else item for item in data]))
    return dict(sorted(counted.items(), key=key))


def random_measurementScheme(qnumber: int, amount: int) -> List[List[int]]:
    return np.random.randint(0, 3, size=(amount, qnumber)).tolist()
