"""Tomography chi-matrix -> Kraus-like gate tensor (reference MPDOSimulator/RealNoise.py).

chi = sum_i d_i u_i u_i^h (general eig, sorted descending, |d_i| > 1e-12 kept),
E_i = sqrt(d_i) sum_j U_ji P_j over the basis {I, X, -i sigma_y, Z}^{x2} (Tools.name2matrix), coefficients
below 1e-12 dropped, accumulated in complex64 whatever the circuit dtype (RealNoise.py:37,147-158), stacked and
permuted to [p0, p1, s0, s1, K] (:168-169). Host side only."""
import os
from typing import Dict, List, Optional, Union

import numpy as np
import torch as tc
from numpy.linalg import eig

from .Tools import gates_list, name2matrix

__all__ = ['czExp_channel', 'cpExp_channel']


def readExpChi(filename: Optional[str] = None):
    if filename is None:
        raise FileNotFoundError('No file found.')
    if '.mat' in filename:
        from scipy.io import loadmat
        return loadmat(filename)['exp']
    if '.npz' in filename:
        return np.load(filename)['chi']
    raise TypeError('Current file-type is not supported.')


def _error_generators(chi, tol: float = 1e-12) -> Dict:
    """{'E_i': {basis name: coefficient}} (or a flat {name: coefficient} when a single generator survives)."""
    vals, vecs = eig(chi)
    order = vals.argsort()[::-1]
    vals, vecs = vals[order], vecs[:, order]
    # the reference counts the eigenvalues above tol and then takes that many LEADING eigenpairs (:91-99)
    kept = range(int(np.sum(np.abs(vals) > tol)))
    names = gates_list(int(np.log10(chi.shape[0]) / np.log10(4)))
    gens = []
    for i in kept:
        coeff = vecs[:, i] * np.sqrt(vals[i])
        gens.append({names[j]: (coeff[j] if np.abs(coeff[j]) >= tol else 0 + 0j) for j in range(len(names))})
    if len(gens) == 1:
        return {k: v for k, v in gens[0].items() if v != 0}
    # the reference shares one index list between all generators, so every generator carries every basis
    # name that is non-zero in any of them; the extra entries are exact zeros and do not change the sum
    used = [n for n in names if any(g[n] != 0 for g in gens)]
    return {f'E_{i}': {n: g[n] for n in used} for i, g in enumerate(gens)}


def noisyTensor(chi, gate_factor: Optional[Dict] = None, dtype=tc.complex64,
                device: Union[int, str] = 'cpu') -> List[tc.Tensor]:
    if gate_factor is None:
        gate_factor = _error_generators(chi)
    nested = any(isinstance(v, dict) for v in gate_factor.values())
    groups = list(gate_factor.values()) if nested else [gate_factor]
    out = []
    for group in groups:
        acc = tc.zeros((2, 2, 2, 2), dtype=dtype, device=device)
        for name, value in group.items():
            acc += tc.reshape(tc.tensor(value, dtype=dtype, device=device) * name2matrix(name), shape=(2, 2, 2, 2))
        out.append(acc)
    return out


def _channel(filename, dtype, device):
    stacked = tc.stack(noisyTensor(readExpChi(filename=filename)))
    return tc.einsum('ijlmn -> jlmni', stacked).to(dtype=dtype, device=device)


def czExp_channel(filename: Optional[str] = None, dtype=tc.complex64, device: Union[int, str] = 'cpu'):
    if filename is None:
        filename = os.path.join(os.path.dirname(__file__), 'data/chi/czDefault.mat')
    return _channel(filename, dtype, device)


def cpExp_channel(filename: Optional[str] = None, dtype=tc.complex64, device: Union[int, str] = 'cpu'):
    if filename is None:
        filename = os.path.join(os.path.dirname(__file__), 'data/chi/cpDefault.mat')
    return _channel(filename, dtype, device)
