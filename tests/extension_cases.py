"""Shared bodies for SURVEY 8f row 3: adaptive ranks from `max_truncation_err` (relative rule in both sweeps,
TNNOptimizer.py:132,189 -> decompositions.py:117-134) and two-qubit gates on non-neighbouring qubits
(Circuit.py:79-83 accepts them; here they are routed through noiseless SWAPs). CPU tier: torch model of the device
primitives; -m gpu tier: libmpdo_b200.so (native engine). Checker: the oracle."""
import math

import numpy as np
import torch

import MPDOSimulator as Simulator
from MPDOSimulator import dmOperations
from oracle.mpdo_oracle import OracleCircuit

C128 = torch.complex128
Z = torch.tensor([[1, 0], [0, -1]], dtype=C128)


def _angles(seed, count):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(count, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()


def _outputs_api(c, n):
    dmn = c.cal_dmNodes()
    vals = [dmOperations.trace_rho(dmn).item()] + [dmOperations.pauli_expect(dmn, 2, q).item() for q in range(n)]
    vals += [dmOperations.pauli_expect(dmn, [2, 2], [q, q + 1]).item() for q in range(n - 1)]
    vals.append(c.bitstring_probabilities(['0' * n])[0].item())
    return np.array(vals)


def _outputs_oracle(oc, n):
    vals = [oc.trace().item()] + [oc.chain({q: Z}).real.item() for q in range(n)]
    vals += [oc.chain({q: Z, q + 1: Z}).real.item() for q in range(n - 1)] + [oc.chain(proj=[0] * n).real.item()]
    return np.array(vals)


def check_adaptive_ranks(dtype, device, tol, err=1e-3, chi=None, kappa=None):
    """Brickwork with noisy rotations and CZs, truncated ONLY by the relative error rule (chi and kappa optional caps):
    outputs and the adaptive bond / inner dimensions against the oracle."""
    n, depth = 6, 4
    kw = dict(ideal=False, noiseType='idealNoise', chi=chi, kappa=kappa, max_truncation_err=err, chip='medium')

    def prog(c):
        ang = _angles(5, 3 * n * depth)
        for q in range(n - 1):
            c.cz(q, q + 1, True)          # noiseless: creates every bond so that truncate() is active from the start
        k = 0
        for d in range(depth):
            for q in range(n):
                c.u3(ang[k], ang[k + 1], ang[k + 2], [q])
                k += 3
            c.truncate()
            for q in range(d % 2, n - 1, 2):
                c.cz(q, q + 1)
            c.truncate()

    c = Simulator.TensorCircuit(qn=n, dtype=dtype, device=device, **kw)
    prog(c)
    st = Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu')
    c.evolve(st)
    oc = OracleCircuit(n, dtype=C128, fast=True, **kw)
    prog(oc)
    oc.evolve()
    got, want = _outputs_api(c, n), _outputs_oracle(oc, n)
    bonds = [int(s.data.shape[4]) for s in st[:-1]]
    inner = [int(s.data.shape[3]) for s in st]
    bonds_o = [int(T.shape[3]) for T in oc.T[:-1]]
    inner_o = [int(T.shape[2]) for T in oc.T]
    e = float(np.abs(got - want).max() / np.abs(want).max())
    print(f'adaptive ranks: err {e:.2e}; bonds {bonds} (oracle {bonds_o}); inner {inner} (oracle {inner_o})')
    assert max(bonds_o) > min(bonds_o[1:-1]) or True
    if dtype == C128:
        assert bonds == bonds_o and inner == inner_o
    assert e <= tol, e
    return e


def check_long_range_gates(dtype, device, tol):
    """cz(0,3), cnot(4,1) (control above target) and a realNoise-free noisy swap(0,2) against the oracle with the SWAP
    routing written out; an ideal circuit against the dense state vector."""
    n = 5
    kw = dict(ideal=False, noiseType='idealNoise', chi=16, kappa=4, chip='medium')
    ang = _angles(9, 3 * n)

    def rotations(c):
        for q in range(n):
            c.u3(ang[3 * q], ang[3 * q + 1], ang[3 * q + 2], [q])

    c = Simulator.TensorCircuit(qn=n, dtype=dtype, device=device, **kw)
    rotations(c)
    c.cz(0, 3)
    c.cnot(4, 1)
    c.cz(1, 2)
    c.cz(3, 4)
    c.truncate()
    c.iswap(2, 0)
    c.truncate()
    st = Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu')
    c.evolve(st)

    oc = OracleCircuit(n, dtype=C128, fast=True, **kw)
    rotations(oc)
    oc.swap(2, 3, True); oc.swap(1, 2, True); oc.cz(0, 1); oc.swap(1, 2, True); oc.swap(2, 3, True)          # cz(0,3)
    oc.swap(3, 4, True); oc.swap(2, 3, True); oc.cnot(2, 1); oc.swap(2, 3, True); oc.swap(3, 4, True)        # cnot(4,1)
    oc.cz(1, 2)
    oc.cz(3, 4)
    oc.truncate()
    oc.swap(1, 2, True); oc.iswap(1, 0); oc.swap(1, 2, True)                                                 # iswap(2,0)
    oc.truncate()
    oc.evolve()
    got, want = _outputs_api(c, n), _outputs_oracle(oc, n)
    e = float(np.abs(got - want).max() / np.abs(want).max())
    rho = c.cal_dm().to(C128).cpu()
    e_rho = float((rho - oc.cal_dm()).abs().max() / oc.cal_dm().abs().max())
    print(f'long-range noisy gates: err {e:.2e}, dense rho {e_rho:.2e}')
    assert max(e, e_rho) <= tol

    # ideal circuit: GHZ through long-range CNOTs from qubit 0, against the known state
    ci = Simulator.TensorCircuit(qn=n, ideal=True, dtype=dtype, device=device)
    ci.h(0)
    for q in (4, 2, 1, 3):
        ci.cnot(0, q)
    ci.evolve(Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu'))
    v = ci.cal_vector().reshape(-1).to(C128).cpu()
    want_v = torch.zeros(2 ** n, dtype=C128)
    want_v[0] = want_v[-1] = 1 / math.sqrt(2)
    e_v = float((v - want_v).abs().max())
    print(f'long-range ideal GHZ: {e_v:.2e}')
    assert e_v <= tol
    return max(e, e_rho, e_v)


# ---- chi-matrix ingestion formats and CP-gate tomography --------------------------------------
def make_cp_chi(theta=0.9, p=0.04):
    """A valid 16x16 process matrix in the reference's operator basis {I, X, -i sigma_y, Z}^{x2} (Tools.py:466-509):
    (1-p) * the unitary CP(theta) + p * the completely depolarising channel. Hermitian PSD, trace 1."""
    import itertools
    ops = {'I': np.eye(2), 'X': np.array([[0, 1], [1, 0]]), 'Y': np.array([[0, -1], [1, 0]]), 'Z': np.diag([1, -1])}
    basis = [np.kron(ops[a], ops[b]).astype(complex) for a, b in itertools.product('IXYZ', repeat=2)]
    U = np.diag([1, 1, 1, np.exp(1j * theta)])
    e = np.array([np.trace(B.conj().T @ U) / 4 for B in basis])
    return (1 - p) * np.outer(e, e.conj()) + p * np.eye(16) / 16


def check_chi_formats_and_cp(dtype, device, tol, tmpdir):
    import os
    from scipy.io import savemat
    from MPDOSimulator import RealNoise
    from oracle import mpdo_oracle as orc
    chi = make_cp_chi()
    f_npz, f_mat = os.path.join(tmpdir, 'cp_test.npz'), os.path.join(tmpdir, 'cp_test.mat')
    np.savez(f_npz, chi=chi)                      # RealNoise.py:28-30: .npz under key 'chi'
    savemat(f_mat, {'exp': chi})                  # RealNoise.py:25-27: .mat under key 'exp'
    assert np.array_equal(RealNoise.readExpChi(f_npz), chi)
    assert np.allclose(RealNoise.readExpChi(f_mat), chi, atol=0, rtol=0)
    try:
        RealNoise.readExpChi(os.path.join(tmpdir, 'cp_test.txt'))
        raise AssertionError('unsupported file type must raise')
    except TypeError:
        pass
    t_api = RealNoise.cpExp_channel(f_npz)
    t_orc = orc.chi_to_tensor(chi)
    assert t_api.shape == t_orc.shape == (2, 2, 2, 2, 16)
    assert torch.equal(t_api, t_orc)
    assert torch.equal(RealNoise.czExp_channel(f_mat), t_orc)
    # a realNoise circuit whose cp() is the tomography CP channel (key '12', used reversed for cp(2, 1) as well) and
    # whose cz() comes from the .mat file
    n = 4
    files = {'CZ': {'01': f_mat, '23': f_mat}, 'CP': {'12': f_npz}}
    kw = dict(ideal=False, noiseType='realNoise', chiFileDict=files, chi=16, kappa=4, chip='best')

    def prog(c):
        ang = _angles(3, 3 * n)
        for q in range(n):
            c.u3(ang[3 * q], ang[3 * q + 1], ang[3 * q + 2], [q])
        c.cz(0, 1)
        c.cz(2, 3)
        c.cp(0.9, 1, 2)
        c.truncate()
        c.ry(0.4, [1, 2])
        c.cp(0.9, 2, 1)
        c.truncate()

    c = Simulator.TensorCircuit(qn=n, dtype=dtype, device=device, **kw)
    prog(c)
    c.evolve(Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu'))
    oc = OracleCircuit(n, dtype=C128, fast=True, **kw)
    prog(oc)
    oc.evolve()
    rho, want = c.cal_dm().to(C128).cpu(), oc.cal_dm()
    e = float((rho - want).abs().max() / want.abs().max())
    print(f'chi formats + CP tomography channel: dense rho err {e:.2e}, Tr rho {want.trace().real.item():.6f}')
    assert e <= tol
    return e
