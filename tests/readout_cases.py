"""Shared bodies of the sampling / readout-family tests. The CPU tier runs them on the torch
model of the device primitives (tests/cpu_prims.py, small shot counts); the -m gpu tier runs the same functions
through libmpdo_b200.so on cuda:0. The checker is the oracle's dense density matrix.

Reference behaviour covered: Circuit.py:297-332 (_conditional_prb), :344-387 (_conditional_batch_sample), :389-446
(sample incl. MeasureX / MeasureY rotations and `reduced=`), :448-467 (randomSample); dmOperations.py:51-70
(trace_rho_rho, trace_rho2), :73-132 (trace_composited_rho / rho2), :135-171 (expect with dense multi-qubit
observables), :174-201 (pauli_expect); Tools.py:528-551 (cal_fidelity), :242-273 (density2prob)."""
import itertools
import math

import numpy as np
import torch

import MPDOSimulator as Simulator
from MPDOSimulator import dmOperations
from oracle.mpdo_oracle import OracleCircuit

C128 = torch.complex128
H = torch.tensor([[1, 1], [1, -1]], dtype=C128) / math.sqrt(2)            # MeasureX (SingleGates.py:247)
MY = torch.tensor([[1, -1j], [1, 1j]], dtype=C128) / math.sqrt(2)         # MeasureY (SingleGates.py:281)
I2 = torch.eye(2, dtype=C128)


def program(c, n, seed, depth=2):
    g = torch.Generator().manual_seed(seed)
    for d in range(depth):
        for q in range(n):
            th, ph, la = (torch.rand(3, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
            c.u3(th, ph, la, [q])
        c.truncate()
        for q in range(d % 2, n - 1, 2):
            c.cz(q, q + 1)
        c.truncate()


def pair(n, seed, dtype, device, chi=8, kappa=3, depth=2):
    """(circuit evolved through the public API on `device`, oracle circuit, dense oracle rho in complex128)."""
    kw = dict(ideal=False, noiseType='idealNoise', chi=chi, kappa=kappa, chip='medium')
    c = Simulator.TensorCircuit(qn=n, dtype=dtype, device=device, **kw)
    program(c, n, seed, depth)
    c.evolve(Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu'))
    oc = OracleCircuit(n, dtype=C128, fast=True, **kw)   # exact mode, small-side LAPACK (tests/test_oracle_fast.py)
    program(oc, n, seed, depth)
    oc.evolve()
    return c, oc, oc.cal_dm()


def kron_all(ops):
    out = ops[0]
    for o in ops[1:]:
        out = torch.kron(out, o)
    return out


def marginal_probs(rho, n, orientation, measured):
    """Outcome distribution of measuring `measured` (ascending) in the bases `orientation` (0 X, 1 Y, 2 Z; one entry
    per measured qubit), all other qubits traced; normalised."""
    rot = {0: H, 1: MY, 2: I2}
    ops = [I2] * n
    for q, o in zip(measured, orientation):
        ops[q] = rot[o]
    U = kron_all(ops)
    p = (U @ rho @ U.mH).diagonal().real.reshape([2] * n)
    traced = [q for q in range(n) if q not in measured]
    if traced:
        p = p.sum(dim=traced)
    p = p.reshape(-1)
    return (p / p.sum()).numpy()


def chi2_ok(counts, probs, shots, nbits):
    """Pearson chi-square of the observed counts against `probs`; bins with expectation < 5 are pooled. Accepts up to
    5 sigma of the chi-square distribution (dof + 5 sqrt(2 dof)): loose enough for a fixed-seed statistical test,
    tight enough to catch a wrong basis, a wrong marginal or a reversed bit order."""
    obs = np.zeros(2 ** nbits)
    for key, v in counts.items():
        bits = ''.join(str(int(b)) for b in key) if not isinstance(key, str) else key
        obs[int(bits, 2)] += v
    exp = probs * shots
    big = exp >= 5
    o = np.append(obs[big], obs[~big].sum())
    e = np.append(exp[big], exp[~big].sum())
    keep = e > 0
    chi2 = float(((o[keep] - e[keep]) ** 2 / e[keep]).sum())
    dof = max(int(keep.sum()) - 1, 1)
    return chi2 <= dof + 5 * math.sqrt(2 * dof), chi2, dof


def check_sampling(dtype, device, shots):
    torch.manual_seed(1234)
    n = 5
    c, oc, rho = pair(n, 11, dtype, device)
    report = {}
    # full register, Z basis
    _, counts = c.sample(shots, _tqdm_disable=True)
    ok, chi2, dof = chi2_ok(counts, marginal_probs(rho, n, [2] * n, list(range(n))), shots, n)
    report['z'] = (chi2, dof)
    assert ok, ('Z', chi2, dof)
    # mixed X / Y / Z bases (MeasureX, MeasureY rotations)
    ori = [0, 1, 2, 0, 1]
    _, counts = c.sample(shots, orientation=ori, _tqdm_disable=True)
    ok, chi2, dof = chi2_ok(counts, marginal_probs(rho, n, ori, list(range(n))), shots, n)
    report['xyz'] = (chi2, dof)
    assert ok, ('XYZ', chi2, dof)
    # reduced register: qubits 1 and 3 are traced, the others measured in X, Z, Y
    measured, ori_r = [0, 2, 4], [0, 2, 1]
    _, counts = c.sample(shots, orientation=ori_r, reduced=[1, 3], _tqdm_disable=True)
    ok, chi2, dof = chi2_ok(counts, marginal_probs(rho, n, ori_r, measured), shots, 3)
    report['reduced'] = (chi2, dof)
    assert ok, ('reduced', chi2, dof)
    assert all(len(k) == 3 for k in counts)
    # only the last qubits measured (leading qubits traced into the initial left environment)
    _, counts = c.sample(shots, reduced=[0, 1], _tqdm_disable=True)
    ok, chi2, dof = chi2_ok(counts, marginal_probs(rho, n, [2, 2, 2], [2, 3, 4]), shots, 3)
    report['tail'] = (chi2, dof)
    assert ok, ('tail', chi2, dof)
    # randomSample: one list of outcomes per scheme
    schemes = [[2, 2, 2, 2, 2], [0, 0, 1, 1, 2]]
    res = c.randomSample(schemes, shots_per_scheme=shots // 2)
    assert len(res) == 2 and all(len(r) == shots // 2 and len(r[0]) == n for r in res)
    for sch, r in zip(schemes, res):
        cnt = {}
        for b in r:
            cnt[tuple(b)] = cnt.get(tuple(b), 0) + 1
        ok, chi2, dof = chi2_ok(cnt, marginal_probs(rho, n, sch, list(range(n))), shots // 2, n)
        assert ok, ('randomSample', sch, chi2, dof)
    # sequential sampler (one conditional chain per shot) and boolean results
    seq = c.sample(64, _tqdm_disable=True, _require_sequential_sample=True, _require_counts=False)
    assert len(seq) == 64 and all(set(s) <= {'0', '1'} and len(s) == n for s in seq)
    bl = c.sample(32, _tqdm_disable=True, sample_string=False, _require_bool_result=True, _require_counts=False)
    assert all(isinstance(v, bool) for v in bl[0])
    # a conditional probability against the dense rho: P(q2 = 1 | q0 = 1, q1 = 0)
    c._prepare_sampling(c.stateNodes, list(range(n)))
    p1 = c._conditional_prb([1, 0])
    pz = marginal_probs(rho, n, [2] * n, list(range(n))).reshape([2] * n)
    want = pz[1, 0, 1].sum() / pz[1, 0].sum()
    # the reference adds GLOBAL_MINIMUM = e*1e-8 to both outcome weights before normalising (Circuit.py:326-331)
    assert abs(p1 - want) < (1e-4 if dtype == torch.complex64 else 1e-6), (p1, want)
    return report


def check_readout_family(dtype, device, tol):
    n = 5
    c0, oc0, rho0 = pair(n, 21, dtype, device)
    c1, oc1, rho1 = pair(n, 22, dtype, device)
    c2, oc2, rho2 = pair(n, 23, dtype, device)
    d0, d1, d2 = c0.cal_dmNodes(), c1.cal_dmNodes(), c2.cal_dmNodes()
    rel = lambda got, want: abs(complex(got) - complex(want)) / abs(complex(want))
    out = {}
    out['trace'] = rel(dmOperations.trace_rho(d0).item(), rho0.trace())
    out['purity'] = rel(dmOperations.trace_rho2(d0).item(), (rho0 @ rho0).trace())
    out['overlap'] = rel(dmOperations.trace_rho_rho(d0, d1).item(), (rho0 @ rho1).trace())
    out['composited'] = rel(dmOperations.trace_composited_rho(d0, d1, d2).item(),
                            sum((r @ r).trace() for r in (rho0, rho1, rho2)))
    mean = (rho0 + rho1 + rho2) / 3
    out['composited2'] = rel(dmOperations.trace_composited_rho2(d0, d1, d2).item(), (mean @ mean).trace())
    # dense multi-qubit observables: a random Hermitian 4x4 on non-neighbouring qubits (1, 3) and an 8x8 on (0, 2, 4)
    g = torch.Generator().manual_seed(5)

    def herm(m):
        a = torch.complex(torch.randn(2 ** m, 2 ** m, generator=g, dtype=torch.float64),
                          torch.randn(2 ** m, 2 ** m, generator=g, dtype=torch.float64))
        return a + a.mH

    def embed(op, qs):
        """op on qubits qs (ascending) -> 2^n x 2^n."""
        m = len(qs)
        rest = [q for q in range(n) if q not in qs]
        full = torch.kron(op, torch.eye(2 ** (n - m), dtype=C128)).reshape([2] * (2 * n))
        order = list(qs) + rest
        inv = [order.index(q) for q in range(n)]
        return full.permute(inv + [n + i for i in inv]).reshape(2 ** n, 2 ** n)

    O2, O3 = herm(2), herm(3)
    vals = dmOperations.expect(d0, [O2, O3, torch.tensor([[1, 0], [0, -1]], dtype=C128)], [[1, 3], [0, 2, 4], 2])
    out['expect2'] = rel(vals[0].item(), (embed(O2, [1, 3]) @ rho0).trace())
    out['expect3'] = rel(vals[1].item(), (embed(O3, [0, 2, 4]) @ rho0).trace())
    out['expect1'] = rel(vals[2].item(), (embed(torch.tensor([[1, 0], [0, -1]], dtype=C128), [2]) @ rho0).trace())
    X, Y, Zm = dmOperations.PAULI_DICT[0].to(C128), dmOperations.PAULI_DICT[1].to(C128), dmOperations.PAULI_DICT[2].to(C128)
    out['pauli_xyz'] = rel(dmOperations.pauli_expect(d0, [0, 1, 2], [0, 2, 3]).item(),
                           (embed(kron_all([X, Y, Zm]), [0, 2, 3]) @ rho0).trace())
    # dense rho, reduced rho, fidelity and probabilities
    dm = c0.cal_dm().to(C128).cpu()
    out['dense'] = float((dm - rho0).abs().max() / rho0.abs().max())
    red = c0.cal_dm(reduced_index=[1, 4]).to(C128).cpu()
    want = rho0.reshape([2] * (2 * n))
    want = torch.einsum('abcdeAbCDe->acdACD', want).reshape(8, 8)
    out['reduced_dense'] = float((red - want).abs().max() / want.abs().max())
    # cal_fidelity zeroes eigenvalues below 1e-10 / 1e-12 (Tools.py:521,545), so F(rho, rho) is not exactly 1: compare
    # with the same function evaluated on the oracle's rho
    fid = Simulator.Tools.cal_fidelity(dm / dm.trace().real, rho0 / rho0.trace().real).item()
    fid_ref = Simulator.Tools.cal_fidelity(rho0 / rho0.trace().real, rho0 / rho0.trace().real).item()
    out['fidelity_self'] = abs(fid - fid_ref)
    on_dev = c0.cal_dm()
    other = c1.cal_dm()
    f_dev = Simulator.Tools.cal_fidelity(on_dev / on_dev.diagonal().sum().real, other / other.diagonal().sum().real)
    f_ref = Simulator.Tools.cal_fidelity(rho0 / rho0.trace().real, rho1 / rho1.trace().real)
    out['fidelity_01'] = abs(f_dev.item() - f_ref.item()) / f_ref.item()
    probs = Simulator.Tools.density2prob(dm, _dict=False)
    out['density2prob'] = float(np.abs(probs - (rho0.diagonal().real / rho0.trace().real).numpy()).max())
    bits = [list(b) for b in itertools.product([0, 1], repeat=n)]
    bp = c0.bitstring_probabilities(bits).cpu().numpy()
    out['bitstrings'] = float(np.abs(bp - rho0.diagonal().real.numpy()).max() / rho0.diagonal().real.max())
    for key, err in out.items():
        lim = max(1e-6, 50 * tol) if key.startswith('fidelity') else tol   # sqrt + eigenvalue thresholds amplify rounding
        assert err <= lim, (key, err, lim)
    return out
