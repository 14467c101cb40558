#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference package from
/root/reference (read-only) on top of oracle/tn_shim (tensornetwork 0.4.6 is not installable offline; the shim
restates the handful of its calls the reference makes and forwards svd/qr to the reference's own patched
decompositions.py). Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin (a) the host-side operand builders (gate tensors, Kraus tensors, chi-matrix tensors) and
(b) the update path end to end (density matrices, traces, expectation values) for circuits small enough that
every decomposition takes the reference's full-SVD branch (numel < 10000), so the results are deterministic.
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'tn_shim'), REF]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import tensornetwork as tn  # noqa: E402  (the shim)

assert tn.__version__.endswith('shim')
import MPDOSimulator as Ref  # noqa: E402  (the reference itself)

assert Ref.__file__.startswith(REF), Ref.__file__
from MPDOSimulator import dmOperations as ref_dm  # noqa: E402
from MPDOSimulator.NoiseChannel import NoiseChannel  # noqa: E402
from MPDOSimulator.RealNoise import czExp_channel  # noqa: E402

tn.set_default_backend('pytorch')
C64, C128 = torch.complex64, torch.complex128
CZ_DEFAULT = os.path.join(REF, 'MPDOSimulator', 'chi', 'czDefault.mat')
CZ_IDEAL = os.path.join(REF, 'MPDOSimulator', 'chi', 'ideal_cz.mat')
out = {}


def npy(t):
    return t.detach().cpu().numpy()


# ---- (a) operands --------------------------------------------------------------------------------
for chip in ('best', 'medium', 'worst'):
    for dt, tag in ((C64, 'c64'), (C128, 'c128')):
        nc = NoiseChannel(chip=chip, dtype=dt, device='cpu')
        for name in ('decayTensor', 'dephasingTensor', 'dpCTensor', 'dpCTensor2', 'apdeCTensor'):
            out[f'noise/{chip}/{tag}/{name}'] = npy(getattr(nc, name))
out['chi/czDefault'] = npy(czExp_channel(filename=CZ_DEFAULT))
out['chi/ideal_cz'] = npy(czExp_channel(filename=CZ_IDEAL))

from MPDOSimulator.QuantumGates import SingleGates, XGates, YGates, ZGates, PhaseGates, DoubleGates  # noqa: E402

ANG = [0.3, 1.1, -2.2]
GATES = [
    (SingleGates, 'IGate', 0), (SingleGates, 'HGate', 0), (SingleGates, 'U1Gate', 1), (SingleGates, 'U3Gate', 3),
    (XGates, 'XGate', 0), (XGates, 'RXGate', 1), (XGates, 'CXGate', 0), (XGates, 'RXXGate', 1),
    (YGates, 'YGate', 0), (YGates, 'RYGate', 1), (YGates, 'CYGate', 0), (YGates, 'RYYGate', 1),
    (ZGates, 'ZGate', 0), (ZGates, 'RZGate', 1), (ZGates, 'CZGate', 0), (ZGates, 'RZZGate', 1),
    (PhaseGates, 'SGate', 0), (PhaseGates, 'SDGGate', 0), (PhaseGates, 'TGate', 0), (PhaseGates, 'PGate', 1),
    (PhaseGates, 'CPGate', 1),
    (DoubleGates, 'IIGate', 0), (DoubleGates, 'CNOTGate', 0), (DoubleGates, 'ISWAPGate', 0),
    (DoubleGates, 'SWAPGate', 0), (DoubleGates, 'PSWAPGate', 1), (DoubleGates, 'XXPlusYYGate', 2),
]
for mod, cls, npar in GATES:
    for dt, tag in ((C64, 'c64'), (C128, 'c128')):
        g = getattr(mod, cls)(*ANG[:npar], None, dtype=dt, device='cpu')
        out[f'gate/{cls}/{tag}'] = npy(g.tensor)
        out[f'gatemeta/{cls}'] = np.array([int(g.single), int(g.variational), g.rank])
        out[f'gatename/{cls}'] = np.array(g.name)
for cls in ('MeasureX', 'MeasureY', 'MeasureZ', 'Reset0', 'Reset1'):
    out[f'gate/{cls}/c128'] = npy(getattr(SingleGates, cls)(dtype=C128, device='cpu').tensor)


# ---- (b) circuits ----------------------------------------------------------------------------------
def brick(c, n, depth, seed, entangler='cz', ghz=True, trunc_after_1q=True, pre_u3=False):
    g = torch.Generator().manual_seed(seed)
    if pre_u3:   # noiseless random rotations first: breaks the GHZ symmetry (no exactly degenerate cuts)
        for q in range(n):
            th, ph, la = (torch.rand(3, generator=g) * 2 * math.pi).tolist()
            c.u3(th, ph, la, [q], True)
    if ghz:
        c.h(0)
        for i in range(n - 1):
            c.cnot(i, i + 1)
        c.truncate()
    for d in range(depth):
        for q in range(n):
            th, ph, la = (torch.rand(3, generator=g) * 2 * math.pi).tolist()
            c.u3(th, ph, la, [q])
        if trunc_after_1q:
            c.truncate()
        for q in range(d % 2, n - 1, 2):
            if entangler == 'rzz':
                c.rzz(float(torch.rand(1, generator=g) * 2 * math.pi), q, q + 1)
            else:
                getattr(c, entangler)(q, q + 1)
        c.truncate()


def run(tag, n, dt, prog, **kw):
    tn.randomized_svd_calls = 0
    c = Ref.TensorCircuit(qn=n, dtype=dt, device='cpu', **kw)
    prog(c)
    state = Ref.Tools.create_ket0Series(n, dtype=dt, device='cpu')
    c.evolve(state)
    out[f'{tag}/randomized_svd_calls'] = np.array(tn.randomized_svd_calls)
    out[f'{tag}/shapes'] = np.array([str((nd.axis_names, tuple(nd.tensor.shape))) for nd in state])
    dmn = c.cal_dmNodes()
    out[f'{tag}/trace'] = npy(ref_dm.trace_rho(dmn))
    out[f'{tag}/trace_rho2'] = npy(ref_dm.trace_rho2(dmn))
    out[f'{tag}/pauli_z'] = np.array([npy(ref_dm.pauli_expect(dmn, 2, q)) for q in range(n)])
    out[f'{tag}/pauli_xy01'] = npy(ref_dm.pauli_expect(dmn, [0, 1], [0, 1]))
    # dmOperations.expect is not exercised: it raises TypeError on torch >= 2.x (`Tensor(0.)`, dmOperations.py:141)
    out[f'{tag}/dm'] = npy(c.cal_dm())
    if kw.get('ideal', True):
        out[f'{tag}/vector'] = npy(c.cal_vector())
    return c


files5 = {'CZ': {f'{i}{i + 1}': CZ_DEFAULT for i in range(4)}, 'CP': {}}


def debug_py(c):  # test/debug.py:35-41
    c.h(0)
    for i in range(4):
        c.cnot(i, i + 1)
    c.truncate()


run('debug_py_c64', 5, C64, debug_py, ideal=False, noiseType='realNoise', chiFileDict=files5, chi=4, kappa=4, chip='best')
run('debug_py_c128', 5, C128, debug_py, ideal=False, noiseType='realNoise', chiFileDict=files5, chi=4, kappa=4, chip='best')
for dt, tag in ((C64, 'c64'), (C128, 'c128')):
    # full-SVD branch everywhere (randomized_svd_calls == 0 is recorded and asserted by the tests)
    run(f'ideal_noise_n4_{tag}', 4, dt, lambda c: brick(c, 4, 2, 21, ghz=False), ideal=False, noiseType='idealNoise',
        chi=4, kappa=2, chip='medium')
    run(f'ideal_noise_n3_{tag}', 3, dt, lambda c: brick(c, 3, 3, 31, pre_u3=True), ideal=False, noiseType='idealNoise',
        chi=4, kappa=2, chip='medium')
    run(f'unified_n4_{tag}', 4, dt, lambda c: brick(c, 4, 2, 22), ideal=False, noiseType='unified', chi=4, kappa=2,
        chip='medium')
    run(f'ideal_n5_{tag}', 5, dt, lambda c: brick(c, 5, 3, 23), ideal=True, chi=4)
    run(f'notrunc_n3_{tag}', 3, dt, lambda c: brick(c, 3, 1, 24), ideal=False, noiseType='idealNoise', chip='worst')
    files4 = {'CZ': {f'{i}{i + 1}': CZ_DEFAULT for i in range(3)}, 'CP': {}}
    run(f'realnoise_rzz_n4_{tag}', 4, dt, lambda c: brick(c, 4, 2, 5, entangler='rzz', ghz=False, trunc_after_1q=False),
        ideal=False, noiseType='realNoise', chiFileDict=files4, chi=2, kappa=1, chip='best')
    files3 = {'CZ': {f'{i}{i + 1}': CZ_DEFAULT for i in range(2)}, 'CP': {}}
    run(f'realnoise_cz_n3_{tag}', 3, dt, lambda c: brick(c, 3, 3, 41, ghz=False, trunc_after_1q=False), ideal=False,
        noiseType='realNoise', chiFileDict=files3, chi=4, kappa=2, chip='best')
    # the reference's randomized branch is taken here (numel >= 10000): loose pin only
    run(f'ideal_noise_n5_{tag}', 5, dt, lambda c: brick(c, 5, 3, 2), ideal=False, noiseType='idealNoise', chi=16,
        kappa=3, chip='medium')

np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
print('wrote', len(out), 'arrays to', os.path.join(HERE, 'reference_golden.npz'))
