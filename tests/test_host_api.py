"""CPU tier: the host-side API (MPDOSimulator.TensorCircuit & friends) driven through the torch-CPU model of
the device primitives (tests/cpu_prims.py, injected through the engine's test hook) and compared with the
oracle. This checks the orchestration / bookkeeping; the CUDA kernels themselves are checked by -m gpu."""
import math
import os

import pytest
import torch

import MPDOSimulator as Simulator
from MPDOSimulator import _engine, dmOperations
from cpu_prims import CpuPrims
from harness import rel
from oracle.mpdo_oracle import OracleCircuit

C64, C128 = torch.complex64, torch.complex128
CHI_DIR = os.path.join(os.path.dirname(Simulator.__file__), 'chi')


@pytest.fixture(autouse=True)
def cpu_model_prims():
    _engine._PRIMS = CpuPrims()   # CPU model of the device primitives, injected by the test
    yield
    _engine._PRIMS = None


def program(c, n, depth, seed, entangler='cz', ghz=True, trunc_after_1q=True):
    g = torch.Generator().manual_seed(seed)
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


def both(n, depth, seed, dtype, **kw):
    prog = {k: kw.pop(k) for k in list(kw) if k in ('entangler', 'ghz', 'trunc_after_1q')}
    circ = Simulator.TensorCircuit(qn=n, dtype=dtype, device='cpu', **kw)
    oc = OracleCircuit(n, dtype=dtype, **kw)
    program(circ, n, depth, seed, **prog)
    program(oc, n, depth, seed, **prog)
    state = Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu')
    circ.evolve(state)
    oc.evolve()
    return circ, oc, state


def test_ideal_circuit_vector_and_ghz():
    n = 5
    circ = Simulator.TensorCircuit(qn=n, ideal=True, dtype=C128, device='cpu')
    circ.h(0)
    for i in range(n - 1):
        circ.cnot(i, i + 1)
    state = Simulator.Tools.create_ket0Series(n, dtype=C128)
    circ.evolve(state)
    v = circ.cal_vector()
    assert v.shape == (2 ** n, 1)
    assert abs(v[0, 0].abs() ** 2 - 0.5) < 1e-12 and abs(v[-1, 0].abs() ** 2 - 0.5) < 1e-12
    probs = Simulator.Tools.density2prob(circ.cal_dm())
    assert abs(probs['0' * n] - 0.5) < 1e-12 and abs(probs['1' * n] - 0.5) < 1e-12
    assert circ.stateNodes is state
    assert state[1].axis_names == ['bond_0_1', 'physics_1', 'bond_1_2']
    assert state[0]['bond_0_1'].dimension == 2


@pytest.mark.parametrize('dtype,tol', [(C128, 1e-10)])
@pytest.mark.parametrize('kw', [
    dict(ideal=False, noiseType='idealNoise', chi=16, kappa=3, chip='medium'),
    dict(ideal=False, noiseType='unified', chi=8, kappa=3, chip='medium'),
    dict(ideal=True, chi=4),
])
def test_noisy_evolution_matches_oracle(dtype, tol, kw):
    circ, oc, _ = both(5, 3, 2, dtype, **kw)
    assert rel(circ.cal_dm().to(C128), oc.cal_dm().to(C128)) < tol
    dm_nodes = circ.cal_dmNodes()
    assert len(dm_nodes) == 10
    assert abs(dmOperations.trace_rho(dm_nodes).item() - oc.trace().item()) < tol
    Z = torch.tensor([[1, 0], [0, -1]], dtype=dtype)
    for q in (0, 3):
        want = oc.chain({q: Z}).real.item()
        assert abs(dmOperations.pauli_expect(dm_nodes, 2, q).item() - want) < tol
        assert abs(dmOperations.expect(dm_nodes, Z, q)[0].item() - want) < tol
    ZZ = torch.kron(Z, Z)
    want = oc.chain({1: Z, 2: Z}).real.item()
    assert abs(dmOperations.expect(dm_nodes, [ZZ], [[1, 2]])[0].item() - want) < tol
    rho = oc.cal_dm().to(C128)
    assert abs(dmOperations.trace_rho2(dm_nodes).item() - torch.trace(rho @ rho).real.item()) < tol
    p = circ.bitstring_probabilities(['00000', '10101'])
    assert abs(p[0].item() - rho[0, 0].real.item()) < tol
    assert abs(p[1].item() - rho[21, 21].real.item()) < tol
    red = circ.cal_dm(reduced_index=[0, 4]).to(C128)
    assert rel(red, oc.rdm([1, 2, 3]).to(C128)) < tol


def test_realnoise_chi_matrix_circuit():
    files = {'CZ': {f'{i}{i + 1}': os.path.join(CHI_DIR, 'czDefault.mat') for i in range(3)}, 'CP': {}}
    kw = dict(ideal=False, noiseType='realNoise', chiFileDict=files, chi=8, kappa=4, chip='best',
              entangler='rzz', ghz=False, trunc_after_1q=False)
    circ, oc, _ = both(4, 2, 5, C128, **kw)
    assert circ.last_stats['noisy_2q_updates'] == oc.stats['updates_2q_noisy'] > 0
    assert rel(circ.cal_dm().to(C128), oc.cal_dm().to(C128)) < 1e-10


def test_ideal_cz_chi_equals_ideal_cz():
    files = {'CZ': {'01': os.path.join(CHI_DIR, 'ideal_cz.mat')}, 'CP': {}}
    circ = Simulator.TensorCircuit(qn=2, ideal=False, noiseType='realNoise', chiFileDict=files, chip='best',
                                   dtype=C128, device='cpu')
    circ.h([0, 1])
    circ.cz(0, 1)
    st = Simulator.Tools.create_ket0Series(2, dtype=C128)
    circ.evolve(st)
    ideal = Simulator.TensorCircuit(qn=2, ideal=True, dtype=C128, device='cpu')
    ideal.h([0, 1])
    ideal.cz(0, 1)
    st2 = Simulator.Tools.create_ket0Series(2, dtype=C128)
    ideal.evolve(st2)
    assert rel(circ.cal_dm().to(C128), ideal.cal_dm().to(C128)) < 1e-6   # chi tensors are built in complex64


def test_parameter_sweep_batch_matches_single_circuits():
    n, B = 4, 3
    g = torch.Generator().manual_seed(3)
    angles = torch.rand(2, n, 3, B, generator=g, dtype=torch.float64) * 2 * math.pi
    kw = dict(qn=n, ideal=False, noiseType='idealNoise', chi=8, kappa=3, chip='medium', dtype=C128, device='cpu')

    def prog(c, ang):
        for d in range(2):
            for q in range(n):
                c.u3(ang[d, q, 0], ang[d, q, 1], ang[d, q, 2], [q])
            c.truncate()
            for q in range(d % 2, n - 1, 2):
                c.cz(q, q + 1)
            c.truncate()

    batched = Simulator.TensorCircuit(**kw)
    prog(batched, angles)
    st = Simulator.Tools.create_ket0Series(n, dtype=C128)
    batched.evolve(st)
    rho_b = batched.cal_dm()
    assert rho_b.shape == (B, 2 ** n, 2 ** n)
    for b in range(B):
        single = Simulator.TensorCircuit(**kw)
        prog(single, angles[..., b].clone())
        s1 = Simulator.Tools.create_ket0Series(n, dtype=C128)
        single.evolve(s1)
        assert rel(rho_b[b], single.cal_dm()) < 1e-10


def test_truncate_is_noop_until_connected_and_errors():
    circ = Simulator.TensorCircuit(qn=3, ideal=False, noiseType='idealNoise', chi=2, kappa=1, chip='medium',
                                   dtype=C128, device='cpu')
    circ.h(0)
    circ.truncate()            # bonds missing: skipped entirely, kappa too (Circuit.py:476-481)
    circ.barrier()
    st = Simulator.Tools.create_ket0Series(3, dtype=C128)
    circ.evolve(st)
    assert st[0].data.shape[3] == 1   # last layer is not truncate -> the closing kappa step ran (kappa = 1)
    with pytest.raises(ValueError):
        Simulator.TensorCircuit(qn=2, ideal=False, noiseType='bogus')
    bad = Simulator.TensorCircuit(qn=2, ideal=True, dtype=C128, device='cpu')
    bad.x(5)
    with pytest.raises(ValueError):
        bad.evolve(Simulator.Tools.create_ket0Series(2, dtype=C128))
    with pytest.raises(ValueError):
        Simulator.TensorCircuit(qn=2).rx(1, 0)          # int angle (AbstractGate.py:42-51)
    var = Simulator.TensorCircuit(qn=2, ideal=False, noiseType='idealNoise', chip='medium', dtype=C128, device='cpu')
    var.rzz(0.3, 0, 1)
    with pytest.raises(ValueError):
        var.evolve(Simulator.Tools.create_ket0Series(2, dtype=C128))


def test_sampling_statistics():
    torch.manual_seed(0)
    n = 3
    circ = Simulator.TensorCircuit(qn=n, ideal=True, dtype=C128, device='cpu')
    circ.h(0)
    circ.cnot(0, 1)
    circ.cnot(1, 2)
    st = Simulator.Tools.create_ket0Series(n, dtype=C128)
    circ.evolve(st)
    samples, counts = circ.sample(2000, _tqdm_disable=True)
    assert set(counts) <= {'000', '111'}
    assert abs(counts['000'] / 2000 - 0.5) < 0.05
    _, cx = circ.sample(500, orientation=[0, 0, 0], _tqdm_disable=True)   # X basis: even parity only
    assert all(k.count('1') % 2 == 0 for k in cx)


def test_cz_pair_fusion_matches_one_split_per_gate(monkeypatch):
    """complex64 realNoise: the two tomography CZs of an rzz are applied as one merge-and-split with the composite
    16 x 16 Kraus tensor (Circuit._fuse_pairs). Same state as one split per gate up to the skipped intermediate
    rank rule (e * 1e-8 absolute), far inside the complex64 tolerance; complex128 circuits never fuse."""
    n = 4
    files = {'CZ': {f'{i}{i + 1}': os.path.join(CHI_DIR, 'czDefault.mat') for i in range(n - 1)}, 'CP': {}}
    kw = dict(ideal=False, noiseType='realNoise', chiFileDict=files, chi=8, kappa=3, chip='best')
    monkeypatch.setenv('MPDO_GROUPING', '0')     # count one split call per pair (same-shape pairs are stacked otherwise)

    def run(dtype):
        c = Simulator.TensorCircuit(qn=n, dtype=dtype, device='cpu', **kw)
        program(c, n, 2, 5, entangler='rzz', ghz=False, trunc_after_1q=False)
        calls = []
        eng = c._engine()
        orig = eng.split_2q
        monkeypatch.setattr(eng, 'split_2q', lambda *a, **k: (calls.append(a[2].shape[-1]), orig(*a, **k))[1])
        c.evolve(Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu'))
        monkeypatch.setattr(eng, 'split_2q', orig)
        return c.cal_dm(), calls

    rho_f, calls_f = run(C64)
    # 3 rzz -> 3 fused splits; the composite 16 x 16 = 256 Kraus operators are rewritten as the <= 16 a two-qubit
    # channel needs (Circuit._compress_kraus: an isometry on an index that is only traced against its conjugate)
    assert len(calls_f) == 3 and all(k <= 16 for k in calls_f)
    monkeypatch.setenv('MPDO_NO_KRAUS_COMPRESSION', '1')
    rho_u, calls_u = run(C64)
    monkeypatch.delenv('MPDO_NO_KRAUS_COMPRESSION')
    assert all(k == 256 for k in calls_u)
    assert rel(rho_f, rho_u) < 2e-6                              # same channel: complex64 rounding only
    monkeypatch.setenv('MPDO_NO_FUSE', '1')
    rho_s, calls_s = run(C64)
    assert len(calls_s) == 2 * len(calls_f) and all(k == 16 for k in calls_s)
    assert rel(rho_f, rho_s) < 2e-5
    monkeypatch.delenv('MPDO_NO_FUSE')
    _, calls_128 = run(C128)
    assert all(k == 16 for k in calls_128)


def test_cz_pair_fusion_with_batched_angles():
    """A parameter sweep through fused pairs: the rz angle between the two CZs of an rzz differs per circuit, so the
    composite Kraus tensor carries the batch axis. Every circuit of the batch equals its own single run."""
    n, B = 3, 2
    files = {'CZ': {f'{i}{i + 1}': os.path.join(CHI_DIR, 'czDefault.mat') for i in range(n - 1)}, 'CP': {}}
    kw = dict(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=8, kappa=3, chip='best', dtype=C64,
              device='cpu')
    g = torch.Generator().manual_seed(11)
    theta = torch.rand(n - 1, B, generator=g, dtype=torch.float64) * 2 * math.pi
    u3 = (torch.rand(n, 3, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()

    def prog(c, th):
        for q in range(n):
            c.u3(*u3[q], [q])
        for q in range(n - 1):
            c.rzz(th[q], q, q + 1)
            c.truncate()

    batched = Simulator.TensorCircuit(**kw)
    prog(batched, [theta[q].clone() for q in range(n - 1)])
    batched.evolve(Simulator.Tools.create_ket0Series(n, dtype=C64))
    rho_b = batched.cal_dm()
    assert rho_b.shape == (B, 2 ** n, 2 ** n)
    for b in range(B):
        single = Simulator.TensorCircuit(**kw)
        prog(single, [float(theta[q, b]) for q in range(n - 1)])
        single.evolve(Simulator.Tools.create_ket0Series(n, dtype=C64))
        assert rel(rho_b[b], single.cal_dm()) < 2e-5


def test_segment_compilation_is_reused_and_follows_the_gates(monkeypatch):
    """truncate() compiles the segment it closes (strands, fused Kraus composites); evolve() reuses the compiled
    segment while the gates are unchanged and recompiles when a gate parameter is modified in place."""
    n = 4
    files = {'CZ': {f'{i}{i + 1}': os.path.join(CHI_DIR, 'czDefault.mat') for i in range(n - 1)}, 'CP': {}}

    def build():
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=8, kappa=4,
                                    chip='best', dtype=C128, device='cpu')
        for q in range(n):
            c.u3(0.3 + q, 0.2, 0.1, [q])
        c.rzz(0.7, 0, 1)
        c.rzz(0.4, 2, 3)
        c.truncate()
        c.rzz(0.9, 1, 2)
        c.truncate()
        return c

    calls = []
    orig = Simulator.TensorCircuit._compile_segment

    def counting(self, ops):
        calls.append(len(ops))
        return orig(self, ops)

    monkeypatch.setattr(Simulator.TensorCircuit, '_compile_segment', counting)
    c = build()
    assert len(calls) == 2 and len(c._compiled) == 2           # compiled at construction
    s1 = Simulator.Tools.create_ket0Series(n, dtype=C128, device='cpu')
    c.evolve(s1)
    assert len(calls) == 2                                      # nothing rebuilt inside evolve
    z1 = c.cal_dm().clone()
    s2 = Simulator.Tools.create_ket0Series(n, dtype=C128, device='cpu')
    c.evolve(s2)
    assert len(calls) == 2 and torch.equal(c.cal_dm(), z1)      # a second evolve is the same computation

    monkeypatch.setenv('MPDO_LAZY_COMPILE', '1')
    lazy = build()
    assert len(calls) == 2 and not lazy._compiled
    s3 = Simulator.Tools.create_ket0Series(n, dtype=C128, device='cpu')
    lazy.evolve(s3)
    assert len(calls) == 4 and torch.equal(lazy.cal_dm(), z1)   # compiled on first use, same numbers
    monkeypatch.delenv('MPDO_LAZY_COMPILE')

    gate = next(g for g in c.layers if getattr(g, 'name', '') == 'U3')
    gate.theta.add_(0.5)                                        # in-place change of a gate parameter
    s4 = Simulator.Tools.create_ket0Series(n, dtype=C128, device='cpu')
    c.evolve(s4)
    assert len(calls) == 5                                      # only the segment holding that gate
    assert not torch.allclose(c.cal_dm(), z1)
