"""CPU tier: the environment form of bondTruncate (steps.Engine.bond_truncate_env: one chain of contractions left to
right, independent factorisations of the bond environments, one sequential right-to-left sweep) against the two-sweep
form (QR sweep + chi sweep, TNNOptimizer.py:72-134) and against the oracle, through the torch-CPU model of the device
primitives. Gauge-invariant comparison: the dense density matrix and the singular values discarded at every bond."""
import pytest
import torch

from cpu_prims import CpuPrims
from harness import brickwork, rel, run_engine
from oracle.mpdo_oracle import OracleCircuit

C64, C128 = torch.complex64, torch.complex128


def _circuit(n, depth, chi, kappa, dtype, seed):
    oc = OracleCircuit(n, ideal=False, noiseType='idealNoise', chip='worst', chi=chi, kappa=kappa, dtype=dtype, fast=True)
    brickwork(oc, n, depth, seed=seed, ghz_prefix=False)
    return oc


@pytest.mark.parametrize('n,depth,chi,kappa', [(5, 3, 4, 2), (6, 3, 6, 2)])
def test_env_sweep_equals_two_sweeps_complex128(n, depth, chi, kappa):
    oc = _circuit(n, depth, chi, kappa, C128, seed=5)
    E1, T1 = run_engine(oc, CpuPrims(), C128, npass=1)
    E2, T2 = run_engine(oc, CpuPrims(), C128, npass=1, env_sweep=True)
    assert [tuple(t.shape) for t in T1] == [tuple(t.shape) for t in T2]
    r1, r2 = E1.dense_rho(T1)[0], E2.dense_rho(T2)[0]
    assert rel(r2, r1) < 1e-8, rel(r2, r1)          # one-pass Gram routes: half of fp64 on the small directions
    oc.evolve()
    assert rel(r2, oc.cal_dm()) < 1e-7, rel(r2, oc.cal_dm())


def test_env_sweep_singular_values_match_two_sweeps():
    """The values each bond discards are the same numbers in both forms (the factor of the environment differs from
    the R of the QR sweep by a unitary only)."""
    oc = _circuit(6, 3, 4, 2, C128, seed=11)
    prims = CpuPrims()
    from MPDOSimulator._engine.steps import Engine
    E, Ts = run_engine(oc, prims, C128, npass=1)
    # one more layer of entanglers so the bonds are wide again, then truncate both ways from the same state
    g = torch.Generator().manual_seed(3)
    cz = torch.diag(torch.tensor([1, 1, 1, -1], dtype=C128)).reshape(1, 2, 2, 2, 2, 1)
    for q in range(0, 5, 2):
        Ts[q], Ts[q + 1] = E.split_2q(Ts[q], Ts[q + 1], cz)
    A, B = [t.clone() for t in Ts], [t.clone() for t in Ts]
    E.qr_left2right(A)
    d1 = E.svd_right2left(A, 3)
    d2 = E.bond_truncate_env(B, 3)
    for x, y in zip(d1, d2):
        assert x.shape == y.shape
        assert x.numel() == 0 or (x - y).abs().max() <= 1e-9 * max(1.0, float(x.abs().max()))
    assert rel(E.dense_rho(B)[0], E.dense_rho(A)[0]) < 1e-8


def test_env_sweep_complex64_inside_fp32_floor():
    oc = _circuit(5, 3, 4, 2, C64, seed=7)
    E1, T1 = run_engine(oc, CpuPrims(), C64)
    E2, T2 = run_engine(oc, CpuPrims(), C64, env_sweep=True)
    r1, r2 = E1.dense_rho(T1)[0], E2.dense_rho(T2)[0]
    exact = _circuit(5, 3, 4, 2, C128, seed=7)
    exact.evolve()
    e1, e2 = rel(r1.to(C128), exact.cal_dm()), rel(r2.to(C128), exact.cal_dm())
    assert e2 < max(2e-5, 3 * e1), (e1, e2)


def test_env_sweep_hands_sites_over_one_by_one():
    """`publish` receives every site exactly once, right to left with site 0 last, and what it receives is what the
    list holds afterwards (the caller drops the tensor a site replaces while the sweep goes on: large states)."""
    oc = _circuit(6, 3, 4, 2, C128, seed=13)
    from MPDOSimulator._engine.steps import Engine
    E, Ts = run_engine(oc, CpuPrims(), C128, npass=1)
    cz = torch.diag(torch.tensor([1, 1, 1, -1], dtype=C128)).reshape(1, 2, 2, 2, 2, 1)
    for q in range(0, 5, 2):
        Ts[q], Ts[q + 1] = E.split_2q(Ts[q], Ts[q + 1], cz)
    A, B = [t.clone() for t in Ts], [t.clone() for t in Ts]
    seen = []
    d1 = E.bond_truncate_env(A, 3)
    d2 = E.bond_truncate_env(B, 3, publish=lambda idx, t: seen.append((idx, t)))
    assert [i for i, _ in seen] == [5, 4, 3, 2, 1, 0]
    for idx, t in seen:
        assert t is B[idx]
    for x, y in zip(A, B):
        assert torch.equal(x, y)
    for x, y in zip(d1, d2):
        assert torch.equal(x, y)


def test_memory_guard_is_a_no_op_off_device_and_on_small_states():
    from MPDOSimulator._engine import strands
    small = [torch.zeros(4, dtype=C64)]
    assert strands.state_bytes(small) == 32
    strands.memory_guard('cpu', small)           # host tensors: returns without touching CUDA
    strands.memory_guard('cuda:0', [])           # empty state
    assert strands.BIG_STATE_BYTES == 1 << 30
