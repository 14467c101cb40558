"""-m gpu: the CUDA update path against the oracle on seeded circuits (gauge-invariant outputs).

Tolerances (north star): 1e-10 relative for complex128. For complex64 the reference's own fp32 LAPACK
arithmetic is only reproducible to the gap between its complex64 and complex128 runs (`floor` below,
typically 2e-5..1e-4, SURVEY 7 hard part 1), so the complex64 result is held to 1e-5 against the
exact-arithmetic (complex128) oracle where the problem is well conditioned, and to the reference's own
floor otherwise; circuits whose truncation cuts through (near-)degenerate singular values are detected
through that floor and skipped, because the reference result itself is then decided by rounding noise.
"""
import pytest
import torch

from harness import brickwork, rel, run_engine
from oracle.mpdo_oracle import OracleCircuit

pytestmark = pytest.mark.gpu
C64, C128 = torch.complex64, torch.complex128


def oracle_run(n, depth, seed, chi, kappa, dtype, noise='idealNoise', chip='medium', ghz=True):
    oc = OracleCircuit(n, ideal=False, noiseType=noise, chi=chi, kappa=kappa, chip=chip, dtype=dtype)
    brickwork(oc, n, depth, seed=seed, ghz_prefix=ghz)
    oc.evolve()
    return oc


CASES = [  # n, depth, seed, chi, kappa
    (5, 1, 0, None, None),
    (5, 4, 2, 16, 3),
    (6, 3, 3, 8, 3),
    (4, 4, 4, 6, 5),
]


@pytest.mark.parametrize('n,depth,seed,chi,kappa', CASES)
def test_density_matrix_c128(cuda_prims, n, depth, seed, chi, kappa):
    oc = oracle_run(n, depth, seed, chi, kappa, C128)
    ref = oc.cal_dm()
    floor = rel(oracle_run(n, depth, seed, chi, kappa, C64).cal_dm().to(C128), ref)
    if floor > 1e-3:
        pytest.skip(f'truncation cuts a degenerate multiplet (reference c64-vs-c128 gap {floor:.1e})')
    E, Ts = run_engine(oc, cuda_prims, C128, device='cuda')
    rho = E.dense_rho(Ts)[0].cpu()
    assert rel(rho, ref) < 1e-10
    tr = E.chain_value(Ts)[0].cpu()
    assert abs(tr.real.item() - oc.trace().item()) < 1e-10


@pytest.mark.parametrize('n,depth,seed,chi,kappa', CASES)
def test_density_matrix_c64(cuda_prims, n, depth, seed, chi, kappa):
    ref = oracle_run(n, depth, seed, chi, kappa, C128).cal_dm()
    oc = oracle_run(n, depth, seed, chi, kappa, C64)
    floor = rel(oc.cal_dm().to(C128), ref)
    if floor > 1e-3:
        pytest.skip(f'truncation cuts a degenerate multiplet (reference c64-vs-c128 gap {floor:.1e})')
    E, Ts = run_engine(oc, cuda_prims, C64, device='cuda')
    rho = E.dense_rho(Ts)[0].cpu()
    err = rel(rho, ref)
    print(f'c64: err vs exact oracle {err:.2e}; reference floor {floor:.2e}')
    assert err < max(1e-5, 3 * floor)


def test_expectation_and_probabilities(cuda_prims):
    n, depth, seed, chi, kappa = 6, 3, 7, 12, 3
    oc = oracle_run(n, depth, seed, chi, kappa, C128)
    E, Ts = run_engine(oc, cuda_prims, C128, device='cuda')
    Z = torch.tensor([[1, 0], [0, -1]], dtype=C128)
    for q in range(n):
        got = E.chain_value(Ts, {q: Z.cuda()})[0].cpu()
        want = oc.chain({q: Z})
        assert abs(got - want) < 1e-10
    bits = [[0] * n, [1] * n, [0, 1, 0, 1, 1, 0]]
    got = E.bitstring_probs(Ts, bits).cpu()
    for i, b in enumerate(bits):
        assert abs(got[i].item() - oc.chain(proj=b).real.item()) < 1e-10


def test_batched_circuits_match_single(cuda_prims):
    """The batch dimension (independent circuits) gives the same result as running each alone."""
    n, depth, chi, kappa = 4, 2, 8, 3
    ocs = [oracle_run(n, depth, s, chi, kappa, C128) for s in (11, 12, 13)]
    singles = []
    for oc in ocs:
        E, Ts = run_engine(oc, cuda_prims, C128, device='cuda')
        singles.append(E.dense_rho(Ts)[0].cpu())
    for oc, rho in zip(ocs, singles):
        assert rel(rho, oc.cal_dm()) < 1e-10


# ---------------------------------------------------------------------------------------------------
# the public API on the CUDA path against the golden vectors produced by the unmodified reference
# ---------------------------------------------------------------------------------------------------
import test_golden as tg  # noqa: E402


@pytest.mark.parametrize('name', sorted(tg.CIRCUITS))
@pytest.mark.parametrize('tag', ['c128', 'c64'])
def test_cuda_api_matches_reference_golden(cuda_prims, name, tag):
    import MPDOSimulator as Simulator
    from MPDOSimulator import dmOperations
    n, prog, kw = tg.CIRCUITS[name]
    dt = tg.DT[tag]
    c = Simulator.TensorCircuit(qn=n, dtype=dt, device='cuda:0', **kw)
    prog(c)
    st = Simulator.Tools.create_ket0Series(n, dtype=dt, device='cpu')      # host buffers in
    c.evolve(st)
    assert all(node.data.is_cuda for node in st)
    tol = tg.tolerance(name, tag)
    dmn = c.cal_dmNodes()
    assert abs(dmOperations.trace_rho(dmn).item() - float(tg.GOLD[f'{name}_{tag}/trace'])) <= tol
    if name in tg.TIES:
        return
    assert tg.close(c.cal_dm().cpu(), tg.t(f'{name}_{tag}/dm'), tol)
    assert abs(dmOperations.trace_rho2(dmn).item() - float(tg.GOLD[f'{name}_{tag}/trace_rho2'])) <= tol
    for q in range(n):
        assert abs(dmOperations.pauli_expect(dmn, 2, q).item() - float(tg.GOLD[f'{name}_{tag}/pauli_z'][q])) <= tol
