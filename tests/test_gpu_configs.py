"""-m gpu: BASELINE configurations against oracle fixtures (tests/golden/config_fixtures.npz, written by
tests/golden/make_cfg_fixtures.py with the oracle in exact complex128 / complex64 and in reference mode).

cfg1 is run at full size (10 qubits, GHZ prefix, depth 10, chi 32, kappa 4); cfg2 as a 6-qubit depth-3 slice with
the same chi-matrix channel, chi 64, kappa 4 (the full 20-qubit circuit takes the oracle tens of minutes).
Quantities are gauge invariant: Tr rho, <Z_q>, <Z_q Z_q+1>, P(0...0), the two-site RDM in the middle.
complex128 is held to the exact oracle; complex64 to max(1e-5, 3 x the oracle's own complex64-vs-complex128 gap)."""
import os

import numpy as np
import pytest
import torch

import bench_configs as bc
import MPDOSimulator as Simulator
from MPDOSimulator import dmOperations

pytestmark = pytest.mark.gpu
FX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'config_fixtures.npz'))
C64, C128 = torch.complex64, torch.complex128


def gpu_quantities(circ, n):
    dmn = circ.cal_dmNodes()
    eng, Ts = circ._engine(), circ._Ts()
    return {
        'trace': np.array(dmOperations.trace_rho(dmn).item()),
        'z': np.array([dmOperations.pauli_expect(dmn, 2, q).item() for q in range(n)]),
        'zz': np.array([dmOperations.pauli_expect(dmn, [2, 2], [q, q + 1]).item() for q in range(n - 1)]),
        'p0': np.array(circ.bitstring_probabilities(['0' * n])[0].item()),
        'rdm_mid': eng.dense_rho(Ts, keep=[n // 2, n // 2 + 1])[0].cpu().numpy(),
    }


def split_ranks(circ):
    ranks = circ.last_stats.get('split_ranks', {})
    return [ranks[k] for k in sorted(ranks)]


def compare(name, got, dt, tie_floor=False, ranks=None):
    """complex128: 1e-10 when the kept rank of every gate split equals the oracle's; a run in which a singular value
    sits within rounding of the reference's absolute rank-rule threshold (e*1e-8) and the kept rank flips is held to
    1e-6 instead (the rule then discards up to 5e-8 of weight more or less) and says so. complex64: max(1e-5, 3 x the
    oracle's own complex64-vs-complex128 gap).
    tie_floor: the circuit has exactly degenerate singular values at its cuts (GHZ symmetry x equally weighted
    depolarizing branches): which 3 of 15 equal branches survive kappa = 4 is decided by rounding noise, the
    reference's own answers spread at the 1e-2 level (precision, SVD branch), and only a 5e-2 sanity bound is
    meaningful; the symmetry-broken variant of the same workload carries the real parity claim."""
    worst = 0.0
    flips = None
    if ranks is not None and f'{name}/c128/exact/split_ranks' in FX:
        want = FX[f'{name}/c128/exact/split_ranks'].tolist()
        flips = sum(int(a != b) for a, b in zip(ranks, want)) if len(ranks) == len(want) else -1
        print(f'{name} {dt}: {flips} rank flips in {len(want)} gate splits')
    for key, val in got.items():
        exact = FX[f'{name}/c128/exact/{key}']
        scale = np.abs(exact).max()
        err = np.abs(val - exact).max() / scale
        gap64 = np.abs(FX[f'{name}/c64/exact/{key}'] - exact).max() / scale
        if dt == 'c128':
            # Two LAPACK formulations of the reference's own algorithm - the plain one (full two-site matrices) and
            # the small-side one (oracle fast=True) - agree only to `form_gap` on these circuits (3e-10 ... 2e-7: the
            # plain formulation decomposes ill-conditioned 72 x 12288-like matrices whose kept singular values span
            # four decades). The CUDA path is held to 1e-10 against the small-side formulation, whose arithmetic it
            # shares, and to the formulation gap against the plain one.
            fast = FX[f'{name}/c128/exact_fast/{key}']
            form_gap = np.abs(fast - exact).max() / scale
            err_fast = np.abs(val - fast).max() / np.abs(fast).max()
            tol = max(1e-10 if flips == 0 else 1e-6, 3 * form_gap)
            if not tie_floor:
                print(f'{name} {dt} {key}: vs small-side formulation {err_fast:.2e} (tol '
                      f'{1e-10 if flips == 0 else 1e-6:.0e}), the two oracle formulations differ by {form_gap:.2e}')
                assert err_fast <= (1e-10 if flips == 0 else 1e-6), (name, dt, key, err_fast)
        else:
            tol = max(1e-5, 3 * gap64)
        if tie_floor:
            spread = max(gap64, np.abs(FX[f'{name}/c128/reference/{key}'] - exact).max() / scale)
            tol = max(tol, 3 * spread, 5e-2)
        print(f'{name} {dt} {key}: rel err {err:.2e} (tol {tol:.1e})')
        assert err <= tol, (name, dt, key, err, tol)
        worst = max(worst, err)
    return worst


def run_cfg1(dtype, tiefree):
    n, depth = 10, 10
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=32, kappa=4, chip='medium',
                                dtype=dtype, device='cuda:0')
    if tiefree:
        pre = bc.angles([7], 3 * n)
        for q in range(n):
            c.u3(float(pre[3 * q, 0]), float(pre[3 * q + 1, 0]), float(pre[3 * q + 2, 0]), [q], True)
    bc.brickwork(c, n, depth, bc.angles([0], bc.n_draws(n, depth, 'cz')), 'cz', prefix_ghz=True)
    st = Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu')
    c.evolve(st)
    return c, n


@pytest.mark.parametrize('dt', ['c128', 'c64'])
def test_cfg1_symmetry_broken(cuda_prims, dt):
    c, n = run_cfg1(C128 if dt == 'c128' else C64, tiefree=True)
    compare('cfg1_tiefree', gpu_quantities(c, n), dt, ranks=split_ranks(c))


@pytest.mark.parametrize('dt', ['c128', 'c64'])
def test_cfg1_full_size(cuda_prims, dt):
    c, n = run_cfg1(C128 if dt == 'c128' else C64, tiefree=False)
    compare('cfg1', gpu_quantities(c, n), dt, tie_floor=True)


@pytest.mark.parametrize('dt', ['c128', 'c64'])
def test_cfg2_slice(cuda_prims, dt):
    n, depth = 6, 3
    dtype = C128 if dt == 'c128' else C64
    files = {'CZ': {f'{i}{i + 1}': bc.chi_file() for i in range(n - 1)}, 'CP': {}}
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=64, kappa=4,
                                chip='best', dtype=dtype, device='cuda:0')
    bc.brickwork(c, n, depth, bc.angles([0], bc.n_draws(n, depth, 'rzz')), 'rzz', trunc_after_1q=False)
    st = Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu')
    c.evolve(st)
    compare('cfg2_n6_d3', gpu_quantities(c, n), dt, ranks=None if dt == 'c64' else split_ranks(c))


def test_batched_sweep_is_deterministic_and_matches_singles(cuda_prims):
    """cfg4 in small: 6 circuits x 8 qubits as one batch, twice (strand concurrency must not change results), and
    against the same circuits run one by one."""
    n, depth, ids = 8, 4, list(range(6))

    def run(id_list):
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=16, kappa=4, chip='medium',
                                    dtype=C128, device='cuda:0')
        bc.brickwork(c, n, depth, bc.angles(id_list, bc.n_draws(n, depth, 'cz')), 'cz')
        st = Simulator.Tools.create_ket0Series(n, dtype=C128, device='cpu')
        c.evolve(st)
        dmn = c.cal_dmNodes()
        return torch.stack([dmOperations.pauli_expect(dmn, 2, q).reshape(-1) for q in range(n)], 1).cpu()

    a, b = run(ids), run(ids)
    # (the kappa step remembers per shape whether its subspace iteration stalled and may take the full decomposition
    # instead on a later call: both are exact solvers, so repeat runs agree to solver tolerance, not bit for bit)
    assert torch.equal(a, b) or (a - b).abs().max().item() < 1e-11
    for i in ids:
        single = run([i])
        assert (a[i] - single[0]).abs().max().item() < 1e-10


def test_cfg2_full_width_fused_pairs_match_one_split_per_gate(cuda_prims, monkeypatch):
    """cfg2 at its full width and truncation parameters (20 qubits, chi 64, kappa 4, czDefault chi-matrix channel on
    every bond), first nine layers (the middle bonds reach chi): the fused CZ pairs of the complex64 path and one split
    per gate (MPDO_NO_FUSE=1), both measured against the complex128 evolution of the same circuit (never fused, fp64
    throughout) on gauge-invariant outputs. This is a sanity check at full width, not the parity claim: there is no
    oracle at 20 qubits, and at this size the chi / kappa cuts run through clusters of nearly equal singular values,
    so fp32-level perturbations (including the summation order of the split-K Gram matrices, which differs from run to
    run) are amplified chaotically - the gate-by-gate path moves between 6e-5 and 6e-4 from run to run. Both paths
    must stay at that fp32 floor (1.5e-3); the claim against the exact oracle is
    tests/test_gpu_big_configs.py::test_cfg2_width10_chi64_saturated."""
    n, depth = 20, 9
    files = {'CZ': {f'{i}{i + 1}': bc.chi_file() for i in range(n - 1)}, 'CP': {}}

    def run(dtype):
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=64, kappa=4,
                                    chip='best', dtype=dtype, device='cuda:0')
        bc.brickwork(c, n, depth, bc.angles([0], bc.n_draws(n, depth, 'rzz')), 'rzz', trunc_after_1q=False)
        c.evolve(Simulator.Tools.create_ket0Series(n, dtype=dtype, device='cpu'))
        dmn = c.cal_dmNodes()
        tr = dmOperations.trace_rho(dmn).item()
        z = np.array([dmOperations.pauli_expect(dmn, 2, q).item() for q in range(n)])
        zz = np.array([dmOperations.pauli_expect(dmn, [2, 2], [q, q + 1]).item() for q in range(0, n - 1, 3)])
        bonds = [int(s.data.shape[4]) for s in c.stateNodes[:-1]]
        return np.concatenate([[tr], z, zz]), bonds

    fused, bonds_f = run(C64)
    monkeypatch.setenv('MPDO_NO_FUSE', '1')
    split, bonds_s = run(C64)
    exact, _ = run(C128)
    assert max(bonds_f) <= 64 and max(bonds_s) <= 64
    assert 0 < exact[0] < 1                              # the tomography channel is not trace preserving
    scale = abs(exact[0])
    err_fused = np.abs(fused - exact).max() / scale
    err_split = np.abs(split - exact).max() / scale
    print(f'bonds {bonds_f}; vs complex128: fused {err_fused:.2e}, one split per gate {err_split:.2e}, '
          f'fused vs split {np.abs(fused - split).max() / scale:.2e}')
    assert err_fused <= 1.5e-3 and err_split <= 1.5e-3
