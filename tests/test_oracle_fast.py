"""CPU tier: the oracle's fast exact formulation (LAPACK on the small side only: _apply_2q_fast,
_svd_right2left_fast) against its plain restatement of the reference (full two-site matrices), complex128, on
gauge-invariant outputs and on the kept rank of every gate split. The fast formulation is what writes the
chi = 64 / 128 / 256 fixtures (tests/golden/make_big_fixtures.py), so it must be the same function."""
import numpy as np
import pytest
import torch

import bench_configs as bc
from oracle.mpdo_oracle import OracleCircuit

Z = torch.tensor([[1, 0], [0, -1]], dtype=torch.complex128)


def outputs(oc, n):
    vals = [oc.trace().item()] + [oc.chain({k: Z}).real.item() for k in range(n)]
    vals += [oc.chain({k: Z, k + 1: Z}).real.item() for k in range(n - 1)] + [oc.chain(proj=[0] * n).real.item()]
    return np.array(vals), oc.rdm([n // 2 - 1, n // 2]).numpy()


@pytest.mark.parametrize('n,depth,chi,kappa,noise,ent,err', [
    (5, 3, 16, 4, 'realNoise', 'rzz', None),
    (4, 4, 6, 3, 'idealNoise', 'cz', None),
    (5, 3, 8, 2, 'idealNoise', 'cnot', None),
    (4, 3, None, 3, 'idealNoise', 'cz', 1e-3),
])
def test_fast_formulation_equals_plain_restatement(n, depth, chi, kappa, noise, ent, err):
    res = []
    for fast in (False, True):
        files = {'CZ': {f'{i}{i + 1}': bc.chi_file() for i in range(n - 1)}, 'CP': {}}
        oc = OracleCircuit(n, ideal=False, noiseType=noise, chiFileDict=files if noise == 'realNoise' else None,
                           chi=chi, kappa=kappa, max_truncation_err=err, chip='medium', dtype=torch.complex128,
                           fast=fast)
        # idealNoise cases start with a chain of noiseless CZs on |0...0> (state unchanged, every bond exists), so that
        # truncation is active from the first layer; otherwise the plain restatement decomposes matrices with
        # un-truncated inner indices and needs minutes. (A GHZ prefix would do the same but makes the kappa cut run
        # through exactly degenerate singular values, where the kept subspace is arbitrary.)
        if noise != 'realNoise':
            for q in range(n - 1):
                oc.cz(q, q + 1, True)
        bc.brickwork(oc, n, depth, bc.angles([3], bc.n_draws(n, depth, ent)), ent,
                     trunc_after_1q=(noise != 'realNoise'))
        oc.evolve()
        res.append((outputs(oc, n), list(oc.stats['split_ranks']), [tuple(T.shape) for T in oc.T]))
    (a, ra), ranks_a, shapes_a = res[0]
    (b, rb), ranks_b, shapes_b = res[1]
    assert ranks_a == ranks_b
    # bond dimensions may differ: the plain two-site SVD pads a bond of true rank < chi back up to chi with zero
    # singular values (Theta has more rows than X); the state is the same
    assert all(x[1:3] == y[1:3] for x, y in zip(shapes_a, shapes_b))
    # Two mathematically identical LAPACK formulations agree to 1e-14 when no truncation cuts near a cluster of
    # singular values and to a few 1e-10 when one does (measured 4.5e-10 on the 5-qubit cnot case: chi = 8 and
    # kappa = 2 cut through nearly equal values, and rounding differences are amplified by 1 / gap): that is the
    # reproducibility floor of the reference's own complex128 arithmetic, not a property of either formulation.
    err = max(np.abs(a - b).max() / np.abs(a).max(), np.abs(ra - rb).max() / np.abs(ra).max())
    print(f'plain vs fast formulation: {err:.2e}')
    assert err < 1e-8
