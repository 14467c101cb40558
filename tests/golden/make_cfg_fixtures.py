#!/usr/bin/env python
"""Oracle outputs for BASELINE config 1 (10-qubit GHZ prefix + depth-10 noisy brickwork, chi = 32, kappa = 4) and a
depth-3 slice of config 2, as fixtures for the -m gpu parity tests (the oracle needs minutes on these; the GPU
box only reads the .npz). Gauge-invariant outputs only: Tr rho, <Z_q>, <Z_q Z_q+1>, P(0...0), a two-site RDM.

    python tests/golden/make_cfg_fixtures.py
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tomography-assisted-mpdo-qcircuit_b200')]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402
from oracle.mpdo_oracle import OracleCircuit  # noqa: E402

Z = torch.tensor([[1, 0], [0, -1]], dtype=torch.complex128)
out = {}
# `--only exact_fast` adds that mode to an existing config_fixtures.npz without re-running the slow plain-formulation runs
ONLY = sys.argv[sys.argv.index('--only') + 1].split(',') if '--only' in sys.argv else None
if ONLY and os.path.exists(os.path.join(HERE, 'config_fixtures.npz')):
    out.update(dict(np.load(os.path.join(HERE, 'config_fixtures.npz'))))


def record(tag, oc, n):
    out[f'{tag}/trace'] = np.array(oc.trace().item())
    out[f'{tag}/z'] = np.array([oc.chain({q: Z}).real.item() for q in range(n)])
    out[f'{tag}/zz'] = np.array([oc.chain({q: Z, q + 1: Z}).real.item() for q in range(n - 1)])
    out[f'{tag}/p0'] = np.array(oc.chain(proj=[0] * n).real.item())
    out[f'{tag}/rdm_mid'] = oc.rdm([n // 2, n // 2 + 1]).to(torch.complex128).numpy()
    out[f'{tag}/split_ranks'] = np.array(oc.stats['split_ranks'], dtype=np.int64)   # to tell rank-rule flips apart


def _kw(mode):
    """'exact_fast' = exact mode in the oracle's small-side formulation (oracle/mpdo_oracle.py fast=True)."""
    return dict(svd_mode='exact', fast=True) if mode == 'exact_fast' else dict(svd_mode=mode)


def cfg1(dtype, mode):
    n, depth = 10, 10
    oc = OracleCircuit(n, ideal=False, noiseType='idealNoise', chi=32, kappa=4, chip='medium', dtype=dtype, **_kw(mode))
    bc.brickwork(oc, n, depth, bc.angles([0], bc.n_draws(n, depth, 'cz')), 'cz', prefix_ghz=True)
    t0 = time.perf_counter()
    oc.evolve()
    return oc, time.perf_counter() - t0


def cfg1_tiefree(dtype, mode):
    """cfg1 with noiseless random rotations in front: the GHZ symmetry makes the 15 equally weighted depolarizing
    Kraus branches exactly degenerate at the kappa = 4 cut (the reference's result is then arbitrary at the 1e-2
    level, see cfg1 exact vs reference below); the rotations lift the degeneracy without changing the workload."""
    n, depth = 10, 10
    oc = OracleCircuit(n, ideal=False, noiseType='idealNoise', chi=32, kappa=4, chip='medium', dtype=dtype, **_kw(mode))
    pre = bc.angles([7], 3 * n)
    for q in range(n):
        oc.u3(float(pre[3 * q, 0]), float(pre[3 * q + 1, 0]), float(pre[3 * q + 2, 0]), [q], True)
    bc.brickwork(oc, n, depth, bc.angles([0], bc.n_draws(n, depth, 'cz')), 'cz', prefix_ghz=True)
    t0 = time.perf_counter()
    oc.evolve()
    return oc, time.perf_counter() - t0


def cfg2_slice(dtype, mode, n=6, depth=3):
    files = {'CZ': {f'{i}{i + 1}': bc.chi_file() for i in range(n - 1)}, 'CP': {}}
    oc = OracleCircuit(n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=64, kappa=4, chip='best',
                       dtype=dtype, **_kw(mode))
    bc.brickwork(oc, n, depth, bc.angles([0], bc.n_draws(n, depth, 'rzz')), 'rzz', trunc_after_1q=False)
    t0 = time.perf_counter()
    oc.evolve()
    return oc, time.perf_counter() - t0


for name, fn, n in (('cfg1', cfg1, 10), ('cfg1_tiefree', cfg1_tiefree, 10), ('cfg2_n6_d3', cfg2_slice, 6)):
    for dtype, dt in ((torch.complex128, 'c128'), (torch.complex64, 'c64')):
        for mode in ('exact', 'exact_fast', 'reference'):
            if name == 'cfg1_tiefree' and mode == 'reference' and dt == 'c64':
                continue
            if mode == 'exact_fast' and dt == 'c64':
                continue
            if ONLY and mode not in ONLY:
                continue
            oc, secs = fn(dtype, mode)
            record(f'{name}/{dt}/{mode}', oc, n)
            out[f'{name}/{dt}/{mode}/seconds'] = np.array(secs)
            print(name, dt, mode, 'trace %.8f' % oc.trace().item(), '%.1f s' % secs, flush=True)
np.savez_compressed(os.path.join(HERE, 'config_fixtures.npz'), **out)
