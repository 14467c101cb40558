#!/usr/bin/env python
"""Oracle fixtures for reduced-width versions of BASELINE configs 2-5 in which the truncation parameters of the full
configuration are ACTIVE (bonds saturate at chi = 64 / 128 / 256, inner indices at kappa = 4 / 8), written with the
oracle in exact mode, fast formulation (oracle/mpdo_oracle.py: _apply_2q_fast, _svd_right2left_fast - LAPACK on the
small side only; tests/test_oracle_fast.py pins it to the plain restatement). The GPU box only reads the .npz.

    python tests/golden/make_big_fixtures.py --case cfg3           # one case -> tests/golden/big_<case>.npz
    python tests/golden/make_big_fixtures.py --merge               # big_*.npz -> big_fixtures.npz

Recorded per case and dtype (gauge-invariant outputs only): Tr rho, <Z_q>, <Z_q Z_q+1>, P(0...0), the two-site RDM in
the middle, bitstring probabilities (cfg5), the kept rank of every gate split (to tell a rank-rule flip from an
arithmetic difference), the final bond dimensions and the oracle's wall time.
"""
import argparse
import glob
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
C64, C128 = torch.complex64, torch.complex128

# name -> parameters of the reduced configuration (SURVEY 8d definitions with fewer qubits / layers)
CASES = {
    'cfg2': dict(n=10, depth=8, chi=64, kappa=4, noise='realNoise', chip='best', ent='rzz', trunc_after_1q=False,
                 ids=[0], dtypes=('c128', 'c64')),
    'cfg3': dict(n=10, depth=12, chi=128, kappa=8, noise='idealNoise', chip='medium', ent='cz', trunc_after_1q=True,
                 ids=[0], dtypes=('c128',)),
    'cfg4': dict(n=16, depth=16, chi=64, kappa=4, noise='idealNoise', chip='medium', ent='cz', trunc_after_1q=True,
                 ids=list(range(8)), dtypes=('c128', 'c64')),
    'cfg5': dict(n=12, depth=12, chi=256, kappa=8, noise='idealNoise', chip='medium', ent='cz', trunc_after_1q=True,
                 ids=[0], dtypes=('c128', 'c64'), bitstrings=64),
}


def bitstrings(n, count):
    g = torch.Generator().manual_seed(99)
    return torch.randint(0, 2, (count, n), generator=g).tolist()


def build(case, cid, dtype, cls=OracleCircuit, **kw):
    p = CASES[case]
    n = p['n']
    files = None
    if p['noise'] == 'realNoise':
        files = {'CZ': {f'{i}{i + 1}': bc.chi_file() for i in range(n - 1)}, 'CP': {}}
    c = cls(n, ideal=False, noiseType=p['noise'], chiFileDict=files, chi=p['chi'], kappa=p['kappa'], chip=p['chip'],
            dtype=dtype, **kw)
    bc.brickwork(c, n, p['depth'], bc.angles([cid], bc.n_draws(n, p['depth'], p['ent'])), p['ent'],
                 trunc_after_1q=p['trunc_after_1q'])
    return c


def record(out, tag, oc, n, nbits=0):
    out[f'{tag}/trace'] = np.array(oc.trace().item())
    out[f'{tag}/z'] = np.array([oc.chain({q: Z}).real.item() for q in range(n)])
    out[f'{tag}/zz'] = np.array([oc.chain({q: Z, q + 1: Z}).real.item() for q in range(n - 1)])
    out[f'{tag}/p0'] = np.array(oc.chain(proj=[0] * n).real.item())
    out[f'{tag}/rdm_mid'] = oc.rdm([n // 2, n // 2 + 1]).to(C128).numpy()
    out[f'{tag}/split_ranks'] = np.array(oc.stats['split_ranks'], dtype=np.int64)
    out[f'{tag}/bonds'] = np.array([int(T.shape[3]) for T in oc.T[:-1]], dtype=np.int64)
    if nbits:
        out[f'{tag}/bitprobs'] = np.array([oc.chain(proj=b).real.item() for b in bitstrings(n, nbits)])


def run_case(case, depth=None, qubits=None):
    p = CASES[case]
    if depth:
        p['depth'] = depth
    if qubits:
        p['n'] = qubits
    out = {f'{case}/params': np.array([p['n'], p['depth'], p['chi'], p['kappa']], dtype=np.int64)}
    for dt in p['dtypes']:
        dtype = C128 if dt == 'c128' else C64
        ids = p['ids']                                          # complex64 runs measure the fp32 floor
        for cid in ids:
            oc = build(case, cid, dtype, svd_mode='exact', fast=True)
            t0 = time.perf_counter()
            oc.evolve()
            secs = time.perf_counter() - t0
            tag = f'{case}/{cid}/{dt}'
            record(out, tag, oc, p['n'], p.get('bitstrings', 0))
            out[f'{tag}/seconds'] = np.array(secs)
            out[f'{tag}/updates'] = np.array(oc.stats['updates_2q_noisy'])
            print(f'{tag}: trace {out[tag + "/trace"]:.10f} bonds {out[tag + "/bonds"].tolist()} '
                  f'max rank {max(oc.stats["split_ranks"])} {secs:.1f} s', flush=True)
    np.savez_compressed(os.path.join(HERE, f'big_{case}.npz'), **out)


def merge():
    out = {}
    for f in sorted(glob.glob(os.path.join(HERE, 'big_cfg*.npz'))):
        out.update(dict(np.load(f)))
    np.savez_compressed(os.path.join(HERE, 'big_fixtures.npz'), **out)
    print(f'{len(out)} arrays -> big_fixtures.npz')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default=None)
    ap.add_argument('--depth', type=int, default=None)
    ap.add_argument('--qubits', type=int, default=None)
    ap.add_argument('--merge', action='store_true')
    a = ap.parse_args()
    if a.case:
        run_case(a.case, a.depth, a.qubits)
    if a.merge:
        merge()
