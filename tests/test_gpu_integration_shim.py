"""-m gpu: INTEGRATION.md route 2 executed - integration/decompositions_b200.py (the reference's S1 seam bound to the
C ABI) against the oracle's restatement of decompositions.py on its full-SVD branch: singular values, kept rank under
every combination of cap / absolute / relative error rule, the truncated reconstruction, and the QR contract."""
import os
import sys

import pytest
import torch

from oracle import mpdo_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C64, C128 = torch.complex64, torch.complex128


@pytest.fixture(scope='module')
def shim(cuda_prims):
    sys.path.insert(0, os.path.join(ROOT, 'integration'))
    import decompositions_b200
    return decompositions_b200


def tensor(shape, dtype, seed, decay=0.7):
    g = torch.Generator().manual_seed(seed)
    real = torch.float32 if dtype == C64 else torch.float64
    t = torch.complex(torch.randn(*shape, generator=g, dtype=real), torch.randn(*shape, generator=g, dtype=real))
    # graded spectrum so that the error rules actually cut somewhere
    flat = t.reshape(shape[0] * shape[1], -1)
    u, s, vh = torch.linalg.svd(flat, full_matrices=False)
    s = decay ** torch.arange(s.numel(), dtype=s.dtype)
    return ((u * s) @ vh).reshape(shape)


@pytest.mark.parametrize('dtype,tol', [(C128, 1e-10), (C64, 2e-5)])
@pytest.mark.parametrize('shape,pivot', [((6, 2, 3, 8), 2), ((4, 2, 5, 3), 3), ((16, 2, 2, 16), 2)])
@pytest.mark.parametrize('cap,err,relative', [(None, None, False), (5, None, False), (None, 2.718281828459045e-8, False),
                                              (7, 1e-2, True), (None, 5e-2, True)])
def test_svd_shim_matches_reference_contract(shim, dtype, tol, shape, pivot, cap, err, relative):
    T = tensor(shape, dtype, seed=3)
    u0, s0, vh0, rest0 = orc.svd(T, pivot, cap, err, relative, mode='exact')
    u, s, vh, rest = shim.svd(torch, T.cuda(), pivot, cap, err, relative)
    assert s.shape == s0.shape and rest.shape == rest0.shape           # same kept rank
    assert u.shape == u0.shape and vh.shape == vh0.shape
    assert (s.cpu() - s0).abs().max() <= tol * s0.abs().max()
    if rest0.numel():
        assert (rest.cpu() - rest0).abs().max() <= tol * s0.abs().max()
    k = s.numel()
    left = 1
    for d in shape[:pivot]:
        left *= d
    rec = (u.reshape(left, k) * s) @ vh.reshape(k, -1)
    rec0 = (u0.reshape(left, k) * s0) @ vh0.reshape(k, -1)
    assert (rec.cpu() - rec0).abs().max() <= 20 * tol * s0.abs().max()  # gauge-invariant: the truncated matrix
    um = u.reshape(left, k).to(C128)
    assert (um.mH @ um - torch.eye(k, device=um.device)).abs().max() <= 50 * tol


@pytest.mark.parametrize('dtype,tol', [(C128, 1e-10), (C64, 2e-5)])
def test_qr_shim_contract(shim, dtype, tol):
    T = tensor((12, 2, 3, 10), dtype, seed=5, decay=0.8)
    q, r = shim.qr(torch, T.cuda(), 3)
    assert q.shape == (12, 2, 3, 10) and r.shape == (10, 10)
    qm = q.reshape(-1, 10).to(C128)
    assert (qm.mH @ qm - torch.eye(10, device=qm.device)).abs().max() <= 50 * tol     # isometry (full rank here)
    rec = qm @ r.to(C128)
    assert (rec.cpu() - T.reshape(-1, 10).to(C128)).abs().max() <= 20 * tol * T.abs().max()
    # same column space as the reference's Householder QR
    q0, r0 = orc.qr(T, 3)
    P0 = q0.reshape(-1, 10).to(C128) @ q0.reshape(-1, 10).to(C128).mH
    P = (qm @ qm.mH).cpu()
    assert (P - P0).abs().max() <= 50 * tol
