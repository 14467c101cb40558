"""The native engine (one C call per step, csrc/engine.cu: Cholesky-QR + preconditioned eigen-solver for complex64)
against the readable Python statement of the same steps (`_engine/steps.py`, eigen route everywhere), on
gauge-invariant quantities: the two engines use different but equally valid gauges (R factors in pivot order vs
eigenvector order), so site tensors are compared through their contractions, never entry by entry."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

C64, C128 = torch.complex64, torch.complex128


def gauss(shape, dtype, seed, dev):
    g = torch.Generator().manual_seed(seed)
    real = torch.float32 if dtype == C64 else torch.float64
    t = torch.complex(torch.randn(*shape, generator=g, dtype=real), torch.randn(*shape, generator=g, dtype=real))
    return (t / math.sqrt(2 * shape[1] * shape[3])).to(dev)


def two_site(Tl, Tr):
    """Theta[b, l, s, a, s', a', r] = sum_m Tl[b,l,s,a,m] Tr[b,m,s',a',r] in complex128."""
    return torch.einsum('blsam,bmtcr->blsatcr', Tl.to(C128), Tr.to(C128))


def rho_site(T):
    """sum over the inner index of T (x) conj(T): what every later contraction sees of a site."""
    T = T.to(C128)
    return torch.einsum('blsar,bmtaq->blsrmtq', T, T.conj())


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


@pytest.fixture(scope='module')
def engines(cuda_prims):
    from MPDOSimulator._engine.native import NativeEngine
    from MPDOSimulator._engine.prims import CudaPrims
    from MPDOSimulator._engine.steps import Engine
    p = cuda_prims
    return {dt: (Engine(p, dt), NativeEngine(p, dt)) for dt in (C64, C128)}


@pytest.mark.parametrize('dt,tol', [(C64, 2e-5), (C128, 1e-10)])
@pytest.mark.parametrize('l,a,r,Bn', [(4, 3, 6, 2), (16, 4, 40, 1), (12, 6, 150, 1)])
def test_qr_step(engines, dt, tol, l, a, r, Bn):
    py, nat = engines[dt]
    Ti, Tn = gauss((Bn, l, 2, a, r), dt, 1, 'cuda'), gauss((Bn, r, 2, 2, 5), dt, 2, 'cuda')
    want = two_site(Ti, Tn)
    for eng in (py, nat):
        Q, Tn2 = eng.qr_step(Ti, Tn)
        assert rel(two_site(Q, Tn2), want) <= tol
        # Q is an isometry on its range: Q^h Q is a projector
        Qm = Q.to(C128).reshape(Bn, -1, r)
        P = Qm.mH @ Qm
        assert rel(P @ P, P) <= 50 * tol


@pytest.mark.parametrize('dt,tol', [(C64, 2e-5), (C128, 1e-10)])
@pytest.mark.parametrize('lp,l,r,chi', [(6, 24, 5, 8), (40, 130, 8, 16), (10, 20, 12, None)])
def test_bond_svd_step(engines, dt, tol, lp, l, r, chi):
    py, nat = engines[dt]
    Bn, ap, a = 2, 2, 3
    # left neighbour must be left-isometric, as after the QR sweep
    X = gauss((Bn, lp, 2, ap, l), dt, 3, 'cuda')
    rows = lp * 2 * ap
    assert rows >= l
    Qm, _ = torch.linalg.qr(X.to(C128).reshape(Bn, rows, l))
    Tl = Qm.reshape(Bn, lp, 2, ap, l).to(dt).contiguous()
    Tr = gauss((Bn, l, 2, a, r), dt, 4, 'cuda') * torch.logspace(0, -3, l, device='cuda').reshape(1, l, 1, 1, 1).to(dt)
    outs = []
    for eng in (py, nat):
        Tl2, Tr2, disc = eng.bond_svd_step(Tl, Tr, chi)
        outs.append(two_site(Tl2, Tr2))
        k = Tl2.shape[-1]
        assert k == (l if chi is None else min(chi, l))
    assert rel(outs[1], outs[0]) <= 10 * tol
    if chi is None:
        assert rel(outs[1], two_site(Tl, Tr)) <= 10 * tol


@pytest.mark.parametrize('dt,tol', [(C64, 2e-5), (C128, 1e-10)])
@pytest.mark.parametrize('Bn,l,a,r,kappa', [(2, 4, 24, 5, 4), (2, 8, 96, 8, 4), (16, 8, 96, 8, 4), (2, 6, 300, 6, 8),
                                            (2, 3, 5, 3, 8)])
def test_kappa_truncate(engines, dt, tol, Bn, l, a, r, kappa):
    # (the native call takes the top-kappa subspace iteration for a >= 256 or batches of 16 and more, the full
    # decomposition otherwise)
    py, nat = engines[dt]
    T = gauss((Bn, l, 2, a, r), dt, 5, 'cuda')
    T = T * torch.logspace(0, -4, a, device='cuda').reshape(1, 1, 1, a, 1).to(dt)   # well separated branches
    got = [rho_site(eng.kappa_truncate(T, kappa)[0]) for eng in (py, nat)]
    assert rel(got[1], got[0]) <= 10 * tol
    if kappa >= a:
        assert rel(got[1], rho_site(T)) <= 10 * tol


@pytest.mark.parametrize('dt,tol', [(C64, 5e-5), (C128, 1e-9)])
@pytest.mark.parametrize('l,a0,m,a1,r,K', [(4, 2, 6, 2, 4, 1), (8, 3, 10, 2, 8, 4), (16, 4, 24, 4, 16, 16)])
def test_split_2q(engines, dt, tol, l, a0, m, a1, r, K):
    py, nat = engines[dt]
    Bn = 1
    Tlo, Thi = gauss((Bn, l, 2, a0, m), dt, 6, 'cuda'), gauss((Bn, m, 2, a1, r), dt, 7, 'cuda')
    g = torch.Generator().manual_seed(9)
    real = torch.float32 if dt == C64 else torch.float64
    G = torch.complex(torch.randn(1, 2, 2, 2, 2, K, generator=g, dtype=real),
                      torch.randn(1, 2, 2, 2, 2, K, generator=g, dtype=real)).to('cuda')
    G = G * torch.logspace(0, -2, K, device='cuda').to(dt)            # one dominant Kraus branch, weak errors
    res = []
    for eng in (py, nat):
        lo, hi = eng.split_2q(Tlo, Thi, G)
        assert lo.shape[-1] == hi.shape[1]
        # trace both inner (Kraus) indices against their conjugates: the only way they are ever used
        Th = two_site(lo, hi)
        res.append(torch.einsum('blsatcr,bmuavcq->blstrmuvq', Th, Th.conj()))
    assert rel(res[1], res[0]) <= 10 * tol
