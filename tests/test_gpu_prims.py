"""-m gpu: every CUDA primitive against its torch-CPU model (tests/cpu_prims.py)."""
import pytest
import torch

from cpu_prims import CpuPrims

pytestmark = pytest.mark.gpu
C64, C128 = torch.complex64, torch.complex128


def rnd(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.float32 if dtype == C64 else torch.float64
    return torch.complex(torch.randn(shape, generator=g, dtype=real), torch.randn(shape, generator=g, dtype=real))


def both(fn, cuda_prims, *tensors):
    cpu = CpuPrims()
    out_c = fn(cpu, *[t.clone() for t in tensors])
    out_g = fn(cuda_prims, *[t.cuda() for t in tensors])
    torch.cuda.synchronize()
    return out_c, out_g


@pytest.mark.parametrize('dt,acc64,tol', [(C64, False, 2e-5), (C64, True, 1e-6), (C128, None, 1e-12)])
@pytest.mark.parametrize('M,N,K,Bn', [(64, 64, 16, 1), (70, 33, 45, 3), (5, 130, 7, 2), (128, 256, 300, 1), (1, 1, 1, 4)])
def test_contract_plain(cuda_prims, dt, acc64, tol, M, N, K, Bn):
    A, B = rnd((Bn, M, K), dt, 1), rnd((Bn, K, N), dt, 2)

    def f(p, A, B):
        C = torch.zeros((Bn, M, N), dtype=dt, device=A.device)
        return p.contract(A, (1, 1, 1), B, (1, 1, 1), C, (1, 1, 1), acc64=acc64)

    c, g = both(f, cuda_prims, A, B)
    assert (g.cpu() - c).abs().max() <= tol * max(1.0, c.abs().max().item())


@pytest.mark.parametrize('dt,tol', [(C64, 2e-5), (C128, 1e-12)])
def test_contract_views_conj_beta(cuda_prims, dt, tol):
    # site-tensor style composite axes: C[b,(l,s,a),j] = sum_r T[b,l,s,a,r] conj(X[b,j,r]) + 0.5*C
    Bn, l, a, r, k = 2, 5, 3, 7, 6
    T, X, C0 = rnd((Bn, l, 2, a, r), dt, 3), rnd((Bn, k, r), dt, 4), rnd((Bn, l, 2, a, k), dt, 5)

    def f(p, T, X, C0):
        return p.contract(T, (1, 3, 1), X.permute(0, 2, 1), (1, 1, 1), C0, (1, 3, 1), conjB=True, alpha=2.0, beta=0.5)

    c, g = both(f, cuda_prims, T, X, C0)
    assert (g.cpu() - c).abs().max() <= tol * c.abs().max().item()


@pytest.mark.parametrize('dt,tol', [(C64, 1e-6), (C128, 1e-12)])
def test_contract_gram_kappa_axis(cuda_prims, dt, tol):
    # Gram over (l,s,r) of the inner axis a (three-level K index), output complex128
    Bn, l, a, r = 2, 6, 10, 9
    T = rnd((Bn, l, 2, a, r), dt, 6)

    def f(p, T):
        Tv = T.permute(0, 1, 2, 4, 3)
        G = torch.zeros((Bn, a, a), dtype=C128, device=T.device)
        return p.contract(Tv.permute(0, 4, 1, 2, 3), (1, 1, 3), Tv, (1, 3, 1), G, (1, 1, 1), conjA=True, acc64=True)

    c, g = both(f, cuda_prims, T)
    assert (g.cpu() - c).abs().max() <= tol * c.abs().max().item()

@pytest.mark.parametrize('dt,tol', [(C64, 1e-6), (C128, 1e-12)])
@pytest.mark.parametrize('n,K,Bn', [(10, 40, 2), (64, 16, 1), (130, 700, 2), (200, 64, 3), (513, 96, 1)])
def test_contract_hermitian_gram(cuda_prims, dt, tol, n, K, Bn):
    """hermitian=1 computes the tiles on/below the diagonal and mirrors them; with and without split-K."""
    X = rnd((Bn, K, n), dt, 8)

    def f(p, X):
        G = torch.zeros((Bn, n, n), dtype=C128, device=X.device)
        return p.contract(X.permute(0, 2, 1), (1, 1, 1), X, (1, 1, 1), G, (1, 1, 1), conjA=True, acc64=True,
                          hermitian=True)

    c, g = both(f, cuda_prims, X)
    g = g.cpu()
    assert (g - c).abs().max() <= tol * c.abs().max().item()
    assert (g - g.mH).abs().max() <= 1e-13 * c.abs().max().item()



def test_contract_splitk_and_broadcast(cuda_prims):
    # long K with few tiles triggers the split-K path; A broadcast over batch with stride 0
    Bn, M, N, K = 2, 8, 8, 5000
    A, B = rnd((1, M, K), C64, 7), rnd((Bn, K, N), C64, 8)

    def f(p, A, B):
        C = torch.zeros((Bn, M, N), dtype=C128, device=A.device)
        return p.contract(A.expand(Bn, M, K), (1, 1, 1), B, (1, 1, 1), C, (1, 1, 1), acc64=True)

    c, g = both(f, cuda_prims, A, B)
    assert (g.cpu() - c).abs().max() <= 1e-6 * c.abs().max().item()


@pytest.mark.parametrize('dt', [C64, C128])
@pytest.mark.parametrize('K,Bg', [(1, 1), (6, 1), (16, 3)])
def test_absorb_1q(cuda_prims, dt, K, Bg):
    Bn, l, a, r = 3, 5, 4, 7
    T, G = rnd((Bn, l, 2, a, r), dt, 9), rnd((Bg, 2, 2, K), dt, 10)
    c, g = both(lambda p, T, G: p.absorb_1q(T, G), cuda_prims, T, G)
    assert g.shape == c.shape
    assert (g.cpu() - c).abs().max() <= (1e-5 if dt == C64 else 1e-13) * c.abs().max().item()


@pytest.mark.parametrize('n,Bn', [(1, 2), (2, 2), (5, 3), (16, 2), (64, 2), (65, 1), (100, 2), (200, 1), (300, 1),
                                  (330, 2), (500, 1)])
def test_eigh_psd(cuda_prims, n, Bn):
    """Complete basis from Jacobi on [G | I] (the complex128 route). n > 256: rows of 2n entries, block pairs held in
    registers (jacobi_persistent_cols_kernel); (330, 2) is two matrices sharing the device."""
    A = rnd((Bn, n, n + 3), C128, 11)
    G = A @ A.mH
    lam_g, Vh_g = cuda_prims.eigh_psd(G.cuda())
    lam_g, Vh_g = lam_g.cpu(), Vh_g.cpu()
    lam_c = torch.linalg.eigvalsh(G).flip(-1)
    assert (lam_g - lam_c).abs().max() <= 1e-12 * lam_c.max()
    eye = torch.eye(n, dtype=C128).expand(Bn, n, n)
    assert (Vh_g @ Vh_g.mH - eye).abs().max() <= 1e-12
    rec = Vh_g.mH @ (lam_g.to(C128)[:, :, None] * Vh_g)
    assert (rec - G).abs().max() <= 1e-12 * G.abs().max()


def test_eigh_psd_rank_deficient_and_degenerate(cuda_prims):
    n, rk = 40, 7
    A = rnd((2, n, rk), C128, 12)
    G = A @ A.mH
    G[1] = torch.eye(n, dtype=C128) * 2.0          # fully degenerate
    lam, Vh = cuda_prims.eigh_psd(G.cuda())
    lam, Vh = lam.cpu(), Vh.cpu()
    assert (lam[0, rk:].abs() <= 1e-12 * lam[0, 0]).all()
    rec = Vh.mH @ (lam.to(C128)[:, :, None] * Vh)
    assert (rec - G).abs().max() <= 1e-12 * G.abs().max()

@pytest.mark.gpu
@pytest.mark.parametrize('n,rk,Bn', [(1, 1, 2), (2, 2, 3), (24, 24, 5), (40, 7, 2), (100, 60, 3), (113, 113, 2),
                                     (130, 90, 2), (256, 256, 1), (400, 300, 1), (512, 512, 1),
                                     (160, 120, 80), (256, 200, 40)])
def test_eigh_psd_rank_revealing(cuda_prims, n, rk, Bn):
    """Pivoted-Cholesky preconditioned route: graded spectrum, numerical rank rk < n, batch. The last two cases
    are batches whose CTA groups cannot all be co-resident: the factorisation and the persistent Jacobi walk them in
    chunks (parameter sweeps, cfg4)."""
    A = rnd((Bn, n, rk), C128, 21)
    A = A * torch.logspace(0, -5, rk, dtype=torch.float64)       # eigenvalues graded over ten decades
    G = A @ A.mH
    G = 0.5 * (G + G.mH)
    lam_g, Vh_g = cuda_prims.eigh_psd(G.cuda(), rank_revealing=True)
    lam_g, Vh_g = lam_g.cpu(), Vh_g.cpu()
    lam_c = torch.linalg.eigvalsh(G).flip(-1).clamp_min(0.0)
    top = lam_c.max()
    assert (lam_g - lam_c).abs().max() <= 1e-12 * top
    # high relative accuracy on the resolved part of the spectrum (the point of preconditioned Jacobi)
    big = lam_c > 1e-9 * top
    assert ((lam_g - lam_c).abs()[big] <= 1e-6 * lam_c[big]).all()
    rec = Vh_g.mH @ (lam_g.to(C128)[:, :, None] * Vh_g)
    assert (rec - G).abs().max() <= 1e-12 * G.abs().max()
    # rows are orthonormal or zero
    gram = Vh_g @ Vh_g.mH
    d = gram.diagonal(dim1=1, dim2=2).real
    assert (((d - 1).abs() <= 1e-10) | (d.abs() <= 1e-300)).all()
    off = gram - torch.diag_embed(gram.diagonal(dim1=1, dim2=2))
    assert off.abs().max() <= 1e-10
    assert (lam_g[:, rk:] <= 1e-12 * top).all()


@pytest.mark.gpu
@pytest.mark.parametrize('n,rk,Bn', [(17, 17, 2), (24, 9, 3), (40, 40, 2), (64, 64, 4), (64, 20, 2), (81, 81, 2), (100, 60, 3), (130, 90, 2), (192, 192, 1), (200, 33, 2), (256, 256, 1),
                                     (256, 200, 12), (160, 120, 40)])
def test_eigh_psd_blocked_factorisation(cuda_prims, n, rk, Bn):
    """precondition = 2: blocked Cholesky without pivoting (rows ordered once by decreasing diagonal, null pivots
    skipped) in front of the Jacobi phase, for Gram matrices of fp32 data. Held to what that use needs: eigenvalues
    to 1e-12 of the largest (the same absolute bound as the pivoted route), reconstruction to 1e-12, orthonormal-or-
    zero rows, the numerical rank; relative accuracy on the resolved spectrum to 1e-4 (1e-6 on the pivoted route)."""
    A = rnd((Bn, n, rk), C128, 23)
    A = A * torch.logspace(0, -5, rk, dtype=torch.float64)
    G = A @ A.mH
    G = 0.5 * (G + G.mH)
    lam_g, Vh_g = cuda_prims.eigh_psd(G.cuda(), rank_revealing=2)
    lam_g, Vh_g = lam_g.cpu(), Vh_g.cpu()
    lam_c = torch.linalg.eigvalsh(G).flip(-1).clamp_min(0.0)
    top = lam_c.max()
    assert (lam_g - lam_c).abs().max() <= 1e-12 * top
    big = lam_c > 1e-9 * top
    assert ((lam_g - lam_c).abs()[big] <= 1e-4 * lam_c[big]).all()
    rec = Vh_g.mH @ (lam_g.to(C128)[:, :, None] * Vh_g)
    assert (rec - G).abs().max() <= 1e-12 * G.abs().max()
    gram = Vh_g @ Vh_g.mH
    d = gram.diagonal(dim1=1, dim2=2).real
    assert (((d - 1).abs() <= 1e-10) | (d.abs() <= 1e-300)).all()
    off = gram - torch.diag_embed(gram.diagonal(dim1=1, dim2=2))
    assert off.abs().max() <= 1e-10
    assert (lam_g[:, rk:] <= 1e-12 * top).all()


@pytest.mark.gpu
def test_eigh_psd_blocked_factorisation_structured(cuda_prims):
    """Cases that defeat a factorisation without pivoting unless null pivots are handled: zero rows in the middle, a tiny
    leading diagonal coupled to a large one, exactly repeated rows, the identity."""
    n = 96
    Gs = []
    G = torch.zeros((n, n), dtype=C128); Gs.append(G)                                   # zero matrix
    Gs.append(torch.eye(n, dtype=C128) * 2.5)                                            # identity
    v = rnd((1, n, 3), C128, 5)[0]; v[10:20] = 0; Gs.append(v @ v.mH)                    # rank 3, zero rows inside
    w = rnd((1, n, n), C128, 6)[0]; w[0] *= 1e-7; Gs.append(w @ w.mH)                    # tiny leading diagonal, coupled
    u = rnd((1, n, 40), C128, 7)[0]; u[50:] = u[:46]; Gs.append(u @ u.mH)                # repeated rows (rank 40)
    G = torch.stack([0.5 * (g + g.mH) for g in Gs])
    lam_g, Vh_g = cuda_prims.eigh_psd(G.cuda(), rank_revealing=2)
    lam_g, Vh_g = lam_g.cpu(), Vh_g.cpu()
    for b in range(G.shape[0]):
        lam_c = torch.linalg.eigvalsh(G[b]).flip(-1).clamp_min(0.0)
        top = max(float(lam_c.max()), 1e-300)
        assert (lam_g[b] - lam_c).abs().max() <= 1e-11 * top, b
        rec = Vh_g[b].mH @ (lam_g[b].to(C128)[:, None] * Vh_g[b])
        assert (rec - G[b]).abs().max() <= 1e-11 * max(float(G[b].abs().max()), 1e-300), b


@pytest.mark.gpu
def test_eigh_psd_rank_revealing_zero_and_identity(cuda_prims):
    G = torch.zeros((3, 16, 16), dtype=C128)
    G[1] = torch.eye(16, dtype=C128) * 3.0
    G[2, 5, 5] = 2.0
    lam, Vh = cuda_prims.eigh_psd(G.cuda(), rank_revealing=True)
    lam, Vh = lam.cpu(), Vh.cpu()
    assert lam[0].abs().max() == 0 and Vh[0].abs().max() == 0
    assert (lam[1] - 3.0).abs().max() <= 1e-14
    assert abs(lam[2, 0] - 2.0) <= 1e-14 and lam[2, 1:].abs().max() == 0
    rec = Vh.mH @ (lam.to(C128)[:, :, None] * Vh)
    assert (rec - G).abs().max() <= 1e-13


@pytest.mark.parametrize('n,rk,Bn', [(1, 1, 2), (3, 2, 2), (24, 24, 4), (60, 31, 3), (84, 84, 1), (85, 85, 2),
                                     (130, 70, 2), (256, 256, 1), (512, 400, 1), (160, 100, 40), (256, 256, 24)])
def test_chol_psd(cuda_prims, n, rk, Bn):
    """Pivoted Cholesky with left inverse: the two identities every Cholesky-QR step of the engine relies on."""
    A = rnd((Bn, n, rk), C128, 31)
    A = A * torch.logspace(0, -4, rk, dtype=torch.float64)
    G = A @ A.mH
    G = 0.5 * (G + G.mH)
    Lh, Linv, rank = cuda_prims.chol_psd(G.cuda(), rel=1e-12)
    Lh, Linv, rank = Lh.cpu(), Linv.cpu(), rank.cpu()
    assert (rank <= rk).all() and (rank >= min(rk, 1)).all()
    assert (Lh.mH @ Lh - G).abs().max() <= 1e-12 * G.abs().max()
    for b in range(Bn):
        r = int(rank[b])
        P = Linv[b] @ Lh[b].mH
        want = torch.zeros(n, dtype=torch.float64)
        want[:r] = 1.0
        assert (P - torch.diag(want).to(C128)).abs().max() <= 1e-8
        assert Lh[b, r:].abs().max() == 0 if r < n else True
        assert Linv[b, r:].abs().max() == 0 if r < n else True
        # pivot order: the diagonal of the factor decreases
        piv = (Lh[b, :r].abs() ** 2).sum(1)
        assert piv[0] >= piv[-1]



@pytest.mark.parametrize('n,m', [(4, 4), (20, 33), (64, 64), (90, 70), (130, 130)])
def test_svd_rows(cuda_prims, n, m):
    L = rnd((2, n, m), C128, 13)
    L[1] *= torch.logspace(0, -9, n, dtype=torch.float64)[:, None]   # graded rows
    Uh, s, Wh = cuda_prims.svd_rows(L.cuda())
    Uh, s, Wh = Uh.cpu(), s.cpu(), Wh.cpu()
    s_c = torch.linalg.svdvals(L)
    k = min(n, m)
    assert (s[:, :k] - s_c).abs().max() <= 1e-12 * s_c.max()
    rec = Uh.mH @ (s.to(C128)[:, :, None] * Wh)
    assert (rec - L).abs().max() <= 1e-12 * L.abs().max()
    eye = torch.eye(n, dtype=C128)
    assert (Uh @ Uh.mH - eye).abs().max() <= 1e-12


def test_rowscale_and_rank_rule(cuda_prims):
    V = rnd((2, 6, 5), C128, 14)
    lam = torch.tensor([[4.0, 1.0, 1e-3, 1e-20, 0.0, 0.0], [9.0, 4.0, 1.0, 0.25, 1e-2, 1e-4]], dtype=torch.float64)
    cpu = CpuPrims()
    for power, tol, mode, dt in [(-0.5, 1e-14, 0, C128), (0.5, 1e-13, 1, C128), (0.25, 0.0, 0, C64)]:
        c = cpu.rowscale(V, lam, 4, power, tol, mode, dt)
        g = cuda_prims.rowscale(V.cuda(), lam.cuda(), 4, power, tol, mode, dt).cpu()
        assert (g - c).abs().max() <= 1e-6 * c.abs().max() if dt == C64 else (g - c).abs().max() <= 1e-13 * c.abs().max()
    s = torch.tensor([[1.0, 0.5, 1e-3, 1e-5, 1e-9, 0.0], [1.0, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4]], dtype=torch.float64)
    for f32 in (True, False):
        for (cap, err, relv, sq) in [(6, 2.718281828459045e-8, False, False), (3, None, True, False),
                                     (6, 1e-3, True, False), (6, 1e-6, True, True)]:
            sc, sg = s.clone(), s.clone().cuda()
            kc = cpu.rank_rule(sc, sq, cap, err, relv, f32)
            kg = cuda_prims.rank_rule(sg, sq, cap, err, relv, f32)
            assert kc == kg, (f32, cap, err, relv, sq, kc, kg)
            if err is not None:
                assert torch.equal(sc, sg.cpu())
