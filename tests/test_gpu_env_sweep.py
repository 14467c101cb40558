"""GPU tier: the environment form of bondTruncate (mpdo_env_sweep + mpdo_bond_env_step through the C ABI) against the
two-sweep form (mpdo_qr_step + mpdo_bond_svd_step), against the same algorithm issued over the Python primitive
wrappers, and against a plain torch complex128 restatement of TNNOptimizer.py:72-134 on random MPDO states whose bonds
are wider than chi. Gauge-invariant comparison: dense density matrix, discarded singular values."""
import pytest
import torch

pytestmark = pytest.mark.gpu
C64, C128 = torch.complex64, torch.complex128


def random_state(bonds, inner, seed, B=1, device='cuda'):
    g = torch.Generator().manual_seed(seed)
    Ts = []
    for i in range(len(bonds) - 1):
        shp = (B, bonds[i], 2, inner[i], bonds[i + 1])
        t = torch.complex(torch.randn(shp, generator=g, dtype=torch.float64), torch.randn(shp, generator=g, dtype=torch.float64))
        Ts.append((t / t.abs().max() / (bonds[i] ** 0.5)).to(device))
    return Ts


def dense_rho(Ts):
    """rho[b, s..., s'...] from site tensors [B,l,s,a,r] (complex128, tiny registers only)."""
    out = []
    for b in range(Ts[0].shape[0]):
        R = torch.ones((1, 1, 1, 1), dtype=C128, device=Ts[0].device)       # [P, P', l, l']
        for T in Ts:
            t = T[b].to(C128)
            R = torch.einsum('pqlm,lsar,mtac->psqtrc', R, t, t.conj())
            P, P2 = R.shape[0] * 2, R.shape[2] * 2
            R = R.reshape(P, P2, R.shape[4], R.shape[5])
        out.append(R[:, :, 0, 0])
    return torch.stack(out)


def reference_bond_truncate(Ts, chi):
    """TNNOptimizer.py:72-134 in complex128 torch: QR sweep, then two-site SVD truncation with sqrt(S) on both sides."""
    Ts = [t.to(C128).clone() for t in Ts]
    n = len(Ts)
    for i in range(n - 1):
        B, l, s, a, r = Ts[i].shape
        Q, R = torch.linalg.qr(Ts[i].reshape(B, l * s * a, r))
        k = Q.shape[2]
        Ts[i] = Q.reshape(B, l, s, a, k)
        Ts[i + 1] = torch.einsum('bkr,brsac->bksac', R, Ts[i + 1])
    disc = []
    for idx in range(n - 1, 0, -1):
        B, l, s, a, r = Ts[idx].shape
        U, S, Vh = torch.linalg.svd(Ts[idx].reshape(B, l, s * a * r), full_matrices=False)
        k = min(chi, S.shape[1])
        disc.append(S[:, k:])
        sq = S[:, :k].sqrt().to(C128)
        Ts[idx] = (sq[:, :, None] * Vh[:, :k]).reshape(B, k, s, a, r)
        Ts[idx - 1] = torch.einsum('blsar,brk->blsak', Ts[idx - 1], U[:, :, :k] * sq[:, None, :])
    return Ts, disc


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


CASES = [([1, 6, 20, 12, 5, 1], [1, 2, 3, 2, 1], 8), ([1, 8, 40, 90, 30, 8, 1], [2, 3, 2, 4, 2, 1], 16),
         ([1, 2, 16, 64, 150, 64, 100, 8, 1], [1, 4, 16, 4, 16, 4, 16, 2], 32)]


@pytest.mark.parametrize('bonds,inner,chi', CASES)
@pytest.mark.parametrize('B', [1, 3])
def test_env_form_matches_two_sweeps_and_reference(cuda_prims, bonds, inner, chi, B):
    from MPDOSimulator._engine.native import NativeEngine
    from MPDOSimulator._engine.steps import Engine
    eng = NativeEngine(cuda_prims, C64)
    state = [t.to(C64) for t in random_state(bonds, inner, seed=sum(bonds) + B, B=B)]
    ref, ref_disc = reference_bond_truncate(state, chi)
    rho_ref = dense_rho(ref) if len(bonds) <= 7 else None

    A = [t.clone() for t in state]
    for i in range(len(A) - 1):
        A[i], A[i + 1] = eng.qr_step(A[i], A[i + 1])
    two_disc = []
    for idx in range(len(A) - 1, 0, -1):
        A[idx - 1], A[idx], d = eng.bond_svd_step(A[idx - 1], A[idx], chi)
        two_disc.append(d)
    Bt = [t.clone() for t in state]
    env_disc = eng.bond_truncate_env(Bt, chi)
    Pt = [t.clone() for t in state]
    py_disc = Engine.bond_truncate_env(Engine(cuda_prims, C64), Pt, chi)
    torch.cuda.synchronize()

    # (the reference's reduced QR / SVD shrink a bond that is wider than its block's rank, the dense build keeps the
    # bond and carries zero directions - compare shapes and singular values between the two device forms only)
    assert [tuple(t.shape) for t in Bt] == [tuple(t.shape) for t in A] == [tuple(t.shape) for t in Pt]
    for d_env, d_two, d_py in zip(env_disc, two_disc, py_disc):
        if d_env.numel() == 0:
            continue
        assert float((d_env - d_two).abs().max()) <= 2e-5, 'discarded singular values: env form vs two sweeps'
        assert float((d_py - d_env).abs().max()) <= 2e-5
    for d_env, d_ref in zip(env_disc, ref_disc):
        m = min(d_env.shape[1], d_ref.shape[1])
        if m:
            assert float((d_env[:, :m].cpu() - d_ref[:, :m].cpu()).abs().max()) <= 2e-5, 'vs the complex128 reference'
    if rho_ref is not None:
        e_env, e_two, e_py = rel(dense_rho(Bt), rho_ref), rel(dense_rho(A), rho_ref), rel(dense_rho(Pt), rho_ref)
        print(f'bonds {bonds} B={B}: env {e_env:.2e}  two-sweep {e_two:.2e}  env over python primitives {e_py:.2e}')
        assert e_env <= max(1e-5, 3 * e_two), (e_env, e_two)
        assert e_py <= max(1e-5, 3 * e_two), (e_py, e_two)


def test_env_form_rank_deficient_bond(cuda_prims):
    """A bond whose environment is numerically singular (the left block has fewer independent directions than the bond
    dimension): the rank-revealing factor drops the null directions, as the Cholesky-QR sweep does."""
    from MPDOSimulator._engine.native import NativeEngine
    eng = NativeEngine(cuda_prims, C64)
    state = [t.to(C64) for t in random_state([1, 2, 24, 10, 1], [1, 1, 2, 1], seed=9)]   # bond 2 has rank <= 4
    ref, _ = reference_bond_truncate(state, 6)
    Bt = [t.clone() for t in state]
    eng.bond_truncate_env(Bt, 6)
    torch.cuda.synchronize()
    assert all(torch.isfinite(torch.view_as_real(t)).all() for t in Bt)
    assert rel(dense_rho(Bt), dense_rho(ref)) <= 1e-5


def test_circuit_level_env_form_against_two_sweeps(cuda_prims, monkeypatch):
    """The same noisy brickwork circuit through TensorCircuit with the environment form forced on and off."""
    import math
    import MPDOSimulator as Simulator
    from MPDOSimulator import dmOperations

    def run():
        g = torch.Generator().manual_seed(21)
        c = Simulator.TensorCircuit(qn=8, ideal=False, noiseType='idealNoise', chi=12, kappa=3, chip='worst',
                                    dtype=C64, device='cuda:0')
        for d in range(6):
            for q in range(8):
                th, ph, la = (torch.rand(3, generator=g) * 2 * math.pi).tolist()
                c.u3(th, ph, la, [q])
            for q in range(d % 2, 7, 2):
                c.cz(q, q + 1)
            c.truncate()
        st = Simulator.Tools.create_ket0Series(8, dtype=C64, device='cpu')
        c.evolve(st)
        return c.cal_dm().to(C128).cpu(), [int(s.data.shape[4]) for s in st[:-1]]

    monkeypatch.setenv('MPDO_ENV_SWEEP', '0')
    rho_two, bonds_two = run()
    monkeypatch.setenv('MPDO_ENV_SWEEP', '1')
    rho_env, bonds_env = run()
    assert bonds_env == bonds_two and max(bonds_env) == 12
    print('circuit-level env form vs two sweeps: %.2e' % rel(rho_env, rho_two))
    assert rel(rho_env, rho_two) <= 5e-5
