"""Dev helper (GPU): the eigen-solver on graded Hermitian PSD matrices of order 300..1024 (the cores / bond Gram
matrices of the chi = 128 / 256 configurations): accuracy and time of both routes. Run once per kernel flavour
(MPDO_JACOBI_NOWIDE=1 selects the one-warp-per-pair persistent kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
C128 = torch.complex128


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def graded(n, lo, B=1):
    g = torch.Generator(device=dev).manual_seed(n)
    A = torch.randn(B, n, n, dtype=C128, device=dev, generator=g)
    Q, _ = torch.linalg.qr(A)
    lam = (lo ** (torch.arange(n, device=dev, dtype=torch.float64) / (n - 1)))
    return ((Q * lam.to(C128)) @ Q.mH).contiguous(), lam


print('knobs', {k: v for k, v in os.environ.items() if k.startswith('MPDO_')})
for n, B in ((300, 1), (384, 1), (512, 1), (768, 1), (1024, 1), (512, 4)):
    G, lam_true = graded(n, 1e-12, B)
    for rr, tol in ((int(os.environ.get('RR', '1')), 1e-10), (False, 1e-15)):
        if not rr and n > 512 and os.environ.get('SKIP_BIG_CLASSIC'):
            continue
        lam, Vh = p.eigh_psd(G, tol, rank_revealing=rr)
        torch.cuda.synchronize()
        rec = (Vh.mH * lam.to(C128).unsqueeze(-2)) @ Vh
        err = (torch.linalg.norm(rec - G) / torch.linalg.norm(G)).item()
        k = int((lam[0] > 0).sum())
        orth = torch.linalg.norm(Vh[0, :k] @ Vh[0, :k].mH - torch.eye(k, dtype=C128, device=dev)).item()
        lerr = ((lam[0] - lam_true).abs().max() / lam_true[0]).item()
        rel = ((lam[0, :k] - lam_true[:k]).abs() / lam_true[:k]).max().item()
        t = timeit(lambda: p.eigh_psd(G, tol, rank_revealing=rr))
        print('n=%4d B=%d %-14s rank %4d  recon %.1e  orth %.1e  lam abs %.1e rel %.1e   %8.3f ms' %
              (n, B, ('chol+jacobi rr=%d' % rr) if rr else 'classic [G|I]', k, err, orth, lerr, rel, t), flush=True)
