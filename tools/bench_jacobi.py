"""Dev helper (GPU): eigen-solver (pivoted Cholesky + one-sided Jacobi) on matrices captured from a cfg2 layer
(tools/data/grams_sel.npz, from tools/dump_grams.py): accuracy and time. Run once per kernel flavour
(MPDO_JACOBI_NOCLUSTER=1 selects the older persistent kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import numpy as np
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'grams_sel.npz'))
for key in data.files:
    G = torch.from_numpy(data[key]).to(dev).unsqueeze(0).contiguous()
    n = G.shape[-1]
    for rr in (True, False):
        lam, Vh = p.eigh_psd(G, 1e-10, rank_revealing=rr)
        rec = (Vh.mH * lam.to(torch.complex128).unsqueeze(-2)) @ Vh
        err = (torch.linalg.norm(rec - G) / torch.linalg.norm(G)).item()
        k = int((lam[0] > 0).sum())
        orth = (torch.linalg.norm(Vh[0, :k] @ Vh[0, :k].mH - torch.eye(k, dtype=torch.complex128, device=dev))).item()
        ref = torch.linalg.eigvalsh(G)[0].flip(0)
        lerr = ((lam[0] - ref).abs().max() / ref[0]).item()
        t = timeit(lambda: p.eigh_psd(G, 1e-10, rank_revealing=rr))
        print('%-16s n=%3d %s  rank %3d  recon %.1e  orth %.1e  lam %.1e   %.3f ms' %
              (key, n, 'chol+jacobi' if rr else 'classic [G|I]', k, err, orth, lerr, t), flush=True)
# batches (cfg4-like): 128 matrices of order 128 / 64
for n, B in ((128, 128), (64, 128), (256, 16)):
    torch.manual_seed(n)
    A = torch.randn(B, n, 2 * n, dtype=torch.complex128, device=dev) * (0.9 ** torch.arange(2 * n, device=dev)).to(torch.complex128)
    G = (A @ A.mH).contiguous()
    lam, Vh = p.eigh_psd(G, 1e-10, rank_revealing=True)
    rec = (Vh.mH * lam.to(torch.complex128).unsqueeze(-2)) @ Vh
    err = (torch.linalg.norm(rec - G) / torch.linalg.norm(G)).item()
    t = timeit(lambda: p.eigh_psd(G, 1e-10, rank_revealing=True), reps=3)
    print('batch %d x n=%d  recon %.1e  %.3f ms' % (B, n, err, t), flush=True)
