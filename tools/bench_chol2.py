"""Dev helper (GPU): pivoted Cholesky (with / without the left inverse) at the orders of a cfg2 layer; run once per
kernel flavour (MPDO_CHOL_NOCLUSTER16=1, MPDO_CHOL_NOSMALL=1 select the older kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
torch.manual_seed(0)
def graded(n, decay):
    A = torch.randn(n, n, dtype=torch.complex128, device=dev)
    Q, _ = torch.linalg.qr(A)
    lam = torch.tensor([max(decay ** i, 1e-30) for i in range(n)], dtype=torch.float64, device=dev)
    return ((Q * lam.to(torch.complex128)) @ Q.mH).contiguous().unsqueeze(0)
def timeit(fn, reps=20):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (32, 64, 80, 96, 128, 160, 192, 224, 256):
    G = graded(n, 0.93)
    t_inv = timeit(lambda: p.chol_psd(G))
    t_eig = timeit(lambda: p.eigh_psd(G, 1e-10, rank_revealing=True))
    t_eig2 = timeit(lambda: p.eigh_psd(G, 1e-10, rank_revealing=2))
    print('n=%4d  chol+inverse %.3f ms   eigh (pivoted chol + jacobi + finalize) %.3f ms   eigh (blocked chol, no pivoting) %.3f ms'
          % (n, t_inv, t_eig, t_eig2), flush=True)
