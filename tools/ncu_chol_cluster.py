"""Dev helper (GPU): one pivoted Cholesky with inverse at n = 192 (chol_cluster_kernel) for an ncu source-level capture."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
torch.manual_seed(0)
n = 192
A = torch.randn(n, n, dtype=torch.complex128, device='cuda')
Q, _ = torch.linalg.qr(A)
lam = torch.tensor([max(0.93 ** i, 1e-30) for i in range(n)], dtype=torch.float64, device='cuda')
G = ((Q * lam.to(torch.complex128)) @ Q.mH).contiguous().unsqueeze(0)
for _ in range(3):
    p.chol_psd(G)
torch.cuda.synchronize()
print('done')
