"""Dev helper (GPU): one launch of each heavy kernel at its steady-state cfg2 size, for `ncu --set full -k ...`.

  chol_kernel + jacobi_persistent_reg_kernel : preconditioned eigen-decomposition of a graded 512 x 512 Gram matrix
  contract_kernel<float2,float2,double2,double> : the kappa-step Gram matrix, 1024 x 1024 over 8192 rows, Hermitian
  contract_kernel<float2,float2,float2,float>   : U^h . Theta projection, 64 x 512 x 131072-ish (fp32 FFMA)
"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
torch.manual_seed(0)
n = 512
A = torch.randn(n, n, dtype=torch.complex128, device=dev)
Q, _ = torch.linalg.qr(A)
lam = torch.tensor([max(0.93 ** i, 1e-30) for i in range(n)], dtype=torch.float64, device=dev)
G = ((Q * lam.to(torch.complex128)) @ Q.mH).contiguous().unsqueeze(0)
for _ in range(2):
    p.eigh_psd(G, 1e-10, rank_revealing=True)
torch.cuda.synchronize()
T = (torch.randn(1, 64, 2, 1024, 64, device=dev) + 1j * torch.randn(1, 64, 2, 1024, 64, device=dev)).to(torch.complex64)
Tv = T.permute(0, 1, 2, 4, 3)
Gk = torch.zeros((1, 1024, 1024), dtype=torch.complex128, device=dev)
for _ in range(2):
    p.contract(Tv.permute(0, 4, 1, 2, 3), (1, 1, 3), Tv, (1, 3, 1), Gk, (1, 1, 1), conjA=True, acc64=True, hermitian=True)
torch.cuda.synchronize()
U = (torch.randn(1, 64, 512, device=dev) + 1j * torch.randn(1, 64, 512, device=dev)).to(torch.complex64)
Th = (torch.randn(1, 512, 2, 4, 64, device=dev) + 1j * torch.randn(1, 512, 2, 4, 64, device=dev)).to(torch.complex64)
out = torch.empty((1, 64, 2, 4, 64), dtype=torch.complex64, device=dev)
for _ in range(2):
    p.contract(U, (1, 1, 1), Th, (1, 1, 3), out, (1, 1, 3))
torch.cuda.synchronize()
print('done')
