"""Dev helper (GPU): one launch of each kernel that matters after round 2, at its steady-state cfg2 / cfg5 size, for
`ncu --set full -k regex:"tc_apply|chol_small|chol_cluster|chol_blocked|jacobi_cluster|jacobi_kernel|contract_kernel"`.

  tc_apply_kernel<128,3> / <256,2>          : complex64 apply 65536 x 256 x 256, default and widest column tile
  chol_cluster_kernel + jacobi_cluster_kernel : preconditioned eigen-decomposition of graded 192 x 192 and 256 x 256
                                              Gram matrices (wide bond of the chi sweep / core of a gate split at chi = 64)
  chol_small_kernel (+ jacobi_kernel<16>)   : Cholesky factor with inverse and eigen-decomposition at n = 64
  contract_kernel (fp64 accumulate)         : Gram matrix of a wide chi-sweep step (200 x 8192 -> 200 x 200, Hermitian)
                                              and the environment step E.T (200 x 200 complex128 times 200 x 8192)
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
torch.manual_seed(0)
C64, C128 = torch.complex64, torch.complex128

def graded(n, decay):
    A = torch.randn(n, n, dtype=C128, device=dev)
    Q, _ = torch.linalg.qr(A)
    lam = torch.tensor([max(decay ** i, 1e-30) for i in range(n)], dtype=torch.float64, device=dev)
    return ((Q * lam.to(C128)) @ Q.mH).contiguous().unsqueeze(0)

A = torch.complex(torch.randn(1, 65536, 256, device=dev), torch.randn(1, 65536, 256, device=dev))
B = torch.complex(torch.randn(1, 256, 256, device=dev), torch.randn(1, 256, 256, device=dev))
out = torch.empty((1, 65536, 256), dtype=C64, device=dev)
for mode in (1, 2):
    p.lib.mpdo_tc_enable(mode)
    for _ in range(2):
        p.contract(A, (1, 1, 1), B, (1, 1, 1), out, (1, 1, 1), conjB=True)
    torch.cuda.synchronize()
p.lib.mpdo_tc_enable(1)
for n, decay in ((192, 0.93), (256, 0.94)):
    G = graded(n, decay)
    for _ in range(2):
        p.eigh_psd(G, 1e-10, rank_revealing=True)
    for _ in range(2):                   # the route of complex64 states: blocked Cholesky without pivoting in front
        p.eigh_psd(G, 1e-10, rank_revealing=2)
    torch.cuda.synchronize()
G64 = graded(64, 0.8)
for _ in range(2):
    p.eigh_psd(G64, 1e-10, rank_revealing=True)
    p.chol_psd(G64)                       # with the left inverse (Cholesky-QR of the sweeps)
torch.cuda.synchronize()
T = torch.complex(torch.randn(1, 200, 8192, device=dev), torch.randn(1, 200, 8192, device=dev))
Gm = torch.empty((1, 200, 200), dtype=C128, device=dev)
E = graded(200, 0.93)
X = torch.empty((1, 200, 8192), dtype=C128, device=dev)
for _ in range(2):
    p.contract(T, (1, 1, 1), T.permute(0, 2, 1), (1, 1, 1), Gm, (1, 1, 1), conjB=True, acc64=True, hermitian=True)
    p.contract(E, (1, 1, 1), T, (1, 1, 1), X, (1, 1, 1))
torch.cuda.synchronize()
# long rows: block pairs in registers, sliced by columns (jacobi_persistent_cols_kernel), order 1024 as in cfg5
G1k = graded(1024, 0.975)
for _ in range(2):
    p.eigh_psd(G1k, 1e-10, rank_revealing=True)
torch.cuda.synchronize()
print('done')
