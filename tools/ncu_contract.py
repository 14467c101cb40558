"""Dev helper (GPU): one launch of the fp64-accumulated contraction at the kappa-Gram and core-apply shapes, for
`ncu --set full -k regex:contract_kernel` (MPDO_DMMA_ALL=1 / MPDO_NO_DMMA=1 select the tile flavour)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
C64, C128 = torch.complex64, torch.complex128
torch.manual_seed(0)
def rnd(*shape, dt=C64):
    real = torch.float32 if dt == C64 else torch.float64
    return torch.complex(torch.randn(*shape, device=dev, dtype=real), torch.randn(*shape, device=dev, dtype=real))
T = rnd(1, 64, 2, 1024, 64); Tv = T.permute(0, 1, 2, 4, 3)
Gk = torch.zeros((1, 1024, 1024), dtype=C128, device=dev)
A = rnd(1, 512, 512, dt=C128); B = rnd(1, 512, 2, 64, 64); Cc = torch.empty((1, 512, 2, 64, 64), dtype=C64, device=dev)
for _ in range(2):
    p.contract(Tv.permute(0, 4, 1, 2, 3), (1, 1, 3), Tv, (1, 3, 1), Gk, (1, 1, 1), conjA=True, acc64=True, hermitian=True)
    p.contract(A, (1, 1, 1), B, (1, 1, 3), Cc, (1, 1, 3))
torch.cuda.synchronize()
print('done')
