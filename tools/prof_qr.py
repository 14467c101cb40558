"""Dev helper (GPU): per-kernel device time inside repeated qr_step calls on a 48x2x4x48 site."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from torch.profiler import ProfilerActivity, profile
from MPDOSimulator._engine.prims import CudaPrims
from MPDOSimulator._engine.native import NativeEngine
dev = 'cuda:0'
E = NativeEngine(CudaPrims(), torch.complex64)
g = torch.Generator().manual_seed(7)
def gauss(*shape):
    return (torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)) / math.sqrt(shape[1] * shape[3] * 2)).to(dev)
for n in (48, 24, 96):
    T1, T2 = gauss(1, n, 2, 4, n), gauss(1, n, 2, 4, n)
    E.qr_step(T1, T2); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            E.qr_step(T1, T2)
        torch.cuda.synchronize()
    print('n =', n)
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=8, max_name_column_width=60))
