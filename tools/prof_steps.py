"""Dev helper (GPU): time the four step functions of both engines on steady-state cfg2-like tensors."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
from MPDOSimulator._engine.steps import Engine
from MPDOSimulator._engine.native import NativeEngine
from MPDOSimulator.RealNoise import czExp_channel
import bench
dev = 'cuda:0'
p = CudaPrims()
g = torch.Generator().manual_seed(7)
def gauss(*shape):
    return (torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)) / math.sqrt(shape[1] * shape[3] * 2)).to(dev)
G = czExp_channel(filename=bench.chi_file()).unsqueeze(0).to(dev).contiguous()
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
for name, E in (('py', Engine(p, torch.complex64)), ('native', NativeEngine(p, torch.complex64))):
    T1, T2 = gauss(1, 48, 2, 4, 48), gauss(1, 48, 2, 4, 48)
    Tbig = gauss(1, 48, 2, 1024, 48)
    Tmid = gauss(1, 48, 2, 4, 64)
    print(name, 'qr_step %.3f ms' % timeit(lambda: E.qr_step(T1, T2)),
          'bond_svd %.3f ms' % timeit(lambda: E.bond_svd_step(T1, T2, 64)),
          'kappa(1024->4) %.3f ms' % timeit(lambda: E.kappa_truncate(Tbig, 4)),
          'split_2q(K=16) %.3f ms' % timeit(lambda: E.split_2q(T1, T2, G), 10), flush=True)
