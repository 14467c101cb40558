"""Dev helper (GPU): times the pivoted Cholesky (with and without the left inverse) and the preconditioned
eigen-solver at the orders the sweeps use."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
torch.manual_seed(0)
def timeit(name, fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print('%-50s %8.3f ms' % (name, e0.elapsed_time(e1) / reps), flush=True)
for n, Bn in [(128, 1), (160, 1), (256, 1), (400, 1), (512, 1), (256, 64)]:
    A = torch.randn(Bn, n, n, dtype=torch.complex128, device=dev)
    Q, _ = torch.linalg.qr(A)
    lam = torch.tensor([max(0.93 ** i, 1e-30) for i in range(n)], dtype=torch.float64, device=dev)
    G = ((Q * lam.to(torch.complex128)) @ Q.mH).contiguous()
    timeit('chol_psd (+inverse) n=%d batch=%d' % (n, Bn), lambda: p.chol_psd(G, rel=1e-15))
    timeit('eigh_psd preconditioned n=%d batch=%d' % (n, Bn), lambda: p.eigh_psd(G, 1e-10, rank_revealing=True))
