"""Dev helper (GPU): times the contraction kernel on the shapes of a steady-state cfg2 layer."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
C64, C128 = torch.complex64, torch.complex128
def rnd(*shape, dt=C64):
    real = torch.float32 if dt == C64 else torch.float64
    return torch.complex(torch.randn(*shape, device=dev, dtype=real), torch.randn(*shape, device=dev, dtype=real))
def timeit(name, fn, flops, reps=20):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%-58s %8.3f ms  %6.2f TFLOP/s' % (name, ms, flops / ms / 1e9), flush=True)
# kappa Gram: a = 1024 over (l,s,r) = 8192, Hermitian, fp64 accumulate
T = rnd(1, 64, 2, 1024, 64); Tv = T.permute(0, 1, 2, 4, 3)
G = torch.zeros((1, 1024, 1024), dtype=C128, device=dev)
timeit('kappa Gram 1024x1024x8192 c64->c128 hermitian', lambda: p.contract(Tv.permute(0, 4, 1, 2, 3), (1, 1, 3), Tv, (1, 3, 1), G, (1, 1, 1), conjA=True, acc64=True, hermitian=True), 4 * 1024 * 1024 * 8192)
timeit('kappa Gram 1024x1024x8192 c64->c128 full', lambda: p.contract(Tv.permute(0, 4, 1, 2, 3), (1, 1, 3), Tv, (1, 3, 1), G, (1, 1, 1), conjA=True, acc64=True), 8 * 1024 * 1024 * 8192)
# bond Gram rows: l = 512 over (s,a,r) = 512
M = rnd(1, 512, 2, 4, 64); G2 = torch.zeros((1, 512, 512), dtype=C128, device=dev)
timeit('bond Gram 512x512x512 c64->c128 hermitian (split-K)', lambda: p.contract(M, (1, 1, 3), M.permute(0, 2, 3, 4, 1), (1, 3, 1), G2, (1, 1, 1), conjB=True, acc64=True, hermitian=True), 4 * 512 * 512 * 512)
# split: T_hi' = core . Qt : [512 x 512] . [512 x (2*64*64)]
A = rnd(1, 512, 512, dt=C128); B = rnd(1, 512, 2, 64, 64); Cc = torch.empty((1, 512, 2, 64, 64), dtype=C64, device=dev)
timeit('apply 512x8192x512 c128.c64->c64 (fp64 acc)', lambda: p.contract(A, (1, 1, 1), B, (1, 1, 3), Cc, (1, 1, 3)), 8 * 512 * 8192 * 512)
A32 = A.to(C64)
timeit('apply 512x8192x512 c64.c64->c64 (fp32)', lambda: p.contract(A32, (1, 1, 1), B, (1, 1, 3), Cc, (1, 1, 3)), 8 * 512 * 8192 * 512)
# kappa projection: Vk [4 x 1024] . T[1024 x 8192]
Vk = rnd(1, 4, 1024); out = torch.empty((1, 64, 2, 4, 64), dtype=C64, device=dev)
timeit('kappa project 4x8192x1024 c64 (fp32)', lambda: p.contract(Vk, (1, 1, 1), T.permute(0, 3, 1, 2, 4), (1, 1, 3), out.permute(0, 3, 1, 2, 4), (1, 1, 3)), 8 * 4 * 8192 * 1024)
# small c128 cores
X = rnd(1, 32, 1024, dt=C128); Gt = rnd(1, 1024, 1024, dt=C128); Z = torch.empty((1, 32, 1024), dtype=C128, device=dev)
timeit('top-k block 32x1024x1024 c128', lambda: p.contract(X, (1, 1, 1), Gt, (1, 1, 1), Z, (1, 1, 1)), 8 * 32 * 1024 * 1024)
