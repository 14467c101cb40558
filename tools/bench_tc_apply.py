"""Dev helper (GPU): the tcgen05 / TMA apply tile against the FFMA tiles it replaces, on the apply shapes of the
chi = 64 / 128 / 256 configurations. Prints ms and algorithmic TFLOP/s (8 M N K) for both, CUDA events, 20 repeats."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
dev = 'cuda:0'
C64 = torch.complex64
def rnd(*shape):
    return torch.complex(torch.randn(*shape, device=dev), torch.randn(*shape, device=dev))
def timeit(fn, reps=20):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rows = []
for (M, K, N) in [(8192, 64, 64), (32768, 128, 128), (65536, 256, 256), (131072, 256, 256), (65536, 512, 256), (65536, 1024, 256),
                  (262144, 64, 64), (97280, 128, 256)]:
    A, B = rnd(1, M, K), rnd(1, K, N)
    out = torch.empty((1, M, N), dtype=C64, device=dev)
    f = lambda: p.contract(A, (1, 1, 1), B, (1, 1, 1), out, (1, 1, 1), conjB=True)
    p.lib.mpdo_tc_enable(1); t_tc = timeit(f)
    p.lib.mpdo_tc_enable(2); t_fast = timeit(f)
    p.lib.mpdo_tc_enable(0); t_ff = timeit(f)
    p.lib.mpdo_tc_enable(1)
    fl = 8.0 * M * K * N
    by = 8.0 * (M * K + K * N + M * N)
    rows.append({'M': M, 'K': K, 'N': N, 'tc_ms': t_tc, 'tc_TFLOPs': fl / t_tc / 1e9, 'tc_fast_ms': t_fast,
                 'tc_fast_TFLOPs': fl / t_fast / 1e9, 'ffma_ms': t_ff, 'ffma_TFLOPs': fl / t_ff / 1e9,
                 'speedup': t_ff / t_tc, 'speedup_fast': t_ff / t_fast, 'tc_GBs': by / t_tc / 1e6})
    print('M=%7d K=%5d N=%4d  tcgen05 %7.3f ms %6.1f TFLOP/s %5.0f GB/s x%.2f | wide tiles %7.3f ms %6.1f TFLOP/s x%.2f | FFMA %7.3f ms %5.1f TFLOP/s' %
          (M, K, N, t_tc, fl / t_tc / 1e9, by / t_tc / 1e6, t_ff / t_tc, t_fast, fl / t_fast / 1e9, t_ff / t_fast, t_ff, fl / t_ff / 1e9), flush=True)
print(json.dumps(rows))
