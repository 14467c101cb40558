"""Dev helper (GPU): where the batched configuration (cfg4: 128 circuits x 16 qubits x depth 16 as one batch) spends
its device time - per-kernel table from the torch profiler plus the library's own per-class launch timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
import bench_configs as bc
import MPDOSimulator as Simulator
from MPDOSimulator import _engine

B = int(os.environ.get('CFG4_B', '128'))
n, depth = bench.CFG4['n'], bench.CFG4['depth']
ang = bc.angles(list(range(B)), bc.n_draws(n, depth, 'cz'))
def build():
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=bench.CFG4['chi'], kappa=bench.CFG4['kappa'], chip='medium', dtype=torch.complex64, device='cuda:0')
    bc.brickwork(c, n, depth, ang, 'cz')
    return c
def run(c):
    st = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64, device='cpu')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c.evolve(st)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
print('rehearsal %.2f s' % run(build()))
c = build()
print('evolve %.2f s' % run(c))
lib = _engine.get_prims().lib
lib.mpdo_timing_enable(1)
c = build()
print('evolve with per-launch events %.2f s' % run(c))
names = {0: 'contract fp32 FFMA', 3: 'contract fp64-acc DMMA', 4: 'tcgen05 apply', 1: 'jacobi', 2: 'cholesky'}
for cls, nm in names.items():
    sec, fl, by, mxs, mxf = (C.c_double() for _ in range(5)); cnt = C.c_int64()
    lib.mpdo_timing_summary(cls, 0.0, C.byref(sec), C.byref(fl), C.byref(by), C.byref(cnt), C.byref(mxs), C.byref(mxf))
    print('%-24s launches %6d  seconds %7.3f  TFLOP/s %7.2f  largest %.1f GFLOP in %.2f ms' % (nm, cnt.value, sec.value, fl.value / max(sec.value, 1e-9) / 1e12, mxf.value / 1e9, 1e3 * mxs.value))
lib.mpdo_timing_enable(0)
from torch.profiler import ProfilerActivity, profile
c = build()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t = run(c)
print('evolve under the profiler %.2f s' % t)
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=80))
