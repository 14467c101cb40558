"""Dev helper (GPU): wall time of the phases of one steady-state cfg2 layer (gates / QR sweep / SVD sweep / kappa)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator
from MPDOSimulator import Circuit as CM, TNNOptimizer as TO

acc = {}
def timed(mod, name):
    fn = getattr(mod, name)
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, wrap)
timed(TO, 'qr_left2right'); timed(TO, 'svd_right2left')
CM.svdKappa_left2right = None
timed(TO, 'svdKappa_left2right'); CM.svdKappa_left2right = TO.svdKappa_left2right
orig_seg = CM.TensorCircuit._run_segment
def seg(self, state, segment):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig_seg(self, state, segment)
    torch.cuda.synchronize(); acc['gates'] = acc.get('gates', 0.0) + time.perf_counter() - t0
CM.TensorCircuit._run_segment = seg

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=14)
circs = []
for d in range(13):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles); circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(10):
    circs[d].evolve(state)
acc.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
for d in range(10, 12):
    circs[d].evolve(state)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print('steady-state per layer ms:', {k: round(1e3 * v / 2, 2) for k, v in acc.items()}, 'total', round(1e3 * tot / 2, 2))
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    circs[12].evolve(state)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=12, max_name_column_width=70))
