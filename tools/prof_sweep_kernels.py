"""Dev helper (GPU): per-kernel device time of the gate phase, the QR sweep and the SVD sweep of one steady-state
cfg2 layer (strands off, torch profiler)."""
import os, sys
os.environ['MPDO_STRANDS'] = '0'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
import MPDOSimulator as Simulator
from MPDOSimulator import TNNOptimizer as TO, Circuit as CM

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=14)
circs = []
for d in range(13):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles); circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(11):
    circs[d].evolve(state)
c = circs[11]
layers = list(c.layers)
seg = [(i, l, c._oqs_list[i]) for i, l in enumerate(layers) if 'truncate' not in l.name.lower()]
def prof(name, fn):
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        fn(); torch.cuda.synchronize()
    print('=====', name)
    print(p.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
c.last_stats = {}
prof('gates', lambda: c._run_segment(state, seg))
prof('qr sweep', lambda: TO.qr_left2right(state))
prof('svd sweep', lambda: TO.svd_right2left(state, max_singular_values=bench.CHI))
prof('kappa', lambda: TO.svdKappa_left2right(state, max_singular_values=bench.KAPPA))
