"""Dev helper (GPU): dump the Hermitian PSD matrices that the eigen-solver sees during one steady-state cfg2 layer
(Python step engine, so that every eigh_psd call passes through the primitive wrapper) into gpurun_out/grams.npz,
for offline convergence studies of Jacobi variants."""
import os, sys
os.environ['MPDO_ENGINE'] = 'py'
os.environ['MPDO_STRANDS'] = '0'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import MPDOSimulator as Simulator
from MPDOSimulator._engine.prims import CudaPrims

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=14)
circs = []
for d in range(12):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI,
                                kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles)
    circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(11):
    circs[d].evolve(state)

dump = {}
orig = CudaPrims.eigh_psd
def hook(self, G, *a, **k):
    if G.shape[-1] >= 48 and G.shape[0] == 1:
        key = 'g%03d_n%d_rr%d' % (len(dump), G.shape[-1], int(k.get('rank_revealing', a[2] if len(a) > 2 else False)))
        dump[key] = G[0].detach().cpu().numpy()
    return orig(self, G, *a, **k)
CudaPrims.eigh_psd = hook
circs[11].evolve(state)
CudaPrims.eigh_psd = orig
os.makedirs('gpurun_out', exist_ok=True)
# keep it small: at most 4 matrices per order
kept, per = {}, {}
for k, v in dump.items():
    nn = v.shape[0]
    per[nn] = per.get(nn, 0) + 1
    if per[nn] <= 3:
        kept[k] = v
np.savez_compressed('gpurun_out/grams.npz', **kept)
print('calls', len(dump), 'kept', len(kept), sorted(per.items()))
