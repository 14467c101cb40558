"""Dev helper (GPU): serial latency of every step call (split_2q, qr_step, bond_svd_step, kappa_truncate) of one
steady-state cfg2 layer, with shapes - strands disabled, a synchronize around every call. Shows what the sequential
sweeps and the per-pair chains cost when nothing overlaps."""
import os, sys, time, collections
os.environ['MPDO_STRANDS'] = '0'
os.environ['MPDO_GROUPING'] = os.environ.get('MPDO_GROUPING', '0')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator
from MPDOSimulator import _engine

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=14)
LAYER = int(os.environ.get('LAYER', '11'))
circs = []
for d in range(13):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles); circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(LAYER):
    circs[d].evolve(state)
eng = _engine.engine_for(torch.complex64)
log = []
def wrap(name):
    fn = getattr(eng, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
        shp = [tuple(x.shape) for x in a if hasattr(x, 'shape')]
        out = [tuple(x.shape) for x in r if hasattr(x, 'shape')] if isinstance(r, tuple) else []
        log.append((name, dt, shp, out))
        return r
    setattr(eng, name, w)
for nm in ('split_2q', 'qr_step', 'bond_svd_step', 'kappa_truncate', 'absorb_1q'):
    wrap(nm)
torch.cuda.synchronize(); t0 = time.perf_counter()
circs[LAYER].evolve(state)
torch.cuda.synchronize(); tot = 1e3 * (time.perf_counter() - t0)
agg = collections.OrderedDict()
for name, dt, shp, out in log:
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += dt
print('serial layer total %.1f ms' % tot, {k: (v[0], round(v[1], 2)) for k, v in agg.items()})
for name, dt, shp, out in log:
    if name != 'absorb_1q':
        print('%-15s %7.3f ms  in %s -> %s' % (name, dt, shp, out[:2]))
