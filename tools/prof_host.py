"""Dev helper (GPU): host-issue time against device time of steady-state cfg2 layers, phase by phase.
For layers 11 and 12 (odd / even brick pattern): wall time of evolve() up to the point where the host has issued
everything (no synchronize), total time with a synchronize, and the same split for the gate phase, the QR sweep,
the chi sweep and the kappa truncation (a synchronize between phases; the phases then cannot overlap, so the sum
is an upper bound of the layer)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator
from MPDOSimulator import Circuit as CircuitMod, TNNOptimizer as Opt

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=16)
t0 = time.perf_counter()
circs = []
for d in range(15):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles); circs.append(c)
print('construction of 15 layer circuits: %.1f ms each' % (1e3 * (time.perf_counter() - t0) / 15))
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(11):
    circs[d].evolve(state)
torch.cuda.synchronize()
snap = [s.data.clone() for s in state]

def restore():
    for s, x in zip(state, snap):
        s.data = x.clone()
    torch.cuda.synchronize()

for rep in range(2):
    restore()
    for d in (11, 12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        circs[d].evolve(state)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print('layer %d: host issue %.1f ms, total %.1f ms' % (d, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))

# phase split
marks = []
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        marks.append((name, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
        return r
    return w
Simulator.TensorCircuit._run_segment = timed('gates', Simulator.TensorCircuit._run_segment)
Opt.qr_left2right = timed('qr sweep', Opt.qr_left2right)
Opt.svd_right2left = timed('chi sweep', Opt.svd_right2left)
Opt.svdKappa_left2right = timed('kappa', Opt.svdKappa_left2right)
CircuitMod.svdKappa_left2right = Opt.svdKappa_left2right
restore()
for d in (11, 12):
    marks.clear()
    circs[d].evolve(state)
    print('layer %d phases (host issue / total ms):' % d, ', '.join('%s %.1f / %.1f' % m for m in marks if m[2] > 0.05))

# per library call, serialised (a synchronize on both sides of every step call of the truncation sweeps)
from MPDOSimulator import _engine
eng = _engine.engine_for(torch.complex64)
acc = {}
orig_call = eng._call
def call(what, fn, *a):
    if what in ('mpdo_split_2q', 'mpdo_kappa_truncate'):
        return orig_call(what, fn, *a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig_call(what, fn, *a)
    torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
    acc.setdefault(what, []).append(dt)
    return r
eng._call = call
restore()
for d in (11, 12):
    acc.clear()
    circs[d].evolve(state)
    print('layer %d sweep calls:' % d, {k: (len(v), round(sum(v), 2), [round(x, 2) for x in v[:24]]) for k, v in acc.items()})
