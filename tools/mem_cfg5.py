"""Dev helper (GPU): where the device memory of a cfg5-shaped circuit (chi = 256, kappa = 8, idealNoise cz brickwork)
goes. Wraps the phases of a layer and prints torch's live / peak bytes per phase.
    N=24 DEPTH=12 python tools/mem_cfg5.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tomography-assisted-mpdo-qcircuit_b200')]
import torch  # noqa: E402
import bench_configs as bc  # noqa: E402
import MPDOSimulator as Simulator  # noqa: E402
from MPDOSimulator import Circuit as Cm, TNNOptimizer as Tm  # noqa: E402

n, depth = int(os.environ.get('N', 24)), int(os.environ.get('DEPTH', 12))
GB = 2.0 ** 30
rows = []


def wrap(mod, name, label):
    fn = getattr(mod, name)

    def inner(*a, **k):
        torch.cuda.reset_peak_memory_stats()
        before = torch.cuda.memory_allocated()
        out = fn(*a, **k)
        rows.append((label, before / GB, torch.cuda.memory_allocated() / GB, torch.cuda.max_memory_allocated() / GB,
                     torch.cuda.memory_reserved() / GB))
        return out
    setattr(mod, name, inner)


wrap(Cm.TensorCircuit, '_run_segment', 'segment')
wrap(Tm, 'bondTruncate', 'bondTruncate')
wrap(Tm, 'svdKappa_left2right', 'kappa')
ang = bc.angles([0], bc.n_draws(n, depth, 'cz'))
c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chiFileDict=None, chi=256, kappa=8,
                            chip='medium', dtype=torch.complex64, device='cuda:0')
bc.brickwork(c, n, depth, ang, 'cz')
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64, device='cpu')
t0 = time.perf_counter()
c.evolve(state)
torch.cuda.synchronize()
print('seconds', round(time.perf_counter() - t0, 2), 'knobs', {k: v for k, v in os.environ.items() if k.startswith('MPDO_')})
print('max bond', max(int(s.data.shape[4]) for s in state), 'split ranks (last layer)',
      sorted(c.last_stats.get('split_ranks', {}).values())[-6:])
print('%-14s %8s %8s %8s %8s' % ('phase', 'before', 'after', 'peak', 'reserved'))
for r in (rows[-8:] if not os.environ.get('ALLROWS') else [max(rows[i:i+6], key=lambda r: r[3]) for i in range(0, len(rows), 6)]):
    print('%-14s %8.2f %8.2f %8.2f %8.2f' % r)
print('site shapes', [tuple(s.data.shape) for s in state][:6])
print('overall peak of any phase %.2f GB' % max(r[3] for r in rows))
free_b, tot = torch.cuda.mem_get_info()
print('device in use %.2f GB' % ((tot - free_b) / GB))
