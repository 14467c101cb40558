"""Dev helper (GPU): find where a batched cfg4 run first produces non-finite numbers."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_configs as bc
import MPDOSimulator as Simulator
from MPDOSimulator import _engine
n, depth, B = 16, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 128
ang = bc.angles(list(range(B)), bc.n_draws(n, depth, 'cz'))
st = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
k = 0
eng = _engine.engine_for(torch.complex64)
for d in range(depth):
    for phase in (0, 1):
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=64, kappa=4, chip='medium', dtype=torch.complex64, device='cuda:0')
        if phase == 0:
            for q in range(n):
                c.u3(ang[k].clone(), ang[k+1].clone(), ang[k+2].clone(), [q]); k += 3
        else:
            for q in range(d % 2, n - 1, 2):
                c.cz(q, q + 1)
        c.truncate()
        c.evolve(st)
        bad = [(i, (~torch.isfinite(torch.view_as_real(s.data)).reshape(s.data.shape[0], -1).all(dim=1)).nonzero().reshape(-1).tolist()) for i, s in enumerate(st)]
        bad = [(i, b) for i, b in bad if b]
        tr = eng.chain_value(c._Ts()).real
        print(d, phase, 'shapes', [tuple(s.data.shape[1:]) for s in st][:4], 'trace min/max %.3e %.3e' % (tr.min().item(), tr.max().item()), 'bad', bad[:3], 'topk', eng.stats.get('topk_iters'), flush=True)
        if bad:
            sys.exit(0)
