"""Dev helper (GPU): which chunk / readout column of the batched cfg4 run is non-finite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_configs as bc
import MPDOSimulator as Simulator
from MPDOSimulator import dmOperations
n, depth, chunk = 16, 16, 128
for start in range(0, 1024, chunk):
    ids = list(range(start, start + chunk))
    ang = bc.angles(ids, bc.n_draws(n, depth, 'cz'))
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=64, kappa=4, chip='medium', dtype=torch.complex64, device='cuda:0')
    bc.brickwork(c, n, depth, ang, 'cz')
    st = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
    c.evolve(st)
    badstate = [i for i, s in enumerate(st) if not torch.isfinite(torch.view_as_real(s.data)).all()]
    dmn = c.cal_dmNodes()
    z = torch.stack([dmOperations.pauli_expect(dmn, 2, q) for q in range(n)], 1)
    zz = torch.stack([dmOperations.pauli_expect(dmn, [2, 2], [q, q + 1]) for q in range(n - 1)], 1)
    p0 = c._engine().chain_value_proj(c._Ts(), [0] * n)
    tr = dmOperations.trace_rho(dmn)
    print(start, 'bad state sites', badstate, 'z finite', bool(torch.isfinite(z).all()), 'zz', bool(torch.isfinite(zz).all()), 'p0', bool(torch.isfinite(p0).all()), 'trace range %.3e %.3e' % (tr.min().item(), tr.max().item()), 'bad circuits', (~torch.isfinite(z).all(dim=1)).nonzero().reshape(-1).tolist()[:5], flush=True)
