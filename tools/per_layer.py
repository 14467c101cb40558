"""Dev helper (GPU): wall time and bond dimensions of every layer of the depth-20 cfg2 circuit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator
n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=20)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
tot = 0
for d in range(20):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    upd = bench.add_layer(c, d, angles)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c.evolve(state)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; tot += dt
    print(d, '%.1f ms' % (1e3 * dt), 'updates', upd, 'bonds', [int(s.data.shape[4]) for s in state[:-1]], flush=True)
print('total %.2f s, %.1f updates/s' % (tot, 380 / tot))
