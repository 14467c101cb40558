"""Dev helper (run on the GPU box): cProfile of one steady-state cfg2 layer, host side."""
import cProfile
import pstats
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=6)
circs = []
for d in range(5):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI,
                                kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles)
    circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(4):
    circs[d].evolve(state)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
circs[4].evolve(state)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(45)
