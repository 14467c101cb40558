"""Dev helper (GPU): a few eigen-decompositions of a captured n = 256 gate-split Gram matrix (cluster Jacobi) for an
ncu source-level capture."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import numpy as np
import torch
from MPDOSimulator._engine.prims import CudaPrims
p = CudaPrims()
data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'grams_sel.npz'))
G = torch.from_numpy(data[os.environ.get('JKEY', 'g001_n256_rr1')]).to('cuda:0').unsqueeze(0).contiguous()
for _ in range(3):
    p.eigh_psd(G, 1e-10, rank_revealing=True)
torch.cuda.synchronize()
print('done')
