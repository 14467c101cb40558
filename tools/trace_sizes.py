"""Dev helper (GPU): histogram of Jacobi problem sizes in one steady-state cfg2 layer (MPDO_TRACE=1)."""
import os, sys, subprocess, collections, re
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, os; sys.path.insert(0, %r)
import torch, bench, MPDOSimulator as S
n=bench.N_QUBITS
files={'CZ':{f'{i}{i+1}':bench.chi_file() for i in range(n-1)},'CP':{}}
ang=bench.layer_angles(0,depth=12)
st=S.Tools.create_ket0Series(n,dtype=torch.complex64)
for d in range(11):
    c=S.TensorCircuit(qn=n,ideal=False,noiseType='realNoise',chiFileDict=files,chi=64,kappa=4,chip='best',dtype=torch.complex64,device='cuda:0')
    bench.add_layer(c,d,ang)
    if d==10: sys.stderr.write('[mpdo] LAYER10\\n'); sys.stderr.flush()
    c.evolve(st)
torch.cuda.synchronize()
''' % root
env = dict(os.environ, MPDO_TRACE='1', MPDO_STRANDS='0')
out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True).stderr
lines = out.split('[mpdo] LAYER10')[-1].splitlines()
hist = collections.Counter()
for l in lines:
    m = re.search(r'jacobi n=(\d+) m=(\d+) mt=(\d+)', l)
    if m: hist[(int(m.group(1)), int(m.group(2)))] += 1
tot = 0
for (n, m), c in sorted(hist.items()):
    cost = c * n * n * (16 * m + 24 * (m + n)) * 5e-9   # ~10 sweeps x n^2/2 pairs, GFLOP
    tot += cost
    print('n=%4d m=%4d count=%4d  ~%.1f GFLOP' % (n, m, c, cost))
print('total ~%.0f GFLOP' % tot)
topk = collections.Counter()
for l in lines:
    m = re.search(r'eigh_topk n=(\d+) k=(\d+) blk=(\d+) B=(\d+) iterations=(\d+)', l)
    if m: topk[tuple(int(x) for x in m.groups())] += 1
for key, c in sorted(topk.items()):
    print('eigh_topk n=%d k=%d blk=%d B=%d iterations=%d: count %d' % (key + (c,)))
