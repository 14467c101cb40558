"""Dev helper (GPU): replays one steady-state cfg2 layer many times from the same state with the wall time of
each phase (gate strands / QR sweep / SVD sweep / kappa) recorded, to find which phase the occasional slow step
comes from."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import MPDOSimulator as Simulator
from MPDOSimulator import Circuit as CM, TNNOptimizer as TO

acc = {}
def timed(mod, name):
    fn = getattr(mod, name)
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, wrap)
for nm in ('qr_left2right', 'svd_right2left', 'svdKappa_left2right'):
    timed(TO, nm)
    if hasattr(CM, nm):
        setattr(CM, nm, getattr(TO, nm))
orig_seg = CM.TensorCircuit._run_segment
def seg(self, state, segment):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig_seg(self, state, segment)
    torch.cuda.synchronize(); acc['gates'] = acc.get('gates', 0.0) + time.perf_counter() - t0
CM.TensorCircuit._run_segment = seg

n = bench.N_QUBITS
files = {'CZ': {f'{i}{i + 1}': bench.chi_file() for i in range(n - 1)}, 'CP': {}}
angles = bench.layer_angles(0, depth=14)
circs = []
for d in range(12):
    c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=bench.CHI, kappa=bench.KAPPA, chip='best', dtype=torch.complex64, device='cuda:0')
    bench.add_layer(c, d, angles); circs.append(c)
state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64)
for d in range(11):
    circs[d].evolve(state)
torch.cuda.synchronize()
snap = [s.data.clone() for s in state]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
import ctypes as C
from MPDOSimulator import _engine
lib = _engine.get_prims().lib
def largest(cls):
    sec, fl, by, mxs, mxf = (C.c_double() for _ in range(5))
    cnt = C.c_int64()
    lib.mpdo_timing_summary(cls, 0.0, C.byref(sec), C.byref(fl), C.byref(by), C.byref(cnt), C.byref(mxs), C.byref(mxf))
    return round(1e3 * mxs.value, 1), round(1e3 * sec.value, 1), cnt.value
TIMING = os.environ.get('OUTLIER_TIMING', '0') == '1'
import threading, collections, traceback
samples = collections.Counter()
sampling = {'on': False, 'stop': False}
def sampler():
    me = threading.get_ident()
    names = {}
    while not sampling['stop']:
        if sampling['on']:
            for t in threading.enumerate():
                names[t.ident] = t.name
            for tid, fr in sys._current_frames().items():
                if tid == me:
                    continue
                st = traceback.extract_stack(fr)[-3:]
                key = (names.get(tid, '?')[:12], ' < '.join('%s:%d:%s' % (os.path.basename(f.filename), f.lineno, f.name) for f in reversed(st)))
                samples[key] += 1
        time.sleep(0.005)
threading.Thread(target=sampler, daemon=True).start()
for it in range(reps):
    for s, x in zip(state, snap):
        s.data = x.clone()
    acc.clear()
    if TIMING:
        lib.mpdo_timing_enable(1)
    samples.clear(); sampling['on'] = True
    ms0 = torch.cuda.memory_stats()
    pn0, ph0 = C.c_int64(), C.c_int64(); lib.mpdo_pool_stats(C.byref(pn0), C.byref(ph0))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    circs[11].evolve(state)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    sampling['on'] = False
    if tot > 0.055 and it > 1:
        agg = collections.Counter()
        for (nm, stk), c in samples.items():
            agg[stk] += c
        print('  OUTLIER stack samples (all threads):')
        for stk, c in agg.most_common(14):
            print('   %5d  %s' % (c, stk))
    extra = ''
    if TIMING:
        extra = ' largest/total/launches: contract %s jacobi %s chol %s' % (largest(0), largest(1), largest(2))
        lib.mpdo_timing_enable(0)
    ms1 = torch.cuda.memory_stats()
    pn1, ph1 = C.c_int64(), C.c_int64(); lib.mpdo_pool_stats(C.byref(pn1), C.byref(ph1))
    dm = {k: ms1[k] - ms0[k] for k in ('num_device_alloc', 'num_device_free', 'num_alloc_retries', 'reserved_bytes.all.current') if ms1[k] != ms0[k]}
    print(it, 'total %.1f' % (1e3 * tot), {k: round(1e3 * v, 1) for k, v in acc.items()}, 'torch', dm, 'pool_MB', (pn1.value - pn0.value) >> 20, extra, flush=True)
