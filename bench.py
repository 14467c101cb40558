#!/usr/bin/env python
"""bench.py - noisy two-qubit-gate updates/sec of the MPDO update path on B200.

Workload (BASELINE.json configs[1], SURVEY 8d "cfg2"): 20-qubit brickwork, complex64, chi = 64, kappa = 4,
noiseType = 'realNoise' with the tomography chi-matrix czDefault.mat on every bond. Layer d =
[u3 on every qubit ; rzz(theta, q, q+1) on bonds q = d mod 2 ; truncate()], and every rzz is rewritten by the
reference API into cx.rz.cx = two chi-matrix CZ updates (K = 16 Kraus terms each) with no truncation in
between. Angles come from torch.Generator().manual_seed(1234 + circuit_id).

A "step" is one such layer on one circuit per GPU. The synthetic input of the bench is the MPDO state after the
first PREROLL = 8 layers of the circuit (bonds saturated at chi; early layers are 5-10x cheaper and would
flatter the number), then W warm-up layers, then exactly K layers are timed with CUDA events between barrier+synchronize pairs (max over ranks). A single MPDO sweep is
sequential and does not shard, so --gpus N runs N independent circuits (replicas; circuit_id = rank), NCCL
only gathers the per-circuit readout at the end of the timed region.

  value  : noisy 2q updates / s, state resident in HBM when the clock starts
  e2e    : same layers through the public API from pinned HOST buffers (H2D of the state, evolve, D2H of the
           state) inside the timed region
  roofline / cpu_baseline: see DESIGN.md section 5.

  value_one_split_per_gate: the same K layers with MPDO_NO_FUSE=1 (one split per chi-matrix CZ, as the reference does;
           `value` fuses the two CZs of an rzz into one split - a documented numerical deviation)
  cfg4   : the batched configuration of BASELINE.json (configs[3]) on every N: circuits sharded `id mod world`
           (MPDOSimulator/_engine/sharding.py), CFG4_PER_RANK circuits per rank evolved as one batch, the 32-wide
           readout rows (16 <Z>, 15 <ZZ>, P(0...0)) gathered with one NCCL all-gather INSIDE the timed region;
           circuits/s device-resident and end to end (angles from pinned host memory, table back to the host)

`--impl reference` times the reference's own formulation on the host cores: the oracle in reference mode
(torch CPU, all threads) on a bounded sample of the same workload - two rzz (four chi-matrix CZ updates) on a
steady-state six-site window plus the QR / chi-SVD / kappa-SVD sweep of the window (5 bonds for 4 updates; the real
layer sweeps 19 bonds for 18-20 updates). --warmup repeats run untimed first. For N > 1 the driver's per-N ratio
divides N GPUs by this ONE CPU process.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'tomography-assisted-mpdo-qcircuit_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')   # see MPDOSimulator/__init__.py (must precede CUDA initialisation)

import torch  # noqa: E402

N_QUBITS, CHI, KAPPA, DEPTH = 20, 64, 4, 20
PREROLL = 8   # layers evolved before warm-up so that the timed layers see saturated bonds (chi everywhere)
METRIC = 'noisy_2q_gate_updates_per_sec'
UNIT = 'updates/s'
WORKLOAD = 'cfg2: 20q U3+RZZ brickwork depth 20, realNoise czDefault chi-matrix (K=16), chi=64, kappa=4, complex64'


def chi_file():
    return os.path.join(PKG, 'MPDOSimulator', 'chi', 'czDefault.mat')


def layer_angles(circuit_id, depth=DEPTH, n=N_QUBITS):
    """Layer-major, qubit-minor Python floats: 3 per u3, then one per rzz of the layer."""
    g = torch.Generator().manual_seed(1234 + circuit_id)
    out = []
    for d in range(depth):
        u = [(torch.rand(3, generator=g, dtype=torch.float64) * 2 * math.pi).tolist() for _ in range(n)]
        z = [float(torch.rand(1, generator=g, dtype=torch.float64) * 2 * math.pi) for _ in range(d % 2, n - 1, 2)]
        out.append((u, z))
    return out


def add_layer(circ, d, angles, n=N_QUBITS):
    u, z = angles[d]
    for q in range(n):
        circ.u3(u[q][0], u[q][1], u[q][2], [q])
    for i, q in enumerate(range(d % 2, n - 1, 2)):
        circ.rzz(z[i], q, q + 1)
    circ.truncate()
    return 2 * len(z)   # noisy 2q updates in this layer


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode, read by a thread. It is started well before the timed region (the first nvidia-smi /
    NVML start-up of a fresh machine takes the driver's locks for a while: measured as 70-210 ms steps in the first
    bench process of a box when it was started at the edge of the timed region) and keeps running through it; only
    the rows read between begin() and stop() are reported."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        t_end = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        rows = [r for t, r in self.rows if t0 <= t <= t_end + 0.25]
        if not rows and self.rows:      # a timed region shorter than the sampling period: the sample nearest to it
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + t_end)))[1]]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle in reference mode on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_sample(repeats=1, threads=None, window=6, warmup=0):
    """`window` = 4: one rzz (= 2 chi-matrix CZ updates) on sites (1, 2) of a four-site steady-state window;
    `window` = 6: rzz on (1, 2) and (3, 4) (4 updates, 5 swept bonds - the sweep-per-update ratio of the real layer
    within 25 %). Gaussian site tensors (re, im ~ N(0,1)/sqrt(2 chi kappa), seed 7; SURVEY 8d micro-benchmark)
    followed by the truncate sweep of the window, oracle in reference mode (randomized SVD branch as in
    decompositions.py:112-115). Returns (updates per repeat, seconds per timed repeat, cores)."""
    from oracle.mpdo_oracle import OracleCircuit
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    pairs = [(1, 2)] if window == 4 else [(1, 2), (3, 4)]
    files = {'CZ': {f'{a}{b}': chi_file() for a, b in pairs}, 'CP': {}}
    times = []
    for rep in range(warmup + repeats):
        oc = OracleCircuit(window, ideal=False, noiseType='realNoise', chiFileDict=files, chi=CHI, kappa=KAPPA,
                           chip='best', dtype=torch.complex64, svd_mode='reference')
        for a, b in pairs:
            oc.rzz(0.7 + 0.1 * rep + 0.05 * a, a, b)
        oc.truncate()
        g = torch.Generator().manual_seed(7)
        scale = 1.0 / math.sqrt(2 * CHI * KAPPA)

        def gauss(l, r):
            return torch.complex(torch.randn(l, 2, KAPPA, r, generator=g), torch.randn(l, 2, KAPPA, r, generator=g)) * scale

        oc.T = [gauss(1, CHI)] + [gauss(CHI, CHI) for _ in range(window - 2)] + [gauss(CHI, 1)]
        oc.bond = [True] * (window - 1)
        oc.inner = [True] * window
        t0 = time.perf_counter()
        oc.run_layers()
        if rep >= warmup:
            times.append(time.perf_counter() - t0)
    return 2 * len(pairs), times, threads


CFG4 = dict(n=16, depth=16, chi=64, kappa=4)
CFG4_PER_RANK = int(os.environ.get('MPDO_BENCH_CFG4_PER_RANK', '128'))   # = 1024 circuits / 8 GPUs
CFG4_WORKLOAD = ('cfg4: independent 16q noisy parameter-sweep circuits (U3 + CZ brickwork depth 16, idealNoise/medium: '
                 'amplitude damping + dephasing + 2q depolarizing), chi=64, kappa=4, complex64')


CFG4_CPU_SAMPLE = ('oracle (reference mode) on brickwork layers 6-7 of circuit 0 (%.1f s), a circuit counted as 8 such '
                   'layer pairs (favours the CPU: the whole circuit measured once in the build container took 171-186 s on '
                   '8 cores = 0.0056 circuits/s, 123 s of it in the layer-1 gate split on un-truncated inner indices)')


def cfg4_cpu_sample(threads=None):
    """Bounded CPU sample of the batched configuration: brickwork layers 6 and 7 (one even + one odd layer: 15 noisy
    CZ updates and 4 truncate sweeps over 16 sites) of circuit 0 in reference mode, starting from the state the
    exact-mode oracle (fast formulation) prepared with layers 0-5. A circuit is counted as 8 such layer pairs:
    circuits/s = 1 / (8 * t). That FAVOURS the CPU: the whole circuit measured once in the build container took
    171-186 s on 8 cores, of which 123 s is the single layer-1 gate split on un-truncated inner indices (the reference
    forms and decomposes the full two-site matrix there) and ~4.7 s each steady-state layer."""
    import bench_configs as bc
    from oracle.mpdo_oracle import OracleCircuit
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    n, depth = CFG4['n'], 8

    def build(mode, fast):
        oc = OracleCircuit(n, ideal=False, noiseType='idealNoise', chi=CFG4['chi'], kappa=CFG4['kappa'], chip='medium',
                           dtype=torch.complex64, svd_mode=mode, fast=fast)
        bc.brickwork(oc, n, depth, bc.angles([0], bc.n_draws(n, CFG4['depth'], 'cz')), 'cz')
        return oc

    def split_layers(oc):
        chunks, cur, ntr = [], [], 0
        for L in oc.layers:
            cur.append(L)
            if L[0] == 'truncate':
                ntr += 1
                if ntr % 2 == 0:
                    chunks.append(cur)
                    cur = []
        return chunks

    prep = build('exact', True)
    chunks = split_layers(prep)
    prep.layers = [L for ch in chunks[:6] for L in ch]
    prep.evolve()
    ref = build('reference', False)
    ref.T, ref.bond, ref.inner = [t.clone() for t in prep.T], list(prep.bond), list(prep.inner)
    ref.layers = [L for ch in split_layers(ref)[6:8] for L in ch]
    t0 = time.perf_counter()
    ref.run_layers()
    secs = time.perf_counter() - t0
    return secs, threads


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    updates, times, cores = cpu_sample(repeats=max(1, args.steps), window=6, warmup=max(0, args.warmup))
    total = sum(times)
    value = updates * len(times) / total
    upd4, times4, _ = cpu_sample(repeats=1, window=4)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'chi': CHI, 'kappa': KAPPA,
                   'note': 'ONE CPU process on all host cores whatever --gpus says: a per-N ratio against this line '
                           'divides N GPUs by one CPU process'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'oracle (reference mode) on two rzz = 4 chi-matrix CZ updates + the truncate sweep '
                                   '(5 bonds, 6 kappa sites) of a 6-site steady-state Gaussian window, per step; '
                                   '%d warm-up repeats untimed' % max(0, args.warmup),
                         'four_site_window_value': upd4 / times4[0]},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if not args.no_cfg4:
        secs4, _ = cfg4_cpu_sample()
        line['cfg4'] = {'metric': 'circuits_per_sec', 'value': 1.0 / (8 * secs4), 'unit': 'circuits/s', 'cores': cores,
                        'kind': 'port', 'workload': CFG4_WORKLOAD,
                        'sample': CFG4_CPU_SAMPLE % secs4}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def cfg4_leg(dev, rank, world, barrier):
    """BASELINE configs[3] on `world` GPUs: circuit ids 0 .. CFG4_PER_RANK*world-1 sharded `id mod world`, each rank
    evolves its shard as ONE batch through the public API, builds the 32-wide readout rows and the ranks exchange them
    with one all-gather (NCCL over NVLink) inside the timed region. Returns per-rank seconds; the caller takes the max."""
    import bench_configs as bc
    import MPDOSimulator as Simulator
    from MPDOSimulator import dmOperations
    from MPDOSimulator._engine.sharding import gather_readout, shard
    n, depth = CFG4['n'], CFG4['depth']
    total = CFG4_PER_RANK * world
    ids = shard(range(total), rank, world)
    ang_host = bc.angles(ids, bc.n_draws(n, depth, 'cz')).pin_memory()     # [draws, B] float64, pinned host memory
    copied = [0]

    def build(ang):
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='idealNoise', chi=CFG4['chi'], kappa=CFG4['kappa'],
                                    chip='medium', dtype=torch.complex64, device=dev)
        bc.brickwork(c, n, depth, ang, 'cz')
        return c

    def run(c):
        st = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64, device='cpu')
        c.evolve(st)
        dmn = c.cal_dmNodes()
        cols = [dmOperations.pauli_expect(dmn, 2, q) for q in range(n)]
        cols += [dmOperations.pauli_expect(dmn, [2, 2], [q, q + 1]) for q in range(n - 1)]
        cols.append(c._engine().chain_value_proj(c._Ts(), [0] * n))
        rows = torch.stack([x.reshape(-1).to(torch.float64) for x in cols], dim=1)     # [B, 32]
        return gather_readout(rows, total), c.last_stats.get('noisy_2q_updates', 0)

    run(build(ang_host))                      # rehearsal of the identical workload (pools, descriptor caches, NCCL)
    import gc
    gc.collect()
    gc.freeze()                               # see b200_arm: no full collection inside a timed region
    barrier()

    def allocator():                          # cudaMalloc / cudaFree calls and out-of-memory retries of torch's allocator
        ms = torch.cuda.memory_stats(dev)     # so far: a leg that needed them explains a slower leg (they synchronise)
        return {'device_alloc': int(ms.get('num_device_alloc', 0)), 'device_free': int(ms.get('num_device_free', 0)),
                'alloc_retries': int(ms.get('num_alloc_retries', 0)),
                'reserved_GB': round(ms.get('reserved_bytes.all.current', 0) / 2 ** 30, 2)}

    alloc_log = {'after_rehearsal': allocator()}
    # device-resident leg: circuit objects (gate operands) built before the clock starts
    c = build(ang_host)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    table, updates = run(c)
    e1.record()
    barrier()
    secs = e0.elapsed_time(e1) * 1e-3
    alloc_log['after_value_leg'] = allocator()
    # end-to-end leg: angles start in pinned host memory, circuits are built, evolved, read out, gathered and the
    # table lands in host memory, all inside the timed region; gate operands uploaded are counted
    orig_dev = Simulator.TensorCircuit._dev

    def counting_dev(self, t):
        out = orig_dev(self, t)
        copied[0] += out.numel() * out.element_size()
        return out

    Simulator.TensorCircuit._dev = counting_dev
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    table2, _ = run(build(ang_host))
    host_table = table2.cpu()
    f1.record()
    barrier()
    Simulator.TensorCircuit._dev = orig_dev
    e2e_secs = f0.elapsed_time(f1) * 1e-3
    alloc_log['after_e2e_leg'] = allocator()
    finite = bool(torch.isfinite(host_table).all())
    same = float((table - table2).abs().max())
    return {'secs': secs, 'e2e_secs': e2e_secs, 'circuits_per_rank': len(ids), 'total': total, 'updates': updates,
            'h2d': copied[0], 'd2h': host_table.numel() * 8, 'finite': finite, 'repeat_diff': same,
            'trace_like_p0_first': float(host_table[0, -1]), 'peak_mem_GB': torch.cuda.max_memory_allocated() / 2 ** 30,
            'allocator': alloc_log}


def b200_arm(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))

    import MPDOSimulator as Simulator
    from MPDOSimulator import _engine
    base = _engine.get_prims()
    lib = base.lib

    n, K, W = N_QUBITS, args.steps, args.warmup
    assert W >= 3, 'timing rules: at least 3 warm-up steps'
    files = {'CZ': {f'{i}{i + 1}': chi_file() for i in range(n - 1)}, 'CP': {}}
    P = PREROLL   # untimed input preparation: layers 0..P-1 saturate every bond at chi (see tools/per_layer.py)
    angles = layer_angles(rank, depth=max(DEPTH, P + W + K))

    def layer_circuit(d):
        c = Simulator.TensorCircuit(qn=n, ideal=False, noiseType='realNoise', chiFileDict=files, chi=CHI,
                                    kappa=KAPPA, chip='best', dtype=torch.complex64, device=dev)
        upd = add_layer(c, d, angles)
        return c, upd

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # long before the timed region (see ClockSampler); rows are filtered to the region
    circuits = [layer_circuit(d) for d in range(P + W + K)]
    state = Simulator.Tools.create_ket0Series(n, dtype=torch.complex64, device='cpu')
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for d in range(P + W):
        circuits[d][0].evolve(state)
    barrier()
    snapshot = [s.data.clone() for s in state]   # the e2e leg replays the same K layers from this state
    # Python's cyclic collector: the ~30 circuit objects built above are a few hundred thousand tracked containers, and a
    # full (generation 2) collection that happens to fall into a timed step walks all of them with the GIL held -
    # measured as single 65-210 ms steps among 30 ms ones. Collect once now and move the survivors to the permanent
    # generation; later collections only see what the steps themselves allocate.
    import gc
    gc.collect()
    gc.freeze()

    if args.profile:   # per-kernel device-time table of one steady-state layer (not a benchmark number)
        from torch.profiler import ProfilerActivity, profile
        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            circuits[P + W][0].evolve(state)
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        with open(args.profile, 'w') as f:
            f.write(f'one layer under the profiler: wall {wall:.3f} s\n')
            f.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=90))
        return

    # ---- last warm-up pass: the K timed layers themselves, run once untimed and rolled back. Every layer has its own
    # ranks and therefore its own scratch shapes; the first pass over a layer grows the stream-ordered pools and fills
    # the descriptor caches (measured: an un-rehearsed pass can take 3x the steady-state time), which is start-up
    # cost of the process, not throughput of the path.
    for d in range(P + W, P + W + K):
        circuits[d][0].evolve(state)
    if world > 1:   # the exchange step too: the first NCCL collective of a process sets the communicator up
        readout = circuits[P + W + K - 1][0].bitstring_probabilities(['0' * n]).to(torch.float64).reshape(1)
        gathered = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, readout)
    barrier()
    for s, snap in zip(state, snapshot):
        s.data = snap.clone()

    # ---- timed region: device-resident state -----------------------------------------------------
    if rank == 0:
        sampler.begin()
    cuprof = os.environ.get('MPDO_BENCH_CUPROF') == '1'   # `ncu --profile-from-start off`: list the timed region only
    if cuprof:
        torch.cuda.cudart().cudaProfilerStart()
    launches0 = base.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    updates = 0
    step_marks = [ev0]
    for d in range(P + W, P + W + K):
        flush.zero_()
        c, upd = circuits[d]
        c.evolve(state)
        updates += upd
        step_marks.append(torch.cuda.Event(enable_timing=True))
        step_marks[-1].record()
    if world > 1:   # the one exchange step of the path: gather the per-circuit readout
        readout = circuits[P + W + K - 1][0].bitstring_probabilities(['0' * n]).to(torch.float64).reshape(1)
        gathered = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, readout)
    ev1.record()
    barrier()
    if cuprof:
        torch.cuda.cudart().cudaProfilerStop()
    launches = base.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    secs = ev0.elapsed_time(ev1) * 1e-3
    import ctypes as _C
    pool_now, pool_high = _C.c_int64(), _C.c_int64()
    lib.mpdo_pool_stats(_C.byref(pool_now), _C.byref(pool_high))
    step_ms = [step_marks[i].elapsed_time(step_marks[i + 1]) for i in range(K)]
    bond_dims = [int(s.data.shape[4]) for s in state[:-1]]

    # ---- roofline leg: the same K layers replayed from the same state with CUDA events around every launch of the
    # library (on the launching stream). Kept out of the `value` / `e2e` passes: ~2 x 1800 event records per step,
    # issued from a dozen strand threads, slow the step itself by 1.5x (measured), which would make the headline a
    # measurement of the instrumentation.
    for s, snap in zip(state, snapshot):
        s.data = snap.clone()
    lib.mpdo_timing_enable(1)
    barrier()
    for d in range(P + W, P + W + K):
        flush.zero_()
        circuits[d][0].evolve(state)
    barrier()

    def timing(cls, min_flops=0.0):
        import ctypes as C
        sec, fl, by, mxs, mxf = (C.c_double() for _ in range(5))
        n = C.c_int64()
        lib.mpdo_timing_summary(cls, float(min_flops), C.byref(sec), C.byref(fl), C.byref(by), C.byref(n),
                                C.byref(mxs), C.byref(mxf))
        return {'seconds': sec.value, 'flops': fl.value, 'bytes': by.value, 'launches': n.value,
                'largest_seconds': mxs.value, 'largest_flops': mxf.value}

    t_c32, t_c64, t_big32, t_big64 = timing(0), timing(3), timing(0, 2e9), timing(3, 2e9)
    t_tc, t_bigtc = timing(4), timing(4, 2e9)            # tcgen05 / TMA applies (csrc/tc_apply.cu)
    t_jacobi, t_chol = timing(1), timing(2)

    def merged(a, b):   # fp32- and fp64-accumulated contraction launches together
        big = a if a['largest_flops'] >= b['largest_flops'] else b
        out = {k: a[k] + b[k] for k in ('seconds', 'flops', 'bytes', 'launches')}
        out['largest_seconds'], out['largest_flops'] = big['largest_seconds'], big['largest_flops']
        return out

    t_contract, t_big = merged(merged(t_c32, t_c64), t_tc), merged(merged(t_big32, t_big64), t_bigtc)
    lib.mpdo_timing_enable(0)

    # ---- e2e: host buffers in, host buffers out, every step -------------------------------------------
    # one pinned staging buffer per site, sized for the largest site tensor the truncated state can have
    cap = CHI * 2 * KAPPA * CHI
    pinned = [torch.empty(cap, dtype=torch.complex64, pin_memory=True) for _ in state]
    shapes = [tuple(s.data.shape) for s in state]
    for s, snap in zip(state, snapshot):
        s.data = snap.clone()
    shapes = [tuple(s.data.shape) for s in state]
    for s, pb in zip(state, pinned):
        pb[:s.data.numel()].copy_(s.data.reshape(-1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d = d2h = 0
    barrier()
    e0.record()
    e2e_updates = 0
    for d in range(P + W, P + W + K):
        flush.zero_()
        for s, pb, shp in zip(state, pinned, shapes):           # H2D: the step's input state from host memory
            nel = math.prod(shp)
            s.data = pb[:nel].to(dev, non_blocking=True).reshape(shp)
            h2d += nel * 8
        c, upd = circuits[d]
        c.evolve(state)
        e2e_updates += upd
        shapes = [tuple(s.data.shape) for s in state]
        for s, pb in zip(state, pinned):                        # D2H: the step's result back to host memory
            nel = s.data.numel()
            pb[:nel].copy_(s.data.reshape(-1), non_blocking=True)
            d2h += nel * 8
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    e2e_secs = e0.elapsed_time(e1) * 1e-3

    # ---- one split per gate: the same K layers the way the reference splits them (MPDO_NO_FUSE=1: every chi-matrix CZ
    # of an rzz gets its own SVD split and rank rule). Rehearsed once (other ranks, other scratch shapes), then timed.
    nofuse_secs = None
    if not args.no_unfused:
        os.environ['MPDO_NO_FUSE'] = '1'
        for timed in (False, True):
            for s, snap in zip(state, snapshot):
                s.data = snap.clone()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            n0.record()
            for d in range(P + W, P + W + K):
                flush.zero_()
                circuits[d][0].evolve(state)
            n1.record()
            barrier()
            if timed:
                nofuse_secs = n0.elapsed_time(n1) * 1e-3
        del os.environ['MPDO_NO_FUSE']

    # ---- cfg4: the batched configuration, sharded over the ranks -----------------------------------------------
    cfg4 = None
    if not args.no_cfg4:
        del state, snapshot, circuits
        torch.cuda.empty_cache()
        cfg4 = cfg4_leg(dev, rank, world, barrier)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    secs, e2e_secs = max_over_ranks(secs), max_over_ranks(e2e_secs)
    if nofuse_secs is not None:
        nofuse_secs = max_over_ranks(nofuse_secs)
    if cfg4 is not None:
        cfg4['secs'], cfg4['e2e_secs'] = max_over_ranks(cfg4['secs']), max_over_ranks(cfg4['e2e_secs'])
    tot_updates, tot_e2e_updates = sum_over_ranks(updates), sum_over_ranks(e2e_updates)
    tot_launches = sum_over_ranks(launches)

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback 1.4 PFLOP/s'
        ct, cf = max(t_contract['seconds'], 1e-30), t_contract['flops']
        roof = {
            'bound': 'tensor',
            'kernel': 'contraction kernels, all launches: tc_apply_kernel (complex64 applies: tcgen05.mma kind::tf32 3xTF32, '
                      'TMA tiles, TMEM accumulators) + contract_kernel (fp64-accumulated Gram matrices and cores on DMMA '
                      'm8n8k4 tiles; short / strided complex64 products on FP32 FFMA tiles)',
            'achieved': cf / ct / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': cf / ct / 1e12 / peak_tf,
            'traffic': None, 'peak_source': peak_src,
            'launches_timed': t_contract['launches'], 'kernel_seconds': ct, 'share_of_step_device_time': None,
            'algorithmic_GB_per_s': t_contract['bytes'] / ct / 1e9,
            'launches_over_2GFLOP': {'n': t_big['launches'],
                                     'TFLOP/s': t_big['flops'] / max(t_big['seconds'], 1e-30) / 1e12},
            'largest_launch': {'GFLOP': t_contract['largest_flops'] / 1e9, 'ms': t_contract['largest_seconds'] * 1e3,
                               'TFLOP/s': t_contract['largest_flops'] / max(t_contract['largest_seconds'], 1e-30) / 1e12},
            'note': 'algorithmic flops = 8*M*N*K per complex contraction (SURVEY 8d), every launch timed with CUDA '
                    'events on its own stream in a replay of the timed layers from the same state (the value / e2e '
                    'passes run without the per-launch events); the denominator is the dense bf16 tensor peak although '
                    'the kernel must deliver fp32/fp64-accurate complex arithmetic (FFMA / fp64 DMMA), most of it fp64-'
                    'accumulated Gram matrices whose own ceiling is the fp64 pipe (largest_launch.frac_of_fp64_pipe). Device time is split between this kernel and the two latency-bound '
                    'factorisation kernels (see factorisation_kernels).',
        }
        # DRAM traffic of the largest contraction launch of a layer (the environment step of a wide site) from the
        # committed `ncu --set full` capture, next to its algorithmic bytes
        tr = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
        if os.path.exists(tr):
            t = json.load(open(tr))
            roof['traffic'] = t['dram_read_bytes'] + t['dram_write_bytes']
            roof['traffic_detail'] = {k: t[k] for k in ('launch', 'dram_read_bytes', 'dram_write_bytes',
                                                        'algorithmic_bytes', 'source')}
        # the same launches against the pipes they actually run on (nominal SIMT peaks at the maximum SM clock:
        # 64 fp64 / 128 fp32 FMA lanes per SM per clock)
        sm_ghz = (peaks.get('sm_max_mhz') or 1965.0) / 1e3
        fp64_peak, fp32_peak = 148 * 64 * 2 * sm_ghz / 1e3, 148 * 128 * 2 * sm_ghz / 1e3

        def pipe(t, tb, peak):
            rate = t['flops'] / max(t['seconds'], 1e-30) / 1e12
            big = tb['flops'] / max(tb['seconds'], 1e-30) / 1e12
            top = t['largest_flops'] / max(t['largest_seconds'], 1e-30) / 1e12
            return {'launches': t['launches'], 'kernel_seconds': t['seconds'], 'TFLOP/s': rate,
                    'launches_over_2GFLOP': {'n': tb['launches'], 'TFLOP/s': big, 'frac_of_pipe': big / peak},
                    'largest_launch': {'GFLOP': t['largest_flops'] / 1e9, 'ms': t['largest_seconds'] * 1e3,
                                       'TFLOP/s': top, 'frac_of_pipe': top / peak},
                    'pipe_peak_TFLOP/s': peak}

        roof['by_pipe'] = {'fp64_accumulated (DMMA m8n8k4 tiles)': pipe(t_c64, t_big64, fp64_peak),
                           'fp32 (FFMA tiles)': pipe(t_c32, t_big32, fp32_peak),
                           # algorithmic rate; the tile executes 3 tf32 MMAs per product (3xTF32), so the tensor pipe
                           # is 3x busier than this figure says. Pipe peak taken as half the measured bf16 peak.
                           'tcgen05 kind::tf32 3xTF32 (TMA + TMEM)': pipe(t_tc, t_bigtc, peak_tf / 2)}
        jt = max(t_jacobi['seconds'], 1e-30)
        ht = max(t_chol['seconds'], 1e-30)
        roof['share_of_step_device_time'] = ct / (ct + jt + ht)
        roof['dominant_by_device_time'] = ('factorisation kernels (see factorisation_kernels): after the work reductions '
                                           'of round 2 the contractions are a few per cent of the layer')
        dominant = {
            'jacobi': {'kernel': 'jacobi_cluster_kernel / jacobi_kernel (one-sided Jacobi on the rows of the Cholesky factor, fp64, '
                                 'cluster-resident in shared memory for 32 < n <= 256)',
                       'launches': t_jacobi['launches'], 'kernel_seconds': jt,
                       'avg_launch_us': 1e6 * jt / max(t_jacobi['launches'], 1),
                       'share_of_timed_device_seconds': jt / (ct + jt + ht),
                       'algorithmic_TFLOP/s': t_jacobi['flops'] / jt / 1e12,
                       'frac': t_jacobi['flops'] / jt / 1e12 / fp64_peak, 'peak': fp64_peak,
                       'flops_convention': 'SURVEY 8d SVD count 4 (6 m n^2 + 20 n^3) per decomposition'},
            'cholesky': {'kernel': 'chol_blocked_kernel (blocked, no pivoting: preconditioner of the eigen-solver) + '
                                   'chol_cluster_kernel / chol_small_kernel (rank-revealing pivoted, with the left inverse: '
                                   'factors of the bond environments)',
                         'launches': t_chol['launches'], 'kernel_seconds': ht,
                         'avg_launch_us': 1e6 * ht / max(t_chol['launches'], 1),
                         'share_of_timed_device_seconds': ht / (ct + jt + ht),
                         'algorithmic_TFLOP/s': t_chol['flops'] / ht / 1e12,
                         'frac': t_chol['flops'] / ht / 1e12 / fp64_peak, 'peak': fp64_peak,
                         'flops_convention': '8 n^3 / 3 for the factor + the same for the left inverse'},
            'bound': 'latency: one cluster barrier per tournament round / pivot step / half panel on the 4-16 SMs of a '
                     'matrix (ncu: profiles/r2_ncu_kernels.md)',
            'note': 'kernel seconds are summed over concurrent streams, so they can exceed the wall time of the step',
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            upd, times, cores = cpu_sample(repeats=3, window=6, warmup=1)
            cpu = {'value': upd * len(times) / sum(times), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': 'oracle (reference mode) on two rzz = 4 chi-matrix CZ updates + the truncate sweep (5 bonds) '
                             'of a 6-site steady-state Gaussian window, 3 repeats after 1 warm-up (%.1f s)' % sum(times)}
        line = {
            'metric': METRIC, 'value': tot_updates / secs, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': 1e3 * secs / K, 'ms_each_step': [round(x, 2) for x in step_ms], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'chi': CHI, 'kappa': KAPPA, 'qubits': n,
                       'step': 'one brickwork layer (20 u3 + 9-10 rzz = 18-20 chi-matrix CZ updates + truncate) per GPU, '
                               'layers %d..%d of the depth-20 circuit (bonds saturated at chi)' % (P + W, P + W + K - 1),
                       'parallelism': f'replicas x{world} (one circuit per GPU)',
                       'l2': 'flushed between steps (256 MB write); per-pair transients are 537 MB > L2',
                       'warmup_detail': '%d pre-roll layers (bonds saturate) + %d warm-up layers + one untimed rehearsal '
                                        'of the %d timed layers, rolled back to the same state' % (P, W, K),
                       'scratch_pool_GiB': {'reserved': round(pool_now.value / 2 ** 30, 2),
                                            'high_water': round(pool_high.value / 2 ** 30, 2)},
                       'bond_dims_after_timed_region': bond_dims},
            'e2e': {'value': tot_e2e_updates / e2e_secs, 'unit': UNIT, 'h2d_bytes_per_step': h2d // K,
                    'd2h_bytes_per_step': d2h // K, 'ms_per_step': 1e3 * e2e_secs / K},
            'gpu_launches': int(tot_launches), 'clocks': clocks, 'roofline': roof, 'factorisation_kernels': dominant,
            'cpu_baseline': cpu,
        }
        if nofuse_secs is not None:
            line['value_one_split_per_gate'] = {
                'value': tot_updates / nofuse_secs, 'unit': UNIT, 'ms_per_step': 1e3 * nofuse_secs / K,
                'note': 'same K layers with MPDO_NO_FUSE=1: one SVD split + rank rule per chi-matrix CZ as in the '
                        'reference (Circuit.py:120-124); `value` fuses the two CZs of an rzz into one split'}
        if cfg4 is not None:
            cps, cps_e2e = cfg4['total'] / cfg4['secs'], cfg4['total'] / cfg4['e2e_secs']
            line['cfg4'] = {
                'metric': 'circuits_per_sec', 'value': cps, 'unit': 'circuits/s', 'n_gpus': world, 'scaling': 'weak',
                'circuits_total': cfg4['total'], 'circuits_per_gpu': cfg4['circuits_per_rank'],
                'seconds': cfg4['secs'], 'noisy_2q_updates_per_circuit': cfg4['updates'],
                'e2e': {'value': cps_e2e, 'unit': 'circuits/s', 'seconds': cfg4['e2e_secs'],
                        'h2d_bytes_per_step': cfg4['h2d'], 'd2h_bytes_per_step': cfg4['d2h']},
                'config': {'workload': CFG4_WORKLOAD, 'sharding': 'circuit id mod world (_engine/sharding.py), one batch of '
                           '%d circuits per GPU = the per-GPU share of the 1024-circuit job on 8 GPUs' % cfg4['circuits_per_rank'],
                           'exchange': 'one all-gather of the [circuits, 32] float64 readout table (16 <Z>, 15 <ZZ>, '
                                       'P(0...0)) inside the timed region',
                           'warmup': 'one untimed rehearsal of the identical batch',
                           'l2': 'the batched site tensors (537 MB per layer sweep) exceed L2'},
                'checks': {'readout_finite': cfg4['finite'], 'repeat_max_abs_diff': cfg4['repeat_diff'],
                           'peak_mem_GB': cfg4['peak_mem_GB'], 'allocator': cfg4['allocator']},
            }
            if world == 1 and not args.no_cpu_baseline:
                secs4, cores4 = cfg4_cpu_sample()
                line['cfg4']['cpu_baseline'] = {
                    'value': 1.0 / (8 * secs4), 'unit': 'circuits/s', 'cores': cores4, 'kind': 'port',
                    'sample': CFG4_CPU_SAMPLE % secs4}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cfg4', action='store_true', help='skip the batched (cfg4) leg')
    ap.add_argument('--no-unfused', action='store_true', help='skip the MPDO_NO_FUSE=1 leg')
    ap.add_argument('--profile', default=None, help='write a per-kernel time table of one steady-state step here')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == '__main__':
    main()
