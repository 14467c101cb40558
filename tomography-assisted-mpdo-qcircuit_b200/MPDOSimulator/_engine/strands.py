"""Concurrent execution of independent pieces of one layer ("strands").

Between two truncate() markers the gates of a brickwork layer fall into groups of qubits that never
interact (a brick pair with its single-qubit rotations, or a lone qubit); likewise the inner-index
truncation treats every site on its own. Each strand is a short chain of small kernels (single-CTA Jacobi
decompositions, skinny contractions) that cannot fill 148 SMs alone, so strands are issued from a few host
threads, each on its own CUDA stream: the kernels of different brick pairs overlap on the device and the
host-side waits (rank read-backs) overlap too. The main stream waits for every side stream afterwards, so
callers see ordinary stream-ordered semantics.
"""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

_POOL = None
_STREAMS = {}
_LOCK = threading.Lock()
MAX_WORKERS = int(os.environ.get('MPDO_STRAND_WORKERS', '12'))   # tuning knob


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=MAX_WORKERS, thread_name_prefix='mpdo-strand')
    return _POOL


PREWARM_SMALL_MB = int(os.environ.get('MPDO_STRAND_PREWARM_MB', '48'))   # per stream; 0 disables
PREWARM_LARGE_MB = int(os.environ.get('MPDO_STRAND_PREWARM_LARGE_MB', '2048'))   # per stream, one splittable block


def _prewarm(device, stream):
    """Fills the caching allocator's small-block pool of `stream` once. Blocks handed between streams are released
    through record_stream events, so whether a strand finds a free small block depends on timing; when it does not,
    the allocator calls cudaMalloc for another 2 MB segment, and a cudaMalloc issued while persistent (barrier)
    kernels of other strands are running was measured to stall every launching thread for 50-400 ms
    (tools/prof_outliers.py: the occasional 2-3x slower step). A few dozen MB of cached segments per stream make
    that path cold in steady state."""
    with torch.cuda.device(device), torch.cuda.stream(stream):
        if PREWARM_SMALL_MB > 0:
            blocks = [torch.empty(1 << 20, dtype=torch.uint8, device=device) for _ in range(PREWARM_SMALL_MB)]
            del blocks
        # The same holds for the large pool (site tensors and gate-split outputs, up to 537 MB each on the headline
        # workload): one big cached block per stream, which the allocator splits on demand, instead of segments
        # that appear one cudaMalloc at a time for hundreds of steps. Capped at 1/64 of the free device memory.
        if PREWARM_LARGE_MB > 0:
            free_b, _ = torch.cuda.mem_get_info(device)
            want = min(PREWARM_LARGE_MB << 20, free_b // 64)
            if want >= (64 << 20):
                big = torch.empty(want, dtype=torch.uint8, device=device)
                del big


def _streams(device, count):
    key = str(device)
    with _LOCK:
        have = _STREAMS.setdefault(key, [])
        if not have:
            _prewarm(device, torch.cuda.current_stream(device))
        while len(have) < count:
            have.append(torch.cuda.Stream(device=device))
            _prewarm(device, have[-1])
        return have[:count]


def run_strands(tasks, device, enabled=True):
    """tasks: list of zero-argument callables, mutually independent. Runs them concurrently on side streams when
    it pays off (CUDA device, more than one task), else in order on the current stream."""
    dev = torch.device(device)
    if os.environ.get('MPDO_STRANDS', '1') == '0':   # debugging knob: everything on the caller's stream
        enabled = False
    if not enabled or len(tasks) < 2 or dev.type != 'cuda':
        for t in tasks:
            t()
        return
    main = torch.cuda.current_stream(dev)
    nworkers = min(len(tasks), MAX_WORKERS)
    streams = _streams(dev, nworkers)
    for s in streams:
        s.wait_stream(main)
    errors = []

    def work(wid):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(streams[wid]):
                for t in tasks[wid::nworkers]:
                    t()
        except BaseException as exc:  # re-raised in the caller's thread
            errors.append(exc)

    futures = [_pool().submit(work, w) for w in range(nworkers)]
    for f in futures:
        f.result()
    for s in streams:
        main.wait_stream(s)
    if errors:
        raise errors[0]


def adopt(tensor):
    """Called inside a strand for every tensor that was produced on another stream: tells the caching allocator
    that the current (side) stream uses it, so that its memory is not recycled under a kernel still reading it."""
    if tensor.is_cuda:
        tensor.record_stream(torch.cuda.current_stream(tensor.device))
    return tensor


def hand_over(tensor, device):
    """A tensor produced on a side stream is about to live on (and later be freed from) the main stream."""
    if tensor.is_cuda:
        tensor.record_stream(torch.cuda.current_stream(torch.device(device)))
    return tensor


BIG_STATE_BYTES = 1 << 30
_TOTAL = {}


def state_bytes(Ts):
    return sum(t.numel() * t.element_size() for t in Ts)


def memory_guard(dev, Ts, need=0):
    """Bounds how far the host runs ahead of the device on LARGE states (chi = 256, ~100 sites: tens of GB). Nothing
    in a truncation sweep synchronises, and tensors that crossed streams (strands) are only returned to the caching
    allocator when the device has passed their last use - with a layer that takes seconds on the device, a layer's
    worth of dead site tensors stays allocated beside the live ones (cfg5 at full size ran out of 178 GB). When the
    state is large and the allocator holds more than a third of the device, wait for the device: the dead tensors
    become reusable, and the host had nothing to do but wait anyway. States below 1 GB (every latency-bound
    workload) never get here."""
    if not Ts or not Ts[0].is_cuda or state_bytes(Ts) < BIG_STATE_BYTES:
        return
    dev = torch.device(dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _TOTAL:
        _TOTAL[idx] = torch.cuda.get_device_properties(idx).total_memory
    if torch.cuda.memory_allocated(idx) + need > _TOTAL[idx] // 3:
        torch.cuda.synchronize(idx)
