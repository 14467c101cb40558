"""Truncation sweeps of the B200 build (reference MPDOSimulator/TNNOptimizer.py): same function names and
argument meaning, operating in place on a list of dense site nodes. Every sweep step is a short sequence
of CUDA launches (see _engine/steps.py); there is no CPU path."""
from typing import List, Optional

from . import _engine
from ._engine.strands import adopt, hand_over, run_strands
from ._node import DenseNode

__all__ = ['qr_left2right', 'svd_left2right', 'svd_right2left', 'svdKappa_left2right', 'bondTruncate',
           'checkConnectivity', 'truncateLayer']


def checkConnectivity(_qubits: List[DenseNode]):
    """True when every neighbouring pair shares a bond (tn.check_connected, TNNOptimizer.py:51-61)."""
    if len(_qubits) <= 1:
        return True
    return all(q.has_right for q in _qubits[:-1])


def _validate_qubit_list(qubits, func_name: str):
    if not isinstance(qubits, list):
        raise TypeError(f'`{func_name}` expects a list of qubit nodes, got {type(qubits).__name__}')


def _engine_of(qubits):
    """Engine for the state's dtype; also brings every site to the common batch size (a parameter-sweep batch
    starts on the sites its batched gates touched, the sweeps need it everywhere)."""
    B = max(q.data.shape[0] for q in qubits)
    for q in qubits:
        if q.data.shape[0] != B:
            q.data = q.data.expand(B, *q.data.shape[1:]).contiguous()
    return _engine.engine_for(qubits[0].data.dtype)


def bondTruncate(_qubits: List[DenseNode], max_singular_values: Optional[int] = None,
                 max_truncation_err: Optional[float] = None, regularization: bool = False):
    """QR sweep left-to-right, then SVD truncation right-to-left (TNNOptimizer.py:72-84)."""
    if max_singular_values is None and max_truncation_err is None and not regularization:
        return None
    if _use_env_form(_qubits, max_singular_values, max_truncation_err):
        eng = _engine_of(_qubits)
        Ts = [q.data for q in _qubits]

        def publish(idx, tensor):       # sites are final one by one, right to left: release what they replace at once
            _qubits[idx].data = tensor

        eng.bond_truncate_env(Ts, max_singular_values, publish=publish)
        for q, t in zip(_qubits, Ts):
            q.data = t
        return None
    qr_left2right(_qubits)
    svd_right2left(_qubits, max_singular_values=max_singular_values, max_truncation_err=max_truncation_err)


ENV_FORM_MAX_BATCH = 4


def _use_env_form(_qubits, max_singular_values, max_truncation_err):
    """The environment form of bondTruncate (_engine/steps.py bond_truncate_env: the QR sweep becomes a chain of
    contractions plus mutually independent factorisations, leaving the right-to-left sweep as the only sequential chain
    of decompositions) is taken for complex64 states with a fixed chi and few circuits per call, where a layer is bound
    by the latency of dependent factorisations. Large batches are bound by throughput instead, and there the fp64
    environment contractions cost more than the fp32 products of the QR sweep. Same results either way
    (tests/test_env_sweep.py, tests/test_gpu_env_sweep.py). MPDO_ENV_SWEEP=0 / 1 forces the choice."""
    import os
    import torch
    if max_singular_values is None or max_truncation_err is not None or len(_qubits) < 2:
        return False
    if not all(q.has_right for q in _qubits[:-1]) or _qubits[0].data.shape[1] != 1:
        return False      # the two-sweep form raises the reference's error for a missing bond
    if _qubits[0].data.dtype != torch.complex64:
        return False
    knob = os.environ.get('MPDO_ENV_SWEEP')
    if knob is not None:
        return knob == '1'
    return max(q.data.shape[0] for q in _qubits) <= ENV_FORM_MAX_BATCH


def qr_left2right(_qubits: List[DenseNode]):
    """T_i = Q R, T_i <- Q, T_{i+1} <- R T_{i+1} for i = 0..n-2 (TNNOptimizer.py:87-108)."""
    _validate_qubit_list(_qubits, "qr_left2right")
    eng = _engine_of(_qubits)
    for i in range(len(_qubits) - 1):
        if not _qubits[i].has_right:
            raise ValueError(f'Axis name bond_{i}_{i + 1} not found')  # the reference indexes it unconditionally
        _qubits[i].data, _qubits[i + 1].data = eng.qr_step(_qubits[i].data, _qubits[i + 1].data)
        _qubits[i]._left_canonical = _qubits[i].data      # svd_right2left relies on this isometry


def svd_right2left(_qubits: List[DenseNode], max_singular_values: Optional[int] = None,
                   max_truncation_err: Optional[float] = None):
    """Two-site SVD truncation for idx = n-1..1, sqrt(S) on both sides (TNNOptimizer.py:111-134). The left
    neighbour is an isometry after qr_left2right, so only T_idx is decomposed (SVD(Q X) = Q SVD(X))."""
    _validate_qubit_list(_qubits, "svd_right2left")
    # SVD(Q X) = Q SVD(X) needs every left neighbour to be the isometry qr_left2right left behind; the reference
    # contracts both sites and is valid for any state, so a state that is not (any more) left-canonical is
    # canonicalised here first instead of being truncated wrongly
    if any(getattr(q, '_left_canonical', None) is not q.data for q in _qubits[:-1]):
        qr_left2right(_qubits)
    eng = _engine_of(_qubits)
    discarded = []
    for idx in range(len(_qubits) - 1, 0, -1):
        left, right = _qubits[idx - 1], _qubits[idx]
        left.data, right.data, d = eng.bond_svd_step(left.data, right.data, max_singular_values, max_truncation_err)
        discarded.append(d)
    return discarded


def svd_left2right(_qubits: List[DenseNode], max_singular_values: int, max_truncation_err: Optional[float] = None):
    """Never called by the reference (dead code, TNNOptimizer.py:137-161); kept for API completeness."""
    raise NotImplementedError('svd_left2right is dead code in the reference and is not part of the B200 build')


def svdKappa_left2right(_qubits: List[DenseNode], max_singular_values: Optional[int] = None,
                        max_truncation_err: Optional[float] = None):
    """Inner-index truncation T <- U S per site (TNNOptimizer.py:164-197)."""
    if max_singular_values is None and max_truncation_err is None:
        return None
    _validate_qubit_list(_qubits, "svdKappa_left2right")
    eng = _engine_of(_qubits)
    todo = []
    for q in _qubits:
        # the reference skips a site by its own inner dimension (:176-181), which counts redundant Kraus operators
        # that the dense build never materialises (DenseNode.ref_inner)
        if max_singular_values is not None and (not q.has_inner or q.nominal_inner() <= max_singular_values):
            continue
        if (max_truncation_err is None and max_singular_values is not None and q.has_inner
                and q.data.shape[3] <= max_singular_values):
            q.ref_inner = None      # nothing to cut: T <- U S would only rotate the inner index (a gauge change)
            continue
        if not q.has_inner:
            raise ValueError(f'Axis name I_{q.index} not found')
        todo.append(q)

    # Sites are independent. Sites whose tensors have the same shape (the bulk of a brickwork layer) are stacked along
    # the batch axis and truncated by ONE launch sequence; the few distinct shapes that remain (chain ends) are issued
    # concurrently, one CUDA stream each.
    import os
    import torch
    cuda = getattr(_engine.get_prims(), 'name', '') == 'cuda'
    groups = {}
    for q in todo:
        # (not with the relative-error rule: its kept rank is per site, and a stacked call pads to the group maximum)
        key = (tuple(q.data.shape) if _engine.grouping_enabled(q.data.shape[0]) and max_truncation_err is None
               else id(q))
        groups.setdefault(key, []).append(q)

    def make(members):
        def task():
            if parallel:
                for q in members:
                    adopt(q.data)
            if len(members) == 1:
                q = members[0]
                q.data, _ = eng.kappa_truncate(q.data, max_singular_values, max_truncation_err)
                q.ref_inner = None
                return
            B = members[0].data.shape[0]
            out, _ = eng.kappa_truncate(torch.cat([q.data for q in members], dim=0), max_singular_values,
                                        max_truncation_err)
            for m, q in enumerate(members):
                q.data = out[m * B:(m + 1) * B]
                q.ref_inner = None
        return task

    device = _qubits[0].data.device
    parallel = cuda and len(groups) > 1
    run_strands([make(members) for members in groups.values()], device, enabled=parallel)
    if parallel:
        for q in todo:
            hand_over(q.data, device)


def _needs_kappa(q: DenseNode, kappa: Optional[int]):
    return not (kappa is not None and (not q.has_inner or q.nominal_inner() <= kappa))


def truncateLayer(_qubits: List[DenseNode], chi: Optional[int] = None, kappa: Optional[int] = None,
                  max_truncation_err: Optional[float] = None, noisy: bool = True):
    """What a `truncate` layer does (Circuit.py:476-481): bondTruncate, then - for a noisy circuit -
    svdKappa_left2right. Same results as calling the two in turn; on a CUDA device the inner-index truncation of site
    i is issued on a side stream as soon as the right-to-left sweep has passed bond (i-1, i) - the site is final from
    then on - so the independent kappa steps can run beside the sequential sweep instead of after it. Opt-in
    (MPDO_KAPPA_PIPELINE=1): measured on B200 (profiles/r2_grouping.md) it does not shorten the cfg2 layer - the sweep's
    factorisations and the kappa Gram contractions compete for the same SMs - and it makes step times less regular."""
    import os
    import torch
    from ._engine.strands import _pool, _streams
    do_bond = not (chi is None and max_truncation_err is None)
    do_kappa = noisy and not (kappa is None and max_truncation_err is None)
    cuda = getattr(_engine.get_prims(), 'name', '') == 'cuda'
    if not (do_bond and do_kappa and cuda and len(_qubits) > 2) or os.environ.get('MPDO_KAPPA_PIPELINE', '0') != '1':
        if do_bond:
            bondTruncate(_qubits, max_singular_values=chi, max_truncation_err=max_truncation_err)
        if do_kappa:
            svdKappa_left2right(_qubits, max_singular_values=kappa, max_truncation_err=max_truncation_err)
        return
    qr_left2right(_qubits)
    eng = _engine_of(_qubits)
    dev = _qubits[0].data.device
    main = torch.cuda.current_stream(dev)
    from ._engine.strands import MAX_WORKERS
    streams = _streams(dev, min(MAX_WORKERS, len(_qubits)))
    futures, used = [], []

    def kappa_task(q, stream, event):
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            stream.wait_event(event)
            adopt(q.data)
            if not q.has_inner:
                raise ValueError(f'Axis name I_{q.index} not found')
            q.data, _ = eng.kappa_truncate(q.data, kappa, max_truncation_err)
            q.ref_inner = None

    def launch(q):
        if not _needs_kappa(q, kappa):
            return
        if max_truncation_err is None and kappa is not None and q.data.shape[3] <= kappa:
            q.ref_inner = None
            return
        ev = torch.cuda.Event()
        ev.record(main)
        stream = streams[len(futures) % len(streams)]
        used.append((q, stream))
        futures.append(_pool().submit(kappa_task, q, stream, ev))

    for idx in range(len(_qubits) - 1, 0, -1):
        left, right = _qubits[idx - 1], _qubits[idx]
        left.data, right.data, _ = eng.bond_svd_step(left.data, right.data, chi, max_truncation_err)
        launch(right)
    launch(_qubits[0])
    errors = []
    for f in futures:
        try:
            f.result()
        except BaseException as exc:   # re-raised below, after every stream has been joined
            errors.append(exc)
    for q, stream in used:
        main.wait_stream(stream)
        hand_over(q.data, dev)
    if errors:
        raise errors[0]
