"""The production engine: each step of the update path is ONE call into libmpdo_b200.so
(mpdo_split_2q, mpdo_qr_step, mpdo_bond_svd_step, mpdo_kappa_truncate; csrc/engine.cu issues the same kernel
sequences that steps.Engine spells out over the Python primitive wrappers). steps.Engine stays the readable
statement of the algorithm, the path for the rarely used relative-error truncation rule, and what the CPU tier
exercises against the oracle; NativeEngine only removes ~50 Python-level launches per step from the hot loop."""
import ctypes as C

import torch

from . import lib as _lib
from .steps import Engine
from .strands import BIG_STATE_BYTES, memory_guard, state_bytes

_DT = {torch.complex64: _lib.MPDO_C64, torch.complex128: _lib.MPDO_C128}
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_int64, C.c_void_p)


def _stream():
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _p(t):
    return C.c_void_p(t.data_ptr())


_CUDA_OOM = 2   # cudaErrorMemoryAllocation

class NativeEngine(Engine):
    def __init__(self, prims, dtype, npass=None):
        super().__init__(prims, dtype, npass)
        self.lib = _lib.load()
        self.dt = _DT[dtype]

    def _call(self, what, fn, *args):
        """One library call; on device-memory exhaustion (the library has already trimmed its own pools) give torch's
        cached blocks back to the driver and retry once. The step functions write nothing the caller keeps before
        their scratch is allocated, so a retry is safe."""
        rc = fn(*args)
        if rc == _CUDA_OOM:
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            self.lib.mpdo_trim_pools()
            rc = fn(*args)
        _lib.check(rc, what)

    def qr_step(self, Ti, Tn):
        Ti, Tn = Ti.contiguous(), Tn.contiguous()
        Bn, l, _, a, r = Ti.shape
        _, _, _, a2, r2 = Tn.shape
        Q, Tn_new = torch.empty_like(Ti), torch.empty_like(Tn)
        self._call('mpdo_qr_step', self.lib.mpdo_qr_step, self.dt, self.npass, Bn, l, a, r, _p(Ti), a2, r2, _p(Tn), _p(Q),
                   _p(Tn_new), _stream())
        return Q, Tn_new

    def bond_svd_step(self, Tl, Tr, chi, max_err=None):
        Tl, Tr = Tl.contiguous(), Tr.contiguous()
        Bn, lp, _, ap, l = Tl.shape
        _, _, _, a, r = Tr.shape
        k = l if chi is None else min(int(chi), l)
        Tl_n = torch.empty((Bn, lp, 2, ap, k), dtype=Tl.dtype, device=Tl.device)
        Tr_n = torch.empty((Bn, k, 2, a, r), dtype=Tr.dtype, device=Tr.device)
        sv = torch.empty((Bn, l), dtype=torch.float64, device=Tr.device)
        kk = C.c_int(k)
        self._call('mpdo_bond_svd_step', self.lib.mpdo_bond_svd_step, self.dt, self.npass, Bn, lp, ap, l, _p(Tl), a, r,
                   _p(Tr), k, -1.0 if max_err is None else float(max_err), C.byref(kk), _p(Tl_n), _p(Tr_n), _p(sv),
                   _stream())
        if kk.value != k:      # the relative-error rule kept fewer: the outputs were written densely with that rank
            k = kk.value
            Tl_n = Tl_n.view(-1)[:Bn * lp * 2 * ap * k].view(Bn, lp, 2, ap, k)
            Tr_n = Tr_n.view(-1)[:Bn * k * 2 * a * r].view(Bn, k, 2, a, r)
        disc = sv[:, k:].clamp_min(0).sqrt() if self.npass == 1 else sv[:, k:]
        return Tl_n, Tr_n, disc

    def bond_truncate_env(self, Ts, chi, publish=None):
        """steps.Engine.bond_truncate_env as two kinds of library calls: mpdo_env_sweep (environment chain on the
        current stream, the factorisations of all bonds on the library's side streams) and one mpdo_bond_env_step per
        bond, right to left. No host synchronisation (except on states of more than 1 GB, see memory_guard).
        `publish(idx, tensor)` hands every finished site to the caller at once, so that the tensor it replaces and the
        C.T product of that bond are released while the sweep goes on (not a layer's worth of them at the end)."""
        n = len(Ts)
        if n < 2:
            return []
        Ts[:] = [t.contiguous() for t in Ts]
        Bn = Ts[0].shape[0]
        dev = Ts[0].device
        big = state_bytes(Ts) >= BIG_STATE_BYTES
        memory_guard(dev, Ts, need=state_bytes(Ts))
        if Ts[0].shape[1] != 1:
            raise ValueError('the first site must have a trivial left bond')
        ls = (C.c_int * n)(*[t.shape[1] for t in Ts])
        as_ = (C.c_int * n)(*[t.shape[3] for t in Ts])
        rs = (C.c_int * n)(*[t.shape[4] for t in Ts])
        Cis = [None] + [torch.empty((Bn, t.shape[1], t.shape[1]), dtype=torch.complex128, device=dev) for t in Ts[1:]]
        Ms = [None] + [torch.empty_like(t) for t in Ts[1:]]
        tp = (C.c_void_p * n)(*[t.data_ptr() for t in Ts])
        cip = (C.c_void_p * n)(*[None if c is None else c.data_ptr() for c in Cis])
        mp = (C.c_void_p * n)(*[None if m is None else m.data_ptr() for m in Ms])
        self._call('mpdo_env_sweep', self.lib.mpdo_env_sweep, self.dt, Bn, n, ls, as_, rs, tp, cip, mp, _stream())
        W, disc = None, []
        # squared singular values of all bonds in one buffer: the discarded ones are clamped and rooted by ONE pair of
        # launches after the sweep instead of one per bond inside the sequential chain
        sv_off = [0] * (n + 1)
        for idx in range(1, n):
            sv_off[idx + 1] = sv_off[idx] + Bn * Ts[idx].shape[1]
        sv_all = torch.empty((max(sv_off[n], 1),), dtype=torch.float64, device=dev)
        for idx in range(n - 1, 0, -1):
            M0 = Ms[idx]
            _, l, _, a, r0 = M0.shape
            rw = r0 if W is None else W.shape[2]
            k = l if chi is None else min(int(chi), l)
            T_n = torch.empty((Bn, k, 2, a, rw), dtype=M0.dtype, device=dev)
            W_n = torch.empty((Bn, l, k), dtype=M0.dtype, device=dev)
            sv = sv_all[sv_off[idx]:sv_off[idx + 1]].view(Bn, l)
            self._call('mpdo_bond_env_step', self.lib.mpdo_bond_env_step, self.dt, Bn, l, a, r0, _p(M0), rw,
                       None if W is None else _p(W), _p(Cis[idx]), k, _p(T_n), _p(W_n), _p(sv), _stream())
            Ts[idx], W = T_n, W_n
            Ms[idx] = Cis[idx] = M0 = None
            if publish is not None:
                publish(idx, T_n)
            disc.append(sv[:, k:])
            if big and (n - idx) % 8 == 0:
                memory_guard(dev, Ts)
        sv_all.clamp_min_(0).sqrt_()      # (the kept values are rooted too; only the discarded ones are handed out)
        T0 = Ts[0]
        out = torch.empty(tuple(T0.shape[:4]) + (W.shape[2],), dtype=T0.dtype, device=dev)
        self.p.contract(T0, (1, 3, 1), W, (1, 1, 1), out, (1, 3, 1))
        Ts[0] = out
        if publish is not None:
            publish(0, out)
        return disc

    def kappa_truncate(self, T, kappa, max_err=None):
        T = T.contiguous()
        Bn, l, _, a, r = T.shape
        k = a if kappa is None else min(int(kappa), a)
        T_n = torch.empty((Bn, l, 2, k, r), dtype=T.dtype, device=T.device)
        disc = torch.empty((Bn,), dtype=torch.float64, device=T.device)
        kk = C.c_int(k)
        self._call('mpdo_kappa_truncate', self.lib.mpdo_kappa_truncate, self.dt, Bn, l, a, r, _p(T), k,
                   -1.0 if max_err is None else float(max_err), C.byref(kk), _p(T_n), _p(disc), _stream())
        if kk.value != k:
            k = kk.value
            T_n = T_n.view(-1)[:Bn * l * 2 * k * r].view(Bn, l, 2, k, r)
        return T_n, disc.unsqueeze(1)

    def split_2q(self, Tlo, Thi, G, max_err=2.718281828459045e-8):
        Tlo, Thi, G = Tlo.contiguous(), Thi.contiguous(), G.contiguous()
        Bn, l, _, a0, m = Tlo.shape
        _, _, _, a1, r = Thi.shape
        Bg, K = G.shape[0], G.shape[-1]
        assert G.dtype == Tlo.dtype == Thi.dtype
        outs = {}

        def alloc(which, count, _user):
            t = torch.empty(int(count), dtype=Tlo.dtype, device=Tlo.device)
            outs[which] = t
            return t.data_ptr()

        cb = ALLOC_FN(alloc)
        k = C.c_int(0)
        rows = (C.c_int * Bn)()
        self._call('mpdo_split_2q', self.lib.mpdo_split_2q, self.dt, self.npass, Bn, l, a0, m, _p(Tlo), a1, r, _p(Thi),
                   Bg, K, _p(G), -1.0 if max_err is None else float(max_err), C.cast(cb, C.c_void_p), None,
                   C.byref(k), rows, _stream())
        kk = k.value
        self.stats['last_rank'] = kk
        self.tls.last_ranks = list(rows)          # kept rank of every batch entry (kk is their maximum)
        return outs[0].view(Bn, l, 2, a0, kk), outs[1].view(Bn, kk, 2, K * a1, r)
