"""TEST INFRASTRUCTURE ONLY - a torch-CPU model of the device primitives in
MPDOSimulator/_engine/prims.py (same call signatures, same semantics, LAPACK instead of the Jacobi
kernels). It lets the CPU-only test tier exercise the host-side orchestration (steps.py, Circuit.py)
without a GPU, and it is the reference the `-m gpu` primitive tests compare each CUDA kernel against.
The product never imports this module: the product back end is CudaPrims and fails loudly without
the CUDA library.
"""
import numpy as np
import torch


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


class CpuPrims:
    name = 'cpu-model'

    def __init__(self):
        self.calls = 0

    def launch_count(self):
        return self.calls

    def contract(self, A, ra, B, rb, Cv, rc, conjA=False, conjB=False, acc64=None, alpha=1.0, beta=0.0,
                 hermitian=False):
        self.calls += 1
        nb, ni, nk = ra
        batch = _prod(A.shape[:nb])
        M, K = _prod(A.shape[nb:nb + ni]), _prod(A.shape[nb + ni:])
        N = _prod(B.shape[rb[0] + rb[1]:])
        assert _prod(B.shape[rb[0]:rb[0] + rb[1]]) == K
        assert _prod(Cv.shape[rc[0]:rc[0] + rc[1]]) == M and _prod(Cv.shape[rc[0] + rc[1]:]) == N
        all64 = A.dtype == B.dtype == Cv.dtype == torch.complex64
        acc = torch.complex128 if (acc64 if acc64 is not None else not all64) else torch.complex64
        a = A.reshape(batch, M, K).to(acc)
        b = B.reshape(batch, K, N).to(acc)
        if conjA:
            a = a.conj()
        if conjB:
            b = b.conj()
        res = alpha * torch.matmul(a, b)
        if beta != 0.0:
            res = res + beta * Cv.reshape(batch, M, N).to(acc)
        Cv.copy_(res.reshape(Cv.shape).to(Cv.dtype))
        return Cv

    def absorb_1q(self, T, G):
        self.calls += 1
        Bn, l, _, a, r = T.shape
        K = G.shape[-1]
        Ge = G.expand(Bn, 2, 2, K)
        out = torch.einsum('bpsg,blsar->blpgar', Ge, T)
        return out.reshape(Bn, l, 2, K * a, r).contiguous()

    def eigh_psd(self, G, tol=1e-15, sweeps=30, rank_revealing=False):
        self.calls += 1
        lam, V = torch.linalg.eigh(G)
        lam = lam.flip(-1).clamp_min(0.0).contiguous()
        Vh = V.flip(-1).mH.contiguous()
        return lam, Vh

    def svd_rows(self, L, tol=1e-15, sweeps=30, zero_tol=1e-300):
        self.calls += 1
        U, s, Wh = torch.linalg.svd(L, full_matrices=False)
        n, m = L.shape[1], L.shape[2]
        if m < n:  # pad to n rows like the device routine (rows beyond rank are zero)
            Bn = L.shape[0]
            Uf, sf, _ = torch.linalg.svd(L, full_matrices=True)
            U = Uf
            s = torch.cat([sf, torch.zeros(Bn, n - m, dtype=sf.dtype)], dim=1)
            Wh = torch.cat([Wh, torch.zeros(Bn, n - m, m, dtype=Wh.dtype)], dim=1)
        Wh = torch.where((s > zero_tol * s[:, :1])[:, :, None], Wh, torch.zeros_like(Wh))
        return U.mH.contiguous(), s.contiguous(), Wh.contiguous()

    def rowscale(self, V, lam, rows, power, tol, mode, dtype):
        self.calls += 1
        lv = lam[:, :rows]
        thr = tol * lam[:, :1]
        if mode == 0:
            ok = (lv > thr) & (lv > 0)
            f = torch.where(ok, lv.clamp_min(1e-300) ** power, torch.zeros_like(lv))
        else:
            lc = torch.maximum(lv, thr)
            f = torch.where(lc > 0, lc.clamp_min(1e-300) ** power, torch.zeros_like(lc))
        return (f[:, :, None] * V[:, :rows, :]).to(dtype).contiguous()

    def rank_rule(self, lam, squared, cap, max_err, relative, f32, zero_tail=True):
        self.calls += 1
        keep = []
        for b in range(lam.shape[0]):
            s = lam[b].clamp_min(0).sqrt() if squared else lam[b]
            num_err = cap
            if max_err is not None and s.numel() > 0:
                if f32:
                    s32 = s.to(torch.float32)
                    t = torch.sqrt(torch.cumsum(s32 ** 2, dim=0))
                    eps = np.float32(max_err) * s32[0] if relative else torch.tensor(max_err, dtype=torch.float32)
                else:
                    t = torch.sqrt(torch.cumsum(s ** 2, dim=0))
                    eps = max_err * s[0] if relative else max_err
                for idx in range(t.shape[0] - 1):
                    if t[-1] - t[idx] <= eps:
                        num_err = idx + 1
                        break
            k = min(cap, num_err)
            keep.append(k)
            if zero_tail:
                lam[b, k:] = 0
        return keep
