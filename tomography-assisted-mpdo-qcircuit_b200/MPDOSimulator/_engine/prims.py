"""Device primitives of the MPDO update path: thin wrappers that turn torch views (device memory and
strides only) into calls of the C ABI in include/mpdo_b200.h. No arithmetic happens in torch here.

`roles = (nb, n1, n2)` tells how the leading dims of a view are grouped: nb batch dims, then n1 dims
forming the first matrix axis, then n2 dims forming the second one. Any permuted / expanded / sliced
view is fine as long as each group collapses to at most three (size, stride) levels.
"""
import ctypes as C

import torch

from . import lib as _lib

_DT = {torch.complex64: _lib.MPDO_C64, torch.complex128: _lib.MPDO_C128}


def _levels(sizes, strides):
    """Collapse (size, stride) pairs, outer -> inner, dropping unit dims and merging contiguous ones."""
    lv = []
    for n, s in zip(sizes, strides):
        n, s = int(n), int(s)
        if n == 1:
            continue
        if lv and lv[-1][1] == s * n:
            lv[-1] = (lv[-1][0] * n, s)
        else:
            lv.append((n, s))
    return lv


def _idxmap(lv):
    m = _lib.IdxMap()
    if len(lv) == 0:
        m.d0, m.d1, m.s0, m.s1, m.s2 = 0, 0, 0, 0, 0
    elif len(lv) == 1:
        m.d0, m.d1, m.s0, m.s1, m.s2 = 0, 0, lv[0][1], 0, 0
    elif len(lv) == 2:
        m.d0, m.d1, m.s0, m.s1, m.s2 = lv[1][0], 0, lv[1][1], lv[0][1], 0
    elif len(lv) == 3:
        m.d0, m.d1, m.s0, m.s1, m.s2 = lv[2][0], lv[1][0], lv[2][1], lv[1][1], lv[0][1]
    else:
        raise ValueError(f'axis group needs more than 3 stride levels: {lv}')
    return m


def split_roles(t, roles):
    nb, n1, n2 = roles
    assert t.dim() == nb + n1 + n2, (t.shape, roles)
    sz, st = list(t.shape), list(t.stride())
    groups = []
    pos = 0
    for n in (nb, n1, n2):
        groups.append((sz[pos:pos + n], st[pos:pos + n]))
        pos += n
    return groups


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


class CudaPrims:
    """The product back end: every call lands in libmpdo_b200.so on the current CUDA stream."""

    name = 'cuda'

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError('MPDOSimulator (B200 build) needs a CUDA device; there is no CPU fallback.')
        self.lib = _lib.load()
        sm = C.c_int()
        smem = C.c_int()
        major = C.c_int()
        minor = C.c_int()
        _lib.check(self.lib.mpdo_device_info(C.byref(sm), C.byref(smem), C.byref(major), C.byref(minor)),
                   'mpdo_device_info')
        self.sm_count = sm.value
        self._desc_cache = {}

    # -- plumbing ---------------------------------------------------------------------------------
    @staticmethod
    def _stream():
        # raw handle of torch's current stream on the current device (thread-local, cheap)
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr())

    def launch_count(self):
        return int(self.lib.mpdo_launch_count())

    # -- contraction --------------------------------------------------------------------------------
    def contract(self, A, ra, B, rb, Cv, rc, conjA=False, conjB=False, acc64=None, alpha=1.0, beta=0.0,
                 hermitian=False):
        """Cv[b,i,j] = alpha * sum_k op(A[b,i,k]) op(B[b,k,j]) + beta * Cv[b,i,j] on strided views.
        hermitian: the caller guarantees a Hermitian result (Gram matrix): half of the tiles are computed."""
        # descriptors depend only on the geometry of the three views: cache them (the same few dozen geometries
        # recur in every layer, and building one costs more host time than launching the kernel)
        key = (A.shape, A.stride(), A.dtype, ra, B.shape, B.stride(), B.dtype, rb, Cv.shape, Cv.stride(), Cv.dtype, rc,
               conjA, conjB, acc64, alpha, beta, hermitian)
        hit = self._desc_cache.get(key)
        if hit is None:
            hit = self._build_desc(A, ra, B, rb, Cv, rc, conjA, conjB, acc64, alpha, beta, hermitian)
            if len(self._desc_cache) < 4096:
                self._desc_cache[key] = hit
        d, ksplit = hit
        if ksplit > 1:
            if beta == 0.0:
                Cv.zero_()
            elif beta != 1.0:
                Cv.mul_(beta)
        _lib.check(self.lib.mpdo_contract(C.byref(d), self._ptr(A), self._ptr(B), self._ptr(Cv), self._stream()),
                   'mpdo_contract')
        return Cv

    def _build_desc(self, A, ra, B, rb, Cv, rc, conjA, conjB, acc64, alpha, beta, hermitian=False):
        (ab, ai, ak), (bb, bk, bj), (cb, ci, cj) = split_roles(A, ra), split_roles(B, rb), split_roles(Cv, rc)
        M, K, N = _prod(ai[0]), _prod(ak[0]), _prod(bj[0])
        batch = _prod(ab[0])
        assert _prod(bk[0]) == K and _prod(ci[0]) == M and _prod(cj[0]) == N, (A.shape, B.shape, Cv.shape)
        assert _prod(bb[0]) == batch and _prod(cb[0]) == batch
        d = _lib.ContractDesc()
        d.M, d.N, d.K, d.batch = M, N, K, batch
        d.dtypeA, d.dtypeB, d.dtypeC = _DT[A.dtype], _DT[B.dtype], _DT[Cv.dtype]
        d.conjA, d.conjB = int(conjA), int(conjB)
        all64 = A.dtype == B.dtype == Cv.dtype == torch.complex64
        d.acc64 = int(acc64 if acc64 is not None else not all64)
        lai, lak = _levels(*ai), _levels(*ak)
        lbk, lbj = _levels(*bk), _levels(*bj)
        d.Ab, d.Ai, d.Ak = _idxmap(_levels(*ab)), _idxmap(lai), _idxmap(lak)
        d.Bb, d.Bk, d.Bj = _idxmap(_levels(*bb)), _idxmap(lbk), _idxmap(lbj)
        d.Cb, d.Ci, d.Cj = _idxmap(_levels(*cb)), _idxmap(_levels(*ci)), _idxmap(_levels(*cj))
        inner = lambda lv: lv[-1][1] if lv else 1 << 60
        d.a_kfast = int(inner(lak) <= inner(lai))
        d.b_jfast = int(inner(lbj) <= inner(lbk))
        tm = (M + 63) // 64
        tiles = (tm * (tm + 1) // 2 if hermitian else tm * ((N + 63) // 64)) * batch
        d.hermitian = int(bool(hermitian))
        ksplit = 1
        if K >= 256 and tiles < self.sm_count and Cv.is_contiguous() and beta in (0.0, 1.0):
            ksplit = max(1, min((K + 63) // 64, (2 * self.sm_count) // max(tiles, 1)))
        d.ksplit = ksplit
        d.alpha, d.beta = float(alpha), float(beta)
        return d, ksplit

    # -- single-qubit absorption --------------------------------------------------------------------
    def absorb_1q(self, T, G):
        """T [B,l,2,a,r] (contiguous), G [Bg,2,2,K] (Bg in {1,B}) -> [B,l,2,K*a,r]."""
        Bn, l, _, a, r = T.shape
        K = G.shape[-1]
        assert T.is_contiguous() and G.is_contiguous() and G.dtype == T.dtype
        out = torch.empty((Bn, l, 2, K * a, r), dtype=T.dtype, device=T.device)
        gstride = 0 if G.shape[0] == 1 else 4 * K
        _lib.check(self.lib.mpdo_absorb_1q(_DT[T.dtype], Bn, l, a, r, K, self._ptr(T), self._ptr(G), gstride,
                                           self._ptr(out), self._stream()), 'mpdo_absorb_1q')
        return out

    # -- small cores --------------------------------------------------------------------------------
    def _decompose(self, L, want_rows, normalize, zero_tol, tol, sweeps):
        Bn, n, m = L.shape
        assert L.dtype == torch.complex128 and L.is_contiguous()
        dev = L.device
        Y = torch.empty((Bn, n, m + n), dtype=torch.complex128, device=dev)
        work = torch.empty((Bn, 48), dtype=torch.int32, device=dev)
        s = torch.empty((Bn, n), dtype=torch.float64, device=dev)
        Z = torch.empty((Bn, n, n), dtype=torch.complex128, device=dev)
        Yn = torch.empty((Bn, n, m), dtype=torch.complex128, device=dev) if want_rows else None
        tol = max(tol, 4.4e-16 * (m ** 0.5))   # rounding level of a length-m dot product
        _lib.check(self.lib.mpdo_decompose_rows(Bn, n, m, self._ptr(L), self._ptr(Y), self._ptr(work), self._ptr(s),
                                                self._ptr(Yn) if want_rows else None, self._ptr(Z), int(normalize),
                                                float(zero_tol), float(tol), int(sweeps), self._stream()),
                   'mpdo_decompose_rows')
        return s, Yn, Z

    def eigh_psd(self, G, tol=1e-15, sweeps=30, rank_revealing=False, rel=1e-15):
        """Hermitian PSD G [B,n,n] (complex128) -> lam [B,n] descending, Vh [B,n,n] with G = Vh^h diag(lam) Vh.
        rank_revealing: pivoted-Cholesky preconditioned route (mpdo_eigh_psd); directions below rel * max diag come
        back as lam = 0 with zero rows of Vh. rank_revealing = 2: the blocked factorisation without pivoting where the
        shape allows (Gram matrices of fp32 data; see include/mpdo_b200.h). Otherwise the complete basis from Jacobi
        on [G | I]."""
        G = G.contiguous()
        Bn, n, _ = G.shape
        assert G.dtype == torch.complex128
        dev = G.device
        nbytes = int(self.lib.mpdo_eigh_psd_scratch_bytes(Bn, n))
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        lam = torch.empty((Bn, n), dtype=torch.float64, device=dev)
        Vh = torch.empty((Bn, n, n), dtype=torch.complex128, device=dev)
        tol = max(tol, 4.4e-16 * (n ** 0.5))
        _lib.check(self.lib.mpdo_eigh_psd(Bn, n, self._ptr(G), self._ptr(scratch), self._ptr(lam), self._ptr(Vh),
                                          int(rank_revealing), float(rel), float(tol), int(sweeps),
                                          self._stream()), 'mpdo_eigh_psd')
        return lam, Vh

    def chol_psd(self, G, rel=1e-14):
        """Rank-revealing pivoted Cholesky of Hermitian PSD G [B,n,n] (complex128): (Lh, Linv, rank) with
        G = Lh^h Lh and Linv . Lh^h = diag(1 x rank, 0) (see mpdo_chol_psd)."""
        G = G.contiguous()
        Bn, n, _ = G.shape
        assert G.dtype == torch.complex128
        dev = G.device
        scratch = torch.empty((int(self.lib.mpdo_chol_psd_scratch_bytes(Bn, n)),), dtype=torch.uint8, device=dev)
        Lh = torch.empty((Bn, n, n), dtype=torch.complex128, device=dev)
        Linv = torch.empty((Bn, n, n), dtype=torch.complex128, device=dev)
        rank = torch.empty((Bn,), dtype=torch.int32, device=dev)
        _lib.check(self.lib.mpdo_chol_psd(Bn, n, self._ptr(G), self._ptr(scratch), self._ptr(Lh), self._ptr(Linv),
                                          self._ptr(rank), float(rel), self._stream()), 'mpdo_chol_psd')
        return Lh, Linv, rank

    def svd_rows(self, L, tol=1e-15, sweeps=30, zero_tol=1e-300):
        """L [B,n,m] (complex128) = Uh^h diag(s) Wh -> (Uh [B,n,n], s [B,n] descending, Wh [B,n,m])."""
        s, Wh, Uh = self._decompose(L.contiguous(), True, 1, zero_tol, tol, sweeps)
        return Uh, s, Wh

    def rowscale(self, V, lam, rows, power, tol, mode, dtype):
        """X[b,j,:] = f(lam[b,j]) V[b,j,:], j < rows (see mpdo_rowscale)."""
        Bn, vrows, cols = V.shape
        assert V.is_contiguous() and V.dtype == torch.complex128 and lam.dtype == torch.float64
        X = torch.empty((Bn, rows, cols), dtype=dtype, device=V.device)
        _lib.check(self.lib.mpdo_rowscale(Bn, rows, cols, vrows, self._ptr(V), self._ptr(lam), lam.stride(0),
                                          float(power), float(tol), int(mode), _DT[dtype], self._ptr(X),
                                          self._stream()), 'mpdo_rowscale')
        return X

    def rank_rule(self, lam, squared, cap, max_err, relative, f32, zero_tail=True):
        """Kept ranks per batch entry (host list; SYNC). lam [B,n] float64, sorted descending."""
        Bn, n = lam.shape
        keep = torch.empty((Bn,), dtype=torch.int32, device=lam.device)
        _lib.check(self.lib.mpdo_rank_rule(Bn, n, self._ptr(lam), lam.stride(0), int(squared), int(cap),
                                           -1.0 if max_err is None else float(max_err), int(bool(relative)),
                                           int(f32), self._ptr(keep), int(zero_tail), self._stream()),
                   'mpdo_rank_rule')
        return keep.cpu().tolist()
