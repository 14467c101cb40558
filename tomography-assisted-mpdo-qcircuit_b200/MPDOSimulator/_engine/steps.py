"""The MPDO noisy-gate update path expressed over the device primitives (prims.py).

State: one dense site tensor T_k[B, l, s, a, r] per qubit (B independent circuits, bond l / r,
physical s = 2, inner "Kraus" index a). Everything the reference does through tensornetwork
(Circuit.py:74-224, TNNOptimizer.py:72-197) becomes a short sequence of batched contractions,
Gram matrices and small fp64 Jacobi decompositions:

* every tall/wide factor is orthogonalised through its Gram matrix (fp64) and a Jacobi
  eigen-decomposition of that small matrix ("Gram-eig"); one pass is exact to ~1e-8 relative, which
  is below complex64 resolution; complex128 states run two passes (the second on the already
  nearly-orthonormal factor), which restores fp64 backward stability, followed by a one-sided Jacobi
  SVD of the small core;
* the right-to-left chi sweep never forms the reference's (l*s*a)^2 two-site matrix: the left
  neighbour is an isometry after the QR sweep, so SVD(Q*X) = Q*SVD(X);
* the two-qubit gate split factors both sites first and decomposes only the (2x) x (2Ky) core.

Singular-vector gauges are not unique; all comparisons with the reference use gauge-invariant
quantities. The inner index order is (new Kraus index major, old index minor) - a pure relabelling of
an index that is only ever traced against its own conjugate.
"""
import os
import threading

import torch

C128 = torch.complex128


class Engine:
    def __init__(self, prims, dtype, npass=None):
        self.p = prims
        self.dtype = dtype
        self.f32 = dtype == torch.complex64
        self.npass = npass if npass is not None else (1 if self.f32 else 2)
        self.null_tol = 1e-14    # relative eigenvalue below which a Gram direction is treated as null
        self.floor_tol = 1e-13   # first-pass floor of the two-pass scheme
        self.jacobi_tol = 1e-10 if self.f32 else 1e-15   # fp32-stored states: see csrc/engine.cu Ctx
        self.stats = {'discarded': []}
        self._omega = {}         # fixed start blocks of the subspace iteration, per (n, block, device)
        self.tls = threading.local()   # per-thread results of the last call (strands run on several host threads)

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    def _empty(self, shape, like, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype, device=like.device)

    def _rr(self, G):
        """Preconditioned (rank-revealing) eigen-solver: complex64 states only (same policy as csrc/engine.cu Ctx;
        MPDO_BATCH_CLASSIC=1 keeps the classic batched Jacobi for batches of 32 and more)."""
        if os.environ.get('MPDO_BATCH_CLASSIC') and G.shape[0] >= 32:
            return False
        return self.npass == 1

    def _gram_cols(self, X, roles):
        """G[b,c,c'] = sum_rows conj(X[b,rows,c]) X[b,rows,c'] for a view X [b | rows | cols]."""
        nb, nr, nc = roles
        dims = list(range(X.dim()))
        At = X.permute(dims[:nb] + dims[nb + nr:] + dims[nb:nb + nr])
        n = 1
        for d in X.shape[nb + nr:]:
            n *= d
        Bn = 1
        for d in X.shape[:nb]:
            Bn *= d
        G = self._empty((Bn, n, n), X, C128)
        self.p.contract(At, (nb, nc, nr), X, (nb, nr, nc), G, (1, 1, 1), conjA=True, acc64=True, hermitian=True)
        return G

    def _gram_rows(self, M, roles):
        """G[b,i,i'] = sum_cols M[b,i,cols] conj(M[b,i',cols]) for a view M [b | rows | cols]."""
        nb, nr, nc = roles
        dims = list(range(M.dim()))
        Mt = M.permute(dims[:nb] + dims[nb + nr:] + dims[nb:nb + nr])
        n = 1
        for d in M.shape[nb:nb + nr]:
            n *= d
        Bn = 1
        for d in M.shape[:nb]:
            Bn *= d
        G = self._empty((Bn, n, n), M, C128)
        self.p.contract(M, (nb, nr, nc), Mt, (nb, nc, nr), G, (1, 1, 1), conjB=True, acc64=True, hermitian=True)
        return G

    def orth_cols(self, X, roles):
        """Tall view X [b | rows | cols] -> (Alast, Xs, R): Q = Alast . Xs^h is an isometry (zero columns
        for numerically null directions) and X = Q . R. Xs, R are [B, n, n] complex128."""
        p = self.p
        nb, nr, nc = roles
        G = self._gram_cols(X, roles)
        lam, Vh = p.eigh_psd(G, self.jacobi_tol, rank_revealing=self._rr(G))
        n = G.shape[1]
        if self.npass == 1:
            Xs = p.rowscale(Vh, lam, n, -0.5, self.null_tol, 0, C128)
            R = p.rowscale(Vh, lam, n, 0.5, self.null_tol, 0, C128)
            return X, Xs, R
        Xs1 = p.rowscale(Vh, lam, n, -0.5, self.floor_tol, 1, C128)
        R1 = p.rowscale(Vh, lam, n, 0.5, self.floor_tol, 1, C128)
        A1 = self._empty(tuple(X.shape), X)                      # same logical shape as the view X
        p.contract(X, roles, Xs1.permute(0, 2, 1), (1, 1, 1), A1, roles, conjB=True)
        G2 = self._gram_cols(A1, roles)
        lam2, Vh2 = p.eigh_psd(G2, self.jacobi_tol)
        Xs2 = p.rowscale(Vh2, lam2, n, -0.5, self.null_tol, 0, C128)
        R2 = p.rowscale(Vh2, lam2, n, 0.5, self.null_tol, 0, C128)
        R = self._empty(R1.shape, X, C128)
        p.contract(R2, (1, 1, 1), R1, (1, 1, 1), R, (1, 1, 1))
        return A1, Xs2, R

    def orth_rows(self, M, roles):
        """Wide view M [b | rows | cols] -> (Mlast, F, Lh): Qt = F . Mlast has orthonormal (or zero) rows and
        M = Lh^h . Qt. F, Lh are [B, l, l] complex128."""
        p = self.p
        nb, nr, nc = roles
        G = self._gram_rows(M, roles)
        lam, Uh = p.eigh_psd(G, self.jacobi_tol, rank_revealing=self._rr(G))
        n = G.shape[1]
        if self.npass == 1:
            F = p.rowscale(Uh, lam, n, -0.5, self.null_tol, 0, C128)
            Lh = p.rowscale(Uh, lam, n, 0.5, self.null_tol, 0, C128)
            return M, F, Lh
        F1 = p.rowscale(Uh, lam, n, -0.5, self.floor_tol, 1, C128)
        R1 = p.rowscale(Uh, lam, n, 0.5, self.floor_tol, 1, C128)
        M1 = self._empty(tuple(M.shape), M)                      # same logical shape as the view M
        p.contract(F1, (1, 1, 1), M, roles, M1, roles)
        G2 = self._gram_rows(M1, roles)
        lam2, Uh2 = p.eigh_psd(G2, self.jacobi_tol)
        F2 = p.rowscale(Uh2, lam2, n, -0.5, self.null_tol, 0, C128)
        R2 = p.rowscale(Uh2, lam2, n, 0.5, self.null_tol, 0, C128)
        Lh = self._empty(R1.shape, M, C128)
        p.contract(R2, (1, 1, 1), R1, (1, 1, 1), Lh, (1, 1, 1))
        return M1, F2, Lh

    def _svd_core(self, Lh):
        """L = Lh^h = U diag(s) Wh. Returns (Uh_L [B,n,n], s [B,n], Wh_L [B,n,n]) with rows sorted by s."""
        Uh_, s, Wh_ = self.p.svd_rows(Lh, self.jacobi_tol)
        return Wh_, s, Uh_

    def svd_wide(self, M, roles):
        """Truncatable SVD of a wide view M [b | rows | cols] = U diag(s) Vh without forming Vh:
        returns (Mlast, sv, squared, right, left) where sv [B,n] holds the singular values (their squares if
        `squared`), right(k, dtype) is the [B,k,n] core with sqrt(S_k) Vh_k = right . Mlast, and
        left(k, dtype)[j, i] = sqrt(s_j) conj(U[i, j]). One-pass mode reads both cores straight off the Gram
        eigen-decomposition; two-pass mode orthogonalises twice and runs the Jacobi SVD on the small core."""
        p = self.p
        if self.npass == 1:
            G = self._gram_rows(M, roles)
            lam, Uh = p.eigh_psd(G, self.jacobi_tol, rank_revealing=self._rr(G))
            right = lambda k, dt: p.rowscale(Uh, lam, k, -0.25, self.null_tol, 0, dt)
            left = lambda k, dt: p.rowscale(Uh, lam, k, 0.25, self.null_tol, 0, dt)
            return M, lam, True, right, left
        Mlast, F, Lh = self.orth_rows(M, roles)
        Uh_L, s, Wh_L = self._svd_core(Lh)

        def right(k, dt):
            WF = torch.empty((F.shape[0], k, F.shape[2]), dtype=dt, device=F.device)
            p.contract(p.rowscale(Wh_L, s, k, 0.5, 0.0, 0, C128), (1, 1, 1), F, (1, 1, 1), WF, (1, 1, 1))
            return WF

        left = lambda k, dt: p.rowscale(Uh_L, s, k, 0.5, 0.0, 0, dt)
        return Mlast, s, False, right, left

    def _keep(self, s, cap, max_err, relative, squared=False):
        """Apply the reference rank rule; returns the common (batch-max) kept rank. s is zero-tailed in place."""
        n = s.shape[1]
        cap = n if cap is None else min(int(cap), n)
        if max_err is None:
            return cap
        keep = self.p.rank_rule(s, squared, cap, max_err, relative, self.f32, True)
        return max(1, max(keep))

    # ------------------------------------------------------------------------------------------
    # R1: single-qubit gate / Kraus absorption          Circuit.py:138-178
    # ------------------------------------------------------------------------------------------
    def absorb_1q(self, T, G):
        """T [B,l,2,a,r], G [Bg,2,2,K] -> [B,l,2,K*a,r]."""
        return self.p.absorb_1q(T.contiguous(), G.contiguous())

    # ------------------------------------------------------------------------------------------
    # R2: two-qubit gate absorption and split           Circuit.py:74-136
    # ------------------------------------------------------------------------------------------
    def split_2q(self, Tlo, Thi, G, max_err=2.718281828459045e-8):
        """Tlo [B,l,2,a0,m], Thi [B,m,2,a1,r], G [Bg,2,2,2,2,K] as [p_lo,p_hi,s_lo,s_hi,g].
        Returns (Tlo' [B,l,2,a0,k], Thi' [B,k,2,K*a1,r]) = (U sqrt(S), sqrt(S) Vh) of the merged two-site
        tensor, rank chosen by the reference rule ||s|| - ||s[:k]|| <= max_err (absolute)."""
        p = self.p
        Bn, l, _, a0, m = Tlo.shape
        _, _, _, a1, r = Thi.shape
        K = G.shape[-1]
        dev = Tlo.device

        # left factor: rows (l,a0), cols (s0,m)
        Xlo = Tlo.permute(0, 1, 3, 2, 4)                        # [B,l,a0,2,m]
        if l * a0 > 2 * m:
            Alo, Xs_lo, Rp = self.orth_cols(Xlo, (1, 2, 2))     # Q' = Alo . Xs_lo^h ; Rp [B,x,(s0,m)]
            x = 2 * m
        else:
            x = l * a0
            Alo, Xs_lo = None, None
            Rp = Xlo.reshape(Bn, x, 2 * m).to(C128)
        # right factor: rows (m,s1), cols (a1,r)
        if self.npass == 1 and a1 * r > 2 * m and not os.environ.get('MPDO_SPLIT_VIA_CORE'):
            return self._split_2q_coreless(Tlo, Thi, G, max_err, Alo, Xs_lo, Rp, x)
        if a1 * r > 2 * m:
            Mhi, F_hi, Lh_hi = self.orth_rows(Thi, (1, 2, 2))   # Qt' = F_hi . Mhi ; Thi = Lh_hi^h . Qt'
            y = 2 * m
        else:
            y = a1 * r
            Mhi, F_hi = None, None
            Lh_hi = None
        # D[b,x,s0,s1,y] = sum_m R'[x,s0,m] L'[m,s1,y]
        D = torch.empty((Bn, x, 2, 2, y), dtype=C128, device=dev)
        Rp5 = Rp.reshape(Bn, x, 2, 1, m).expand(Bn, x, 2, 2, m).permute(0, 2, 3, 1, 4)     # [b,s0,s1 | x | m]
        if Lh_hi is not None:
            # L'[(m,s1),y] = conj(Lh_hi[y,(m,s1)])
            Lv = Lh_hi.reshape(Bn, y, m, 1, 2).expand(Bn, y, m, 2, 2).permute(0, 3, 4, 2, 1)  # [b,s0,s1 | m | y]
            p.contract(Rp5, (3, 1, 1), Lv, (3, 1, 1), D.permute(0, 2, 3, 1, 4), (3, 1, 1), conjB=True)
        else:
            Lv = Thi.reshape(Bn, m, 1, 2, y).expand(Bn, m, 2, 2, y).permute(0, 2, 3, 1, 4)    # [b,s0,s1 | m | y]
            p.contract(Rp5, (3, 1, 1), Lv, (3, 1, 1), D.permute(0, 2, 3, 1, 4), (3, 1, 1))
        # Cm[b,x,p0,p1,g,y] = sum_{s0,s1} G[p0,p1,s0,s1,g] D[b,x,s0,s1,y]
        Gp = G.permute(0, 1, 2, 5, 3, 4).contiguous().to(C128)                               # [Bg,p0,p1,g,s0,s1]
        Cm = torch.empty((Bn, x, 2, 2, K, y), dtype=C128, device=dev)
        GpE = Gp.reshape(Gp.shape[0], 1, 4 * K, 4).expand(Bn, x, 4 * K, 4)
        p.contract(GpE, (2, 1, 1), D.reshape(Bn, x, 4, y), (2, 1, 1), Cm.reshape(Bn, x, 4 * K, y), (2, 1, 1))
        # SVD of the core as rows (x,p0) x cols (p1,g,y)
        Cv = Cm.reshape(Bn, 2 * x, 2 * K * y)
        nrow, ncol = 2 * x, 2 * K * y
        if nrow <= ncol:
            Mlast, sv, sq, right, left = self.svd_wide(Cv, (1, 1, 1))
            k = self._keep(sv, None, max_err, False, squared=sq)
            UL = left(k, C128)                                   # [B,k,(x,p0)] = sqrt(s_j) conj(U[(x,p0), j])
            Zc = torch.empty((Bn, k, 2, K, y), dtype=C128 if F_hi is not None else self.dtype, device=dev)
            p.contract(right(k, C128), (1, 1, 1), Mlast, (1, 1, 1), Zc.reshape(Bn, k, ncol), (1, 1, 1))
        else:
            # tall core (rare: tiny right factor): decompose the column side instead
            Alast, Xs, R = self.orth_cols(Cv, (1, 1, 1))         # Cv = (Alast Xs^h) R, R [B,ncol,ncol]
            Uh_, s, Wh_ = p.svd_rows(R, self.jacobi_tol)          # R = Uh_^h diag(s) Wh_
            k = self._keep(s, None, max_err, False)
            # U_L = Q . Uh_^h  ->  UL rows: sqrt(s_j) conj(U_L[:, j])
            XU = torch.empty((Bn, k, ncol), dtype=C128, device=dev)
            # conj(U_L[i,j]) = sum_{c,d} Uh_[j,c] Xs[c,d] conj(Alast[i,d])
            p.contract(p.rowscale(Uh_, s, k, 0.5, 0.0, 0, C128), (1, 1, 1), Xs, (1, 1, 1), XU, (1, 1, 1))
            UL = torch.empty((Bn, k, nrow), dtype=C128, device=dev)
            p.contract(XU, (1, 1, 1), Alast.permute(0, 2, 1), (1, 1, 1), UL, (1, 1, 1), conjB=True)
            Zc = p.rowscale(Wh_, s, k, 0.5, 0.0, 0, C128 if F_hi is not None else self.dtype).reshape(Bn, k, 2, K, y)
        self.stats['last_rank'] = k
        self.tls.last_ranks = None       # per-entry ranks are only reported by the native engine

        # Tlo'[b,l,p0,a0,j] = sum_x Q'[(l,a0),x] conj(UL[j,(x,p0)])
        Tlo_n = self._empty((Bn, l, 2, a0, k), Tlo)
        ULv = UL.reshape(Bn, k, x, 2)
        if Alo is not None:
            # W[b,c,(p0,j)] = sum_x conj(Xs_lo[x,c]) conj(UL[j,x,p0]);   Tlo' = Alo . W
            W = torch.empty((Bn, 2 * m, 2, k), dtype=self.dtype, device=dev)
            p.contract(Xs_lo.permute(0, 2, 1), (1, 1, 1), ULv.permute(0, 2, 3, 1), (1, 1, 2), W.reshape(Bn, 2 * m, 2 * k),
                       (1, 1, 1), conjA=True, conjB=True)
            p.contract(Alo, (1, 2, 2), W.reshape(Bn, 2, m, 2, k), (1, 2, 2), Tlo_n.permute(0, 1, 3, 2, 4), (1, 2, 2))
        else:
            Tlo_n.copy_(ULv.reshape(Bn, k, l, a0, 2).permute(0, 2, 4, 3, 1).conj())
        # Thi'[b,j,p1,(g,a1),r] = sum_y Zc[j,p1,g,y] Qt'[y,(a1,r)]
        if F_hi is not None:
            # Zc . F_hi -> [B,k,2,K,(m,s1)] then times Mhi [b | m,s1 | a1,r]
            ZF = torch.empty((Bn, k * 2 * K, 2 * m), dtype=self.dtype, device=dev)
            p.contract(Zc.reshape(Bn, k * 2 * K, y), (1, 1, 1), F_hi, (1, 1, 1), ZF, (1, 1, 1))
            Thi_n = self._empty((Bn, k, 2, K * a1, r), Thi)
            p.contract(ZF.reshape(Bn, k * 2 * K, m, 2), (1, 1, 2), Mhi, (1, 2, 2),
                       Thi_n.reshape(Bn, k * 2 * K, a1, r), (1, 1, 2))
        else:
            Thi_n = Zc.reshape(Bn, k, 2, K * a1, r)
            if Thi_n.dtype != self.dtype:
                Thi_n = Thi_n.to(self.dtype)
        return Tlo_n, Thi_n

    def _split_2q_coreless(self, Tlo, Thi, G, max_err, Alo, Xs_lo, Rp, x):
        """complex64 states, wide right factor (csrc/engine.cu mpdo_split_2q, same sequence): the (2x) x (2Ky) core is
        never formed. With Lam = T_hi T_hi^h over (a1,r) and Gam[p0 s0 s1; p0' s0' s1'] = sum_{p1,g} G conj(G),
          C C^h[(x,p0),(x',p0')] = sum R'[x,s0,m] conj(R'[x',s0',m']) Lam[(m,s1),(m',s1')] Gam[p0 s0 s1; p0' s0' s1'],
        and sqrt(S) Vh is rebuilt from the site itself:
          T_hi'[j,p1,(g,a1),r] = sum_{m,s1} Y[j,p1,g,m,s1] T_hi[m,s1,a1,r],
          Y = sum_{x,p0,s0} W[j,x,p0] R'[x,s0,m] G[p0,p1,s0,s1,g],  W = S^-1/4 U^h."""
        p = self.p
        Bn, l, _, a0, m = Tlo.shape
        _, _, _, a1, r = Thi.shape
        Bg, K = G.shape[0], G.shape[-1]
        dev = Tlo.device
        Lam = self._gram_rows(Thi, (1, 2, 2))                                   # [B, (m,s1), (m',s1')]
        T1 = torch.empty((Bn, 2, x * 2, 2 * m), dtype=C128, device=dev)         # [b, s1, (x,s0), (m',s1')]
        p.contract(Rp.reshape(Bn, 1, x * 2, m).expand(Bn, 2, x * 2, m), (2, 1, 1),
                   Lam.reshape(Bn, m, 2, 2 * m).permute(0, 2, 1, 3), (2, 1, 1), T1, (2, 1, 1))
        T2s = torch.empty((Bn, x, x, 2, 2, 2, 2), dtype=C128, device=dev)       # [b, x, x', s0, s1, s0', s1']
        p.contract(T1.reshape(Bn, 2, x, 2, m, 2).permute(0, 5, 1, 2, 3, 4), (2, 3, 1),
                   Rp.reshape(Bn, 1, x, 2, m).expand(Bn, 2, x, 2, m).permute(0, 1, 4, 2, 3), (2, 1, 2),
                   T2s.permute(0, 6, 4, 1, 3, 2, 5), (2, 3, 2), conjB=True)
        Gc = G.to(C128).contiguous()                                            # [bg, p0, p1, s0, s1, g]
        Gam = torch.empty((Bg, 8, 8), dtype=C128, device=dev)
        p.contract(Gc.permute(0, 1, 3, 4, 2, 5), (1, 3, 2), Gc.permute(0, 2, 5, 1, 3, 4), (1, 2, 3), Gam, (1, 1, 1),
                   conjB=True)
        GG = torch.empty((Bn, x, 2, x, 2), dtype=C128, device=dev)              # [b, (x,p0), (x',p0')]
        GamV = Gam.reshape(Bg, 2, 2, 2, 2, 2, 2).permute(0, 1, 4, 2, 3, 5, 6)
        if Bg == 1:
            GamV = GamV.expand(Bn, 2, 2, 2, 2, 2, 2)
        p.contract(GamV, (1, 2, 4), T2s.reshape(Bn, x * x, 16).permute(0, 2, 1), (1, 1, 1),
                   GG.permute(0, 2, 4, 1, 3), (1, 2, 2))
        GGm = GG.reshape(Bn, 2 * x, 2 * x)
        lam, Uh = p.eigh_psd(GGm, self.jacobi_tol, rank_revealing=self._rr(GGm))
        k = self._keep(lam, None, max_err, False, squared=True)
        self.stats['last_rank'] = k
        self.tls.last_ranks = None
        UL = p.rowscale(Uh, lam, k, 0.25, self.null_tol, 0, C128)               # sqrt(s_j) conj(U[(x,p0), j])
        Wr = p.rowscale(Uh, lam, k, -0.25, self.null_tol, 0, C128)
        V = torch.empty((Bn, k, 2, 2, m), dtype=C128, device=dev)               # [b, j, p0, s0, m]
        p.contract(Wr.reshape(Bn, k, x, 2).permute(0, 3, 1, 2), (2, 1, 1),
                   Rp.reshape(Bn, 1, x, 2 * m).expand(Bn, 2, x, 2 * m), (2, 1, 1), V.permute(0, 2, 1, 3, 4), (2, 1, 2))
        Gq = Gc.permute(0, 2, 5, 4, 1, 3).contiguous()                          # [bg, p1, g, s1, p0, s0]
        Y = torch.empty((Bn, k, 2, K, m, 2), dtype=self.dtype, device=dev)      # [b, j, p1, g, m, s1]
        GqV = Gq.reshape(Bg, 1, 4 * K, 4).expand(Bn, k, 4 * K, 4)
        p.contract(GqV, (2, 1, 1), V.reshape(Bn, k, 4, m), (2, 1, 1), Y.permute(0, 1, 2, 3, 5, 4), (2, 3, 1))
        Tlo_n = self._empty((Bn, l, 2, a0, k), Tlo)
        ULv = UL.reshape(Bn, k, x, 2)
        if Alo is not None:
            W = torch.empty((Bn, 2 * m, 2, k), dtype=self.dtype, device=dev)
            p.contract(Xs_lo.permute(0, 2, 1), (1, 1, 1), ULv.permute(0, 2, 3, 1), (1, 1, 2), W.reshape(Bn, 2 * m, 2 * k),
                       (1, 1, 1), conjA=True, conjB=True)
            p.contract(Alo, (1, 2, 2), W.reshape(Bn, 2, m, 2, k), (1, 2, 2), Tlo_n.permute(0, 1, 3, 2, 4), (1, 2, 2))
        else:
            Tlo_n.copy_(ULv.reshape(Bn, k, l, a0, 2).permute(0, 2, 4, 3, 1).conj())
        Thi_n = self._empty((Bn, k, 2, K * a1, r), Thi)
        p.contract(Y.reshape(Bn, k * 2 * K, m, 2), (1, 1, 2), Thi, (1, 2, 2), Thi_n.reshape(Bn, k * 2 * K, a1, r), (1, 1, 2))
        return Tlo_n, Thi_n

    # ------------------------------------------------------------------------------------------
    # R4: one step of the left-to-right QR sweep        TNNOptimizer.py:87-108
    # ------------------------------------------------------------------------------------------
    def qr_step(self, Ti, Tn):
        """Ti [B,l,2,a,r] -> isometry Q (same shape); Tn [B,r,2,a',r'] <- R . Tn."""
        p = self.p
        Bn, l, _, a, r = Ti.shape
        Alast, Xs, R = self.orth_cols(Ti, (1, 3, 1))
        Q = self._empty(Ti.shape, Ti)
        p.contract(Alast, (1, 3, 1), Xs.to(self.dtype).permute(0, 2, 1), (1, 1, 1), Q, (1, 3, 1), conjB=True)
        Tn_new = self._empty(Tn.shape, Tn)
        p.contract(R.to(self.dtype), (1, 1, 1), Tn, (1, 1, 3), Tn_new, (1, 1, 3))
        return Q, Tn_new

    # ------------------------------------------------------------------------------------------
    # R5: one step of the right-to-left bond (chi) truncation      TNNOptimizer.py:111-134
    # ------------------------------------------------------------------------------------------
    def bond_svd_step(self, Tl, Tr, chi, max_err=None):
        """Tl [B,l',2,a',l] (left-isometric), Tr [B,l,2,a,r]. Theta = Tl.Tr is truncated to chi singular
        values (and optionally to the relative error max_err); returns (Tl.U sqrt(S), sqrt(S) Vh, s_discarded)."""
        p = self.p
        Bn, l, _, a, r = Tr.shape
        Mlast, sv, sq, right, left = self.svd_wide(Tr, (1, 1, 3))
        k = self._keep(sv, chi, max_err, True, squared=sq)
        disc = sv[:, k:].clamp_min(0).sqrt() if sq else sv[:, k:].clone()
        # Tr' = sqrt(S_k) Vh_k = right . Mlast
        Tr_n = self._empty((Bn, k, 2, a, r), Tr)
        p.contract(right(k, self.dtype), (1, 1, 1), Mlast, (1, 1, 3), Tr_n, (1, 1, 3))
        # Tl' = Tl . U[:, :k] sqrt(S_k):  U[r,j] sqrt(s_j) = conj(UL[j,r])
        UL = left(k, self.dtype)
        Tl_n = self._empty(tuple(Tl.shape[:4]) + (k,), Tl)
        p.contract(Tl, (1, 3, 1), UL.permute(0, 2, 1), (1, 1, 1), Tl_n, (1, 3, 1), conjB=True)
        return Tl_n, Tr_n, disc

    # ------------------------------------------------------------------------------------------
    # R4 + R5 through left environments (complex64 states): the same truncation, one sequential sweep instead of two
    # ------------------------------------------------------------------------------------------
    def env_step(self, E, T):
        """E [B,l,l] (complex128) = A^h A of the block left of the bond, T [B,l,2,a,r] ->
        E'[r,r'] = sum conj(T[l,s,a,r]) E[l,l'] T[l',s,a,r'] (the same matrix the QR sweep forms as the Gram matrix of
        R.T, TNNOptimizer.py:87-108, without the factorisation in between)."""
        p = self.p
        Bn, l, _, a, r = T.shape
        X = torch.empty((Bn, l, 2, a, r), dtype=C128, device=T.device)
        p.contract(E, (1, 1, 1), T, (1, 1, 3), X, (1, 1, 3))
        En = torch.empty((Bn, r, r), dtype=C128, device=T.device)
        p.contract(T.permute(0, 4, 1, 2, 3), (1, 1, 3), X, (1, 3, 1), En, (1, 1, 1), conjA=True)
        return En

    def env_factor(self, E):
        """E = C^h C and the left inverse Ci (Ci . C^h = 1 on the numerical range): the R of the QR sweep up to a
        unitary from the left, which no later quantity depends on."""
        p = self.p
        if hasattr(p, 'chol_psd') and not os.environ.get('MPDO_NO_CHOLQR'):
            try:
                Lh, Linv, _ = p.chol_psd(E, self.null_tol)
                return Lh, Linv
            except RuntimeError:      # shape not schedulable: eigen route
                pass
        lam, Vh = p.eigh_psd(E, self.jacobi_tol, rank_revealing=self._rr(E))
        n = E.shape[1]
        return (p.rowscale(Vh, lam, n, 0.5, self.null_tol, 0, C128), p.rowscale(Vh, lam, n, -0.5, self.null_tol, 0, C128))

    def bond_env_step(self, M0, W, Ci, chi):
        """One step of the right-to-left chi truncation on the un-canonicalised state. M0 [B,l,2,a,r0] = C . T is the
        site right of the bond times the factor of the bond's left environment, W [B,r0,r] what the previous step left
        for the right index (None at the last site), Ci the left inverse of the factor. With M = M0 . W the two-site
        matrix of the reference (TNNOptimizer.py:111-134) is Q . M for an isometry Q, so its truncated SVD is
        Q U | S | V^h with M = U S V^h: returns (sqrt(S_k) V_k^h [B,k,2,a,r], W' = C^+ U_k sqrt(S_k) [B,l,k],
        discarded values)."""
        p = self.p
        Bn, l, _, a, _ = M0.shape
        if W is not None:
            M = self._empty((Bn, l, 2, a, W.shape[2]), M0)
            p.contract(M0, (1, 3, 1), W, (1, 1, 1), M, (1, 3, 1))
        else:
            M = M0
        r = M.shape[4]
        G = self._gram_rows(M, (1, 1, 3))
        lam, Uh = p.eigh_psd(G, self.jacobi_tol, rank_revealing=self._rr(G))
        k = l if chi is None else min(int(chi), l)
        disc = lam[:, k:].clamp_min(0).sqrt()
        T_n = self._empty((Bn, k, 2, a, r), M0)
        p.contract(p.rowscale(Uh, lam, k, -0.25, self.null_tol, 0, self.dtype), (1, 1, 1), M, (1, 1, 3), T_n, (1, 1, 3))
        UL = p.rowscale(Uh, lam, k, 0.25, self.null_tol, 0, C128)               # sqrt(s_j) conj(U[i, j])
        W_n = self._empty((Bn, l, k), M0)                                        # sum_i conj(Ci[i,i']) conj(UL[j,i])
        p.contract(Ci.permute(0, 2, 1), (1, 1, 1), UL.permute(0, 2, 1), (1, 1, 1), W_n, (1, 1, 1), conjA=True, conjB=True)
        return T_n, W_n, disc

    def bond_truncate_env(self, Ts, chi, publish=None):
        """bondTruncate (QR sweep + chi sweep, TNNOptimizer.py:72-134) for complex64 states with a fixed chi, in place
        on the list of site tensors. The left-to-right pass only chains the environments E_i (contractions); their
        factorisations are independent of one another; the right-to-left pass is the only sequential chain of
        decompositions left. Same singular values and the same sqrt(S) | sqrt(S) split at every bond as the two-sweep
        form (the factors differ from its R by unitaries that cancel)."""
        p = self.p
        n = len(Ts)
        if n < 2:
            return []
        Bn = Ts[0].shape[0]
        E = torch.ones((Bn, 1, 1), dtype=C128, device=Ts[0].device)
        if Ts[0].shape[1] != 1:
            raise ValueError('the first site must have a trivial left bond')
        factors = [None] * n
        for i in range(n - 1):
            E = self.env_step(E, Ts[i])
            C, Ci = self.env_factor(E)
            M0 = self._empty(tuple(Ts[i + 1].shape), Ts[i + 1])      # C . T of the site right of the bond: independent
            p.contract(C.to(self.dtype), (1, 1, 1), Ts[i + 1], (1, 1, 3), M0, (1, 1, 3))     # of the sweep that follows
            factors[i + 1] = (M0, Ci)
        W, disc = None, []
        for idx in range(n - 1, 0, -1):
            M0, Ci = factors[idx]
            Ts[idx], W, d = self.bond_env_step(M0, W, Ci, chi)
            factors[idx] = M0 = Ci = None
            if publish is not None:     # the caller takes the finished site at once (and drops the one it replaces)
                publish(idx, Ts[idx])
            disc.append(d)
        T0 = Ts[0]
        out = self._empty(tuple(T0.shape[:4]) + (W.shape[2],), T0)
        p.contract(T0, (1, 3, 1), W, (1, 1, 1), out, (1, 3, 1))
        Ts[0] = out
        if publish is not None:
            publish(0, out)
        return disc

    # ------------------------------------------------------------------------------------------
    # R3: inner-index (kappa) truncation                 TNNOptimizer.py:164-197
    # ------------------------------------------------------------------------------------------
    def eigh_topk(self, G, k, max_iter=None):
        """Leading k eigenpairs of Hermitian PSD G [B,n,n] (complex128) by block subspace iteration with
        Rayleigh-Ritz, iterated until the residuals |G v - theta v| of the k kept pairs are at rounding level
        (this is an exact solver run to convergence, unlike the reference's three un-orthonormalised power
        steps; it replaces the full Jacobi decomposition when k << n). Returns (theta [B,k], Vt [B,k,n]) with
        Vt[j,:] the components of eigenvector j (not conjugated), or None when it does not converge."""
        p = self.p
        Bn, n, _ = G.shape
        blk = min(n, max(2 * k, k + 28))
        tol = 1e-10 if self.f32 else 1e-12
        # one copy per CUDA stream: strands run this concurrently on different streams
        stream_id = torch.cuda.current_stream(G.device).cuda_stream if G.is_cuda else 0
        key = (n, blk, str(G.device), stream_id)
        if key not in self._omega:
            g = torch.Generator().manual_seed(20261017)
            om = torch.complex(torch.randn(blk, n, generator=g, dtype=torch.float64),
                               torch.randn(blk, n, generator=g, dtype=torch.float64))
            self._omega[key] = om.to(G.device)
        Yr = self._omega[key].unsqueeze(0).expand(Bn, blk, n)
        Gt = G.permute(0, 2, 1)
        Zr = torch.empty((Bn, blk, n), dtype=C128, device=G.device)
        p.contract(Yr, (1, 1, 1), Gt, (1, 1, 1), Zr, (1, 1, 1))
        if max_iter is None:   # short leash: the caller's full decomposition is the better policy for stalled clusters
            max_iter = 6 if self.f32 else 12
        prev_worst = float('inf')
        for it in range(max_iter):
            # orthonormalise the rows of Zr
            H = self._gram_rows(Zr, (1, 1, 1))
            lam, Uh = p.eigh_psd(H, self.jacobi_tol, rank_revealing=self._rr(H))
            Yr = torch.empty((Bn, blk, n), dtype=C128, device=G.device)
            p.contract(p.rowscale(Uh, lam, blk, -0.5, self.null_tol, 0, C128), (1, 1, 1), Zr, (1, 1, 1), Yr, (1, 1, 1))
            Zr = torch.empty((Bn, blk, n), dtype=C128, device=G.device)
            p.contract(Yr, (1, 1, 1), Gt, (1, 1, 1), Zr, (1, 1, 1))
            # Rayleigh-Ritz in the block: Bm = Zr Yr^h (Hermitian PSD), rotate both to the Ritz basis
            Bm = torch.empty((Bn, blk, blk), dtype=C128, device=G.device)
            p.contract(Zr, (1, 1, 1), Yr.permute(0, 2, 1), (1, 1, 1), Bm, (1, 1, 1), conjB=True)
            theta, Wh = p.eigh_psd(Bm, self.jacobi_tol, rank_revealing=self._rr(Bm))
            Yn = torch.empty_like(Yr)
            Zn = torch.empty_like(Zr)
            p.contract(Wh, (1, 1, 1), Yr, (1, 1, 1), Yn, (1, 1, 1))
            p.contract(Wh, (1, 1, 1), Zr, (1, 1, 1), Zn, (1, 1, 1))
            Yr, Zr = Yn, Zn
            # residual rows of the kept pairs: Zr[:k] - theta * Yr[:k]
            Rr = Zr[:, :k, :].clone()
            p.contract(p.rowscale(torch.eye(k, dtype=C128, device=G.device).expand(Bn, k, k).contiguous(), theta, k,
                                  1.0, 0.0, 0, C128), (1, 1, 1), Yr[:, :k, :], (1, 1, 1), Rr, (1, 1, 1), alpha=-1.0, beta=1.0)
            Rg = self._gram_rows(Rr, (1, 1, 1))
            res = Rg.diagonal(dim1=1, dim2=2).real.clamp_min(0).sqrt()
            worst = (res.max(dim=1).values / theta[:, 0].clamp_min(1e-300)).max().item()   # SYNC
            self.stats['topk_iters'] = it + 1
            if worst <= tol:
                return theta[:, :k].contiguous(), Yr[:, :k, :].contiguous()
            if it >= 1 and worst > 0.05 * prev_worst:   # stalled: the cut sits inside a cluster wider than the block
                return None
            prev_worst = worst
        return None

    def kappa_truncate(self, T, kappa, max_err=None):
        """T [B,l,2,a,r] -> U.S over the inner index: [B,l,2,k,r], k = min(kappa, a) (and the relative rule)."""
        p = self.p
        Bn, l, _, a, r = T.shape
        Tv = T.permute(0, 1, 2, 4, 3)                           # [b | l,s,r | a]
        G = self._gram_cols(Tv, (1, 3, 1))                      # G = T^h T over (l,s,r)
        if max_err is None and kappa is not None and a >= 64 and a >= 8 * kappa:
            top = self.eigh_topk(G, kappa)
            if top is not None:
                theta, Vt = top
                tr = G.diagonal(dim1=1, dim2=2).real.sum(dim=1)
                disc = (tr - theta.sum(dim=1)).clamp_min(0).sqrt().unsqueeze(1)   # norm of what was discarded
                T_n = self._empty((Bn, l, 2, kappa, r), T)
                # T'[l,s,j,r] = sum_a V[a,j] T[l,s,a,r]
                p.contract(Vt.to(self.dtype), (1, 1, 1), T.permute(0, 3, 1, 2, 4), (1, 1, 3),
                           T_n.permute(0, 3, 1, 2, 4), (1, 1, 3))
                return T_n, disc
        lam, Vh = p.eigh_psd(G, self.jacobi_tol, rank_revealing=self._rr(G))
        k = self._keep(lam, kappa, max_err, True, squared=True)
        disc = lam[:, k:].clamp_min(0).sqrt()
        # T'[l,s,j,r] = sum_a conj(Vh[j,a]) T[l,s,a,r]
        Vk = Vh[:, :k, :].to(self.dtype)
        T_n = self._empty((Bn, l, 2, k, r), T)
        p.contract(Vk, (1, 1, 1), T.permute(0, 3, 1, 2, 4), (1, 1, 3), T_n.permute(0, 3, 1, 2, 4), (1, 1, 3), conjA=True)
        return T_n, disc

    # ------------------------------------------------------------------------------------------
    # truncate layer = bondTruncate + svdKappa          Circuit.py:476-481
    # ------------------------------------------------------------------------------------------
    def qr_left2right(self, Ts):
        for i in range(len(Ts) - 1):
            Ts[i], Ts[i + 1] = self.qr_step(Ts[i], Ts[i + 1])

    def svd_right2left(self, Ts, chi, max_err=None):
        disc = []
        for i in range(len(Ts) - 1, 0, -1):
            Ts[i - 1], Ts[i], d = self.bond_svd_step(Ts[i - 1], Ts[i], chi, max_err)
            disc.append(d)
        return disc

    def svd_kappa(self, Ts, kappa, max_err=None, has_inner=None):
        for i, T in enumerate(Ts):
            if kappa is not None and ((has_inner is not None and not has_inner[i]) or T.shape[3] <= kappa):
                continue
            Ts[i], _ = self.kappa_truncate(T, kappa, max_err)

    # ------------------------------------------------------------------------------------------
    # R7: readout by transfer matrices                   Circuit.py:226-332, dmOperations.py
    # ------------------------------------------------------------------------------------------
    def transfer(self, L, T, Tc=None, op=None):
        """L [B,l,l'] (c128) -> L'[r,r'] = sum L[l,l'] (O.T)[l,s',a,r] conj(Tc[l',s',a,r']),
        (O.T)[s'] = sum_s O[s',s] T[s]; O = identity if op is None; Tc = T if None."""
        p = self.p
        Tc = T if Tc is None else Tc
        Bn, l, _, a, r = T.shape
        X = torch.empty((Bn, Tc.shape[1], 2, a, r), dtype=C128, device=T.device)
        p.contract(L.permute(0, 2, 1), (1, 1, 1), T, (1, 1, 3), X, (1, 1, 3))
        if op is not None:
            X = p.absorb_1q(X, op.reshape(-1, 2, 2, 1).to(C128).contiguous())
        Ln = torch.empty((Bn, r, Tc.shape[4]), dtype=C128, device=T.device)
        p.contract(X.permute(0, 4, 1, 2, 3), (1, 1, 3), Tc, (1, 3, 1), Ln, (1, 1, 1), conjB=True)
        return Ln

    def transfer_proj(self, L, T, bit):
        """Same with both copies projected on |bit><bit| (bit may differ per batch entry: LongTensor [B])."""
        p = self.p
        Bn, l, _, a, r = T.shape
        nL = L.shape[0]
        Ln = torch.empty((nL, r, r), dtype=C128, device=T.device)
        bits = torch.as_tensor(bit, device='cpu').reshape(-1).expand(nL) if not isinstance(bit, int) else None
        groups = [(None, bit)] if bits is None else [((bits == v).nonzero().reshape(-1), v) for v in (0, 1)]
        for idx, v in groups:
            if idx is not None and idx.numel() == 0:
                continue
            Lg = L if idx is None else L.index_select(0, idx.to(L.device))
            n = Lg.shape[0]
            Tb = T[:, :, v]                                      # [B,l,a,r]
            if Bn == 1:
                Tb = Tb.expand(n, l, a, r)
            elif idx is not None:
                Tb = Tb.index_select(0, idx.to(T.device))
            X = torch.empty((n, l, a, r), dtype=C128, device=T.device)
            p.contract(Lg.permute(0, 2, 1), (1, 1, 1), Tb, (1, 1, 2), X, (1, 1, 2))
            out = torch.empty((n, r, r), dtype=C128, device=T.device)
            p.contract(X.permute(0, 3, 1, 2), (1, 1, 2), Tb, (1, 2, 1), out, (1, 1, 1), conjB=True)
            if idx is None:
                Ln = out
            else:
                Ln.index_copy_(0, idx.to(L.device), out)
        return Ln

    def chain_value(self, Ts, ops=None, Tcs=None):
        """Tr(prod_k O_k rho) for the un-normalised MPDO (ops: {site: 2x2 tensor}); returns [B] complex128."""
        ops = ops or {}
        Bn = Ts[0].shape[0]
        L = torch.ones((Bn, 1, 1), dtype=C128, device=Ts[0].device)
        for k, T in enumerate(Ts):
            L = self.transfer(L, T, None if Tcs is None else Tcs[k], ops.get(k))
        return L.reshape(Bn)

    def chain_overlap(self, Ts0, Ts1):
        """Tr(rho_0 rho_1) of two un-normalised MPDOs on the same register -> [B] complex128 (O(chi^5))."""
        p = self.p
        Bn = Ts0[0].shape[0]
        dev = Ts0[0].device
        E = torch.ones((Bn, 1, 1, 1, 1), dtype=C128, device=dev)     # [b, l0, l0', l1, l1']
        for A, T1 in zip(Ts0, Ts1):
            _, l0, _, a0, r0 = A.shape
            _, l1, _, a1, r1 = T1.shape
            X1 = torch.empty((Bn, l0, l1, l1, 2, a0, r0), dtype=C128, device=dev)
            p.contract(E.permute(0, 2, 3, 4, 1), (1, 3, 1), A, (1, 1, 3), X1, (1, 3, 3))
            X2 = torch.empty((Bn, l1, l1, 2, r0, 2, r0), dtype=C128, device=dev)
            p.contract(X1.permute(0, 2, 3, 4, 6, 1, 5), (1, 4, 2), A.permute(0, 1, 3, 2, 4), (1, 2, 2), X2, (1, 4, 2),
                       conjB=True)
            X3 = torch.empty((Bn, l1, 2, r0, r0, a1, r1), dtype=C128, device=dev)
            p.contract(X2.permute(0, 2, 3, 4, 6, 1, 5), (1, 4, 2), T1, (1, 2, 2), X3, (1, 4, 2))
            En = torch.empty((Bn, r0, r0, r1, r1), dtype=C128, device=dev)
            p.contract(X3.permute(0, 3, 4, 6, 1, 2, 5), (1, 3, 3), T1, (1, 3, 1), En, (1, 3, 1), conjB=True)
            E = En
        return E.reshape(Bn)

    def chain_value_proj(self, Ts, bits):
        """<b|rho_c|b> of ONE bitstring b (list of 0/1) for every circuit c of the batch -> [B] float64."""
        Bn = Ts[0].shape[0]
        L = torch.ones((Bn, 1, 1), dtype=C128, device=Ts[0].device)
        for k, T in enumerate(Ts):
            L = self.transfer_proj(L, T, int(bits[k]))
        return L.reshape(Bn).real

    def bitstring_probs(self, Ts, bits):
        """<b|rho|b> for bitstrings bits [NB, n] (0/1 ints) of one circuit (B = 1) -> [NB] float64."""
        bits = torch.as_tensor(bits).reshape(-1, len(Ts))
        nb = bits.shape[0]
        L = torch.ones((nb, 1, 1), dtype=C128, device=Ts[0].device)
        for k, T in enumerate(Ts):
            L = self.transfer_proj(L, T, bits[:, k])
        return L.reshape(nb).real

    def transfer_right(self, R, T, op=None):
        """R [B,r,r'] (c128) -> R'[l,l'] = sum (O.T)[l,s',a,r] R[r,r'] conj(T[l',s',a,r'])."""
        p = self.p
        Bn, l, _, a, r = T.shape
        Tk = T if op is None else p.absorb_1q(T.contiguous(), op.reshape(-1, 2, 2, 1).to(T.dtype).contiguous())
        X = torch.empty((Bn, l, 2, a, R.shape[2]), dtype=C128, device=T.device)
        p.contract(Tk, (1, 3, 1), R, (1, 1, 1), X, (1, 3, 1))
        Rn = torch.empty((Bn, l, l), dtype=C128, device=T.device)
        p.contract(X, (1, 1, 3), T.permute(0, 2, 3, 4, 1), (1, 3, 1), Rn, (1, 1, 1), conjB=True)
        return Rn

    def inner(self, L, R):
        """sum_{r,r'} L[b,r,r'] R[b,r,r'] -> [B] complex128 (closing a left and a right environment)."""
        Bn = L.shape[0]
        out = torch.empty((Bn, 1, 1), dtype=C128, device=L.device)
        self.p.contract(L.reshape(Bn, 1, -1), (1, 1, 1), R.reshape(Bn, -1, 1), (1, 1, 1), out, (1, 1, 1))
        return out.reshape(Bn)

    def dense_rho(self, Ts, keep=None):
        """Dense un-normalised (reduced) rho [B, 2^m, 2^m] over the sites in `keep` (default all; small m only)."""
        p = self.p
        Bn = Ts[0].shape[0]
        n = len(Ts)
        keep = list(range(n)) if keep is None else list(keep)
        Rm = torch.ones((Bn, 1, 1, 1), dtype=C128, device=Ts[0].device)  # [B, P, l, l']
        for k, T in enumerate(Ts):
            _, l, _, a, r = T.shape
            P = Rm.shape[1]
            # Y[b,P,l',s,a,r] = sum_l R[b,P,l,l'] T[b,l,s,a,r]
            Y = torch.empty((Bn, P, l, 2, a, r), dtype=C128, device=T.device)
            p.contract(Rm.permute(0, 1, 3, 2), (1, 2, 1), T, (1, 1, 3), Y.reshape(Bn, P, l, 2 * a * r), (1, 2, 1))
            if k in keep:
                # R'[b,P,s,s',r,r'] = sum_{l',a} Y[b,P,l',s,a,r] conj(T[b,l',s',a,r'])
                Rn = torch.empty((Bn, P, 2, 2, r, r), dtype=C128, device=T.device)
                p.contract(Y.permute(0, 1, 3, 5, 2, 4), (1, 3, 2), T.permute(0, 1, 3, 2, 4), (1, 2, 2),
                           Rn.permute(0, 1, 2, 4, 3, 5), (1, 3, 2), conjB=True)
                Rm = Rn.reshape(Bn, P * 4, r, r)
            else:
                # R'[b,P,r,r'] = sum_{l',s,a} Y[b,P,l',s,a,r] conj(T[b,l',s,a,r'])
                Rn = torch.empty((Bn, P, r, r), dtype=C128, device=T.device)
                p.contract(Y.permute(0, 1, 5, 2, 3, 4), (1, 2, 3), T, (1, 3, 1), Rn, (1, 2, 1), conjB=True)
                Rm = Rn
        m = len(keep)
        rho = Rm.reshape([Bn] + [2, 2] * m)
        perm = [0] + [1 + 2 * i for i in range(m)] + [2 + 2 * i for i in range(m)]
        return rho.permute(perm).reshape(Bn, 2 ** m, 2 ** m)

    def dense_vector(self, Ts):
        """Dense state vector [B, 2^n] of an ideal circuit (inner dims all 1)."""
        p = self.p
        Bn = Ts[0].shape[0]
        V = torch.ones((Bn, 1, 1), dtype=C128, device=Ts[0].device)
        for T in Ts:
            _, l, _, a, r = T.shape
            assert a == 1, 'state vector readout needs an ideal (noise-free) state'
            Vn = torch.empty((Bn, V.shape[1], 2 * r), dtype=C128, device=T.device)
            p.contract(V, (1, 1, 1), T.reshape(Bn, l, 2 * r), (1, 1, 1), Vn, (1, 1, 1))
            V = Vn.reshape(Bn, -1, r)
        return V.reshape(Bn, -1)
