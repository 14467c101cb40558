"""Dev helper (CPU): numpy models of one-sided Jacobi variants on pivoted-Cholesky factors, to count sweeps / rounds
on real matrices (gpurun_out/grams.npz from tools/dump_grams.py) before writing kernels."""
import sys
import numpy as np


def pivoted_cholesky(G, rel=1e-14):
    n = G.shape[0]
    A = G.copy()
    perm = np.arange(n)
    L = np.zeros((n, n), dtype=complex)
    d = A.diagonal().real.copy()
    dmax = d.max()
    r = 0
    for k in range(n):
        j = k + int(np.argmax(d[k:]))
        if d[j] <= rel * dmax:
            break
        if j != k:
            A[[k, j], :] = A[[j, k], :]
            A[:, [k, j]] = A[:, [j, k]]
            L[[k, j], :] = L[[j, k], :]
            d[[k, j]] = d[[j, k]]
            perm[[k, j]] = perm[[j, k]]
        L[k, k] = np.sqrt(d[k])
        L[k + 1:, k] = (A[k + 1:, k] - L[k + 1:, :k] @ L[k, :k].conj()) / L[k, k]
        d[k + 1:] -= np.abs(L[k + 1:, k]) ** 2
        r += 1
    return L[:, :r], perm, r


def rr_pair(np_, r, q):
    w = np_ - 1
    if q == 0:
        return w, (r - w if r >= w else r)
    p0 = (r + q) % w
    p1 = (r - q + w) % w
    return p0, p1


def rot(Y, i, j, tol, stats):
    x, y = Y[i], Y[j]
    a = np.vdot(x, x).real
    b = np.vdot(y, y).real
    g = np.vdot(y, x)   # sum x conj(y)
    g2 = abs(g) ** 2
    if not (a * b > 0) or not (g2 > tol * tol * a * b):
        return 0
    d = 0.5 * (b - a)
    ad = abs(d)
    h = np.sqrt(d * d + g2)
    ru = 1.0 / np.sqrt(2 * h * (h + ad))
    c = (h + ad) * ru
    sg = ru if d >= 0 else -ru
    s = sg * g
    xn = c * x - s * y
    yn = np.conj(s) * x + c * y
    Y[i], Y[j] = xn, yn
    return 1


def flat_jacobi(Y, b, tol=1e-10, max_sweeps=30):
    """The current kernel's ordering: blocks of b rows, tournament over blocks, intra-block pairs at round 0."""
    n = Y.shape[0]
    nb = (n + b - 1) // b
    nbp = (nb + 1) & ~1
    for sw in range(max_sweeps):
        cnt = 0
        for rnd in range(nbp - 1):
            for q in range(nbp // 2):
                I, J = rr_pair(nbp, rnd, q)
                if rnd == 0:
                    for blk in (I, J):
                        rows = [blk * b + t for t in range(b) if blk * b + t < n]
                        for i in range(len(rows)):
                            for j in range(i + 1, len(rows)):
                                cnt += rot(Y, rows[i], rows[j], tol, None)
                for step in range(b):
                    for t in range(b):
                        i, j = I * b + t, J * b + (t + step) % b
                        if i < n and j < n:
                            cnt += rot(Y, i, j, tol, None)
        print('  flat sweep %d rotations %d' % (sw, cnt), flush=True)
        if cnt == 0:
            return sw + 1
    return max_sweeps


def inner_eig(S, tol, max_inner, cross_only_b=None):
    """Two-sided cyclic Jacobi on Hermitian S; returns W (rows transform like rows of Y: Y' = W Y), #inner sweeps,
    #rotations."""
    m = S.shape[0]
    S = S.copy()
    W = np.eye(m, dtype=complex)
    tot = 0
    mp = (m + 1) & ~1
    for isw in range(max_inner):
        cnt = 0
        for step in range(mp - 1):
            for q in range(mp // 2):
                p0, p1 = rr_pair(mp, step, q)
                if p0 >= m or p1 >= m:
                    continue
                i, j = min(p0, p1), max(p0, p1)
                a, bq, g = S[i, i].real, S[j, j].real, S[i, j]
                g2 = abs(g) ** 2
                if not (a * bq > 0) or not (g2 > tol * tol * a * bq):
                    continue
                d = 0.5 * (bq - a)
                ad = abs(d)
                h = np.sqrt(d * d + g2)
                ru = 1.0 / np.sqrt(2 * h * (h + ad))
                c = (h + ad) * ru
                sg = ru if d >= 0 else -ru
                s = sg * g
                R = np.array([[c, -s], [np.conj(s), c]])
                S[[i, j], :] = R @ S[[i, j], :]
                S[:, [i, j]] = S[:, [i, j]] @ R.conj().T
                W[[i, j], :] = R @ W[[i, j], :]
                cnt += 1
        tot += cnt
        if cnt == 0:
            return W, isw + 1, tot
    return W, max_inner, tot


def block_jacobi(Y, b, tol=1e-10, max_sweeps=30, max_inner=30):
    n = Y.shape[0]
    nb = (n + b - 1) // b
    nbp = (nb + 1) & ~1
    for sw in range(max_sweeps):
        cnt = 0
        inner_tot = 0
        inner_max = 0
        for rnd in range(nbp - 1):
            for q in range(nbp // 2):
                I, J = rr_pair(nbp, rnd, q)
                rows = [I * b + t for t in range(b) if I * b + t < n] + [J * b + t for t in range(b) if J * b + t < n]
                if len(rows) < 2:
                    continue
                Z = Y[rows]
                S = Z @ Z.conj().T
                W, isw, c = inner_eig(S, tol, max_inner)
                if c:
                    Y[rows] = W @ Z
                cnt += c
                inner_tot += isw
                inner_max = max(inner_max, isw)
        print('  block(b=%d) sweep %d rotations %d  inner sweeps avg %.2f max %d' %
              (b, sw, cnt, inner_tot / max(1, (nbp - 1) * (nbp // 2)), inner_max), flush=True)
        if cnt == 0:
            return sw + 1
    return max_sweeps


def check(G, perm, Y):
    lam = (np.abs(Y) ** 2).sum(1)
    order = np.argsort(-lam)
    V = Y[order] / np.sqrt(lam[order])[:, None]
    # rows of Y are in the pivoted coordinates
    Gp = G[np.ix_(perm, perm)]
    rec = (V.conj().T * lam[order]) @ V
    # Y = W L^h -> Y^h Y = L L^h = Gp; rows v_j^h ... eigenvectors of Gp are columns conj? check both
    e1 = np.linalg.norm(rec - Gp) / np.linalg.norm(Gp)
    rec2 = (V.T * lam[order]) @ V.conj()
    e2 = np.linalg.norm(rec2 - Gp) / np.linalg.norm(Gp)
    return min(e1, e2), lam[order]


if __name__ == '__main__':
    path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/grams.npz'
    mode = sys.argv[2] if len(sys.argv) > 2 else 'all'
    data = np.load(path)
    seen = set()
    for key in data.files:
        G = data[key]
        n = G.shape[0]
        if n in seen:
            continue
        seen.add(n)
        L, perm, r = pivoted_cholesky(G)
        print(key, 'n', n, 'rank', r, 'cond(L) ~ %.1e' % (abs(L[0, 0]) / abs(L[r - 1, r - 1])), flush=True)
        Y0 = L.conj().T.copy()   # r rows of length n
        if mode in ('all', 'flat'):
            Y = Y0.copy()
            print(' flat b=8 sweeps', flat_jacobi(Y, 8), 'err %.1e' % check(G, perm, Y)[0], flush=True)
        for b in (8, 16, 32):
            if mode in ('all', 'block', 'block%d' % b):
                Y = Y0.copy()
                print(' block b=%d sweeps' % b, block_jacobi(Y, b), 'err %.1e' % check(G, perm, Y)[0], flush=True)
