"""ORACLE - test infrastructure, not product code.

A dense torch-CPU restatement of the MPDOSimulator noisy-gate update path of
WeiguoMa/Tomography-assisted-MPDO-QCircuit. Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import it; the product package never does.

Parity status: the reference ships no tests or golden vectors for this path (test/debug.py asserts
nothing), so parity is unpinned by the reference's own tests. The oracle is instead pinned against the
UNMODIFIED reference modules executed in the build container on top of oracle/tn_shim (a restatement of
the tensornetwork==0.4.6 calls the reference makes; tensornetwork itself is not installable offline):
tests/golden/make_golden.py produced the fixtures under tests/golden/ that tests/test_oracle.py checks.

Every function cites the reference lines it follows (paths relative to /root/reference).

State layout: one tensor T_k[l, s, a, r] per qubit (left bond, physical, inner/Kraus, right bond),
size-1 axes where the reference node has no such axis, plus flags recording which axes "exist"
(the reference's truncate() is a no-op until every bond exists: Circuit.py:477, TNNOptimizer.py:51-61).
"""
import itertools
import os

import numpy as np
import torch

GLOBAL_MINIMUM = 2.718281828459045 * 1e-8  # Circuit.py:22


# =============================================================================================
# decompositions.py restated (the numerical contract of every split)
# =============================================================================================
def _randomized_svd(M, n_components, n_oversamples=7, n_iter='auto'):
    """decompositions.py:23-48 - power iterations without re-orthonormalisation, CPU probe matrix
    drawn from a fresh default-seeded generator."""
    m, n = M.shape
    rng = torch.Generator()
    Q = torch.randn(m, n_components + n_oversamples, dtype=M.dtype, generator=rng)
    if n_iter == 'auto':
        n_iter = 3 if m >= n else 2
    for _ in range(n_iter):
        Q = M @ (M.T.conj() @ Q)
    Q, _ = torch.linalg.qr(Q)
    B = Q.T.conj() @ M
    u, s, vh = torch.linalg.svd(B, full_matrices=False)
    return Q @ u, s, vh


def svd(tensor, pivot_axis, max_singular_values=None, max_truncation_error=None, relative=False, mode='exact'):
    """decompositions.py:51-146. mode='reference' keeps the randomized branch (numel >= 10000 and a
    rank cap, :112-115); mode='exact' always takes the full LAPACK SVD."""
    left_dims = tensor.shape[:pivot_axis]
    right_dims = tensor.shape[pivot_axis:]
    M = tensor.reshape((-1, int(np.prod(right_dims)) if len(right_dims) else 1))
    if M.numel() < 10000 or max_singular_values is None or mode == 'exact':
        u, s, vh = torch.linalg.svd(M, full_matrices=False)
    else:
        u, s, vh = _randomized_svd(M, n_components=max_singular_values)
    if max_singular_values is None:
        max_singular_values = s.numel()
    num_err = max_singular_values
    if max_truncation_error is not None:
        trunc_errs = torch.sqrt(torch.cumsum(s ** 2, dim=0))
        eps = max_truncation_error * s[0] if relative else max_truncation_error
        for idx in range(trunc_errs.shape[0] - 1):
            if trunc_errs[-1] - trunc_errs[idx] <= eps:
                num_err = idx + 1
                break
    keep = min(max_singular_values, num_err)
    s = s.to(M.dtype)
    s_rest = s[keep:]
    s = s[:keep]
    u = u[:, :keep]
    vh = vh[:keep, :]
    dim_s = s.size(0)                                  # may be < keep when the cap exceeds len(s)
    u = u.reshape(*left_dims, dim_s)
    vh = vh.reshape(dim_s, *right_dims)
    return u, s, vh, s_rest


def qr(tensor, pivot_axis):
    """decompositions.py:149-195 with non_negative_diagonal=False (how tn.split_node_qr calls it)."""
    left_dims = list(tensor.shape)[:pivot_axis]
    right_dims = list(tensor.shape)[pivot_axis:]
    M = tensor.reshape(int(np.prod(left_dims)), int(np.prod(right_dims)))
    q, r = torch.linalg.qr(M)
    c = q.shape[1]
    return q.reshape(left_dims + [c]), r.reshape([c] + right_dims)


# =============================================================================================
# operand builders (host side of the reference)
# =============================================================================================
CHIPS = {  # ChipInfo.py:44-87
    'best': dict(gateTime=30, bath_rate=0., decay_rate=0.0, dephasing_rate=0.0, T1=2e11, T2=2e10, dpc=11e-4),
    'medium': dict(gateTime=1, bath_rate=0.01, decay_rate=0.98, dephasing_rate=0.02, T1=2e11, T2=2e10, dpc=5e-2),
    'worst': dict(gateTime=30, bath_rate=0., decay_rate=0.0, dephasing_rate=0.0, T1=2e2, T2=2e1, dpc=11e-2),
}


def noise_tensors(chip, dtype):
    """NoiseChannel.py:19-46,51-91,144-170: decayTensor [2,2,2], dephasingTensor [2,2,3],
    dpCTensor2 [2,2,2,2,16]."""
    c = CHIPS[chip or 'worst']
    pd = 1 - np.exp(-c['bath_rate'] * c['decay_rate'] * c['gateTime'])
    decay = torch.tensor([[[1, 0], [0, np.sqrt(1 - pd)]], [[0, np.sqrt(pd)], [0, 0]]], dtype=dtype).permute(1, 2, 0)
    pp = 1 - np.exp(-c['bath_rate'] * c['dephasing_rate'] * c['gateTime'])
    sp, s1p = np.sqrt(pp), np.sqrt(1 - pp)
    deph = torch.tensor([[[s1p, 0], [0, s1p]], [[sp, 0], [0, 0]], [[0, 0], [0, sp]]], dtype=dtype).permute(1, 2, 0)
    paulis = [torch.tensor(m, dtype=dtype) for m in
              ([[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]])]

    def dpc(p, qn):
        probs = [np.sqrt(1 - (4 ** qn - 1) * p / (4 ** qn))] + [np.sqrt(p / (4 ** qn))] * (4 ** qn - 1)
        diag = torch.diag(torch.tensor(probs, dtype=dtype))
        ops = paulis
        for _ in range(qn - 1):
            ops = [torch.kron(e, b) for e in ops for b in paulis]
        t = torch.einsum('ij, jfk -> fki', diag, torch.stack(ops))
        return t.reshape([2] * (2 * qn) + [t.shape[-1]])

    return dict(decay=decay, dephasing=deph, dpc1=dpc(c['dpc'], 1), dpc2=dpc(c['dpc'], 2))


def read_chi(filename):
    """RealNoise.py:22-32."""
    if '.mat' in filename:
        from scipy.io import loadmat
        return loadmat(filename)['exp']
    if '.npz' in filename:
        return np.load(filename)['chi']
    raise TypeError('Current file-type is not supported.')


def chi_to_tensor(chi):
    """RealNoise.py:35-170 (noisyTensor + czExp_channel): chi -> [p0,p1,s0,s1,K] in complex64.
    eig (general, not eigh) sorted descending, |lambda| > 1e-12 kept, coefficients below 1e-12 zeroed,
    basis {I, X, -i*sigma_y, Z}^{x2} (Tools.py:466-509), accumulation in complex64."""
    s, u = np.linalg.eig(chi)
    idx = s.argsort()[::-1]
    s, u = s[idx], u[:, idx]
    eff = [v for v in s if np.abs(v) > 1e-12]
    names = [''.join(g) for g in itertools.product('IXYZ', repeat=2)]
    ops = {
        'I': torch.eye(2, dtype=torch.complex64),
        'X': torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64),
        'Y': -1j * torch.tensor([[0, -1j], [1j, 0]], dtype=torch.complex64),
        'Z': torch.tensor([[1, 0], [0, -1]], dtype=torch.complex64),
    }
    tensors = []
    for i in range(len(eff)):
        coeff = [u[j, i] * np.sqrt(s[i]) for j in range(len(names))]
        coeff = [0 + 0j if np.abs(cj) < 1e-12 else cj for cj in coeff]
        step = torch.zeros((2, 2, 2, 2), dtype=torch.complex64)
        for name, cj in zip(names, coeff):
            if cj != 0:
                step += torch.reshape(torch.tensor(cj, dtype=torch.complex64) * torch.kron(ops[name[0]], ops[name[1]]),
                                      (2, 2, 2, 2))
        tensors.append(step)
    return torch.einsum('ijlmn -> jlmni', torch.stack(tensors))


def _c(x, dtype):
    """AbstractGate.py:42-51: a float / Tensor angle is cast to the complex circuit dtype first."""
    if isinstance(x, torch.Tensor):
        return x.to(dtype=dtype)
    if isinstance(x, float):
        return torch.tensor(x, dtype=dtype)
    raise ValueError(f'Invalid type for gate parameter: {type(x)}')


def gate_matrix(name, params, dtype):
    """QuantumGates/*.py tensors. Returns (tensor, single, variational)."""
    t = lambda data: torch.tensor(data, dtype=dtype)
    sq2 = np.sqrt(2)
    P = [_c(p, dtype) for p in params]
    e, cos, sin = torch.exp, torch.cos, torch.sin
    if name == 'I': return t([[1, 0], [0, 1]]), True, False
    if name == 'H': return t([[1, 1], [1, -1]]) / sq2, True, False                      # SingleGates.py:64
    if name == 'X': return t([[0, 1], [1, 0]]), True, False
    if name == 'Y': return t([[0, -1j], [1j, 0]]), True, False
    if name == 'Z': return t([[1, 0], [0, -1]]), True, False
    if name == 'S': return t([[1, 0], [0, 1j]]), True, False
    if name == 'Sdg': return t([[1, 0], [0, -1j]]), True, False
    if name == 'T': return t([[1, 0], [0, (1 + 1j) / sq2]]), True, False
    if name == 'P': return t([[1, 0], [0, e(P[0] * 1j)]]), True, True
    if name == 'U1': return t([[1, 0], [0, e(1j * P[0])]]), True, True
    if name == 'U3':                                                                     # SingleGates.py:179-185
        th, ph, la = P
        lM, pM, tC, tS = e(1j * la), e(1j * ph), cos(th / 2), sin(th / 2)
        return t([[tC, -lM * tS], [pM * tS, lM * pM * tC]]), True, True
    if name == 'RX':
        tC, tS = cos(P[0] / 2), sin(P[0] / 2)
        return t([[tC, -1j * tS], [-1j * tS, tC]]), True, True
    if name == 'RY':
        tC, tS = cos(P[0] / 2), sin(P[0] / 2)
        return t([[tC, -tS], [tS, tC]]), True, True
    if name == 'RZ': return t([[e(-1j * P[0] / 2), 0], [0, e(1j * P[0] / 2)]]), True, True
    if name == 'MeasureX': return 1 / sq2 * t([[1, 1], [1, -1]]), True, False
    if name == 'MeasureY': return 1 / sq2 * t([[1, -1j], [1, 1j]]), True, False
    if name == 'Reset0': return t([[1, 0], [0, 0]]), True, False
    if name == 'Reset1': return t([[0, 0], [0, 1]]), True, False
    r4 = lambda data: t(data).reshape(2, 2, 2, 2)
    if name == 'II': return r4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]), False, False
    if name in ('CX', 'CNOT'): return r4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), False, False
    if name == 'CY': return r4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]]), False, False
    if name == 'CZ': return r4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]]), False, False
    if name == 'SWAP': return r4([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]), False, False
    if name == 'ISWAP': return r4([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]), False, False
    if name == 'CP': return r4([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, e(1j * P[0])]]), False, True
    if name == 'RZZ':
        mi, pl = e(-1j * P[0] / 2), e(1j * P[0] / 2)
        return r4([[mi, 0, 0, 0], [0, pl, 0, 0], [0, 0, pl, 0], [0, 0, 0, mi]]), False, True
    if name == 'RXX':
        tC, tS = cos(P[0] / 2), sin(P[0] / 2)
        return r4([[tC, 0, 0, -1j * tS], [0, tC, -1j * tS, 0], [0, -1j * tS, tC, 0], [-1j * tS, 0, 0, tC]]), False, True
    if name == 'RYY':
        tC, tS = cos(P[0] / 2), sin(P[0] / 2)
        return r4([[tC, 0, 0, -tS], [0, tC, tS, 0], [0, -tS, tC, 0], [tS, 0, 0, tC]]), False, True
    raise KeyError(name)


# =============================================================================================
# the circuit
# =============================================================================================
class OracleCircuit:
    """Dense restatement of TensorCircuit (Circuit.py) + QuantumCircuit builders (AbstractCircuit.py)."""

    def __init__(self, qn, ideal=True, noiseType='no', chiFileDict=None, chi=None, kappa=None,
                 max_truncation_err=None, chip=None, dtype=torch.complex64, svd_mode='exact', fast=False):
        self.qn, self.ideal, self.noiseType = qn, ideal, noiseType.lower()
        self.chi, self.kappa, self.max_truncation_err = chi, kappa, max_truncation_err
        self.dtype, self.svd_mode = dtype, svd_mode
        self.unified = self.realNoise = self.idealNoise = False
        self.layers = []            # ('gate', name, tensor, single, variational, ideal, oqs) | ('truncate',) | ('barrier',)
        self.noise_cache = {}       # Circuit.py:93-96,146-151: keyed by gate name only
        self.cz_tensors = {}
        self.cp_tensors = {}
        if not ideal:
            if self.noiseType not in ['unified', 'realnoise', 'idealnoise']:
                raise ValueError(f'Unsupported noise type: {self.noiseType}')  # Circuit.py:54-55
            self.noise = noise_tensors(chip, dtype)
            self.unified = self.noiseType == 'unified'
            self.realNoise = self.noiseType == 'realnoise'
            self.idealNoise = self.noiseType == 'idealnoise'
            if self.realNoise:                                                  # AbstractCircuit.py:66-75
                files = chiFileDict if chiFileDict is not None else {'CZ': {}, 'CP': {}}
                cache = {}
                for key, fn in files.get('CZ', {}).items():
                    if fn not in cache:
                        cache[fn] = chi_to_tensor(read_chi(fn)).to(dtype)
                    self.cz_tensors[key] = cache[fn]
                for key, fn in files.get('CP', {}).items():                       # cpExp_channel, RealNoise.py:173-179
                    if fn not in cache:
                        cache[fn] = chi_to_tensor(read_chi(fn)).to(dtype)
                    self.cp_tensors[key] = cache[fn]
        self.T = None
        self.bond = None
        self.inner = None
        self.stats = {'updates_2q_noisy': 0, 'split_ranks': []}
        # fast=True: same mathematics, smaller LAPACK calls (see _apply_2q_fast / _svd_right2left_fast); only
        # meaningful with svd_mode='exact'. tests/test_oracle_fast.py pins it to the plain restatement.
        self.fast = bool(fast) and svd_mode == 'exact'

    # ---- builders (AbstractCircuit.py:156-532), subset used by the parity suites ----
    def _add(self, name, params, oqs, ideal):
        tensor, single, var = gate_matrix(name, params, self.dtype)
        self.layers.append(('gate', name, tensor, single, var, ideal, list(oqs)))

    def _1q(self, name, params, oqs, ideal):
        self._add(name, params, [oqs] if isinstance(oqs, int) else oqs, ideal)

    def i(self, oqs, _ideal=None): self._1q('I', [], oqs, _ideal)
    def h(self, oqs, _ideal=None): self._1q('H', [], oqs, _ideal)
    def x(self, oqs, _ideal=None): self._1q('X', [], oqs, _ideal)
    def y(self, oqs, _ideal=None): self._1q('Y', [], oqs, _ideal)
    def z(self, oqs, _ideal=None): self._1q('Z', [], oqs, _ideal)
    def s(self, oqs, _ideal=None): self._1q('S', [], oqs, _ideal)
    def sdg(self, oqs, _ideal=None): self._1q('Sdg', [], oqs, _ideal)
    def t(self, oqs, _ideal=None): self._1q('T', [], oqs, _ideal)
    def p(self, theta, oqs, _ideal=None): self._1q('P', [theta], oqs, _ideal)
    def u1(self, theta, oqs, _ideal=None): self._1q('U1', [theta], oqs, _ideal)
    def u3(self, theta, phi, lam, oqs, _ideal=None): self._1q('U3', [theta, phi, lam], oqs, _ideal)
    def rx(self, theta, oqs, _ideal=None): self._1q('RX', [theta], oqs, _ideal)
    def ry(self, theta, oqs, _ideal=None): self._1q('RY', [theta], oqs, _ideal)
    def rz(self, theta, oqs, _ideal=None): self._1q('RZ', [theta], oqs, _ideal)

    def cz(self, q0, q1, _ideal=None):                                          # AbstractCircuit.py:313-330
        if not self.realNoise or _ideal:
            self._add('CZ', [], [q0, q1], _ideal)
        else:
            tensor = self.cz_tensors.get(f'{q0}{q1}')
            if tensor is None:
                tensor = self.cz_tensors.get(f'{q1}{q0}')
            if tensor is None:
                raise FileNotFoundError('oracle: no chi file for this pair (the reference falls back to a cwd path)')
            self.layers.append(('gate', 'CZEXP', tensor, False, False, False, [q0, q1]))

    def _cx(self, name, q0, q1, _ideal):                                         # AbstractCircuit.py:291-301,332-342
        if not self.realNoise or _ideal:
            self._add(name, [], [q0, q1], _ideal)
        else:
            self.ry(-torch.tensor(np.pi / 2), q1, True)   # float32 tensor angle, as in the reference
            self.cz(q0, q1)
            self.ry(torch.tensor(np.pi / 2), q1, True)

    def cx(self, q0, q1, _ideal=None): self._cx('CX', q0, q1, _ideal)
    def cnot(self, q0, q1, _ideal=None): self._cx('CNOT', q0, q1, _ideal)
    def cy(self, q0, q1, _ideal=None):
        if not self.realNoise or _ideal:
            self._add('CY', [], [q0, q1], _ideal)
        else:
            raise NotImplementedError('EXPCYGate is not implemented yet.')
    def swap(self, q0, q1, _ideal=None): self._add('SWAP', [], [q0, q1], _ideal)
    def iswap(self, q0, q1, _ideal=None): self._add('ISWAP', [], [q0, q1], _ideal)
    def ii(self, q0, q1, _ideal=None): self._add('II', [], [q0, q1], _ideal)
    def cp(self, theta, q0, q1, _ideal=None):                                    # AbstractCircuit.py:404-422
        if not self.realNoise or _ideal:
            self._add('CP', [theta], [q0, q1], _ideal)
        else:
            tensor = self.cp_tensors.get(f'{q0}{q1}')
            if tensor is None:
                tensor = self.cp_tensors.get(f'{q1}{q0}')
            if tensor is None:
                raise FileNotFoundError('oracle: no CP chi file for this pair (the reference default path does not exist)')
            self.layers.append(('gate', 'CPEXP', tensor, False, False, False, [q0, q1]))

    def rzz(self, theta, q0, q1, _ideal=None):                                   # AbstractCircuit.py:263-275
        if not self.realNoise or _ideal:
            self._add('RZZ', [theta], [q0, q1], _ideal)
        else:
            self.cx(q0, q1)
            self.rz(theta, q1, True)
            self.cx(q0, q1)

    def rxx(self, theta, q0, q1, _ideal=None):                                   # AbstractCircuit.py:231-245
        if not self.realNoise or _ideal:
            self._add('RXX', [theta], [q0, q1], _ideal)
        else:
            self.h([q0, q1], True); self.cx(q0, q1); self.rz(theta, q1, True); self.cx(q0, q1); self.h([q0, q1], True)

    def ryy(self, theta, q0, q1, _ideal=None):                                   # AbstractCircuit.py:247-261
        if not self.realNoise or _ideal:
            self._add('RYY', [theta], [q0, q1], _ideal)
        else:
            self.rx(torch.tensor(np.pi / 2), [q0, q1], True); self.cx(q0, q1); self.rz(theta, q1, True)
            self.cx(q0, q1); self.rx(-torch.tensor(np.pi / 2), [q0, q1], True)

    def truncate(self): self.layers.append(('truncate',))
    def barrier(self): self.layers.append(('barrier',))

    # ---- gate application -------------------------------------------------------------------
    def _apply_1q(self, name, G, var, ideal, oqs):
        """Circuit.py:138-178."""
        noisy = (self.idealNoise or self.unified) and not ideal
        if noisy:
            def build():
                return torch.einsum('nlm, ljk, ji -> nimk', self.noise['decay'], self.noise['dephasing'],
                                    G).reshape((2, 2, -1))
            G = build() if var else self.noise_cache.setdefault(name, build())
        for q in oqs:
            T = self.T[q]
            if noisy:
                l, _, a, r = T.shape
                T = torch.einsum('psg,lsar->lpagr', G, T).reshape(l, 2, a * G.shape[-1], r)
                self.inner[q] = True
            else:
                T = torch.einsum('ps,lsar->lpar', G, T)
            self.T[q] = T

    def _apply_2q(self, name, G, var, ideal, oqs):
        """Circuit.py:74-136. The noise index goes to the higher-numbered qubit (:88-89)."""
        if self.fast:
            return self._apply_2q_fast(name, G, var, ideal, oqs)
        lo, hi = min(oqs), max(oqs)
        if hi != lo + 1:
            raise NotImplementedError('oracle: two-qubit gates must act on neighbouring qubits')
        g_noise = (self.idealNoise and not ideal) or self.realNoise
        if g_noise and not self.realNoise:
            if var:  # 5-index tensor but 4 axis names -> tn.Node raises
                raise ValueError('variational two-qubit gates cannot carry idealNoise in the reference')
            G = self.noise_cache.setdefault(name, torch.einsum('ijklp, klmn -> ijmnp', self.noise['dpc2'], G))
        if self.realNoise and G.dim() != 5:
            raise ValueError('realNoise mode needs a 5-index two-qubit gate tensor (CZEXP / CPEXP)')
        if G.dim() == 4:
            G = G.unsqueeze(-1)
        if oqs[0] != lo:  # axes are named by _oqs order (control first): bring to [p_lo,p_hi,s_lo,s_hi,g]
            G = G.permute(1, 0, 3, 2, 4)
        Tl, Th = self.T[lo], self.T[hi]
        l, _, a0, m = Tl.shape
        _, _, a1, r = Th.shape
        K = G.shape[-1]
        theta = torch.einsum('lxam,mybr,PQxyg->lPaQbgr', Tl, Th, G).reshape(l, 2, a0, 2, a1 * K, r)
        u, s, vh, _ = svd(theta, 3, None, GLOBAL_MINIMUM, False, mode='exact')  # no rank cap: always the full SVD
        sq = torch.sqrt(s)
        self.T[lo] = u * sq
        self.T[hi] = sq.reshape(-1, 1, 1, 1) * vh
        self.bond[lo] = True
        self.stats['split_ranks'].append(int(s.numel()))
        if K > 1 or g_noise:
            self.inner[hi] = True
        if g_noise:
            self.stats['updates_2q_noisy'] += 1

    def _gate_operand_2q(self, name, G, var, ideal, oqs):
        """The operand part of _apply_2q (Circuit.py:84-99): [p_lo,p_hi,s_lo,s_hi,g] and the noise flag."""
        lo = min(oqs)
        g_noise = (self.idealNoise and not ideal) or self.realNoise
        if g_noise and not self.realNoise:
            if var:
                raise ValueError('variational two-qubit gates cannot carry idealNoise in the reference')
            G = self.noise_cache.setdefault(name, torch.einsum('ijklp, klmn -> ijmnp', self.noise['dpc2'], G))
        if self.realNoise and G.dim() != 5:
            raise ValueError('realNoise mode needs a 5-index two-qubit gate tensor (CZEXP / CPEXP)')
        if G.dim() == 4:
            G = G.unsqueeze(-1)
        if oqs[0] != lo:
            G = G.permute(1, 0, 3, 2, 4)
        return G, g_noise

    def _apply_2q_fast(self, name, G, var, ideal, oqs):
        """Circuit.py:74-136 with the (l*2*a0) x (2*a1*K*r) two-site matrix never formed: Theta =
        (Q1 x 1) . C . (1 x Q2h) with Q1, Q2h the LAPACK QR / LQ isometries of T_lo[(l,a0),(s0,m)] and
        T_hi[(m,s1),(a1,r)], so Theta and the core C [(x,p0),(p1,g,y)] have the same singular values and
        SVD(Theta) = (Q1 x 1) SVD(C) (1 x Q2h). Same full LAPACK SVD, same rank rule, on the small matrix."""
        lo, hi = min(oqs), max(oqs)
        if hi != lo + 1:
            raise NotImplementedError('oracle: two-qubit gates must act on neighbouring qubits')
        G, g_noise = self._gate_operand_2q(name, G, var, ideal, oqs)
        Tl, Th = self.T[lo], self.T[hi]
        l, _, a0, m = Tl.shape
        _, _, a1, r = Th.shape
        K = G.shape[-1]
        X = Tl.permute(0, 2, 1, 3).reshape(l * a0, 2 * m)
        if l * a0 > 2 * m:
            Q1, R1 = torch.linalg.qr(X)                       # [(l,a0),x], [x,(s0,m)]
        else:
            Q1, R1 = None, X
        x = R1.shape[0]
        Y = Th.reshape(m * 2, a1 * r)
        if a1 * r > 2 * m:
            Q2, R2 = torch.linalg.qr(Y.mH)                    # Y^h = Q2 R2  ->  Y = R2^h Q2^h
            L2, Q2h = R2.mH, Q2.mH                            # [(m,s1),y], [y,(a1,r)]
        else:
            L2, Q2h = Y, None
        y = L2.shape[1]
        C = torch.einsum('xam,mby,PQabg->xPQgy', R1.reshape(x, 2, m), L2.reshape(m, 2, y), G)
        u, s, vh, _ = svd(C, 2, None, GLOBAL_MINIMUM, False, mode='exact')
        sq = torch.sqrt(s)
        k = s.numel()
        u = u * sq                                            # [x,P,k]
        vh = sq.reshape(-1, 1, 1, 1) * vh                     # [k,Q,g,y]
        if Q1 is not None:
            Tl_n = (Q1 @ u.reshape(x, 2 * k)).reshape(l, a0, 2, k).permute(0, 2, 1, 3)
        else:
            Tl_n = u.reshape(l, a0, 2, k).permute(0, 2, 1, 3)
        if Q2h is not None:
            Th_n = (vh.reshape(k * 2 * K, y) @ Q2h).reshape(k, 2, K, a1, r)
        else:
            Th_n = vh.reshape(k, 2, K, a1, r)
        self.T[lo] = Tl_n.contiguous()
        self.T[hi] = Th_n.permute(0, 1, 3, 2, 4).reshape(k, 2, a1 * K, r).contiguous()   # inner = (a1, g) as in _apply_2q
        self.bond[lo] = True
        self.stats['split_ranks'].append(int(k))
        if K > 1 or g_noise:
            self.inner[hi] = True
        if g_noise:
            self.stats['updates_2q_noisy'] += 1

    # ---- truncation sweeps --------------------------------------------------------------------
    def _connected(self):
        return self.qn <= 1 or all(self.bond)                                    # TNNOptimizer.py:51-61

    def _qr_left2right(self):
        """TNNOptimizer.py:87-108."""
        for i in range(self.qn - 1):
            q, r = qr(self.T[i], 3)
            self.T[i] = q
            self.T[i + 1] = torch.tensordot(r, self.T[i + 1], dims=([1], [0]))

    def _svd_right2left(self):
        """TNNOptimizer.py:111-134 (tn.split_node puts sqrt(S) on both sides)."""
        if self.fast:
            return self._svd_right2left_fast()
        for i in range(self.qn - 1, 0, -1):
            theta = torch.tensordot(self.T[i - 1], self.T[i], dims=([3], [0]))
            u, s, vh, rest = svd(theta, 3, self.chi, self.max_truncation_err, True, mode=self.svd_mode)
            sq = torch.sqrt(s)
            self.T[i - 1] = u * sq
            self.T[i] = sq.reshape(-1, 1, 1, 1) * vh
            self.stats.setdefault('discarded', []).append(rest)

    def _svd_right2left_fast(self):
        """TNNOptimizer.py:111-134 right after :87-108: the left neighbour T_{i-1} is the Q of a LAPACK QR
        (an isometry over its right bond), so SVD(T_{i-1} . X) = T_{i-1} . SVD(X) with X = T_i as
        [l | (s,a,r)]: the same singular values, the same rank decision, the same sqrt(S) | sqrt(S) split,
        without forming the (l's'a') x (s a r) two-site matrix."""
        for i in range(self.qn - 1, 0, -1):
            X = self.T[i]
            u, s, vh, rest = svd(X, 1, self.chi, self.max_truncation_err, True, mode='exact')
            sq = torch.sqrt(s)
            self.T[i - 1] = torch.tensordot(self.T[i - 1], u * sq, dims=([3], [0]))
            self.T[i] = sq.reshape(-1, 1, 1, 1) * vh
            self.stats.setdefault('discarded', []).append(rest)

    def _svd_kappa(self):
        """TNNOptimizer.py:164-197: T <- U.S over the inner index, Vh dropped."""
        if self.kappa is None and self.max_truncation_err is None:
            return
        for k in range(self.qn):
            T = self.T[k]
            if self.kappa is not None and (not self.inner[k] or T.shape[2] <= self.kappa):
                continue
            l, _, a, r = T.shape
            M = T.permute(0, 1, 3, 2)
            u, s, _, _ = svd(M, 3, self.kappa, self.max_truncation_err, True, mode=self.svd_mode)
            self.T[k] = (u * s).permute(0, 1, 3, 2).contiguous()

    def evolve(self, state=None):
        """Circuit.py:469-491. state: list of (2,) tensors (default |0...0>)."""
        if state is None:
            state = [torch.tensor([1, 0], dtype=self.dtype) for _ in range(self.qn)]
        self.T = [s.to(self.dtype).reshape(1, 2, 1, 1) for s in state]
        self.bond = [False] * (self.qn - 1)
        self.inner = [False] * self.qn
        return self.run_layers()

    def run_layers(self):
        """The layer loop of evolve on the current self.T / self.bond / self.inner (lets a caller start from a
        prepared mid-circuit window)."""
        for layer in self.layers:
            if layer[0] == 'truncate':
                if self._connected():
                    if not (self.chi is None and self.max_truncation_err is None):  # TNNOptimizer.py:79-80
                        self._qr_left2right()
                        self._svd_right2left()
                    if not self.ideal:
                        self._svd_kappa()
            elif layer[0] == 'barrier':
                pass
            else:
                _, name, G, single, var, ideal, oqs = layer
                if name == 'MeasureZ':
                    continue
                if max(oqs) >= self.qn:
                    raise ValueError(f'Qubit index out of range, max index is Q{max(oqs)}.')
                (self._apply_1q if single else self._apply_2q)(name, G, var, ideal, oqs)
        if not self.ideal and self.layers and self.layers[-1][0] != 'truncate':
            self._svd_kappa()
        return self.T

    # ---- readout (Circuit.py:226-332, dmOperations.py) -------------------------------------------
    def cal_dm(self):
        """Dense un-normalised rho (small n)."""
        R = torch.ones((1, 1, 1), dtype=self.dtype)  # [P, l, l']
        for T in self.T:
            R = torch.einsum('Pab,asxr,btxq->Pstrq', R, T, T.conj())
            R = R.reshape(-1, R.shape[-2], R.shape[-1])
        n = self.qn
        rho = R.reshape([2, 2] * n)
        perm = [2 * i for i in range(n)] + [2 * i + 1 for i in range(n)]
        return rho.permute(perm).reshape(2 ** n, 2 ** n)

    def cal_vector(self):
        if not self.ideal:
            raise ValueError('Noisy circuit cannot be represented by state vector efficiently.')
        V = torch.ones((1, 1), dtype=self.dtype)
        for T in self.T:
            V = torch.einsum('Pa,asr->Psr', V, T[:, :, 0, :]).reshape(-1, T.shape[-1])
        return V.reshape(-1, 1)

    def chain(self, ops=None, proj=None):
        """Tr(prod O_k rho) (ops {site: 2x2}) or <b|rho|b> (proj = list of bits); complex scalar."""
        ops = ops or {}
        L = torch.ones((1, 1), dtype=self.dtype)
        for k, T in enumerate(self.T):
            if proj is not None:
                Tb = T[:, proj[k]]
                L = torch.einsum('ab,axr,bxq->rq', L, Tb, Tb.conj())
            else:
                O = ops.get(k)
                Tk = T if O is None else torch.einsum('ts,lsar->ltar', O.to(self.dtype), T)
                L = torch.einsum('ab,asxr,bsxq->rq', L, Tk, T.conj())
        return L.reshape(())

    def trace(self):
        return self.chain().real

    def rdm(self, sites):
        """Reduced density matrix on `sites` (sorted list), everything else traced."""
        R = torch.ones((1, 1, 1), dtype=self.dtype)
        for k, T in enumerate(self.T):
            if k in sites:
                R = torch.einsum('Pab,asxr,btxq->Pstrq', R, T, T.conj()).reshape(-1, T.shape[-1], T.shape[-1])
            else:
                R = torch.einsum('Pab,asxr,bsxq->Prq', R, T, T.conj())
        m = len(sites)
        rho = R.reshape([2, 2] * m)
        perm = [2 * i for i in range(m)] + [2 * i + 1 for i in range(m)]
        return rho.permute(perm).reshape(2 ** m, 2 ** m)
