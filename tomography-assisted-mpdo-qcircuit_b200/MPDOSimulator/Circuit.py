"""TensorCircuit of the B200 build: the reference's evolve / gate absorption / readout / sampling surface
(MPDOSimulator/Circuit.py) over dense site tensors T_k[B, l, s, a, r] and the CUDA kernels in
libmpdo_b200.so. The host side stays Python; nothing here computes on the CPU except tiny gate operands.

Behaviour kept from the reference on purpose: `truncate()` is a no-op until every
neighbouring pair shares a bond; the noise-tensor cache is keyed by gate name; in idealNoise mode a
variational two-qubit gate raises; in unified mode two-qubit gates carry no noise and in realNoise mode
single-qubit gates carry none; the rank of a gate split follows ||s|| - ||s[:k]|| <= e * 1e-8.
Extensions: gate angles may be 1-D tensors (B circuits evolved as one batch), non-neighbouring two-qubit
gates are routed through noiseless SWAPs (the reference accepts them and then breaks in truncate), and
cal_dm(reduced_index=[...]) really traces those qubits (the reference call raises)."""
import os
from typing import Any, Dict, List, Optional, Tuple, Union

import torch as tc
from torch.linalg import LinAlgError  # noqa: F401  (part of the reference's error surface)

from . import _engine
from ._node import DenseNode, replicate_nodes
from ._engine.strands import adopt, hand_over, memory_guard, run_strands
from .AbstractCircuit import QuantumCircuit
from .NoiseChannel import NoiseChannel
from .QuantumGates.AbstractGate import QuantumGate
from .QuantumGates.SingleGates import MeasureX, MeasureY
from .TNNOptimizer import bondTruncate, svdKappa_left2right, checkConnectivity, truncateLayer  # noqa: F401
from .Tools import count_item
from .dmOperations import DmNodes

GLOBAL_MINIMUM = 2.718281828459045 * 1e-8
MAX_ATTEMPTS = 4


class TensorCircuit(QuantumCircuit):
    def __init__(self, qn: int, ideal: bool = True, noiseType: str = 'no',
                 chiFileDict: Optional[Dict[str, Dict[str, Any]]] = None,
                 chi: Optional[int] = None, kappa: Optional[int] = None,
                 max_truncation_err: Optional[float] = None, chip: Optional[str] = None,
                 dtype: Optional = tc.complex64, device: Optional[Union[str, int]] = None):
        super(TensorCircuit, self).__init__(
            noiseFiles=chiFileDict, chi=chi, kappa=kappa, max_truncation_err=max_truncation_err,
            dtype=dtype, device=device
        )
        self.qnumber = qn
        self.ideal = ideal
        self.noiseType = noiseType.lower()
        self.Noise = None
        if not ideal:
            if self.noiseType not in ['unified', 'realnoise', 'idealnoise']:
                raise ValueError(f'Unsupported noise type: {self.noiseType}')
            self.Noise = NoiseChannel(chip=chip, dtype=dtype, device='cpu')
            self.unified = self.noiseType == 'unified'
            self.realNoise = self.noiseType == 'realnoise'
            self.idealNoise = self.noiseType == 'idealnoise'
            if self.realNoise:
                self._load_exp_tensors()
        self.last_stats = {}
        self._compiled = {}      # segment (tuple of layer indices) -> (stamp, programs, noisy 2q count)
        self._segment_start = 0

    # ------------------------------------------------------------------------------------------------
    # gate operands (host) -> device
    # ------------------------------------------------------------------------------------------------
    def _engine(self):
        return _engine.engine_for(self.dtype)

    def _dev(self, t: tc.Tensor) -> tc.Tensor:
        return t.to(device=self.device, dtype=self.dtype).contiguous()

    @staticmethod
    def _match_batch(node: DenseNode, Bg: int):
        if Bg > 1 and node.data.shape[0] == 1:
            node.data = node.data.expand(Bg, *node.data.shape[1:]).contiguous()
        elif Bg > 1 and node.data.shape[0] != Bg:
            raise ValueError(f'batch mismatch: state has {node.data.shape[0]} circuits, gate has {Bg}')

    @staticmethod
    def _compress_kraus(G: tc.Tensor) -> tc.Tensor:
        """G [B, (operator axes), K] -> the same channel written with at most d = (product of the operator axes)
        Kraus operators. The Kraus index is only ever traced against its own conjugate (Circuit.py:240-242), so any
        isometry on it leaves sum_g G_g rho G_g^h - and every later sweep / truncation, which see the index through
        unitarily invariant quantities only - unchanged: G -> U S from the thin SVD of the d x K matrix. A noisy
        single-qubit gate needs 4 operators instead of 6 (decay x dephasing), a fused pair of tomography CZs 16
        instead of 256: the inner index of the site tensors shrinks by the same factor before any kernel runs."""
        K = G.shape[-1]
        B = G.shape[0]
        d = G[0].numel() // K
        if K <= d or os.environ.get('MPDO_NO_KRAUS_COMPRESSION'):
            return G
        U, S, _ = tc.linalg.svd(G.reshape(B, d, K).to(tc.complex128), full_matrices=False)
        keep = max(1, int((S > 1e-15 * S[:, :1]).sum(dim=1).max()))
        out = (U * S.unsqueeze(1))[:, :, :keep]
        return out.reshape(*G.shape[:-1], keep).to(G.dtype)

    def _single_operand(self, gate: QuantumGate) -> Tuple[tc.Tensor, bool]:
        """[Bg, 2, 2, K] operand of a single-qubit gate and whether it carries noise (reference :141-155)."""
        noisy = (self.idealNoise or self.unified) and not gate.ideal
        U = gate.tensor
        if not noisy:
            G = U.reshape(-1, 2, 2, 1) if U.dim() in (2, 3) and U.shape[-2:] == (2, 2) else None
            if G is None:
                raise ValueError(f'single-qubit gate tensor of shape {tuple(U.shape)} cannot be applied without noise axes')
            return G, False

        def build():
            Ub = U.reshape(-1, 2, 2)
            return self._compress_kraus(tc.einsum('nlm, ljk, bji -> bnimk', self.Noise.decayTensor,
                                                  self.Noise.dephasingTensor, Ub).reshape(Ub.shape[0], 2, 2, -1))

        # the second value is truthy for a noisy gate and carries the number of Kraus operators of the reference's own
        # construction (2 decay x 3 dephasing), which DenseNode.ref_inner keeps track of
        kref = int(self.Noise.decayTensor.shape[-1] * self.Noise.dephasingTensor.shape[-1])
        if gate.variational:
            return build(), kref
        return self.noiseTensorDict.setdefault(gate.name, build()), kref

    def _double_operand(self, gate: QuantumGate, _oqs: List[int]) -> Tuple[tc.Tensor, bool]:
        """[Bg, 2, 2, 2, 2, K] operand in (lo, hi) qubit order and whether it adds an inner index (:84-99)."""
        noisy = (self.idealNoise and not gate.ideal) or self.realNoise
        G = gate.tensor
        if noisy and not self.realNoise:
            if gate.variational:
                raise ValueError('A variational two-qubit gate cannot carry idealNoise (the reference builds a '
                                 '5-index tensor with 4 axis names here); use cz/cx/cnot/swap/iswap.')
            base = G.reshape(-1, 2, 2, 2, 2)
            G = self.noiseTensorDict.setdefault(
                gate.name, tc.einsum('ijklp, bklmn -> bijmnp', self.Noise.dpCTensor2, base))
            noisy = int(G.shape[-1])          # truthy, and the reference's Kraus count (see _single_operand)
        elif self.realNoise:
            if G.dim() != 5:
                raise ValueError('realNoise mode needs a (2,2,2,2,K) two-qubit gate tensor (CZEXP / CPEXP).')
            G = G.unsqueeze(0)
            noisy = int(G.shape[-1])
        else:
            if G.shape[-4:] != (2, 2, 2, 2):
                raise ValueError(f'two-qubit gate tensor of shape {tuple(G.shape)} cannot be applied without a noise axis')
            G = G.reshape(-1, 2, 2, 2, 2, 1)
        if _oqs[0] > _oqs[1]:   # gate axes are ordered by _oqs (control first): bring to (lo, hi)
            G = G.permute(0, 2, 1, 4, 3, 5)
        return G, noisy

    # ------------------------------------------------------------------------------------------------
    # gate absorption
    # ------------------------------------------------------------------------------------------------
    def _apply_two_qubits_gate(self, _qNodes: List[DenseNode], _qubits, gate: QuantumGate, _oqs: List[int],
                               _tag=None):
        """Merge both sites with the (noisy) gate and split back by SVD (reference :74-136)."""
        if len(_oqs) != 2 or _oqs[0] == _oqs[1]:
            raise ValueError('Invalid operating qubits for a two-qubit gate.')
        lo, hi = min(_oqs), max(_oqs)
        G, noisy = self._double_operand(gate, _oqs)
        G = self._dev(G)
        if hi != lo + 1:
            # Non-neighbouring qubits: the reference accepts the gate but its truncate() then indexes a
            # bond that does not exist (TNNOptimizer.py:94-95). Here the higher qubit is routed next to the lower one
            # by noiseless SWAPs, the gate acts on (lo, lo+1), and the SWAPs are undone: the state on the register is
            # the one the long-range gate defines, every bond is a nearest-neighbour bond, and truncate() works.
            # The Kraus index of the gate stays on site lo+1 (it is only ever traced against its own conjugate).
            swap = self._dev(tc.tensor([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
                                       dtype=self.dtype).reshape(1, 2, 2, 2, 2, 1))
            for k in range(hi - 1, lo, -1):
                self._pair_op(_qNodes, k, k + 1, swap, False, None)
            self._pair_op(_qNodes, lo, lo + 1, G, noisy, _tag)
            for k in range(lo + 1, hi):
                self._pair_op(_qNodes, k, k + 1, swap, False, None)
            return
        self._pair_op(_qNodes, lo, hi, G, noisy, _tag)

    def _pair_op(self, _qNodes: List[DenseNode], lo: int, hi: int, G: tc.Tensor, noisy: bool, _tag=None):
        """Batch bookkeeping + merge-and-split of one neighbouring pair."""
        for q in (lo, hi):
            self._match_batch(_qNodes[q], G.shape[0])
        if _qNodes[lo].data.shape[0] != _qNodes[hi].data.shape[0]:
            B = max(_qNodes[lo].data.shape[0], _qNodes[hi].data.shape[0])
            self._match_batch(_qNodes[lo], B)
            self._match_batch(_qNodes[hi], B)
        self._merge_split(_qNodes, lo, hi, G, noisy, _tag)

    def _merge_split(self, _qNodes: List[DenseNode], lo: int, hi: int, G: tc.Tensor, noisy: bool, _tag=None):
        """Theta = T_lo . T_hi . G, split back with the reference rank rule; the noise index goes to `hi`.
        The kept rank (batch maximum) is recorded in last_stats['split_ranks'][layer index]."""
        eng = self._engine()
        before = _qNodes[hi].nominal_inner()
        _qNodes[lo].data, _qNodes[hi].data = eng.split_2q(_qNodes[lo].data, _qNodes[hi].data, G, GLOBAL_MINIMUM)
        if _tag is not None:
            self.last_stats.setdefault('split_ranks', {})[_tag] = int(_qNodes[lo].data.shape[4])
        _qNodes[lo].has_right = True
        _qNodes[hi].has_left = True
        if noisy:
            _qNodes[hi].ref_inner = before * int(noisy)
            _qNodes[hi].has_inner = True

    # Largest composite Kraus count a fused pair may carry (two tomography CZs: 16 x 16).
    FUSE_MAX_K = 256

    def _fuse_pairs(self, chain: list) -> list:
        """Fuses `2q gate A ; noiseless 1q gates on the same two qubits ; 2q gate B` (same pair, no truncate in
        between - e.g. the two CZEXP of a realNoise rzz decomposition, reference AbstractCircuit.py:263-275) into ONE
        merge-and-split with the composite Kraus tensor
            G[p0,p1,s0,s1,(gB,gA)] = sum B[p0,p1,t0,t1,gB] M0[t0,u0] M1[t1,u1] A[u0,u1,s0,s1,gA].
        The two-site tensor after gate B is the same either way; what is skipped is the reference's intermediate
        SVD split, i.e. one application of its rank rule ||s|| - ||s[:k]|| <= e*1e-8 (Circuit.py:120-124 ->
        decompositions.py:120-134; NOT a bound on the tail norm: it drops tails up to sqrt(2 ||s|| e*1e-8) ~ 2e-4
        relative) and, in complex64, one more fp32 rounding of both site tensors. This is a documented numerical
        deviation from the reference: tests/test_gpu_big_configs.py measures the fused and the gate-by-gate path
        against the exact oracle on the same chi = 64 circuit (both inside the complex64 fp32 floor), bench.py
        reports both, and complex128 circuits, circuits with a relative truncation error and MPDO_NO_FUSE=1 keep
        one split per gate. Entries of the returned list are either the
        original (index, gate, oqs) tuples or ('fused', lo, hi, G, noisy, layer index)."""
        if (self.dtype != tc.complex64 or self.max_truncation_err is not None or not self.realNoise
                or os.environ.get('MPDO_NO_FUSE')):
            return chain

        def is_2q(op):
            return isinstance(op[1], QuantumGate) and not op[1].single and len(op[2]) == 2

        out, i = [], 0
        while i < len(chain):
            op = chain[i]
            if not is_2q(op) or abs(op[2][0] - op[2][1]) != 1:
                out.append(op)
                i += 1
                continue
            lo, hi = min(op[2]), max(op[2])
            j, mids = i + 1, []
            while j < len(chain):
                g, oqs = chain[j][1], chain[j][2]
                if (isinstance(g, QuantumGate) and g.single and g.name != 'MeasureZ' and set(oqs) <= {lo, hi}
                        and not ((self.idealNoise or self.unified) and not g.ideal)):
                    mids.append(chain[j])
                    j += 1
                else:
                    break
            if j >= len(chain) or not is_2q(chain[j]) or {min(chain[j][2]), max(chain[j][2])} != {lo, hi}:
                out.append(op)
                i += 1
                continue
            A, noisyA = self._double_operand(op[1], op[2])
            Bt, noisyB = self._double_operand(chain[j][1], chain[j][2])
            if A.shape[-1] * Bt.shape[-1] > self.FUSE_MAX_K:
                out.append(op)
                i += 1
                continue
            eye = tc.eye(2, dtype=tc.complex128).reshape(1, 2, 2)
            M = {lo: eye, hi: eye}
            for _, g, oqs in mids:
                U = self._single_operand(g)[0].to(tc.complex128)[..., 0]      # [Bg, 2, 2] (p, s)
                for q in oqs:
                    M[q] = U @ M[q]
            tot = tc.einsum('bpqtvg, btu, bvw, buwxyh -> bpqxygh', Bt.to(tc.complex128), M[lo], M[hi],
                            A.to(tc.complex128))
            tot = self._compress_kraus(tot.reshape(*tot.shape[:5], -1))
            kref = (int(noisyA) or 1) * (int(noisyB) or 1) if (noisyA or noisyB) else False
            out.append(('fused', lo, hi, tot, kref, chain[j][0]))
            i = j + 1
        return out

    def _apply_single_qubit_gate(self, _qNodes: List[DenseNode], _qubits, gate: QuantumGate,
                                 _oqs: Union[int, List[int]]):
        """Contract the (noisy) gate into each operating site; the Kraus index joins the inner index (:138-178)."""
        G, noisy = self._single_operand(gate)
        G = self._dev(G)
        eng = self._engine()
        for q in _oqs:
            self._match_batch(_qNodes[q], G.shape[0])
            before = _qNodes[q].nominal_inner()
            _qNodes[q].data = eng.absorb_1q(_qNodes[q].data, G)
            if noisy:
                _qNodes[q].ref_inner = before * int(noisy)
                _qNodes[q].has_inner = True

    def _add_gate(self, _qubits: List[DenseNode], _layer_num: int, _oqs: List[int], _gate: Optional = None):
        if not isinstance(_qubits, List):
            raise TypeError('Qubit must be a list of nodes.')
        if not isinstance(_oqs, List):
            raise TypeError('Operating qubits must be a list.')
        if _oqs[0] is None:
            return None
        if max(_oqs) >= self.qnumber:
            raise ValueError(f'Qubit index out of range, max index is Q{max(_oqs)}.')
        gate = _gate or self.layers[_layer_num]
        if not gate or gate.name == 'MeasureZ':
            return None
        if not isinstance(gate, QuantumGate):
            raise TypeError(f'Gate must be a QuantumGate, current type is {type(gate)}.')
        if not gate.single:
            self._apply_two_qubits_gate(_qubits, None, gate, _oqs, _tag=_layer_num)
        else:
            self._apply_single_qubit_gate(_qubits, None, gate, _oqs)

    # ------------------------------------------------------------------------------------------------
    # evolution
    # ------------------------------------------------------------------------------------------------
    def evolve(self, state: List[DenseNode]):
        """Run every layer on `state` in place (reference :469-491). Host-resident state tensors are uploaded to
        the circuit's device first; afterwards `state` (aliased as self.stateNodes) lives on the device."""
        if not isinstance(state, list):
            raise TypeError('state must be a list of nodes')
        self._initState = replicate_nodes(state)
        for node in state:
            node.data = node.data.to(device=self.device, dtype=self.dtype)
        self._dm, self._dmNodes, self._vector = None, None, None
        self.last_stats = {'noisy_2q_updates': 0}

        layers = list(self.layers)   # nn.Sequential indexing is O(n) per access
        segment = []                 # gates since the last truncate: independent qubit groups run concurrently
        for _i, layer in enumerate(layers):
            name = layer.name.lower()
            if 'truncate' in name:
                self._run_segment(state, segment)
                memory_guard(self.device, [s.data for s in state])      # large states only (cfg5), see strands.py
                if checkConnectivity(state):
                    truncateLayer(state, chi=self.chi, kappa=self.kappa, max_truncation_err=self.max_truncation_err,
                                  noisy=not self.ideal)
                    memory_guard(self.device, [s.data for s in state])
            elif 'barrier' in name:
                pass
            else:
                segment.append((_i, layer, self._oqs_list[_i]))
        self._run_segment(state, segment)

        if not self.ideal and layers and 'truncate' not in layers[-1].name:
            svdKappa_left2right(state, max_singular_values=self.kappa, max_truncation_err=self.max_truncation_err)
        self._stateNodes = state

    # ------------------------------------------------------------------------------------------------
    # segment compilation (host): strands, primitive steps, fused Kraus composites, device operands
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _stamp(ops):
        """Identity + in-place version of every tensor a gate of the segment holds: a compiled segment is reused
        only while its gates are the ones it was compiled from."""
        out = [(os.environ.get('MPDO_NO_FUSE'), os.environ.get('MPDO_NO_KRAUS_COMPRESSION'))]   # knobs read by _program
        for _, g, oqs in ops:
            out.append((id(g), tuple(oqs)))
            for v in vars(g).values():
                if isinstance(v, tc.Tensor):
                    out.append((id(v), v._version))
        return tuple(out)

    def _compile_segment(self, ops: list):
        """Host side of one segment (the gates between two truncate markers): split the gates into strands of qubits
        that do not interact inside the segment, resolve every strand into primitive steps on gate operands
        (_program: Kraus composites of fused pairs, minimal Kraus representation, SWAP routing) and place the
        operands on the circuit's device. Pure operand building, as the reference's gate constructors and its
        noiseTensorDict cache do - done once per segment when `truncate()` closes it (or on first use) and reused by
        every evolve() while the gates are unchanged. Returns (stamp, programs, noisy two-qubit gate count)."""
        for _, _, oqs in ops:
            if not isinstance(oqs, List):
                raise TypeError('Operating qubits must be a list.')
            if max(oqs) >= self.qnumber:
                raise ValueError(f'Qubit index out of range, max index is Q{max(oqs)}.')
        parent = list(range(self.qnumber))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for _, _, oqs in ops:
            for q in range(min(oqs), max(oqs) + 1):   # a long-range gate is routed through the qubits in between
                parent[find(q)] = find(oqs[0])
        strands = {}
        for op in ops:
            strands.setdefault(find(op[2][0]), []).append(op)
        noisy2q = sum(1 for _, g, _ in ops if isinstance(g, QuantumGate) and not g.single
                      and ((self.idealNoise and not g.ideal) or self.realNoise))
        programs = [self._program(chain) for chain in strands.values()]
        if tc.device(self.device).type == 'cuda' and tc.cuda.is_available():
            uploaded = {}

            def up(G):      # one device tensor per distinct operand VALUE: _run_group recognises equal operands of
                key = (tuple(G.shape), G.dtype, G.contiguous().numpy().tobytes())   # stacked strands by identity
                if key not in uploaded:
                    uploaded[key] = self._dev(G)
                return uploaded[key]

            programs = [(qubits, [st[:2] + (up(st[2]),) + st[3:] if st[0] == '1q' else st[:3] + (up(st[3]),) + st[4:]
                                  for st in steps]) for qubits, steps in programs]
        return self._stamp(ops), programs, noisy2q

    def _segment_ops(self, first: int, last: int):
        import itertools
        mods = itertools.islice(self.layers._modules.values(), first, last)   # nn.Sequential indexing is O(n)
        return [(i, g, self._oqs_list[i]) for i, g in zip(range(first, last), mods)
                if self._oqs_list[i] and self._oqs_list[i][0] is not None
                and 'barrier' not in getattr(g, 'name', '').lower()]

    def truncate(self):
        """Add a truncation layer; the segment it closes is compiled here (host operand building belongs to circuit
        construction, as in the reference, not to evolve). Construction never raises for it: a segment that cannot be
        compiled raises from evolve(), where the reference raises."""
        first = getattr(self, '_segment_start', 0)
        super().truncate()
        last = len(self._oqs_list) - 1
        self._segment_start = last + 1
        if os.environ.get('MPDO_LAZY_COMPILE'):
            return
        try:
            ops = self._segment_ops(first, last)
            if ops:
                self._compiled[tuple(i for i, _, _ in ops)] = self._compile_segment(ops)
        except Exception:   # noqa: BLE001  (evolve() compiles again and raises there)
            pass

    def _run_segment(self, state: List[DenseNode], segment: list):
        """Apply the gates collected since the last truncate. Gates are grouped into strands of qubits that do
        not interact inside the segment (order inside a strand is the circuit order); strands are independent,
        so they are issued concurrently (one CUDA stream each, _engine/strands.py)."""
        if not segment:
            return
        ops = [(i, g, oqs) for i, g, oqs in segment if oqs and oqs[0] is not None]
        segment.clear()
        key = tuple(i for i, _, _ in ops)
        entry = self._compiled.get(key)
        if entry is None or entry[0] != self._stamp(ops):
            entry = self._compiled[key] = self._compile_segment(ops)
        _, programs, noisy2q = entry
        self.last_stats['noisy_2q_updates'] = self.last_stats.get('noisy_2q_updates', 0) + noisy2q

        # Strands are issued concurrently, one CUDA stream each; strands whose step sequences and tensor shapes coincide
        # (the bulk brick pairs of a layer) are stacked along the batch axis and run as ONE launch sequence when the
        # circuit batch is small (_engine.grouping_enabled: a third fewer launches and 5 % per cfg2 layer; for large
        # batches one stream serialises the Gram contractions of one pair behind the factorisations of another and
        # separate streams win - cfg4 8.4 vs 12.9 circuits/s, profiles/r2_grouping.md).
        cuda = getattr(_engine.get_prims(), 'name', '') == 'cuda'
        batch = max([s.data.shape[0] for s in state] +
                    [st[-3 if st[0] == '2q' else -2].shape[0] for _, steps in programs for st in steps])
        grouped = _engine.grouping_enabled(batch)
        groups = {}
        for prog in programs:
            key = self._signature(state, prog) if grouped else id(prog)
            groups.setdefault(key, []).append(prog)
        parallel = cuda and len(groups) > 1

        def make(members):
            def task():
                if parallel:
                    for qubits, _ in members:
                        for q in qubits:
                            adopt(state[q].data)
                self._run_group(state, members)
            return task

        tasks = [make(members) for members in groups.values()]
        run_strands(tasks, self.device, enabled=parallel)
        if parallel:
            for qubits, _ in programs:
                for q in qubits:
                    hand_over(state[q].data, self.device)

    _SWAP = tc.tensor([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])

    def _program(self, chain: list):
        """Resolve one strand (gates in circuit order on a set of qubits that touch nothing else in this segment) into
        primitive steps on host operands: ('1q', q, G [Bg,2,2,K], noisy) and ('2q', lo, lo+1, G [Bg,2,2,2,2,K], noisy,
        tag). Returns (sorted qubits, steps)."""
        steps = []
        for op in self._fuse_pairs(chain):
            if op[0] == 'fused':
                _, lo, hi, G, noisy, tag = op
                steps.append(('2q', lo, hi, G, noisy, tag))
                continue
            i, gate, oqs = op
            if not gate or gate.name == 'MeasureZ':
                continue
            if not isinstance(gate, QuantumGate):
                raise TypeError(f'Gate must be a QuantumGate, current type is {type(gate)}.')
            if gate.single:
                G, noisy = self._single_operand(gate)
                steps.extend(('1q', q, G, noisy) for q in oqs)
                continue
            if len(oqs) != 2 or oqs[0] == oqs[1]:
                raise ValueError('Invalid operating qubits for a two-qubit gate.')
            lo, hi = min(oqs), max(oqs)
            G, noisy = self._double_operand(gate, oqs)
            swap = self._SWAP.to(self.dtype).reshape(1, 2, 2, 2, 2, 1)
            for k in range(hi - 1, lo, -1):          # long-range gate: route through noiseless SWAPs (see
                steps.append(('2q', k, k + 1, swap, False, None))      # _apply_two_qubits_gate)
            steps.append(('2q', lo, lo + 1, G, noisy, i))
            for k in range(lo + 1, hi):
                steps.append(('2q', k, k + 1, swap, False, None))
        qubits = sorted({q for st in steps for q in st[1:(2 if st[0] == '1q' else 3)]})
        return qubits, steps

    @staticmethod
    def _signature(state, prog):
        """Two strands with the same signature run the same kernels on tensors of the same shapes."""
        qubits, steps = prog
        rel = {q: j for j, q in enumerate(qubits)}
        shapes = tuple(tuple(state[q].data.shape[1:]) for q in qubits)
        flags = tuple((state[q].has_left, state[q].has_right, state[q].has_inner) for q in qubits)
        seq = tuple((st[0],) + tuple(rel[q] for q in st[1:(2 if st[0] == '1q' else 3)]) + (tuple(st[-3 if st[0] == '2q' else -2].shape[1:]),)
                    for st in steps)
        return shapes, flags, seq

    def _run_group(self, state: List[DenseNode], members: list):
        """Execute strands with identical signatures as one batched launch sequence: member m's circuits occupy rows
        m*B .. (m+1)*B of every stacked tensor. A single member runs in place (no stacking)."""
        eng = self._engine()
        M = len(members)
        qubits0, steps0 = members[0]
        if not steps0:
            return
        Bmax = max([state[q].data.shape[0] for qs, _ in members for q in qs] +
                   [st[-3 if st[0] == '2q' else -2].shape[0] for _, sts in members for st in sts])
        stacked = []
        for j in range(len(qubits0)):
            parts = []
            for qs, _ in members:
                node = state[qs[j]]
                self._match_batch(node, Bmax)
                parts.append(node.data)
            stacked.append(parts[0] if M == 1 else tc.cat(parts, dim=0))
        rel = {q: j for j, q in enumerate(qubits0)}
        flags = [dict(l=state[q].has_left, r=state[q].has_right, i=state[q].has_inner, k=1) for q in qubits0]
        nominal = [[state[qs[j]].nominal_inner() for j in range(len(qubits0))] for qs, _ in members]
        ranks = {}
        for t, st in enumerate(steps0):
            Gs = [sts[t][-3 if st[0] == '2q' else -2] for _, sts in members]
            if all(g.shape[0] == 1 for g in Gs) and all(g is Gs[0] or (not g.is_cuda and tc.equal(g, Gs[0])) for g in Gs[1:]):
                G = Gs[0]
            else:
                G = tc.cat([g.expand(Bmax, *g.shape[1:]) for g in Gs], dim=0)
            G = self._dev(G)
            if st[0] == '1q':
                j = rel[st[1]]
                stacked[j] = eng.absorb_1q(stacked[j], G)
                flags[j]['i'] = flags[j]['i'] or bool(st[3])
                flags[j]['k'] *= int(st[3]) or 1
            else:
                jl, jh = rel[st[1]], rel[st[2]]
                stacked[jl], stacked[jh] = eng.split_2q(stacked[jl], stacked[jh], G, GLOBAL_MINIMUM)
                flags[jl]['r'] = flags[jh]['l'] = True
                flags[jh]['i'] = flags[jh]['i'] or bool(st[4])
                flags[jh]['k'] *= int(st[4]) or 1
                per_row = getattr(eng.tls, 'last_ranks', None)
                for m, (_, sts) in enumerate(members):
                    tag = sts[t][5]
                    if tag is not None:
                        ranks[tag] = (int(max(per_row[m * Bmax:(m + 1) * Bmax])) if per_row is not None
                                      else int(stacked[jl].shape[4]))
        if ranks:
            self.last_stats.setdefault('split_ranks', {}).update(ranks)
        for m, (qs, _) in enumerate(members):
            for j, q in enumerate(qs):
                node = state[q]
                node.data = stacked[j] if M == 1 else stacked[j][m * Bmax:(m + 1) * Bmax]
                node.has_left, node.has_right, node.has_inner = flags[j]['l'], flags[j]['r'], flags[j]['i']
                if flags[j]['k'] > 1:
                    node.ref_inner = nominal[m][j] * flags[j]['k']

    def forward(self, state: List[DenseNode]):
        self.evolve(state)

    # ------------------------------------------------------------------------------------------------
    # readout
    # ------------------------------------------------------------------------------------------------
    def _Ts(self, nodes=None):
        nodes = self._stateNodes if nodes is None else nodes
        if nodes is None:
            raise RuntimeError('evolve() the circuit first')
        B = max(n.data.shape[0] for n in nodes)
        return [n.data if n.data.shape[0] == B else n.data.expand(B, *n.data.shape[1:]).contiguous() for n in nodes]

    def _create_dmNodes(self, _stateNodes: Optional[List[DenseNode]] = None, reduced_index: Optional[List] = None):
        nodes = self._stateNodes if _stateNodes is None else _stateNodes
        dm = DmNodes(nodes, reduced=reduced_index or [])
        return dm.state_nodes, dm.conj_nodes

    def cal_dmNodes(self, reduced_index: Optional[List] = None):
        """The un-contracted density operator: n ket-side nodes followed by n conjugated ones. As in the
        reference (Circuit.py:244 passes reduced_index into another slot) no qubit is traced here."""
        self._dmNodes = DmNodes(self._stateNodes, reduced=[])
        return self._dmNodes

    def cal_dm(self, reduced_index: Optional[List] = None, _replicate_require: bool = False):
        """Dense, un-normalised density matrix of the qubits not in reduced_index (small registers only)."""
        reduced_index = reduced_index or []
        if self._dmNodes is None or _replicate_require:
            self._dmNodes = DmNodes(self._stateNodes, reduced=[])
        eng = self._engine()
        keep = [i for i in range(self.qnumber) if i not in reduced_index]
        dm = eng.dense_rho(self._Ts(), keep=keep).to(self.dtype)
        self._dm = dm[0] if dm.shape[0] == 1 else dm
        return self._dm

    def cal_vector(self):
        if not self.ideal:
            raise ValueError('Noisy circuit cannot be represented by state vector efficiently.')
        v = self._engine().dense_vector(self._Ts()).to(self.dtype)
        self._vector = v[0].reshape(-1, 1) if v.shape[0] == 1 else v.unsqueeze(-1)
        return self._vector

    def bitstring_probabilities(self, bitstrings, normalize: bool = False) -> tc.Tensor:
        """<b|rho|b> for a batch of full-register bitstrings (list of lists / strings of 0/1) by one
        transfer-matrix chain (the product of the reference's conditional probabilities, Circuit.py:297-332)."""
        bits = [[int(c) for c in b] if isinstance(b, str) else list(b) for b in bitstrings]
        eng = self._engine()
        probs = eng.bitstring_probs(self._Ts(), bits)
        if normalize:
            probs = probs / eng.chain_value(self._Ts())[0].real
        return probs

    # ---- sampling (reference :297-467) --------------------------------------------------------------
    def _conditional_prb(self, _history: List[Union[int, bool]]) -> float:
        """P(next measured qubit = 1 | history) on the prepared sampling state."""
        ctx = self._nodes4samples
        L = ctx['L0']
        for j, bit in enumerate(_history):
            L = self._project(L, j, int(bit))
        return self._p1_from_left(L, len(_history))

    def _project(self, L, pos, bit):
        """Left environment after measured qubit number `pos` gave `bit`, carried on through the traced
        (unmeasured) qubits that follow it up to the next measured one."""
        ctx = self._nodes4samples
        eng, Ts, measured = ctx['eng'], ctx['Ts'], ctx['measured']
        q = measured[pos]
        L = eng.transfer_proj(L, Ts[q], bit)
        stop = measured[pos + 1] if pos + 1 < len(measured) else q + 1
        for k in range(q + 1, stop):
            L = eng.transfer(L, Ts[k])
        return L

    def _p1_from_left(self, L, pos):
        ctx = self._nodes4samples
        eng, Ts, Rs = ctx['eng'], ctx['Ts'], ctx['R']
        q = ctx['measured'][pos]
        vals = []
        for bit in (0, 1):
            Lb = eng.transfer_proj(L, Ts[q], bit)
            vals.append(eng.inner(Lb, Rs[q + 1])[0].real.item())
        probs = tc.tensor(vals, dtype=tc.float64) + GLOBAL_MINIMUM
        if (probs < 0).sum() > 0:
            raise RuntimeError(f"State is illegal, and your probability distribution is {probs}.")
        return (probs / probs.sum())[-1].item()

    def _prepare_sampling(self, nodes, measured: List[int]):
        """Environments for sampling the qubits in `measured` (ascending); every other qubit is traced inside the
        transfer-matrix chain (the reference connects its physical legs, Circuit.py:314 / dmOperations.py:18-37).
        R[k] = trace of sites k..n-1 as [1, l, l'] (complex128); L0 = trace of the sites left of the first
        measured qubit."""
        eng = self._engine()
        Ts_all = self._Ts(nodes)
        assert Ts_all[0].shape[0] == 1, 'sampling works on a single circuit'
        n = len(Ts_all)
        if not measured or sorted(set(measured)) != list(measured) or measured[-1] >= n:
            raise ValueError('measured qubits must be a non-empty ascending list of distinct qubit indices')
        R = [None] * (n + 1)
        R[n] = tc.ones((1, 1, 1), dtype=tc.complex128, device=Ts_all[0].device)
        for k in range(n - 1, measured[0], -1):
            R[k] = eng.transfer_right(R[k + 1], Ts_all[k])
        L0 = tc.ones((1, 1, 1), dtype=tc.complex128, device=Ts_all[0].device)
        for k in range(measured[0]):
            L0 = eng.transfer(L0, Ts_all[k])
        self._nodes4samples = {'eng': eng, 'Ts': Ts_all, 'R': R, 'L0': L0, 'measured': list(measured)}
        self._indices4samples = measured

    def _conditional_batch_sample(self, shots: int, _sampleLength: int, _bool: bool = False,
                                  _tqdm_disable: bool = False) -> List[List[int]]:
        """Breadth-first over outcome prefixes: one conditional probability per distinct prefix, shot counts
        split by Bernoulli draws (reference :344-387). Left environments are carried per prefix group."""
        dtype = tc.bool if _bool else tc.int
        sequences = tc.zeros((shots, _sampleLength), dtype=dtype)
        ctx = self._nodes4samples
        groups = [(ctx['L0'], 0, shots)]
        for pos in range(_sampleLength):
            nxt = []
            for L, start, length in groups:
                if length == 0:
                    continue
                p1 = self._p1_from_left(L, pos)
                if p1 < GLOBAL_MINIMUM:
                    k1 = 0
                elif p1 > 1 - GLOBAL_MINIMUM:
                    k1 = length
                else:
                    k1 = int(tc.bernoulli(tc.full((length,), p1)).sum())
                k0 = length - k1
                sequences[start:start + k1, pos] = True if _bool else 1
                if pos < _sampleLength - 1:
                    if k1 > 0:
                        nxt.append((self._project(L, pos, 1), start, k1))
                    if k0 > 0:
                        nxt.append((self._project(L, pos, 0), start + k1, k0))
            groups = nxt
        return sequences.tolist()

    def _conditional_sample(self, shots: int, _sampleLength: int, _tqdm_disable: bool = False) -> List[List[int]]:
        out = []
        for _ in range(shots):
            choices = []
            for _j in range(_sampleLength):
                p1 = self._conditional_prb(choices)
                choices.append(int(tc.multinomial(tc.tensor([1 - p1, p1]), num_samples=1).item()))
            out.append(choices)
        return out

    def sample(self, shots: Optional[int] = None, orientation: Optional[List[int]] = None,
               reduced: Optional[List[int]] = None, sample_string: bool = True, _tqdm_disable: bool = False,
               _require_sequential_sample: bool = False, _require_bool_result: bool = False,
               _require_counts: bool = True, _stateNodes4Sample: Optional[List[DenseNode]] = None):
        """Sample measurement outcomes (orientation per measured qubit: 0 = X, 1 = Y, 2 = Z)."""
        shots = shots or 1024
        reduced = reduced or []
        ori_list = [q for q in range(self.qnumber) if q not in reduced]
        length = len(ori_list)
        orientation = orientation or [2] * length
        if len(orientation) != length:
            raise ValueError("Length of orientation must match the sample length. Check reduced or unmeasured qubits.")
        nodes = replicate_nodes(self._stateNodes if _stateNodes4Sample is None else _stateNodes4Sample)
        for val, cls in [(0, MeasureX), (1, MeasureY)]:
            idx = [ori_list[i] for i, v in enumerate(orientation) if v == val]
            if idx:
                self._add_gate(nodes, 0, idx, cls(dtype=self.dtype, device='cpu'))
        self._prepare_sampling(nodes, ori_list)
        if not _tqdm_disable:
            print('Sample Direction:\n (scheme-' + ''.join(self._projectors_string[v] for v in orientation) + ')')
        if _require_sequential_sample:
            bitstrings = self._conditional_sample(shots, length, _tqdm_disable)
        else:
            bitstrings = self._conditional_batch_sample(shots, length, _bool=_require_bool_result,
                                                        _tqdm_disable=_tqdm_disable)
        if sample_string and not _require_bool_result:
            bitstrings = [''.join(map(str, b)) for b in bitstrings]
        self._samples = bitstrings
        if _require_counts:
            self._counts = count_item(bitstrings)
            return self._samples, self._counts
        return self._samples

    def randomSample(self, measurement_schemes: List[List[int]], shots_per_scheme: int = 1024,
                     reduced: Optional[List[int]] = None, _tqdm_disable: bool = False,
                     _require_sequential_sample: bool = False, _require_bool_result: bool = False,
                     _stateNodes4Sample: Optional[List[DenseNode]] = None):
        return [
            self.sample(shots=shots_per_scheme, orientation=scheme, reduced=reduced, sample_string=False,
                        _tqdm_disable=True, _require_sequential_sample=_require_sequential_sample,
                        _require_bool_result=_require_bool_result, _require_counts=False,
                        _stateNodes4Sample=_stateNodes4Sample)
            for scheme in measurement_schemes
        ]
