"""Circuit construction API of the B200 build.

Mirrors the reference's MPDOSimulator/AbstractCircuit.py: same constructor state (:35-64), the same 35 builder
methods with identical argument order (:156-532), the same realNoise gate decompositions
(cx/cnot -> ry(-pi/2).CZEXP.ry(pi/2), rzz -> cx.rz.cx, rxx, ryy; :231-342), and the same headline strings.
Pure host work: a builder only records a gate module and its operating qubits; the arithmetic happens in
TensorCircuit.evolve on the device."""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

from torch import Tensor, complex64, nn, pi, tensor

from .RealNoise import czExp_channel, cpExp_channel
from .Tools import select_device

CHIFILENAMES = {
    'CZ': {'01': './MPDOSimulator/chi/czDefault.mat'},
    'CP': {}
}


def _as_list(oqs):
    return [oqs] if isinstance(oqs, int) else oqs


def _num(x):
    return x.item() if isinstance(x, Tensor) and x.numel() == 1 else (float(x.reshape(-1)[0]) if isinstance(x, Tensor) else x)


class QuantumCircuit(ABC, nn.Module):
    """Abstract circuit: layer list, noise-model switches, readout caches."""

    def __init__(self, noiseFiles: Optional[Dict[str, Dict[str, Any]]] = None,
                 chi: Optional[int] = None, kappa: Optional[int] = None,
                 max_truncation_err: Optional[float] = None,
                 dtype=complex64, device: Union[str, int] = 'cpu'):
        super(QuantumCircuit, self).__init__()
        self.device = select_device(device)
        self.dtype = dtype
        self.chi, self.kappa, self.max_truncation_err = chi, kappa, max_truncation_err

        self.layers = nn.Sequential()
        self._oqs_list = []

        self.noiseTensorDict = {}
        self.unified, self.realNoise, self.idealNoise = False, False, False
        self.noiseFiles = noiseFiles if noiseFiles is not None else CHIFILENAMES

        self._initState = None
        self._vector = None
        self._stateNodes, self._dm, self._dmNodes = None, None, None
        self._samples, self._counts = None, None

        self._sequence = 0
        self._projectors_string = ['X', 'Y', 'Z']
        self._nodes4samples, self._indices4samples = None, None

    def _load_exp_tensors(self):
        """chi-matrix files -> (2,2,2,2,K) tensors, keyed by the qubit-pair string (reference :66-75). Built on
        the host in complex64 and upcast only here, as the reference does."""
        self._cz_expTensors, self._cp_expTensors = {}, {}
        cache = {}
        for kind, loader, store in (('CZ', czExp_channel, self._cz_expTensors),
                                    ('CP', cpExp_channel, self._cp_expTensors)):
            for key, filename in self.noiseFiles.get(kind, {}).items():
                if (kind, filename) not in cache:
                    cache[(kind, filename)] = loader(filename=filename).to(dtype=self.dtype)
                store[key] = cache[(kind, filename)]

    @dataclass
    class Group:
        history: List[int]
        start: int
        length: int

    dm = property(lambda self: self._dm)
    samples = property(lambda self: self._samples)
    counts = property(lambda self: self._counts)
    initial_state = property(lambda self: self._initState)
    stateNodes = property(lambda self: self._stateNodes)
    dmNodes = property(lambda self: self._dmNodes)
    vector = property(lambda self: self._vector)

    @abstractmethod
    def cal_vector(self):
        pass

    @abstractmethod
    def cal_dm(self):
        pass

    @abstractmethod
    def evolve(self, state):
        pass

    # ------------------------------------------------------------------------------------------------
    def _add_module(self, _gate: nn.Module, oqs: List, headline: str):
        self._oqs_list.append(oqs)
        self.layers.add_module(headline + f'-S{self._sequence}', _gate)
        self._sequence += 1

    def _gate_kwargs(self):
        return dict(dtype=self.dtype, device='cpu')   # gate tensors are host operands; uploaded when applied

    def _fixed(self, module, cls_name, label, oqs, _ideal):
        import importlib
        cls = getattr(importlib.import_module(f'.QuantumGates.{module}', __package__), cls_name)
        self._add_module(cls(_ideal, **self._gate_kwargs()), oqs, f"{label}{oqs}|None")

    def _param(self, module, cls_name, label, oqs, _ideal, *params):
        import importlib
        cls = getattr(importlib.import_module(f'.QuantumGates.{module}', __package__), cls_name)
        self._add_module(cls(*params, _ideal, **self._gate_kwargs()), oqs,
                         f"{label}{oqs}|({_num(params[0]):.3f})".replace('.', ';'))

    def _iter_add_module(self, _gate_list: List, oqs_list: List, _transpile: bool = False):
        for _gate, _oq in zip(_gate_list, oqs_list):
            if _gate.para is None:
                _headline = f"{_gate.name}{_oq}|None-TRANS"
            else:
                _headline = f"{_gate.name}{_oq}|({_num(_gate.para):.3f})-TRANS".replace('.', ';')
            self._add_module(_gate, _oq, _headline)

    # ---- single-qubit gates ---------------------------------------------------------------------------
    def i(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('SingleGates', 'IGate', 'I', _as_list(oqs), _ideal)

    def h(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('SingleGates', 'HGate', 'H', _as_list(oqs), _ideal)

    def x(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('XGates', 'XGate', 'X', _as_list(oqs), _ideal)

    def y(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('YGates', 'YGate', 'Y', _as_list(oqs), _ideal)

    def z(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('ZGates', 'ZGate', 'Z', _as_list(oqs), _ideal)

    def s(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('PhaseGates', 'SGate', 'S', _as_list(oqs), _ideal)

    def sdg(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('PhaseGates', 'SDGGate', 'SDG', _as_list(oqs), _ideal)

    def t(self, oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._fixed('PhaseGates', 'TGate', 'T', _as_list(oqs), _ideal)

    def rx(self, theta: Union[Tensor, float], oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._param('XGates', 'RXGate', 'RX', _as_list(oqs), _ideal, theta)

    def ry(self, theta: Union[Tensor, float], oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._param('YGates', 'RYGate', 'RY', _as_list(oqs), _ideal, theta)

    def rz(self, theta: Union[Tensor, float], oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._param('ZGates', 'RZGate', 'RZ', _as_list(oqs), _ideal, theta)

    def p(self, theta: Union[Tensor, float], oqs: Union[List, int], _ideal: Optional[bool] = None):
        self._param('PhaseGates', 'PGate', 'P', _as_list(oqs), _ideal, theta)

    def u1(self, theta: Union[Tensor, float], oqs: List, _ideal: Optional[bool] = None):
        self._param('SingleGates', 'U1Gate', 'U1', _as_list(oqs), _ideal, theta)

    def u2(self, phi: Union[Tensor, float], lam: Union[Tensor, float], oqs: List, _ideal: Optional[bool] = None):
        from .QuantumGates.SingleGates import U2Gate
        oqs = _as_list(oqs)
        _headline = f"U2{oqs}|(P{_num(phi):3f})-(L{_num(lam):3f})".replace('.', ';')
        self._add_module(U2Gate(phi, lam, _ideal, **self._gate_kwargs()), oqs, _headline)

    def u3(self, theta: Union[Tensor, float], phi: Union[Tensor, float], lam: Union[Tensor, float],
           oqs: List, _ideal: Optional[bool] = None):
        from .QuantumGates.SingleGates import U3Gate
        oqs = _as_list(oqs)
        _headline = f"U3{oqs}|(T{_num(theta):3f})-(P{_num(phi):3f})-(L{_num(lam):3f})".replace('.', ';')
        self._add_module(U3Gate(theta, phi, lam, _ideal, **self._gate_kwargs()), oqs, _headline)

    def arbSingle(self, data: Tensor, oqs: List, _ideal: Optional[bool] = None):
        from .QuantumGates.SingleGates import ArbSingleGate
        oqs = _as_list(oqs)
        self._add_module(ArbSingleGate(data, _ideal, **self._gate_kwargs()), oqs, f"ArbS{oqs}|None")

    # ---- two-qubit gates ------------------------------------------------------------------------------
    def _exp_decomposed(self, _ideal):
        """True when a two-qubit gate has to be rewritten over the tomography CZ (realNoise, not forced ideal)."""
        return self.realNoise and not _ideal

    def rxx(self, theta: Union[Tensor, float], oq0: int, oq1: int, _ideal: Optional[bool] = None):
        oqs = [oq0, oq1]
        if not self._exp_decomposed(_ideal):
            self._param('XGates', 'RXXGate', 'RXX', oqs, _ideal, theta)
        else:
            self.h(oqs, True)
            self.cx(oq0, oq1)
            self.rz(theta, oq1, True)
            self.cx(oq0, oq1)
            self.h(oqs, True)

    def ryy(self, theta: Union[Tensor, float], oq0: int, oq1: int, _ideal: Optional[bool] = None):
        oqs = [oq0, oq1]
        if not self._exp_decomposed(_ideal):
            self._param('YGates', 'RYYGate', 'RYY', oqs, _ideal, theta)
        else:
            self.rx(tensor(pi / 2), oqs, True)        # float32 tensor angle, exactly as the reference passes it
            self.cx(oq0, oq1)
            self.rz(theta, oq1, True)
            self.cx(oq0, oq1)
            self.rx(-tensor(pi / 2), oqs, True)

    def rzz(self, theta: Union[Tensor, float], oq0: int, oq1: int, _ideal: Optional[bool] = None):
        oqs = [oq0, oq1]
        if not self._exp_decomposed(_ideal):
            self._param('ZGates', 'RZZGate', 'RZZ', oqs, _ideal, theta)
        else:
            self.cx(oq0, oq1)
            self.rz(theta, oq1, True)
            self.cx(oq0, oq1)

    def xx_yy(self, theta: Union[Tensor, float], beta: Union[Tensor, float],
              control: int, target: int, _ideal: Optional[bool] = None):
        from .QuantumGates.DoubleGates import XXPlusYYGate
        oqs = [control, target]
        _headline = f"XXPlusYYGate{oqs}|(P{_num(theta):3f})-(L{_num(beta):3f})".replace('.', ';')
        self._add_module(XXPlusYYGate(theta, beta, _ideal, **self._gate_kwargs()), oqs, _headline)

    def _cx_like(self, module, cls_name, label, oq0, oq1, _ideal):
        if not self._exp_decomposed(_ideal):
            self._fixed(module, cls_name, label, [oq0, oq1], _ideal)
        else:
            self.ry(-tensor(pi / 2), oq1, True)
            self.cz(oq0, oq1)
            self.ry(tensor(pi / 2), oq1, True)

    def cx(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        self._cx_like('XGates', 'CXGate', 'CX', oq0, oq1, _ideal)

    def cnot(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        self._cx_like('DoubleGates', 'CNOTGate', 'CNOT', oq0, oq1, _ideal)

    def cy(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        if not self._exp_decomposed(_ideal):
            self._fixed('YGates', 'CYGate', 'CY', [oq0, oq1], _ideal)
        else:
            raise NotImplementedError("EXPCYGate is not implemented yet.")

    def _exp_gate(self, cls_name, label, table, oq0, oq1):
        import importlib
        cls = getattr(importlib.import_module('.QuantumGates.NoiseGates', __package__), cls_name)
        _tensor = table.get(f'{oq0}{oq1}')
        if _tensor is None:
            _tensor = table.get(f'{oq1}{oq0}')
        self._add_module(cls(_tensor, **self._gate_kwargs()), [oq0, oq1], f"{label}{[oq0, oq1]}|None")

    def cz(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        if not self._exp_decomposed(_ideal):
            self._fixed('ZGates', 'CZGate', 'CZ', [oq0, oq1], _ideal)
        else:
            self._exp_gate('CZEXPGate', 'CZEXP', self._cz_expTensors, oq0, oq1)

    def cp(self, theta: Optional[Union[Tensor, float]], oq0: int, oq1: int, _ideal: Optional[bool] = None):
        if not self._exp_decomposed(_ideal):
            self._param('PhaseGates', 'CPGate', 'CP', [oq0, oq1], _ideal, theta)
        else:
            self._exp_gate('CPEXPGate', 'CPEXP', self._cp_expTensors, oq0, oq1)

    def swap(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        self._fixed('DoubleGates', 'SWAPGate', 'SWAP', [oq0, oq1], _ideal)

    def iswap(self, oq0: int, oq1: int, _ideal: Optional[bool] = None):
        self._fixed('DoubleGates', 'ISWAPGate', 'ISWAP', [oq0, oq1], _ideal)

    def pswap(self, theta: Union[float, Tensor], oq0: int, oq1: int, _ideal: Optional[bool] = None):
        self._param('DoubleGates', 'PSWAPGate', 'PSWAP', [oq0, oq1], _ideal, theta)

    def ii(self, oq1: int, oq2: int, _ideal: Optional[bool] = None):
        self._fixed('DoubleGates', 'IIGate', 'II', [oq1, oq2], _ideal)

    def arbDouble(self, data: Tensor, oq1: int, oq2: int, _ideal: Optional[bool] = None):
        from .QuantumGates.DoubleGates import ArbDoubleGate
        oqs = [oq1, oq2]
        self._add_module(ArbDoubleGate(data, _ideal, **self._gate_kwargs()), oqs, f"ArbD{oqs}|None")

    # ---- markers, resets, measurements -------------------------------------------------------------------
    def truncate(self):
        """Add a truncation layer (bond chi sweep, then inner kappa truncation)."""
        from .QuantumGates.AbstractGate import Truncate
        self._oqs_list.append([None])
        self.layers.append(Truncate())

    def barrier(self):
        from .QuantumGates.AbstractGate import Barrier
        self._oqs_list.append([None])
        self.layers.append(Barrier())

    def reset0(self, oqs: Union[List, int]):
        from .QuantumGates.SingleGates import Reset0
        oqs = _as_list(oqs)
        self._add_module(Reset0(**self._gate_kwargs()), oqs, f"Reset;0{oqs}|None")

    def reset1(self, oqs: Union[List, int]):
        from .QuantumGates.SingleGates import Reset1
        oqs = _as_list(oqs)
        self._add_module(Reset1(**self._gate_kwargs()), oqs, f"Reset;1{oqs}|None")

    def measure(self, oqs: Union[List, int], orientations: Optional[Union[List, int]] = None):
        from .QuantumGates.SingleGates import MeasureX, MeasureY, MeasureZ
        oqs = _as_list(oqs)
        orientations = orientations if orientations is not None else [2] * len(oqs)
        orientations = _as_list(orientations)
        table = {0: MeasureX, 1: MeasureY, 2: MeasureZ}
        for oq, ori in zip(oqs, orientations):
            if ori not in table:
                raise ValueError("Orientation beyond the settings.")
            self._add_module(table[ori](**self._gate_kwargs()), [oq], f"Measure;{oq}|Orientation;{ori}")
