"""Gate base classes of the B200 build. Same public surface as the reference's
MPDOSimulator/QuantumGates/AbstractGate.py:13-99 (Truncate / Barrier markers, QuantumGate with
name / tensor / rank / dimension / single / variational / ideal / para).

Gate tensors are operands of the device kernels; they are tiny, so they are assembled on the host with
the same arithmetic as the reference (angles cast to the complex circuit dtype first, then cos/sin/exp
in complex arithmetic) and uploaded when applied. One extension: a gate parameter may be a 1-D tensor
of B angles (a parameter sweep); `.tensor` then carries a leading batch axis [B, ...] and the circuit
is evolved as B independent circuits in one batched launch sequence."""
from abc import ABC, abstractmethod
from typing import Optional

import torch
from torch import Tensor, nn, complex64


class Truncate(nn.Module):
    def __init__(self):
        super().__init__()
        self.name = 'truncate'

    def forward(self, *args, **kwargs):
        pass


class Barrier(nn.Module):
    def __init__(self):
        super().__init__()
        self.name = 'barrier'

    def forward(self, *args, **kwargs):
        pass


def assemble(rows, dtype, device='cpu'):
    """Stack a nested list of scalars / 0-d / 1-d tensors into [..., n, n] (batch axis first if any)."""
    flat = [torch.tensor(x, dtype=dtype) if not isinstance(x, Tensor) else x.to(dtype) for r in rows for x in r]
    batch = max([x.numel() for x in flat])
    flat = [x.reshape(-1).expand(batch) if x.numel() in (1, batch) else x for x in flat]
    n = len(rows)
    out = torch.stack(flat, dim=-1).reshape(batch, n, len(rows[0]))
    out = out if batch > 1 else out[0]
    return out.to(device)


class QuantumGate(ABC, nn.Module):
    """Base class for quantum gates (reference AbstractGate.py:32-99)."""

    def __init__(self, ideal: Optional[bool] = True, dtype=complex64, device: str = 'cpu'):
        super(QuantumGate, self).__init__()
        self._para = None
        self._ideal = ideal
        self.dtype, self.device = dtype, device

    def _check_Para_Tensor(self, *parameters):
        def convert(param):
            if isinstance(param, Tensor):
                if param.requires_grad:
                    # the reference's gates carry requires_grad through torch ops (e.g. ZGates.py:65-68); the CUDA
                    # update path has no backward pass, so refuse rather than return gradient-free results
                    raise NotImplementedError(
                        'autograd through the CUDA update path is not implemented: gate parameters must not '
                        'require grad (detach() them, or differentiate by parameter shift with a batched sweep)')
                return param.detach().to(dtype=self.dtype, device='cpu')
            if isinstance(param, float):
                return torch.tensor(param, dtype=self.dtype)
            raise ValueError(f"Invalid type for gate parameter: {type(param)}")

        params = [convert(p) for p in parameters]
        return params if len(parameters) > 1 else params[0]

    def _m(self, rows):
        return assemble(rows, self.dtype, self.device)

    def _m4(self, rows):
        t = assemble(rows, self.dtype, self.device)
        return t.reshape(tuple(t.shape[:-2]) + (2, 2, 2, 2))

    @property
    @abstractmethod
    def name(self):
        pass

    @property
    @abstractmethod
    def tensor(self):
        pass

    @property
    @abstractmethod
    def rank(self):
        pass

    @property
    @abstractmethod
    def dimension(self):
        pass

    @property
    @abstractmethod
    def single(self) -> bool:
        pass

    @property
    @abstractmethod
    def variational(self) -> bool:
        pass

    @property
    def para(self):
        return self._para

    @para.setter
    def para(self, para):
        self._para = para

    @property
    def ideal(self) -> Optional[bool]:
        return self._ideal

    @ideal.setter
    def ideal(self, value: Optional[bool]):
        if not isinstance(value, bool):
            raise ValueError("Value must be bool.")
        self._ideal = value


def _gate_class(cls_name, gate_name, single, variational, matrix, n_params=0, doc=None):
    """Build a gate class with the reference constructor signature:
    fixed gates (ideal=None, dtype, device); parametric gates (p1[, p2, p3], ideal=None, dtype, device)."""

    def __init__(self, *args, **kwargs):
        params = list(args[:n_params])
        rest = list(args[n_params:])
        names = ['ideal', 'dtype', 'device']
        opts = {'ideal': None, 'dtype': complex64, 'device': 'cpu'}
        for k, v in zip(names, rest):
            opts[k] = v
        for k in list(kwargs):
            if k in opts:
                opts[k] = kwargs.pop(k)
        pnames = self._param_names
        for pn in pnames[len(params):]:
            if pn in kwargs:
                params.append(kwargs.pop(pn))
        if kwargs or len(params) != n_params:
            raise TypeError(f'{cls_name}: bad arguments')
        QuantumGate.__init__(self, ideal=opts['ideal'], dtype=opts['dtype'], device=opts['device'])
        if n_params:
            conv = self._check_Para_Tensor(*params)
            conv = conv if n_params > 1 else [conv]
            for pn, v in zip(pnames, conv):
                setattr(self, pn, v)
            self.para = conv if n_params > 1 else conv[0]

    def tensor(self):
        vals = [getattr(self, pn) for pn in self._param_names]
        rows = matrix(*vals)
        return self._m(rows) if single else self._m4(rows)

    ns = {
        '__init__': __init__,
        '__doc__': doc or f'{gate_name} gate.',
        '_param_names': [],
        'name': property(lambda self: gate_name),
        'tensor': property(tensor),
        'rank': property(lambda self: 2 if single else 4),
        'dimension': property(lambda self: [2, 2] if single else [[2, 2], [2, 2]]),
        'single': property(lambda self: single),
        'variational': property(lambda self: variational),
    }
    return type(cls_name, (QuantumGate,), ns)


def make_gate(cls_name, gate_name, single, variational, matrix, params=()):
    cls = _gate_class(cls_name, gate_name, single, variational, matrix, n_params=len(params))
    cls._param_names = list(params)
    return cls
