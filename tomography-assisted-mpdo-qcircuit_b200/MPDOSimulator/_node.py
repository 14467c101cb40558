"""Dense site node: the B200 build's stand-in for the reference's `tn.Node` state elements.

The reference state is a list of tensornetwork nodes whose axes are addressed by name
(`physics_k`, `bond_i_j`, `I_k`; README.md:40-48). Here every site owns one dense tensor
`data[B, l, s, a, r]` (batch of independent circuits, left bond, physical, inner/Kraus, right bond) with
size-1 axes where the reference node has no such axis; flags record which axes exist so that
`axis_names`, `tensor` and `node[axis_name]` answer as the reference nodes do."""
from typing import List

import torch


class Axis:
    """What `node[axis_name]` returns: just enough of tn.Edge for user code (name, dimension)."""

    def __init__(self, node, name, dim_index):
        self.node1, self.name, self._dim = node, name, dim_index

    @property
    def dimension(self):
        return int(self.node1.data.shape[self._dim])

    def is_dangling(self):
        return 'bond' not in self.name

    def __repr__(self):
        return f'Axis({self.name}, dim={self.dimension})'


class DenseNode:
    def __init__(self, data: torch.Tensor, index: int, name: str = None, has_left=False, has_right=False,
                 has_inner=False):
        if data.dim() == 1:
            data = data.reshape(1, 1, data.shape[0], 1, 1)
        assert data.dim() == 5
        self.data = data
        self.index = index
        self.name = name if name is not None else f'qubit_{index}'
        self.has_left, self.has_right, self.has_inner = has_left, has_right, has_inner
        # Inner dimension the REFERENCE node would have (None: same as data.shape[3]). Gate operands are applied with
        # the minimal number of Kraus operators (Circuit._compress_kraus), but the reference decides whether a site
        # takes part in the inner-index truncation by dim(I_k) > kappa on its own, redundant, dimension
        # (TNNOptimizer.py:176-181), so that number is carried along.
        self.ref_inner = None

    # --- reference-node look-alikes -------------------------------------------------------------
    def _axes(self):
        k = self.index
        out = []
        if self.has_left:
            out.append((f'bond_{k - 1}_{k}', 1))
        out.append((f'physics_{k}', 2))
        if self.has_inner:
            out.append((f'I_{k}', 3))
        if self.has_right:
            out.append((f'bond_{k}_{k + 1}', 4))
        return out

    @property
    def axis_names(self) -> List[str]:
        return [n for n, _ in self._axes()]

    @property
    def edges(self):
        return [Axis(self, n, d) for n, d in self._axes()]

    def __getitem__(self, key):
        for n, d in self._axes():
            if n == key:
                return Axis(self, n, d)
        raise ValueError(f'Axis name {key} not found for node {self.name}')

    @property
    def tensor(self) -> torch.Tensor:
        """The site tensor with exactly the reference's axes (batch axis first when B > 1)."""
        dims = [d for _, d in self._axes()]
        t = self.data
        shape = ([t.shape[0]] if t.shape[0] > 1 else []) + [t.shape[d] for d in dims]
        return t.reshape(shape)

    def set_tensor(self, tensor: torch.Tensor):
        dims = [d for _, d in self._axes()]
        batched = tensor.dim() == len(dims) + 1
        full = [tensor.shape[0] if batched else 1, 1, 1, 1, 1]
        for pos, d in enumerate(dims):
            full[d] = tensor.shape[pos + (1 if batched else 0)]
        self.data = tensor.reshape(full)

    def set_name(self, name):
        self.name = name

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def nominal_inner(self) -> int:
        return int(self.data.shape[3]) if self.ref_inner is None else int(self.ref_inner)

    def copy(self):
        node = DenseNode(self.data.clone(), self.index, self.name, self.has_left, self.has_right, self.has_inner)
        node.ref_inner = self.ref_inner
        return node

    def __repr__(self):
        return f'DenseNode({self.name}, axes={self.axis_names}, data={tuple(self.data.shape)}, {self.data.dtype})'


def replicate_nodes(nodes):
    """Deep copy of a state (the reference uses tn.replicate_nodes to keep a state before reuse)."""
    return [n.copy() for n in nodes]
