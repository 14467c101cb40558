"""ORACLE INFRASTRUCTURE - a minimal restatement of the tensornetwork==0.4.6 API surface that
WeiguoMa/Tomography-assisted-MPDO-QCircuit calls on its noisy-gate update path (requirements.txt:4 pins the
package; it is not vendored and cannot be installed offline). It exists so that the UNMODIFIED reference modules
under /root/reference can be executed in the build container to produce golden vectors
(tests/golden/make_golden.py). It is not part of the product and is never imported by it.

Restated semantics (tensornetwork 0.4.6, network_components.py / network_operations.py / contractors):
  Node(tensor, name, axis_names)   one Edge per axis, edges named after the axis names
  e1 ^ e2, connect                  replace two dangling edges by one shared edge
  contract(edge)                    tensordot over that edge; result axes = node1's remaining, then node2's;
                                    the new node keeps the Edge objects (and their names), axis_names reset to '0','1',..
  contract_between(n1, n2)          tensordot over all shared edges, same axis rule
  contractors.optimal / auto        contract the given nodes in place (pairwise contract_between), final transpose to
                                    output_edge_order unless ignore_edge_order
  flatten_edges([e...])             move the edges' axes last (given order) and merge them row-major
  split_node                        transpose to (left, right), backend.svd -> u*sqrt(s), sqrt(s)*vh, discarded s
  split_node_qr / split_node_full_svd   same with backend.qr / (u, diag(s), vh)
  replicate_nodes, check_connected, NodeCollection
The pytorch backend's svd / qr are the reference's own patched decompositions.py (README.md:17), imported
from the reference tree.
"""
import itertools
from typing import List

import numpy as np
import torch

try:
    import decompositions as _dec      # /root/reference/decompositions.py (the file the README says to install)
except ImportError:                    # pragma: no cover
    _dec = None

__version__ = '0.4.6-shim'
randomized_svd_calls = 0   # how often the reference's randomized branch (decompositions.py:112-115) was taken


def _svd(tensor, pivot, max_singular_values, max_truncation_err, relative):
    global randomized_svd_calls
    if tensor.numel() >= 10000 and max_singular_values is not None:
        randomized_svd_calls += 1
    return _dec.svd(torch, tensor, pivot, max_singular_values, max_truncation_err, relative)

_UNNAMED_EDGE = '__unnamed_edge__'
_UNNAMED_NODE = '__unnamed_node__'
_collection_stack = []


def set_default_backend(name):
    if name != 'pytorch':
        raise ValueError('the shim only restates the pytorch backend')


class Edge:
    def __init__(self, node1, axis1, name=None, node2=None, axis2=None):
        self.node1, self.axis1, self.node2, self.axis2 = node1, axis1, node2, axis2
        self.name = name if name else _UNNAMED_EDGE

    def set_name(self, name):
        if not isinstance(name, str):
            raise TypeError('Edge name should be str type')
        self.name = name

    def is_dangling(self):
        return self.node2 is None

    def is_trace(self):
        return self.node1 is self.node2

    def get_nodes(self):
        return [self.node1, self.node2]

    @property
    def dimension(self):
        return int(self.node1.tensor.shape[self.axis1])

    def update_axis(self, old_axis, old_node, new_axis, new_node):
        if self.node1 is old_node and self.axis1 == old_axis:
            self.node1, self.axis1 = new_node, new_axis
        elif self.node2 is old_node and self.axis2 == old_axis:
            self.node2, self.axis2 = new_node, new_axis
        else:
            raise ValueError('edge is not attached at that axis')

    def disconnect(self, edge1_name=None, edge2_name=None):
        if self.is_dangling():
            raise ValueError(f'Cannot break dangling edge {self.name}.')
        e1 = Edge(self.node1, self.axis1, name=edge1_name)
        e2 = Edge(self.node2, self.axis2, name=edge2_name)
        self.node1.edges[self.axis1] = e1
        self.node2.edges[self.axis2] = e2
        return e1, e2

    def __xor__(self, other):
        return connect(self, other)

    def __repr__(self):
        return f'Edge({self.name})'


class AbstractNode:
    pass


class Node(AbstractNode):
    def __init__(self, tensor, name=None, axis_names=None, backend=None):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor)
        self.tensor = tensor
        self.name = name if name else _UNNAMED_NODE
        nd = tensor.dim()
        if axis_names is not None:
            if len(axis_names) != nd:
                raise ValueError(f'axis_names is not the same length as the tensor shape. {len(axis_names)} vs {nd}')
            if len(set(axis_names)) != len(axis_names):
                raise ValueError('Not all axis names are unique')
            self._axis_names = list(axis_names)
            self.edges = [Edge(self, i, name=n) for i, n in enumerate(axis_names)]
        else:
            self._axis_names = [str(i) for i in range(nd)]
            self.edges = [Edge(self, i) for i in range(nd)]
        if _collection_stack:
            c = _collection_stack[-1]
            c.append(self) if isinstance(c, list) else c.add(self)

    # axis names -------------------------------------------------------------------------------
    @property
    def axis_names(self):
        return self._axis_names

    @axis_names.setter
    def axis_names(self, names):
        if len(names) != len(self.edges):
            raise ValueError('Expected {} names, only got {}.'.format(len(self.edges), len(names)))
        self._axis_names = list(names)

    def get_axis_number(self, axis):
        if isinstance(axis, int):
            return axis
        try:
            return self._axis_names.index(axis)
        except ValueError:
            raise ValueError(f"Axis name '{axis}' not found for node '{self.name}'")

    def __getitem__(self, key):
        return self.edges[self.get_axis_number(key)]

    def get_edge(self, key):
        return self[key]

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def set_tensor(self, tensor):
        self.tensor = tensor

    def set_name(self, name):
        self.name = name

    def get_all_dangling(self):
        return [e for e in self.edges if e.is_dangling()]

    def reorder_edges(self, edge_order):
        if set(map(id, edge_order)) != set(map(id, self.edges)):
            raise ValueError('Given edges do not match the node edges')
        perm = [next(i for i, e in enumerate(self.edges) if e is want) for want in edge_order]
        self.reorder_axes(perm)
        return self

    def reorder_axes(self, perm):
        self.tensor = self.tensor.permute(perm) if self.tensor.dim() else self.tensor
        old_edges, old_names = self.edges, self._axis_names
        self.edges = [old_edges[p] for p in perm]
        self._axis_names = [old_names[p] for p in perm]
        for new_axis, (e, p) in enumerate(zip(self.edges, perm)):
            if e.is_trace():
                raise NotImplementedError('trace edges are not needed by the reference path')
            e.update_axis(p, self, new_axis, self) if not (e.node1 is self and e.axis1 == new_axis and p == new_axis) else None
        return self

    def fresh_edges(self, axis_names=None):
        axis_names = axis_names or [str(i) for i in range(len(self.edges))]
        self.edges = [Edge(self, i, name=n) for i, n in enumerate(axis_names)]
        self._axis_names = list(axis_names)

    def __repr__(self):
        return f'Node({self.name}, {self._axis_names}, {tuple(self.tensor.shape)})'


class NodeCollection:
    def __init__(self, container):
        self.container = container

    def __enter__(self):
        _collection_stack.append(self.container)

    def __exit__(self, *exc):
        _collection_stack.pop()


# -------------------------------------------------------------------------------------------------------
def connect(edge1, edge2, name=None):
    for e in (edge1, edge2):
        if not e.is_dangling():
            raise ValueError(f'Edge {e.name} is not a dangling edge.')
    if edge1 is edge2:
        raise ValueError('Cannot connect an edge to itself.')
    if edge1.dimension != edge2.dimension:
        raise ValueError(f'Cannot connect edges of unequal dimension. {edge1.dimension} vs {edge2.dimension}')
    n1, a1, n2, a2 = edge1.node1, edge1.axis1, edge2.node1, edge2.axis1
    new = Edge(n1, a1, name=name, node2=n2, axis2=a2)
    n1.edges[a1] = new
    n2.edges[a2] = new
    return new


def get_shared_edges(node1, node2):
    return [e for e in node1.edges if not e.is_dangling() and
            ((e.node1 is node1 and e.node2 is node2) or (e.node1 is node2 and e.node2 is node1))]


def _merge(node1, node2, shared, name):
    """tensordot over `shared`; remaining axes: node1's in order, then node2's. Edge objects move to the new node."""
    ax1 = [e.axis1 if e.node1 is node1 else e.axis2 for e in shared]
    ax2 = [e.axis2 if e.node1 is node1 else e.axis1 for e in shared]
    if node1 is node2:
        raise NotImplementedError('trace contraction is not needed by the reference path')
    t = torch.tensordot(node1.tensor, node2.tensor, dims=(ax1, ax2)) if shared else \
        torch.tensordot(node1.tensor, node2.tensor, dims=0)
    new = Node(t, name=name)
    pos = 0
    new_edges = []
    for nd, skip in ((node1, set(ax1)), (node2, set(ax2))):
        for ax, e in enumerate(nd.edges):
            if ax in skip:
                continue
            e.update_axis(ax, nd, pos, new)
            new_edges.append(e)
            pos += 1
    new.edges = new_edges
    node1.fresh_edges(node1.axis_names)
    node2.fresh_edges(node2.axis_names)
    return new


def contract(edge, name=None, axis_names=None):
    if edge.is_dangling():
        raise ValueError('Attempting to contract dangling edge')
    return _merge(edge.node1, edge.node2, [edge], name)


def contract_between(node1, node2, name=None, allow_outer_product=False, output_edge_order=None, axis_names=None):
    shared = get_shared_edges(node1, node2)
    if not shared and not allow_outer_product:
        raise ValueError(f'No edges found between nodes {node1.name} and {node2.name} and allow_outer_product=False.')
    new = _merge(node1, node2, shared, name)
    if output_edge_order is not None:
        new.reorder_edges(list(output_edge_order))
    return new


def get_all_edges(nodes):
    seen, out = set(), []
    for n in nodes:
        for e in n.edges:
            if id(e) not in seen:
                seen.add(id(e))
                out.append(e)
    return out


def get_subgraph_dangling(nodes):
    ids = set(map(id, nodes))
    return [e for e in get_all_edges(nodes) if e.is_dangling() or id(e.node1) not in ids or id(e.node2) not in ids]


def _contract_all(nodes, output_edge_order=None, ignore_edge_order=False):
    nodes = list(nodes)
    if not ignore_edge_order:
        dangling = get_subgraph_dangling(nodes)
        if output_edge_order is None:
            output_edge_order = dangling
            if len(output_edge_order) > 1:
                raise ValueError('The final node after contraction has more than one remaining edge. '
                                 'In this case `output_edge_order` has to be provided.')
        if set(map(id, output_edge_order)) != set(map(id, dangling)):
            raise ValueError('output edges are not equal to the remaining non-contracted edges of the final node.')
    while len(nodes) > 1:
        best = None
        for i, j in itertools.combinations(range(len(nodes)), 2):
            shared = get_shared_edges(nodes[i], nodes[j])
            if not shared:
                continue
            dims = np.prod([e.dimension for e in shared], dtype=np.float64)
            size = nodes[i].tensor.numel() * nodes[j].tensor.numel() / (dims * dims)
            if best is None or size < best[0]:
                best = (size, i, j)
        if best is None:                      # disconnected pieces: outer product
            i, j = 0, 1
        else:
            _, i, j = best
        new = contract_between(nodes[i], nodes[j], allow_outer_product=True)
        nodes = [n for k, n in enumerate(nodes) if k not in (i, j)] + [new]
    final = nodes[0]
    if not ignore_edge_order and len(final.edges) > 0:
        final.reorder_edges(list(output_edge_order))
    return final


class _Contractors:
    @staticmethod
    def optimal(nodes, output_edge_order=None, memory_limit=None, ignore_edge_order=False):
        return _contract_all(nodes, output_edge_order, ignore_edge_order)

    auto = greedy = branch = optimal


contractors = _Contractors()


def flatten_edges(edges, new_edge_name=None):
    if not edges:
        raise ValueError('At least 1 edge must be given.')
    if not all(e.is_dangling() for e in edges) or len({id(e.node1) for e in edges}) != 1:
        raise NotImplementedError('the reference only flattens dangling edges of one node')
    node = edges[0].node1
    back = [e.axis1 for e in edges]
    front = [i for i in range(len(node.edges)) if i not in back]
    node.reorder_axes(front + back)
    shape = list(node.tensor.shape)
    node.tensor = node.tensor.reshape(shape[:len(front)] + [int(np.prod(shape[len(front):]))])
    new_edge = Edge(node, len(front), name=new_edge_name)
    node.edges = node.edges[:len(front)] + [new_edge]
    node._axis_names = [str(i) for i in range(len(node.edges))]
    return new_edge


def _split_prepare(node, left_edges, right_edges):
    node.reorder_edges(list(left_edges) + list(right_edges))
    lnames = [node.axis_names[i] for i in range(len(left_edges))]
    rnames = [node.axis_names[len(left_edges) + i] for i in range(len(right_edges))]
    return lnames, rnames


def _attach(new_node, edges, offset, old_node):
    for i, e in enumerate(edges):
        old_axis = e.axis1 if e.node1 is old_node else e.axis2
        e.update_axis(old_axis, old_node, i + offset, new_node)
        new_node.edges[i + offset] = e


def split_node(node, left_edges, right_edges, max_singular_values=None, max_truncation_err=None, relative=False,
               left_name=None, right_name=None, edge_name=None):
    lnames, rnames = _split_prepare(node, left_edges, right_edges)
    u, s, vh, trun = _svd(node.tensor, len(left_edges), max_singular_values, max_truncation_err, relative)
    sq = torch.sqrt(s)
    u_s = u * sq
    vh_s = sq.reshape([-1] + [1] * (vh.dim() - 1)) * vh
    cname = edge_name if edge_name else _UNNAMED_EDGE
    left = Node(u_s, name=left_name, axis_names=lnames + [cname])
    right = Node(vh_s, name=right_name, axis_names=[cname] + rnames)
    _attach(left, list(left_edges), 0, node)
    _attach(right, list(right_edges), 1, node)
    connect(left.edges[-1], right.edges[0], name=edge_name)
    node.fresh_edges(node.axis_names)
    return left, right, trun


def split_node_qr(node, left_edges, right_edges, left_name=None, right_name=None, edge_name=None):
    lnames, rnames = _split_prepare(node, left_edges, right_edges)
    q, r = _dec.qr(torch, node.tensor, len(left_edges))
    cname = edge_name if edge_name else _UNNAMED_EDGE
    left = Node(q, name=left_name, axis_names=lnames + [cname])
    right = Node(r, name=right_name, axis_names=[cname] + rnames)
    _attach(left, list(left_edges), 0, node)
    _attach(right, list(right_edges), 1, node)
    connect(left.edges[-1], right.edges[0], name=edge_name)
    node.fresh_edges(node.axis_names)
    return left, right


def split_node_full_svd(node, left_edges, right_edges, max_singular_values=None, max_truncation_err=None,
                        relative=False, left_name=None, middle_name=None, right_name=None, left_edge_name=None,
                        right_edge_name=None):
    lnames, rnames = _split_prepare(node, left_edges, right_edges)
    u, s, vh, trun = _svd(node.tensor, len(left_edges), max_singular_values, max_truncation_err, relative)
    ln = left_edge_name if left_edge_name else '__left_svd_edge__'
    rn = right_edge_name if right_edge_name else '__right_svd_edge__'
    left = Node(u, name=left_name, axis_names=lnames + [ln])
    mid = Node(torch.diag(s), name=middle_name, axis_names=[ln, rn])
    right = Node(vh, name=right_name, axis_names=[rn] + rnames)
    _attach(left, list(left_edges), 0, node)
    _attach(right, list(right_edges), 1, node)
    connect(left.edges[-1], mid.edges[0], name=left_edge_name)
    connect(mid.edges[1], right.edges[0], name=right_edge_name)
    node.fresh_edges(node.axis_names)
    return left, mid, right, trun


def replicate_nodes(nodes, conjugate=False):
    nodes = list(nodes)
    copies = {}
    for n in nodes:
        t = n.tensor.conj() if conjugate else n.tensor
        copies[id(n)] = Node(t, name=n.name, axis_names=list(n.axis_names))
    for e in get_all_edges(nodes):
        if e.is_dangling():
            copies[id(e.node1)].edges[e.axis1].name = e.name if e.name else _UNNAMED_EDGE
            continue
        in1, in2 = id(e.node1) in copies, id(e.node2) in copies
        if in1 and in2:
            connect(copies[id(e.node1)].edges[e.axis1], copies[id(e.node2)].edges[e.axis2], name=e.name)
    return [copies[id(n)] for n in nodes]


def check_connected(nodes):
    nodes = list(nodes)
    if not nodes:
        return
    ids = {id(n): n for n in nodes}
    seen, stack = {id(nodes[0])}, [nodes[0]]
    while stack:
        n = stack.pop()
        for e in n.edges:
            if e.is_dangling():
                continue
            other = e.node2 if e.node1 is n else e.node1
            if id(other) in ids and id(other) not in seen:
                seen.add(id(other))
                stack.append(other)
    if len(seen) != len(ids):
        raise ValueError('Non-connected graph')
