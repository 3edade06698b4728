"""Host container for a tensor-network state (the role of `ITensorNetwork{V}`, src/itensornetwork.jl:51-55).

Only the memory layout matters at the boundary: tensors[v] is an ndarray with axes
[site, bond to inc[v][0], bond to inc[v][1], ...] (the order `random_tensornetwork` uses, test/utils.jl:32-36).
"""
import math

import numpy as np

from .graphs import NamedGraph


class ITensorNetwork:
    def __init__(self, graph: NamedGraph, tensors, dtype=None):
        self.graph = graph
        self.dtype = np.dtype(dtype if dtype is not None else tensors[0].dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.complex64), np.dtype(np.complex128)), \
            "Float32 / Float64 / ComplexF32 / ComplexF64 only"
        self.tensors = [np.asarray(t, dtype=self.dtype) for t in tensors]
        for v, t in enumerate(self.tensors):
            assert t.ndim == 1 + graph.degree(v), f"tensor {v} must have axes [site, bonds...]"
        for e, (u, v) in enumerate(graph.edges):
            du = self.tensors[u].shape[1 + graph.inc[u].index(e)]
            dv = self.tensors[v].shape[1 + graph.inc[v].index(e)]
            assert du == dv, f"bond dimension mismatch on edge {e}"

    def copy(self):
        return ITensorNetwork(self.graph, [t.copy() for t in self.tensors], self.dtype)

    def edge_dim(self, e):
        u, _ = self.graph.edges[e]
        return self.tensors[u].shape[1 + self.graph.inc[u].index(e)]

    def siteinds(self, v):
        return self.tensors[v].shape[0]


def random_tensornetwork(seed, dtype, graph, link_space=1, d=2):
    """iid N(0,1) / CN(0,1) entries (test/utils.jl:23-38; Julia randn(ComplexF64) has unit variance)."""
    rng = np.random.default_rng(seed)
    chis = link_space if isinstance(link_space, (list, tuple, np.ndarray)) else [link_space] * graph.ne
    ts = []
    for v in range(graph.nv):
        shape = (d,) + tuple(int(chis[e]) for e in graph.inc[v])
        if np.dtype(dtype).kind == "c":
            t = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / math.sqrt(2.0)
        else:
            t = rng.standard_normal(shape)
        ts.append(t.astype(dtype))
    return ITensorNetwork(graph, ts, dtype)


def productstate(graph, states, dtype=np.complex128, d=2):
    """Bond-dimension-1 product state; states[v] is a length-d vector or a basis index."""
    ts = []
    for v in range(graph.nv):
        s = states[v] if not callable(states) else states(v)
        vec = np.zeros(d, dtype=dtype)
        if np.isscalar(s):
            vec[int(s)] = 1.0
        else:
            vec[:] = s
        ts.append(vec.reshape((d,) + (1,) * graph.degree(v)))
    return ITensorNetwork(graph, ts, dtype)
