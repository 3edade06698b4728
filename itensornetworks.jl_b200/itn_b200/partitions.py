"""Generalised partitions: several sites per BP partition.

The reference partitions the three-layer <psi|psi> network with `partitioned_vertices` (src/caches/beliefpropagationcache.jl:20-35);
the default is one site per partition, the tests also group whole columns so that the quotient graph becomes a chain and BP
exact (test/test_expect.jl:22-39, test/test_apply.jl:38-43).  A partition of k sites is a SUPER-SITE: contracting its ket tensors
over the edges inside the partition gives one tensor whose site index is the fused site indices of its members and whose bond
to a neighbouring partition is the fused bundle of edges between the two.  BP over partitions is then exactly BP over the
super-site network, which the engine runs unchanged.  The contractions inside a partition run on the device (`itn_tensordot`);
this module only does index bookkeeping and axis permutations.
"""
import ctypes as C

import numpy as np

from ._lib import check, i32, lib
from .graphs import NamedGraph
from .network import ITensorNetwork


def tensordot(a, b, axes_a, axes_b, ctx):
    """numpy.tensordot(a, b, (axes_a, axes_b)) computed by the engine (k_tensordot)."""
    dtype = np.result_type(a.dtype, b.dtype)
    dtype = np.dtype(np.complex128 if dtype.kind == "c" else np.float64)
    fa = np.asfortranarray(a.astype(dtype, copy=False))
    fb = np.asfortranarray(b.astype(dtype, copy=False))
    free_a = [i for i in range(a.ndim) if i not in axes_a]
    free_b = [i for i in range(b.ndim) if i not in axes_b]
    shape = tuple(a.shape[i] for i in free_a) + tuple(b.shape[i] for i in free_b)
    out = np.empty(shape, dtype=dtype, order="F")
    _, pa = i32(list(a.shape) or [1])
    _, pb = i32(list(b.shape) or [1])
    _, xa = i32(list(axes_a) or [0])
    _, xb = i32(list(axes_b) or [0])
    check(lib().itn_tensordot(ctx.h, 1 if dtype.kind == "c" else 0, fa.ctypes.data_as(C.c_void_p), a.ndim, pa,
                              fb.ctypes.data_as(C.c_void_p), b.ndim, pb, len(axes_a), xa, xb, out.ctypes.data_as(C.c_void_p)))
    return out


class PartitionMap:
    """Where the sites of the original network live in the super-site network."""

    def __init__(self, groups, group_of, pos, site_dims):
        self.groups, self.group_of, self.pos, self.site_dims = groups, group_of, pos, site_dims

    def lift_operator(self, v, o):
        """d x d operator on site v -> operator on the fused site index of its partition (first member fastest)."""
        g, p = self.group_of[v], self.pos[v]
        out = np.ones((1, 1), dtype=np.asarray(o).dtype)
        for q, d in enumerate(self.site_dims[g]):
            out = np.kron(np.asarray(o) if q == p else np.eye(d, dtype=out.dtype), out)
        return out


def partition_plan(graph, groups):
    """Quotient graph of `graph` under `groups` (list of vertex lists): edges between partitions, and for each the
    bundle of original edges it fuses (ascending edge id on both sides)."""
    group_of = {}
    for gi, grp in enumerate(groups):
        for v in grp:
            assert v not in group_of, f"vertex {v} appears in two partitions"
            group_of[int(v)] = gi
    assert len(group_of) == graph.nv, "partitioned_vertices must cover every vertex"
    bundles = {}
    for e, (u, v) in enumerate(graph.edges):
        a, b = group_of[u], group_of[v]
        if a != b:
            bundles.setdefault((min(a, b), max(a, b)), []).append(e)
    qedges = sorted(bundles)
    return group_of, qedges, bundles


def partitioned_network(psi: ITensorNetwork, groups, ctx):
    """Super-site network of `psi` for the partition `groups`; returns (ITensorNetwork, PartitionMap)."""
    g = psi.graph
    groups = [[int(v) for v in grp] for grp in groups]
    group_of, qedges, bundles = partition_plan(g, groups)
    qg = NamedGraph(len(groups), qedges)
    tensors, site_dims = [], []
    for gi, grp in enumerate(groups):
        # running tensor with labelled axes: ("s", v) site indices, ("e", e) open edges
        t, labels = None, []
        for v in grp:
            tv = psi.tensors[v]
            lv = [("s", v)] + [("e", e) for e in g.inc[v]]
            if t is None:
                t, labels = tv, lv
                continue
            shared = [lab for lab in lv if lab[0] == "e" and lab in labels]
            ax_t = [labels.index(lab) for lab in shared]
            ax_v = [lv.index(lab) for lab in shared]
            t = tensordot(t, tv, ax_t, ax_v, ctx)
            labels = [lab for lab in labels if lab not in shared] + [lab for lab in lv if lab not in shared]
        # target order: site indices of the members (first member fastest), then one fused axis per quotient edge
        order = [("s", v) for v in grp]
        fused_shape = [int(np.prod([psi.tensors[v].shape[0] for v in grp]))]
        for qe in qg.inc[gi]:
            es = bundles[qedges[qe]]
            order += [("e", e) for e in es]
            fused_shape.append(int(np.prod([psi.edge_dim(e) for e in es])))
        assert sorted(order) == sorted(labels), "internal edges must be contracted, external ones kept"
        t = np.transpose(t, [labels.index(lab) for lab in order])
        tensors.append(np.ascontiguousarray(t.reshape(fused_shape, order="F")))
        site_dims.append([psi.tensors[v].shape[0] for v in grp])
    pos = {v: p for grp in groups for p, v in enumerate(grp)}
    return ITensorNetwork(qg, tensors, tensors[0].dtype), PartitionMap(groups, group_of, pos, site_dims)
