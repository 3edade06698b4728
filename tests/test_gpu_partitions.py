"""GPU parity for generalised partitions (several sites per BP partition).

Restates test/test_expect.jl:22-39 (a grid grouped by column: the quotient graph is a chain, BP is exact) and the
non-product-environment error of test/test_apply.jl:38-64 on the engine, and checks the super-site network and its BP
messages against the oracle's own construction on a partition whose quotient graph still has a loop."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import assert_messages_close, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    return E.Context(0)


def host_net(net):
    g = E.NamedGraph(net.graph.nv, net.graph.edges)
    return E.ITensorNetwork(g, [t.copy() for t in net.tensors], net.dtype)


def columns(g):
    cols = {}
    for v, c in enumerate(g.coords):
        cols.setdefault(c[0], []).append(v)
    return [cols[k] for k in sorted(cols)]


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_device_tensordot_matches_numpy(ctx, dtype):
    rng = np.random.default_rng(3)

    def rnd(shape):
        t = rng.standard_normal(shape)
        return (t + 1j * rng.standard_normal(shape)).astype(dtype) if np.dtype(dtype).kind == "c" else t.astype(dtype)

    cases = [((2, 3, 4), (4, 5), [2], [0]), ((2, 3, 4), (3, 2, 6), [1, 0], [0, 1]), ((3, 2), (4,), [], []),
             ((5,), (5,), [0], [0]), ((2, 3, 2, 3), (3, 2, 2), [1, 2], [0, 1])]
    for sa, sb, xa, xb in cases:
        a, b = rnd(sa), rnd(sb)
        got = E.tensordot(a, b, xa, xb, ctx)
        ref = np.tensordot(a, b, axes=(xa, xb))
        assert got.shape == ref.shape and rel_err(got, ref) < 1e-13


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("dims", [(2, 2), (3, 3)])
def test_expect_with_column_partition_is_exact(ctx, dtype, dims):
    g = O.grid_graph(dims)
    net = O.random_network(g, 2, dtype=dtype, seed=1234)
    groups = columns(g)
    sz = E.expect(host_net(net), "Sz", alg="bp", ctx=ctx, cache_construction_kwargs={"partitioned_vertices": groups},
                  cache_update_kwargs={"maxiter": 20})
    for v in range(g.nv):
        assert abs(sz[v] - O.exact_expect1(net, v, 0.5 * O.PAULI_Z)) < TOL
    # one site per partition on the same loopy graph is NOT exact (the partition is what makes the difference)
    sz1 = E.expect(host_net(net), "Sz", alg="bp", ctx=ctx, cache_update_kwargs={"maxiter": 20})
    assert max(abs(sz1[v] - sz[v]) for v in range(g.nv)) > 1e-6


def test_block_partition_matches_oracle_on_a_loopy_quotient_graph(ctx):
    g = O.grid_graph((4, 4))
    net = O.random_network(g, 2, dtype=np.complex128, seed=7)
    blocks = {}
    for v, c in enumerate(g.coords):
        blocks.setdefault((c[0] // 2, c[1] // 2), []).append(v)
    groups = [blocks[k] for k in sorted(blocks)]
    coarse, group_of = O.partition_network(net, groups)
    assert not coarse.graph.is_tree()
    bpc = E.BeliefPropagationCache(host_net(net), ctx=ctx, partitioned_vertices=groups)
    assert bpc.graph.edges == coarse.graph.edges
    for gi in range(len(groups)):
        assert rel_err(bpc.factor(gi), coarse.tensors[gi]) < 1e-13  # device contraction of the partition = oracle einsum
    seq = O.parallel_edge_sequence(coarse.graph)
    ref, _, _ = O.bp_update(coarse, O.identity_messages(coarse), seq=seq, groups=O.synchronous_groups(seq), maxiter=5)
    E.update(bpc, maxiter=5, edge_sequence=[[e] for e in seq], inplace=True)
    assert_messages_close(bpc, ref, TOL)
    ez = E.expect(bpc, "Z")
    for v in range(g.nv):
        gi = group_of[v]
        lifted = O.lift_operator([2] * len(groups[gi]), groups[gi].index(v), O.PAULI_Z)
        assert abs(ez[v] - O.expect1(coarse, ref, gi, lifted)) < TOL
    # environment(bpc, verts): the messages into the partitions containing verts, over fused bond indices
    env = E.environment(bpc, [0])
    assert all(m.shape == (4, 4) for _, m in env) and len(env) == 2
    # apply needs a product environment (src/apply.jl:119-125, test/test_apply.jl:64)
    with pytest.raises(E.ITNError):
        E.apply(np.eye(4).reshape(2, 2, 2, 2), bpc, g.edges[0], maxdim=2)
