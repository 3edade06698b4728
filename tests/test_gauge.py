"""Tree gauge (gauge_walk / tree_gauge / tree_orthogonalize, src/abstractitensornetwork.jl:387-420) and
apply(...; ortho = true) (src/apply.jl:109-111, 130-132).

The gauge is not unique (the reference's Householder R versus the engine's Hermitian factor differ by a unitary on the
bond), so -- as in the reference's own test (test/test_itensornetwork.jl:143-156: the orthogonalised network contracts to
the same tensor) -- the comparison set is gauge invariant: the contracted state, the isometry property of every tensor
that was walked over, and for the gate the singular values (on a tree with ortho = true they are the EXACT Schmidt
coefficients of gate.psi, a known answer that comes from neither implementation), truncation error and the new state.
The oracle half of every check runs on the CPU; the engine half needs a GPU."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import make_pair, rel_err

DTYPES = [np.float64, np.complex128]
TREES = {
    "chain6": lambda: O.chain_graph(6),
    "comb3x3": lambda: O.comb_tree_graph(3, 3),
    "random10": lambda: O.random_tree_graph(10, seed=3),
}


def isometry_defect(t, k):
    """Distance of Q^H Q from a projector of rank min(rows, chi), for the tensor t read as a matrix Q = (everything
    else) x (bond axis k): an isometry when the matrix is tall, a partial isometry when a leaf has fewer rows than bond
    states (d = 2 < chi = 3)."""
    m = np.moveaxis(t, k, -1).reshape(-1, t.shape[k])
    p = m.conj().T @ m
    return np.linalg.norm(p @ p - p) + abs(np.trace(p).real - min(m.shape))


def exact_schmidt(net, e, gate, dtype):
    """Singular values of gate.psi across the bond e of a tree, by brute force on the state vector."""
    g = net.graph
    u, v = g.edges[e]
    new = O.exact_apply2(net, e, gate)  # state vector after the gate, axes = sites
    # vertices on u's side of the tree once e is removed
    side, stack = {u}, [u]
    while stack:
        x = stack.pop()
        for f in g.inc[x]:
            y = g.other(f, x)
            if f != e and y not in side:
                side.add(y)
                stack.append(y)
    left = sorted(side)
    right = [w for w in range(g.nv) if w not in side]
    m = np.transpose(new, left + right).reshape(int(np.prod([new.shape[w] for w in left])), -1)
    return np.linalg.svd(m, compute_uv=False)


def same_spectrum(a, b, tol=1e-10, tail_tol=1e-10):
    """The common leading part agrees, whatever one list holds beyond the other's length is zero.  `tail_tol` applies to
    singular values below 1e-6 of the largest one: the engine takes the R factors from the bond environment (a Gram
    matrix), which resolves singular values only down to sqrt(eps) * sigma_max -- an exactly
    rank-deficient theta comes back with 1e-9 * sigma_max where the QR route of the reference has 1e-16."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    n = min(len(a), len(b))
    scale = max(a[0], b[0])
    big = np.maximum(a[:n], b[:n]) >= 1e-6 * scale
    d = np.abs(a[:n] - b[:n])
    return (np.all(d[big] < tol * scale) and np.all(d[~big] < tail_tol * scale)
            and np.all(np.abs(a[n:]) < tail_tol * scale) and np.all(np.abs(b[n:]) < tail_tol * scale))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", sorted(TREES))
def test_oracle_tree_orthogonalize(name, dtype):
    g = TREES[name]()
    net = O.random_network(g, 3, dtype=dtype, seed=11)
    root = g.nv // 2
    seq = O.tree_gauge_sequence(g, root)
    assert len(seq) == g.nv - 1 and seq[-1][1] == root
    out = O.tree_orthogonalize(net, root)
    assert rel_err(O._state_vector(out), O._state_vector(net)) < 1e-12
    for (a, b) in seq:  # every tensor the walk left behind is an isometry towards the root
        assert isometry_defect(out.tensors[a], 1 + g.slot(a, g.eid[(a, b)])) < 1e-12
    # the host mirror plans the same walk
    assert E.graphs.tree_gauge_sequence(E.NamedGraph(g.nv, g.edges), root) == seq
    # ortho = true makes the simple update exact on a tree: singular values = Schmidt coefficients of gate.psi
    gate = O.random_unitary(4, seed=5, dtype=dtype).reshape(2, 2, 2, 2)
    e = g.inc[root][0]
    new, info = O.apply2_ortho(net, e, gate)
    sv = exact_schmidt(net, e, gate, dtype)
    assert same_spectrum(info["svals"], sv)
    assert rel_err(O._state_vector(new), O.exact_apply2(net, e, gate)) < 1e-10


@pytest.fixture(scope="module")
def ctx():
    return E.Context(0)


def engine_network(bpc, net):
    return O.Network(net.graph, [np.asarray(bpc.factor(v)) for v in range(net.graph.nv)], net.dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", sorted(TREES))
def test_tree_orthogonalize_on_the_device(ctx, name, dtype):
    g = TREES[name]()
    net, psi = make_pair(g, 3, dtype, seed=11)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    root = g.nv // 2
    out = E.tree_orthogonalize(bpc, root)
    got = engine_network(out, net)
    assert rel_err(O._state_vector(got), O._state_vector(net)) < 1e-12
    for (a, b) in O.tree_gauge_sequence(g, root):
        assert isometry_defect(got.tensors[a], 1 + g.slot(a, g.eid[(a, b)])) < 1e-12
    # out of place: the input cache still holds the original tensors
    assert all(np.array_equal(bpc.factor(v), net.tensors[v]) for v in range(g.nv))
    # a region of two vertices: the edge inside the region is not walked (edge_sequence_between_regions)
    u, w = g.edges[g.inc[root][0]]
    out2 = engine_network(E.tree_gauge(bpc, [u, w]), net)
    assert rel_err(O._state_vector(out2), O._state_vector(net)) < 1e-12
    assert (u, w) not in O.tree_gauge_sequence(g, [u, w]) and (w, u) not in O.tree_gauge_sequence(g, [u, w])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_gauge_walk_ill_conditioned_and_rank_deficient(ctx, dtype):
    g = O.chain_graph(5)
    net, psi = make_pair(g, 4, dtype, seed=2)
    # graded bond: condition number 1e6 on the matrix that is QR-factorised (two passes restore Q^H Q = 1 to eps)
    net.tensors[0] = (net.tensors[0] * np.logspace(0, -6, 4)[None, :]).astype(dtype)
    # rank-deficient bond: the last two bond states of vertex 4 are empty
    net.tensors[4][:, 2:] = 0
    psi = E.ITensorNetwork(E.NamedGraph(g.nv, g.edges), [t.copy() for t in net.tensors], dtype)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    out = engine_network(E.gauge_walk(bpc, [(0, 1), (4, 3)]), net)
    assert rel_err(O._state_vector(out), O._state_vector(net)) < 1e-10
    assert isometry_defect(out.tensors[0], 1) < 1e-10
    q = np.moveaxis(out.tensors[4], 1, -1).reshape(-1, 4)
    p = q.conj().T @ q  # partial isometry: a projector of rank 2
    assert rel_err(p @ p, p) < 1e-10 and abs(np.trace(p).real - 2) < 1e-10
    with pytest.raises(E.ITNError, match="Edge not in graph"):
        E.gauge_walk(bpc, [(0, 2)])


@pytest.mark.gpu
def test_gauge_walk_on_a_loopy_graph_keeps_the_state(ctx):
    # "treating the network as a tree spanned by a spanning tree" (abstractitensornetwork.jl:407-418)
    g = O.grid_graph((3, 3))
    net, psi = make_pair(g, 2, np.complex128, seed=4)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    out = engine_network(E.tree_gauge(bpc, 4), net)
    assert rel_err(O._state_vector(out), O._state_vector(net)) < 1e-12
    ref = O.tree_orthogonalize(net, 4)
    assert rel_err(O._state_vector(ref), O._state_vector(net)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", sorted(TREES))
def test_apply_with_ortho_matches_the_oracle_and_the_exact_schmidt_values(ctx, name, dtype):
    g = TREES[name]()
    net, psi = make_pair(g, 3, dtype, seed=11)
    gate = O.random_unitary(4, seed=5, dtype=dtype).reshape(2, 2, 2, 2)
    root = g.nv // 2
    e = g.inc[root][0]
    v1, v2 = g.edges[e]
    for maxdim in (None, 2):
        ref, info = O.apply2_ortho(net, e, gate, maxdim=maxdim)
        seen = {}
        bpc = E.BeliefPropagationCache(psi, ctx=ctx)  # identity messages = the reference's default `envs = ITensor[]`
        out = E.apply(gate, bpc, (v1, v2), maxdim=maxdim, ortho=True, callback=lambda **kw: seen.update(kw))
        assert len(seen["singular_values"]) == info["newdim"]
        assert same_spectrum(seen["singular_values"], info["svals"], tail_tol=1e-7)
        assert abs(seen["truncation_error"] - info["truncerr"]) < 1e-10
        sv = exact_schmidt(net, e, gate, dtype)
        assert same_spectrum(seen["singular_values"], sv[:info["newdim"]] if maxdim else sv, tail_tol=1e-7)
        assert rel_err(O._state_vector(engine_network(out, ref)), O._state_vector(ref)) < 1e-10
    # one-site gate with ortho = true (apply.jl:108-116): the state is gate.psi whatever the gauge
    g1 = O.random_unitary(2, seed=9, dtype=dtype)
    out1 = E.apply(g1, E.BeliefPropagationCache(psi, ctx=ctx), (root,), ortho=True)
    assert rel_err(O._state_vector(engine_network(out1, net)), O._state_vector(O.apply1(net, root, g1))) < 1e-12
