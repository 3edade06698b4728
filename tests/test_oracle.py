"""Pin the oracle to the properties the reference's own tests assert for the hot path.

Each test names the reference test it restates (there are no golden vectors in the reference).
"""
import numpy as np
import pytest

from oracle import itn_oracle as O

EPS = np.finfo(np.float64).eps
DTYPES = [np.float64, np.complex128]


@pytest.mark.parametrize("dtype", DTYPES)
def test_bp_fixed_point_3x3(dtype):
    # test/test_belief_propagation.jl:18-55 (3x3 grid, chi=2, maxiter=25, tol=eps)
    g = O.grid_graph((3, 3))
    net = O.random_network(g, 2, dtype=dtype)
    msgs = O.identity_messages(net)
    msgs, iters, _ = O.bp_update(net, msgs, maxiter=40, tol=EPS)
    for (u, v) in g.edges:
        for (a, b) in ((u, v), (v, u)):
            new = O.updated_message(net, msgs, a, b)
            assert O.message_diff(new, msgs[(a, b)]) < 10 * EPS
            assert msgs[(a, b)].dtype == dtype


@pytest.mark.parametrize("dtype", DTYPES)
def test_rdm2_psd(dtype):
    # test/test_belief_propagation.jl:64-91
    g = O.grid_graph((3, 3))
    net = O.random_network(g, 2, dtype=dtype)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), maxiter=40, tol=EPS)
    e = g.eid[(4, 5)]
    rho = O.rdm2(net, msgs, e)
    assert rho.shape == (4, 4)
    w = np.linalg.eigvals(rho)
    assert np.all(np.abs(w.imag) <= 1e-7)
    assert np.all(w.real >= -1e-7)
    assert abs(np.trace(rho) - 1) < 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_network_scalar(dtype):
    # test/test_belief_propagation.jl:93-99
    g = O.grid_graph((3, 1))
    net = O.random_network(g, 2, dtype=dtype)
    net.tensors[0] = 0 * net.tensors[0]
    seq = O.default_edge_sequence(g)
    msgs = {}
    # tree: no initial messages needed; forest-cover order guarantees availability
    for (v, w) in seq:
        msgs[(v, w)] = O.updated_message(net, msgs, v, w)
    assert O.scalar(net, msgs) == 0


@pytest.mark.parametrize("dtype", DTYPES)
def test_bp_exact_on_tree(dtype):
    # test/test_expect.jl:12-20, test/test_inner.jl:13-48, test/test_forms.jl:59-70
    g = O.random_tree_graph(7, seed=3)
    net = O.random_network(g, 2, dtype=dtype)
    seq = O.default_edge_sequence(g)
    assert sorted(seq) == sorted(list(g.edges) + [(v, u) for u, v in g.edges])
    msgs = {}
    for (v, w) in seq:  # one sweep in forest-cover order is exact on a tree
        msgs[(v, w)] = O.updated_message(net, msgs, v, w)
    z = O.scalar(net, msgs)
    assert np.allclose(z, O.exact_norm_sqr(net), rtol=1e-12)
    for v in range(g.nv):
        assert np.allclose(O.expect1(net, msgs, v, O.PAULI_Z), O.exact_expect1(net, v, O.PAULI_Z), atol=1e-12)
    for e in range(g.ne):
        assert np.allclose(O.rdm2(net, msgs, e), O.exact_rdm2(net, e), atol=1e-12)


def test_edge_sequence_covers_every_directed_edge_once():
    # src/edge_sequences.jl:32-47
    for g in (O.grid_graph((4, 4)), O.grid_graph((3, 3, 3)), O.heavy_hex_eagle_graph()):
        seq = O.default_edge_sequence(g)
        assert len(seq) == 2 * g.ne
        assert len(set(seq)) == 2 * g.ne


def test_heavy_hex_census():
    g = O.heavy_hex_eagle_graph()
    degs = np.bincount([g.degree(v) for v in range(g.nv)])
    assert g.nv == 127 and g.ne == 144
    assert degs[1] == 2 and degs[2] == 89 and degs[3] == 36
    cols = O.edge_coloring(g)
    assert sum(len(c) for c in cols) == 144
    for c in cols:
        vs = [x for e in c for x in g.edges[e]]
        assert len(vs) == len(set(vs))


@pytest.mark.parametrize("dtype", DTYPES)
def test_rescale_scalars_are_one(dtype):
    # test/test_normalize.jl:40-66
    g = O.grid_graph((3, 2))
    net = O.random_network(g, 2, dtype=dtype)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), maxiter=20)
    net2, msgs2 = O.rescale(net, msgs)
    zv, ze = O.region_scalars(net2, msgs2)
    assert np.allclose(zv, 1.0) and np.allclose(ze, 1.0)
    assert np.allclose(O.scalar(net2, msgs2), 1.0)
    # re-running BP on the rescaled state gives norm 1
    m3, _, _ = O.bp_update(net2, O.identity_messages(net2), maxiter=20)
    assert np.allclose(O.scalar(net2, m3), 1.0)


def test_rescale_tree_exact():
    # test/test_normalize.jl:16-28
    g = O.comb_tree_graph(2, 3)
    net = O.random_network(g, 2, dtype=np.float64)
    seq = O.default_edge_sequence(g)
    msgs = {}
    for (v, w) in seq:
        msgs[(v, w)] = O.updated_message(net, msgs, v, w)
    net2, _ = O.rescale(net, msgs)
    assert np.allclose(O.exact_norm_sqr(net2), 1.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [2, 3, 5, 10])
def test_map_eigvals(dtype, n):
    # test/test_map_eigvals.jl:7-34
    rng = np.random.default_rng(1234)
    a = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if dtype == np.complex128 else 0)
    p = (a @ a.conj().T).astype(dtype)
    sq = O.map_eigvals(np.sqrt, p)
    inv = O.map_eigvals(lambda x: 1 / x, p)
    isq = O.map_eigvals(lambda x: 1 / np.sqrt(x), p)
    assert np.allclose(sq @ sq.conj().T, p)
    assert np.allclose(inv @ p, np.eye(n))
    assert np.allclose(isq @ sq, np.eye(n))


def test_truncate_spectrum_semantics():
    p = np.array([0.5, 0.3, 0.15, 0.04, 0.01])
    assert O.truncate_spectrum(p) == (5, 0.0)
    n, err = O.truncate_spectrum(p, maxdim=3)
    assert n == 3 and np.isclose(err, 0.05)
    n, err = O.truncate_spectrum(p, cutoff=0.011)
    assert n == 4 and np.isclose(err, 0.01)
    n, err = O.truncate_spectrum(p, cutoff=0.05)
    assert n == 3 and np.isclose(err, 0.05)
    n, err = O.truncate_spectrum(p, cutoff=10.0)
    assert n == 1  # never below mindim
    n, err = O.truncate_spectrum(np.array([1.0, 1e-3, -1e-18]), cutoff=1e-12)
    assert n == 2


@pytest.mark.parametrize("dtype", DTYPES)
def test_simple_update_identity_gate_is_gauge_only(dtype):
    g = O.grid_graph((2, 2))
    net = O.random_network(g, 2, dtype=dtype)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), maxiter=30)
    e = 0
    gate = np.eye(4).reshape(2, 2, 2, 2).astype(dtype)
    new, info = O.simple_update_bp(net, msgs, e, gate)
    assert info["truncerr"] == 0
    psi0, psi1 = O._state_vector(net), O._state_vector(new)
    assert np.allclose(psi0, psi1, atol=1e-10)


@pytest.mark.parametrize("dtype", DTYPES)
def test_simple_update_exact_on_tree_without_truncation(dtype):
    # apply.jl:33-95: with exact (tree) environments and no truncation the gate is applied exactly
    g = O.chain_graph(4)
    net = O.random_network(g, 3, dtype=dtype)
    msgs = {}
    for (v, w) in O.default_edge_sequence(g):
        msgs[(v, w)] = O.updated_message(net, msgs, v, w)
    gate = O.random_unitary(4, seed=5, dtype=dtype).reshape(2, 2, 2, 2)
    e = 1
    new, info = O.simple_update_bp(net, msgs, e, gate)
    assert info["newdim"] == 6 and info["truncerr"] == 0
    assert np.allclose(O._state_vector(new), O.exact_apply2(net, e, gate), atol=1e-10)


def test_simple_update_truncation_reports_error():
    # test/test_apply.jl:52-64: truncerr != 0 with maxdim = chi, fidelity > 0
    g = O.grid_graph((2, 2))
    net = O.random_network(g, 2, dtype=np.complex128)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), maxiter=20)
    gate = O.random_unitary(4, seed=7).reshape(2, 2, 2, 2)
    e = g.eid[(3, 1)]
    new, info = O.simple_update_bp(net, msgs, e, gate, maxdim=2, normalize=True)
    assert info["truncerr"] != 0 and info["newdim"] == 2
    exact = O.exact_apply2(net, e, gate)
    got = O._state_vector(new)
    f = np.vdot(got, exact) / np.sqrt(np.vdot(exact, exact) * np.vdot(got, got))
    assert abs(f) ** 2 > 0.5
    for v in g.edges[e]:
        assert np.isclose(np.linalg.norm(new.tensors[v]), 1.0)


def test_synchronous_vs_sequential_same_fixed_point():
    g = O.grid_graph((3, 3))
    net = O.random_network(g, 2, dtype=np.complex128)
    seq = O.parallel_edge_sequence(g)
    m_seq, _, _ = O.bp_update(net, O.identity_messages(net), seq=O.default_edge_sequence(g), maxiter=200, tol=1e-30)
    m_syn, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=400, tol=1e-30)
    for k in m_seq:
        assert O.message_diff(m_seq[k], m_syn[k]) < 1e-12


# ---- bilinear forms: test/test_inner.jl:12-48 ("Inner products, BP vs exact comparison") ----
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_inner_bp_equals_exact_on_trees(dtype):
    g = O.random_tree_graph(7, seed=4)
    x = O.random_network(g, 2, dtype=dtype, seed=1234)
    y = O.random_network(g, [2, 3, 2, 3, 2, 3], dtype=dtype, seed=4321)
    net = O.bilinear_network(x, y)
    msgs, _, _ = O.bp_update(net, {}, seq=O.default_edge_sequence(g), maxiter=1)
    exact = O.exact_inner(x, y)
    assert abs(O.scalar(net, msgs) - exact) < 1e-10 * abs(exact)
    # three layers <x|A|y> (test_inner.jl:41-48) with a random operator network of bond dimension 2
    rng = np.random.default_rng(9)
    ops = [rng.standard_normal((2, 2) + (2,) * len(g.inc[v])).astype(dtype) for v in range(g.nv)]
    a = O.Network(g, ops, dtype)
    net3 = O.bilinear_network(x, O.apply_operator_network(a, y))
    msgs3, _, _ = O.bp_update(net3, {}, seq=O.default_edge_sequence(g), maxiter=1)
    exact3 = O.exact_inner_operator(x, a, y)
    assert abs(O.scalar(net3, msgs3) - exact3) < 1e-10 * abs(exact3)


def test_bilinear_with_equal_layers_is_the_quadratic_form():
    g = O.grid_graph((3, 3))
    psi = O.random_network(g, 2, dtype=np.complex128, seed=3)
    seq = O.parallel_edge_sequence(g)
    m_q, _, _ = O.bp_update(psi, O.identity_messages(psi), seq=seq, groups=O.synchronous_groups(seq), maxiter=4)
    both = O.bilinear_network(psi, psi)
    m_b, _, _ = O.bp_update(both, O.identity_messages(both), seq=seq, groups=O.synchronous_groups(seq), maxiter=4)
    assert max(np.abs(m_q[k] - m_b[k]).max() for k in m_q) < 1e-14


# ---- multi-site partitions: test/test_expect.jl:22-39 (group by column to make BP exact) ----
@pytest.mark.parametrize("dims", [(2, 2), (3, 3), (3, 2)])
def test_column_partition_makes_bp_exact(dims):
    g = O.grid_graph(dims)
    net = O.random_network(g, 2, dtype=np.complex128, seed=1234)
    cols = {}
    for v, c in enumerate(g.coords):
        cols.setdefault(c[0], []).append(v)
    groups = [cols[k] for k in sorted(cols)]
    coarse, group_of = O.partition_network(net, groups)
    assert coarse.graph.is_tree()
    msgs, _, _ = O.bp_update(coarse, {}, seq=O.default_edge_sequence(coarse.graph), maxiter=1)
    sz = 0.5 * O.PAULI_Z
    for v in range(g.nv):
        gi = group_of[v]
        lifted = O.lift_operator([2] * len(groups[gi]), groups[gi].index(v), sz)
        assert abs(O.expect1(coarse, msgs, gi, lifted) - O.exact_expect1(net, v, sz)) < 1e-12


# ---- test/test_forms.jl:62-75: on a tree the BP environment of a site equals the exact one ----
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_bp_environment_of_a_site_is_exact_on_a_tree(dtype):
    import string
    g = O.grid_graph((1, 4))  # the 1 x 4 chain of test_forms.jl
    net = O.random_network(g, [2, 3, 2], dtype=dtype, seed=1234)
    msgs, _, _ = O.bp_update(net, {}, seq=O.default_edge_sequence(g), maxiter=1)
    v = 1
    # exact environment of v in <psi|psi>: every other ket and bra tensor contracted, the bonds of v left open
    letters = iter(string.ascii_letters)
    site = [next(letters) for _ in range(g.nv)]
    kb = [next(letters) for _ in range(g.ne)]
    bb = [next(letters) for _ in range(g.ne)]
    ops, subs = [], []
    for u in range(g.nv):
        if u == v:
            continue
        ops += [net.tensors[u], net.tensors[u].conj()]
        subs += [site[u] + "".join(kb[e] for e in g.inc[u]), site[u] + "".join(bb[e] for e in g.inc[u])]
    out = "".join(kb[e] + bb[e] for e in g.inc[v])
    exact = np.einsum(",".join(subs) + "->" + out, *ops)
    # BP environment = outer product of the incoming messages M_{u->v}[a, a']
    bp = np.ones(())
    for e in g.inc[v]:
        bp = np.multiply.outer(bp, msgs[(g.other(e, v), v)])
    exact = exact / np.linalg.norm(exact)
    bp = bp / np.linalg.norm(bp)
    phase = np.vdot(bp, exact)  # BP messages are normalised individually: compare up to one scalar
    assert abs(abs(phase) - 1) < 1e-12 and np.linalg.norm(exact - phase * bp) < 1e-12


def test_partition_plan_quotient_graph():
    # host bookkeeping of multi-site partitions: a 3 x 3 grid grouped by column is a 3-vertex chain whose bonds fuse
    # the three horizontal edges between neighbouring columns (ascending edge id on both sides)
    import itn_b200 as E
    g = E.named_grid((3, 3))
    cols = {}
    for v, c in enumerate(g.names):
        cols.setdefault(c[0], []).append(v)
    groups = [cols[k] for k in sorted(cols)]
    group_of, qedges, bundles = E.partition_plan(g, groups)
    assert qedges == [(0, 1), (1, 2)]
    assert all(len(bundles[q]) == 3 and bundles[q] == sorted(bundles[q]) for q in qedges)
    assert sorted(group_of) == list(range(9)) and [group_of[v] for v in groups[2]] == [2, 2, 2]
    with pytest.raises(AssertionError):
        E.partition_plan(g, [[0, 1], [1, 2, 3, 4, 5, 6, 7, 8]])
