"""GPU parity for bilinear form networks: inner(phi, psi; alg = "bp"), loginner and inner(phi, A, psi; alg = "bp")
(src/inner.jl:100-171, BilinearFormNetwork src/formnetworks/bilinearformnetwork.jl:23-94).

Restates test/test_inner.jl:12-48 (BP equals the exact inner product on a tree, two and three layers) on the engine and
adds loopy-graph parity against the oracle: with an explicit bra layer the messages are general (non-Hermitian) matrices,
so this also checks that no kernel on the path silently assumes bra = conj(ket).  Tolerance 1e-10 (north_star)."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import assert_messages_close, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10
DTYPES = [np.float64, np.complex128]


@pytest.fixture(scope="module")
def ctx():
    return E.Context(0)


def host_net(net):
    g = E.NamedGraph(net.graph.nv, net.graph.edges)
    return E.ITensorNetwork(g, [t.copy() for t in net.tensors], net.dtype)


def random_operator_network(g, dtype, chi=2, d=2, seed=9):
    rng = np.random.default_rng(seed)
    ops = []
    for v in range(g.nv):
        shape = (d, d) + (chi,) * len(g.inc[v])
        t = rng.standard_normal(shape)
        if np.dtype(dtype).kind == "c":
            t = t + 1j * rng.standard_normal(shape)
        ops.append(t.astype(dtype))
    return O.Network(g, ops, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_inner_on_tree_equals_exact(ctx, dtype):
    # test/test_inner.jl:12-39: uniform tree, chi = 2; here also with unequal bond dimensions in the two states
    g = O.random_tree_graph(8, seed=5)
    x = O.random_network(g, 2, dtype=dtype, seed=1234)
    y = O.random_network(g, [2, 3, 2, 3, 2, 3, 2], dtype=dtype, seed=4321)
    exact = O.exact_inner(x, y)
    got = E.inner(host_net(x), host_net(y), alg="bp", ctx=ctx)
    got_log = np.exp(E.loginner(host_net(x), host_net(y), alg="bp", ctx=ctx))
    assert abs(got - exact) < TOL * abs(exact)
    assert abs(got_log - exact) < TOL * abs(exact)
    # <x|x> through the bilinear route equals norm_sqr through the quadratic one
    nx = E.inner(host_net(x), host_net(x), ctx=ctx)
    assert abs(nx - O.exact_norm_sqr(x)) < TOL * abs(nx)


@pytest.mark.parametrize("dtype", DTYPES)
def test_three_layer_inner_on_tree_equals_exact(ctx, dtype):
    # test/test_inner.jl:41-48: <x|A|y> with an operator network A (there: the TTN of a Heisenberg Hamiltonian)
    g = O.random_tree_graph(6, seed=2)
    x = O.random_network(g, 2, dtype=dtype, seed=11)
    y = O.random_network(g, 3, dtype=dtype, seed=12)
    a = random_operator_network(g, dtype)
    exact = O.exact_inner_operator(x, a, y)
    got = E.inner(host_net(x), host_net(y), operator=a.tensors, alg="bp", ctx=ctx)
    assert abs(got - exact) < TOL * abs(exact)


def test_operator_layer_is_contracted_on_the_device(ctx):
    # <phi|A|psi> (src/inner.jl:154-171, bilinearformnetwork.jl:23-42): the operator and ket tensors of a vertex are one
    # partition; the engine contracts them (itn_tensordot) and the result equals the oracle's fused-bond restatement
    g = O.random_tree_graph(7, seed=3)
    phi = O.random_network(g, [2, 3, 2, 1, 2, 3], dtype=np.complex128, seed=1)
    psi = O.random_network(g, [3, 2, 3, 2, 2, 1], dtype=np.complex128, seed=2)
    rng = np.random.default_rng(0)
    ops = [rng.standard_normal((2, 2) + (2,) * len(g.inc[v])) + 0j for v in range(g.nv)]
    l0 = ctx.launch_count()
    ket, bra = E.inner_network(host_net(phi), host_net(psi), ops, ctx=ctx)
    assert ctx.launch_count() - l0 >= g.nv  # one device contraction per vertex
    ref = O.bilinear_network(phi, O.apply_operator_network(O.Network(g, ops, np.complex128), psi))
    for v in range(g.nv):
        assert ket.tensors[v].shape == bra.tensors[v].shape == ref.tensors[v].shape
        assert np.linalg.norm(ket.tensors[v] - ref.tensors[v]) < 1e-13 * np.linalg.norm(ref.tensors[v])
        assert np.array_equal(bra.tensors[v], ref.bra[v])


CASES = [("grid3x3_chi3", (3, 3), 3), ("grid4x4_chi2", (4, 4), 2), ("cubic3_chi2", (3, 3, 3), 2)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,dims,chi", CASES, ids=[c[0] for c in CASES])
def test_bilinear_sweeps_match_oracle_on_loopy_graphs(ctx, dtype, name, dims, chi):
    g = O.grid_graph(dims)
    phi = O.random_network(g, chi, dtype=dtype, seed=21)
    psi = O.random_network(g, chi, dtype=dtype, seed=22)
    net = O.bilinear_network(phi, psi)
    seq = O.parallel_edge_sequence(g)
    ref, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=4)
    ket, bra = E.inner_network(host_net(phi), host_net(psi))
    bpc = E.BeliefPropagationCache(ket, ctx=ctx, bra=bra, messages="identity")
    E.update(bpc, maxiter=4, edge_sequence=[[e] for e in seq], inplace=True)
    assert_messages_close(bpc, ref, TOL)
    # non-Hermitian by construction: the test would not notice a conj(ket) close otherwise
    m = bpc.message(seq[0])
    assert np.linalg.norm(m - m.conj().T) > 1e-3 * np.linalg.norm(m)
    zv, ze = O.region_scalars(net, ref)
    assert rel_err(E.vertex_scalars(bpc), zv) < TOL
    assert rel_err(E.edge_scalars(bpc), ze) < TOL
    ls = O.logscalar(net, ref)
    assert abs(np.exp(E.logscalar(bpc)) - np.exp(ls)) < TOL * abs(np.exp(ls))
    # sequential (Gauss-Seidel) schedule, one sweep from identical inputs
    seq2 = O.default_edge_sequence(g)
    ref2, _, _ = O.bp_update(net, ref, seq=seq2, maxiter=1)
    out = E.update(bpc, maxiter=1, edge_sequence=seq2)
    assert_messages_close(out, ref2, TOL)
    # updated_message and the clone keep the bra layer
    um = E.updated_message(out.copy(), seq2[3])
    assert rel_err(um, O.updated_message(net, ref2, *seq2[3])) < TOL


def test_bra_equal_to_ket_reproduces_the_quadratic_form(ctx):
    g = O.grid_graph((3, 4))
    psi = O.random_network(g, 3, dtype=np.complex128, seed=5)
    seq = [[e] for e in O.parallel_edge_sequence(g)]
    q = E.update(E.BeliefPropagationCache(host_net(psi), ctx=ctx), maxiter=3, edge_sequence=seq)
    b = E.update(E.BeliefPropagationCache(host_net(psi), ctx=ctx, bra=host_net(psi)), maxiter=3, edge_sequence=seq)
    for e in O.parallel_edge_sequence(g):
        assert rel_err(b.message(e), q.message(e)) < 1e-13


def test_tile_path_is_bypassed_for_bilinear_forms(ctx):
    # degree-4, chi = 16 vertices qualify for the DMMA tile kernels, which close with conj(ket): a bra layer must
    # route them to the shape-generic kernels
    g = O.grid_graph((4, 4))
    phi = O.random_network(g, 16, dtype=np.complex128, seed=31)
    psi = O.random_network(g, 16, dtype=np.complex128, seed=32)
    net = O.bilinear_network(phi, psi)
    seq = O.parallel_edge_sequence(g)
    ref, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
    bpc = E.BeliefPropagationCache(host_net(psi), ctx=ctx, bra=host_net(phi), messages="identity")
    E.update(bpc, maxiter=2, edge_sequence=[[e] for e in seq], inplace=True)
    assert_messages_close(bpc, ref, TOL)


def test_observables_and_gates_are_rejected_on_bilinear_forms(ctx):
    g = O.grid_graph((2, 2))
    psi = O.random_network(g, 2, dtype=np.complex128, seed=1)
    phi = O.random_network(g, 2, dtype=np.complex128, seed=2)
    bpc = E.BeliefPropagationCache(host_net(psi), ctx=ctx, bra=host_net(phi), messages="identity")
    E.update(bpc, maxiter=2, inplace=True)
    for call in (lambda: E.expect(bpc, "Z"), lambda: E.rescale(bpc),
                 lambda: E.apply(np.eye(4).reshape(2, 2, 2, 2), bpc, g.edges[0], maxdim=2)):
        with pytest.raises(E.ITNError) as err:
            call()
        assert err.value.code == 6  # ITN_EUNSUPPORTED
