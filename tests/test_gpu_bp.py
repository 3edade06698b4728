"""GPU parity: the CUDA engine (through the C ABI) against the oracle on identical seeded inputs.

Tolerance: north_star asks for 1e-10 relative in Float64 on messages, norms and observables at a
fixed iteration count with the identical schedule; the asserts below use 1e-10 (typical error 1e-14)."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import assert_messages_close, make_pair, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10
DTYPES = [np.float64, np.complex128]
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module", params=[0, 1, 2], ids=["auto", "generic_dmma", "generic_fma"])
def ctx(request):
    c = E.Context(0)
    c.set_path(request.param)
    yield c


def sync_seq(g):
    return [[e] for e in list(g.edges) + [(v, u) for (u, v) in g.edges]]


@pytest.mark.parametrize("dtype", DTYPES)
def test_tensor_roundtrip(ctx, dtype):
    net, psi = make_pair(O.grid_graph((3, 3)), [2, 3, 2, 4, 3, 2, 2, 3, 4, 2, 3, 2], dtype)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    for v in range(psi.graph.nv):
        assert np.array_equal(bpc.factor(v), net.tensors[v])
    ident = bpc.message((0, 1))
    assert np.array_equal(ident, np.eye(ident.shape[0]))


CASES = [
    ("grid4x4_chi2", lambda: O.grid_graph((4, 4)), 2),           # BASELINE config 1
    ("grid3x3_chi3", lambda: O.grid_graph((3, 3)), 3),
    ("cubic3_chi2", lambda: O.grid_graph((3, 3, 3)), 2),         # degree-6 vertices (config 5 shape)
    ("grid5x4_ragged", lambda: O.grid_graph((5, 4)), None),      # non-uniform bond dims
    ("heavyhex_chi3", O.heavy_hex_eagle_graph, 3),               # config 3 graph
]


def _chis(g, chi, seed=7):
    if chi is not None:
        return chi
    rng = np.random.default_rng(seed)
    return [int(x) for x in rng.integers(1, 5, size=g.ne)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,mk,chi", CASES, ids=[c[0] for c in CASES])
def test_synchronous_sweeps_match_oracle(ctx, dtype, name, mk, chi):
    g = mk()
    net, psi = make_pair(g, _chis(g, chi), dtype)
    seq = O.parallel_edge_sequence(g)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=5)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    info = {}
    E.update(bpc, maxiter=5, edge_sequence=[[e] for e in seq], inplace=True, info=info)
    assert info["iterations"] == 5
    assert_messages_close(bpc, msgs, TOL)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,mk,chi", CASES[:4], ids=[c[0] for c in CASES[:4]])
def test_sequential_sweeps_match_oracle_every_iteration(ctx, dtype, name, mk, chi):
    g = mk()
    net, psi = make_pair(g, _chis(g, chi), dtype)
    seq = O.default_edge_sequence(g)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    _, _, _, hist = O.bp_update(net, O.identity_messages(net), seq=seq, maxiter=3, return_history=True)
    msgs = O.identity_messages(net)
    for it in range(3):
        # one Gauss-Seidel sweep of the oracle from the engine's current messages: BP on these random networks is not
        # contractive (on cubic3_chi2 complex a rounding-level difference grows ~300x per sweep, measured with three
        # different kernel families, tools/dbg_cubic.py), so every iteration is compared from identical inputs ...
        ref, _, _ = O.bp_update(net, msgs, seq=seq, maxiter=1)
        E.update(bpc, maxiter=1, edge_sequence=seq, inplace=True)
        assert_messages_close(bpc, ref, TOL)
        msgs = {k: bpc.message(k) for k in ref}
        # ... and against the oracle's own history while the amplification leaves room (first two sweeps)
        if it < 2:
            assert_messages_close(bpc, hist[it], 1e-9)


@pytest.mark.parametrize("dtype", DTYPES)
def test_convergence_and_fixed_point(ctx, dtype):
    # test/test_belief_propagation.jl:18-55 on the engine
    g = O.grid_graph((3, 3))
    net, psi = make_pair(g, 2, dtype)
    seq = O.default_edge_sequence(g)
    msgs, it_o, diff_o = O.bp_update(net, O.identity_messages(net), seq=seq, maxiter=40, tol=EPS)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    info = {}
    out = E.update(bpc, maxiter=40, tol=EPS, edge_sequence=seq, info=info)
    assert abs(info["iterations"] - it_o) <= 1
    res = E.message_residuals(out)
    assert np.all(res < 10 * EPS)
    for k, m in msgs.items():
        assert O.message_diff(out.message(k), m) < 1e-12
    # out-of-place: the input cache still holds identity messages
    assert np.array_equal(bpc.message((0, 1)), np.eye(2))
    # updated_message agrees with the oracle's
    um = E.updated_message(out, (4, 5))
    assert rel_err(um, O.updated_message(net, out.messages(), 4, 5)) < TOL


@pytest.mark.parametrize("dtype", DTYPES)
def test_tree_without_initial_messages(ctx, dtype):
    g = O.random_tree_graph(9, seed=5)
    net, psi = make_pair(g, 3, dtype)
    seq = O.default_edge_sequence(g)
    msgs = {}
    for (v, w) in seq:
        msgs[(v, w)] = O.updated_message(net, msgs, v, w)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx, messages="default")
    out = E.update(bpc)  # default maxiter = 1 on a tree, forest-cover order
    assert_messages_close(out, msgs, TOL)
    assert rel_err(E.scalar(out), O.exact_norm_sqr(net)) < TOL
    sz = E.expect(out, "Sz")
    for v in range(g.nv):
        assert abs(sz[v] - O.exact_expect1(net, v, 0.5 * O.PAULI_Z)) < TOL


def test_errors_mirror_reference(ctx):
    g = O.grid_graph((2, 2))
    net, psi = make_pair(g, 2, np.float64)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    with pytest.raises(E.ITNError, match="number of iterations"):
        E.update(bpc)  # loopy graph, no maxiter (abstractbeliefpropagationcache.jl:315-317)
    with pytest.raises(E.ITNError, match="not an edge"):
        E.update(bpc, maxiter=1, edge_sequence=[(0, 3)])
    nomsg = E.BeliefPropagationCache(psi, ctx=ctx, messages=None)
    with pytest.raises(E.ITNError, match="does not exist"):
        E.update(nomsg, maxiter=1)
    with pytest.raises(E.ITNError):
        bpc.set_factor(0, np.zeros((2, 3, 3)))


@pytest.mark.parametrize("dtype", DTYPES)
def test_scalars_rescale_observables(ctx, dtype):
    g = O.grid_graph((4, 3))
    net, psi = make_pair(g, 3, dtype)
    seq = O.parallel_edge_sequence(g)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=12)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=12, edge_sequence=[[e] for e in seq])
    zv, ze = E.scalar_factors_quotient(bpc)
    zv_o, ze_o = O.region_scalars(net, msgs)
    assert rel_err(zv, zv_o) < TOL and rel_err(ze, ze_o) < TOL
    assert abs(E.logscalar(bpc) - O.logscalar(net, msgs)) < 1e-9
    # expect / rdm2 (src/expect.jl:5-19, test_belief_propagation.jl:64-91)
    ez = E.expect(bpc, "Z")
    for v in range(g.nv):
        assert abs(ez[v] - O.expect1(net, msgs, v, O.PAULI_Z)) < TOL
    rdms = E.rdm2(bpc, list(range(g.ne)))
    for e in range(g.ne):
        assert rel_err(rdms[e], O.rdm2(net, msgs, e)) < TOL
    zz = E.expect2(bpc, [0, 5], "Z", "Z")
    for i, e in enumerate([0, 5]):
        assert abs(zz[i] - O.expect2(net, msgs, e, O.PAULI_Z, O.PAULI_Z)) < TOL
    # rescale (test/test_normalize.jl:40-66)
    net2, msgs2 = O.rescale(net, msgs)
    r = E.rescale(bpc)
    zv2, ze2 = E.scalar_factors_quotient(r)
    assert np.allclose(zv2, 1.0, atol=1e-12) and np.allclose(ze2, 1.0, atol=1e-12)
    for v in range(g.nv):
        assert rel_err(r.factor(v), net2.tensors[v]) < TOL
    assert_messages_close(r, msgs2, TOL)
    assert abs(E.scalar(r) - 1.0) < 1e-10
    # rescale(bpc; verts) (abstractbeliefpropagationcache.jl:349-395): messages of every edge, tensors of the listed sites only
    sub = [0, 3, g.nv - 1]
    net3, msgs3 = O.rescale(net, msgs, verts=sub)
    r3 = E.rescale(bpc, verts=sub)
    assert_messages_close(r3, msgs3, TOL)
    zv3, ze3 = E.scalar_factors_quotient(r3)
    assert np.allclose(ze3, 1.0, atol=1e-12) and np.allclose(zv3[sub], 1.0, atol=1e-12)
    for v in range(g.nv):
        assert rel_err(r3.factor(v), net3.tensors[v]) < TOL
        if v not in sub:
            assert np.array_equal(r3.factor(v), net.tensors[v])  # untouched
    r0 = E.rescale(bpc, verts=[])  # messages only
    assert_messages_close(r0, msgs3, TOL)


@pytest.mark.parametrize("dtype", DTYPES)
def test_apply1(ctx, dtype):
    g = O.grid_graph((3, 2))
    net, psi = make_pair(g, 2, dtype)
    gate = O.random_unitary(2, seed=3, dtype=dtype)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    out = E.apply(gate, bpc, (4,), normalize=True)
    ref = O.apply1(net, 4, gate, normalize=True)
    assert rel_err(out.factor(4), ref.tensors[4]) < TOL
    assert np.array_equal(bpc.factor(4), net.tensors[4])


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_network(ctx, dtype):
    # test/test_belief_propagation.jl:93-99
    g = O.grid_graph((3, 1))
    net, psi = make_pair(g, 2, dtype)
    psi.tensors[0] = 0 * psi.tensors[0]
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx, messages="default"))
    assert E.scalar(bpc) == 0


def test_midsize_parity(ctx):
    # 10x10 chi=4 complex: every degree bucket of a square lattice
    g = O.grid_graph((10, 10))
    net, psi = make_pair(g, 4, np.complex128)
    seq = O.parallel_edge_sequence(g)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=4)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=4, edge_sequence=[[e] for e in seq])
    assert_messages_close(bpc, msgs, TOL)


@pytest.mark.parametrize("dtype", DTYPES)
def test_dmma_fast_path_chi16(dtype):
    # degree-4, chi=16 vertices run on the DMMA kernels (itn_fast.cu); boundary vertices on the generic ones.
    g = O.grid_graph((5, 5))
    net, psi = make_pair(g, 16, dtype)
    seq = O.parallel_edge_sequence(g)
    msgs, _, diff_o = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=3,
                                  tol=0.0)
    c = E.Context(0)
    bpc = E.BeliefPropagationCache(psi, ctx=c)
    l0 = c.launch_count()
    info = {}
    E.update(bpc, maxiter=3, tol=0.0, edge_sequence=[[e] for e in seq], inplace=True, info=info)
    assert_messages_close(bpc, msgs, TOL)
    assert abs(info["mean_diff"] - diff_o) < 1e-12
    # the generic kernels give the same answer (second opinion on the device)
    for mode in (1, 2):  # shape-generic DMMA kernels, plain FMA kernels
        c2 = E.Context(0)
        c2.set_path(mode)
        b2 = E.BeliefPropagationCache(psi, ctx=c2)
        E.update(b2, maxiter=3, edge_sequence=[[e] for e in seq], inplace=True)
        for k in msgs:
            assert rel_err(bpc.message(k), b2.message(k)) < 1e-12
    # observables downstream of the fast path
    ez = E.expect(bpc, "Z")
    for v in (0, 6, 12):
        assert abs(ez[v] - O.expect1(net, msgs, v, O.PAULI_Z)) < TOL


def test_synchronous_sequence_with_a_repeated_edge(ctx):
    # the reference accepts arbitrary edge lists: a synchronous sequence that lists a directed edge twice computes it
    # twice from the pre-sweep messages (same result); the tile path must not take such a vertex (two staged buffers
    # for one slot), the generic path handles every listed update
    g = O.grid_graph((4, 4))
    net, psi = make_pair(g, 16, np.complex128)
    seq = O.parallel_edge_sequence(g)
    dup = [e for e in seq if e[0] == 5][:1]  # vertex 5 has degree 4
    seq2 = seq + dup
    msgs, _, diff_o = O.bp_update(net, O.identity_messages(net), seq=seq2, groups=O.synchronous_groups(seq2), maxiter=2, tol=0.0)
    info = {}
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=2, tol=0.0, edge_sequence=[[e] for e in seq2], info=info)
    assert_messages_close(bpc, msgs, TOL)
    assert abs(info["mean_diff"] - diff_o) < 1e-12


GROUPED = [
    ("grid3x3_pairs", lambda: O.grid_graph((3, 3)), 3, 2),
    ("grid4x4_by_vertex", lambda: O.grid_graph((4, 4)), 2, 0),
    ("cubic2_triples", lambda: O.grid_graph((2, 2, 2)), 2, 3),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,mk,chi,gsize", GROUPED, ids=[c[0] for c in GROUPED])
def test_grouped_schedule_with_multi_edge_groups(ctx, dtype, name, mk, chi, gsize):
    # update_iteration(alg, bpc, edge_groups) (abstractbeliefpropagationcache.jl:294-308, intended semantics): every group
    # is a SEQUENTIAL pass over its edges starting from the pre-sweep messages; the results of all groups are written at
    # the end of the sweep; the diff is divided by the number of groups (:319-321)
    g = mk()
    net, psi = make_pair(g, chi, dtype)
    seq = O.default_edge_sequence(g)  # an order in which later edges of a group read earlier ones
    if gsize == 0:   # one group per source vertex
        by_v = {}
        for e in seq:
            by_v.setdefault(e[0], []).append(e)
        groups = list(by_v.values())
    else:
        groups = [seq[i:i + gsize] for i in range(0, len(seq), gsize)]
    flat = [e for grp in groups for e in grp]
    ptr = np.cumsum([0] + [len(grp) for grp in groups])
    ranges = [(int(ptr[i]), int(ptr[i + 1])) for i in range(len(groups))]
    for it in (1, 3):
        msgs, _, diff_o = O.bp_update(net, O.identity_messages(net), seq=flat, groups=ranges, maxiter=it, tol=0.0)
        info = {}
        bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=it, tol=0.0, edge_sequence=[list(grp) for grp in groups],
                       info=info)
        assert_messages_close(bpc, msgs, TOL)
        assert abs(info["mean_diff"] - diff_o) < 1e-12, (info["mean_diff"], diff_o)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("dims,chi", [((6, 6), 16), ((4, 4), 3), ((9, 8), 16)], ids=["grid6_chi16", "grid4_chi3", "grid9x8_chi16"])
def test_deferred_upload_first_sweep_overlaps_copy(dtype, dims, chi, monkeypatch):
    # ITN_HOST_DEFERRED (include/itn_b200.h): the constructor only registers the host tensors, the first synchronous
    # sweep runs vertex chunk by vertex chunk behind the host -> device copy.  Same messages as the eager path (bit for
    # bit: identical kernels on identical data) and as the oracle.
    g = O.grid_graph(dims)
    net, psi = make_pair(g, chi, dtype)
    seq = O.parallel_edge_sequence(g)
    sync = [[e] for e in seq]
    c = E.Context(0)
    eager = E.BeliefPropagationCache(psi, ctx=c)
    E.update(eager, maxiter=2, edge_sequence=sync, inplace=True)
    lazy = E.BeliefPropagationCache(psi, ctx=c, defer_upload=True)
    assert lazy._host_refs is not None
    E.update(lazy, maxiter=2, edge_sequence=sync, inplace=True)
    assert lazy._host_refs is None
    for k in [(u, v) for (u, v) in g.edges] + [(v, u) for (u, v) in g.edges]:
        assert np.array_equal(lazy.message(k), eager.message(k))
    if g.nv <= 36:
        msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
        assert_messages_close(lazy, msgs, TOL)
    for v in range(g.nv):
        assert np.array_equal(lazy.factor(v), net.tensors[v])
    # consumers other than update() flush the pending tensors first
    lazy2 = E.BeliefPropagationCache(psi, ctx=c, defer_upload=True)
    assert np.array_equal(lazy2.factor(3), net.tensors[3])
    ez = E.expect(E.update(lazy2, maxiter=2, edge_sequence=sync), "Z")
    ez0 = E.expect(eager, "Z")
    assert max(abs(ez[v] - ez0[v]) for v in range(g.nv)) < 1e-13
    # sequential schedule on deferred tensors
    lazy3 = E.BeliefPropagationCache(psi, ctx=c, defer_upload=True)
    e3 = E.update(E.BeliefPropagationCache(psi, ctx=c), maxiter=1)
    E.update(lazy3, maxiter=1, inplace=True)
    assert np.array_equal(lazy3.message(g.edges[0]), e3.message(g.edges[0]))


@pytest.mark.parametrize("eltype", [np.float32, np.complex64])
def test_single_precision_element_types_are_preserved(eltype):
    # test_belief_propagation.jl:18-55 and test_normalize.jl:52-66 run Float32 / ComplexF32 networks and pin that the
    # element type survives; the engine widens at the boundary, computes in double and narrows what it hands back
    g = O.grid_graph((3, 3))
    net = O.random_network(g, 2, dtype=np.complex128 if np.dtype(eltype).kind == "c" else np.float64, seed=1234)
    eg = E.NamedGraph(g.nv, g.edges)
    psi = E.ITensorNetwork(eg, [t.astype(eltype) for t in net.tensors], eltype)
    bpc = E.BeliefPropagationCache(psi, ctx=E.Context(0))
    seq = [[e] for e in O.parallel_edge_sequence(g)]
    bpc = E.update(bpc, maxiter=25, tol=float(np.finfo(eltype).eps), edge_sequence=seq)
    for e in g.edges:
        assert bpc.message(e).dtype == np.dtype(eltype)
        assert E.message_diff(E.updated_message(bpc, e), bpc.message(e)) < 10 * np.finfo(eltype).eps
    assert bpc.factor(0).dtype == np.dtype(eltype)
    r = E.rescale(bpc)
    zv, ze = E.scalar_factors_quotient(r)
    assert zv.dtype == np.dtype(eltype)
    assert np.allclose(zv, 1.0, atol=1e-5) and np.allclose(ze, 1.0, atol=1e-5)
    out = E.normalize(psi, cache=bpc)
    assert out.dtype == np.dtype(eltype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_environment_on_a_tree_matches_the_exact_environment(ctx, dtype):
    # test/test_forms.jl:62-75: environment(qf, state_vertices(qf, [v]); alg = "bp", update_cache = true) on a tree
    from util import exact_site_environment
    g = O.comb_tree_graph(3, 3)
    net, psi = make_pair(g, 2, dtype)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx, messages="default"))  # tree: no initial messages, one sweep
    for v in (0, 4, 8):
        env = dict(E.environment(bpc, [v]))
        bp = np.ones(())
        for e in g.inc[v]:
            bp = np.multiply.outer(bp, env[(g.other(e, v), v)])
        exact = exact_site_environment(net, v)
        exact, bp = exact / np.linalg.norm(exact), bp / np.linalg.norm(bp)
        phase = np.vdot(bp, exact)
        assert abs(abs(phase) - 1) < TOL and np.linalg.norm(exact - phase * bp) < TOL
