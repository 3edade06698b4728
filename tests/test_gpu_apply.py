"""GPU parity for the simple-update gate path (src/apply.jl:9-146) and map_eigvals (apply.jl:21-25).

Site tensors after a two-site gate are defined only up to a gauge on the updated bond (SVD phases; the
reference's own qr/svd have the same freedom), so the comparison set is the gauge-invariant one of
SURVEY.md A.8: kept dimension, singular values, truncation error, the contracted pair A1'.A2', and the
downstream BP messages / observables.  Tolerance 1e-10 (north_star)."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import make_pair, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10
DTYPES = [np.float64, np.complex128]


@pytest.fixture(scope="module")
def ctx():
    return E.Context(0)


def pair_tensor(tensors, g, e):
    u, v = g.edges[e]
    ku, kv = g.inc[u].index(e), g.inc[v].index(e)
    return np.tensordot(tensors[u], tensors[v], axes=([1 + ku], [1 + kv]))


def bp_both(net, psi, ctx, iters):
    seq = O.parallel_edge_sequence(net.graph)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=iters)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=iters, edge_sequence=[[e] for e in seq])
    return msgs, bpc


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [2, 3, 5, 10, 16])
def test_map_eigvals(ctx, dtype, n):
    # test/test_map_eigvals.jl:7-34
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if np.dtype(dtype).kind == "c" else 0)
    p = (a @ a.conj().T).astype(dtype)
    sq = E.map_eigvals("sqrt", p, ctx=ctx)
    isq = E.map_eigvals("invsqrt", p, ctx=ctx)
    inv = E.map_eigvals("inv", p, ctx=ctx)
    assert rel_err(sq @ sq.conj().T, p) < 1e-11
    assert rel_err(inv @ p, np.eye(n)) < 1e-8
    assert rel_err(isq @ sq, np.eye(n)) < 1e-9
    cut = 10 * np.finfo(np.float64).eps
    assert rel_err(E.map_eigvals("sqrt", p, cutoff=cut, ctx=ctx), O.map_eigvals(np.sqrt, p, cutoff=cut)) < 1e-10
    # batch + diagonal short-circuit (map_diag, apply.jl:22)
    d = np.diag(rng.uniform(0.5, 2.0, n)).astype(dtype)
    out = E.map_eigvals("invsqrt", np.stack([p, d]), ctx=ctx)
    assert rel_err(out[0], isq) < 1e-13
    assert rel_err(out[1], np.diag(1.0 / np.sqrt(np.diagonal(d)))) < 1e-15
    # rank-deficient input: eigenvalues below the cutoff are dropped (pseudo-inverse semantics)
    b = rng.standard_normal((n, max(1, n // 2))) + 0j
    q = (b @ b.conj().T).astype(dtype) if np.dtype(dtype).kind == "c" else (b.real @ b.real.T)
    ref = O.map_eigvals(lambda x: 1.0 / np.sqrt(x), q, cutoff=1e-12)
    assert rel_err(E.map_eigvals("invsqrt", q, cutoff=1e-12, ctx=ctx), ref) < 1e-8


def test_map_eigvals_indefinite_hermitian(ctx):
    # map_eigvals(f, A; ishermitian = true) is advertised for any Hermitian A (src/apply.jl:9-25), not only PSD ones:
    # a +-lambda pair mixes in the one-sided Jacobi SVD (H = [[0, 1], [1, 0]] has V = 1), which the engine detects and
    # redoes through H + shift * 1 (csrc/itn_linalg.cu, itn_dev_map_eigvals)
    def ref(f, h):
        w, v = np.linalg.eigh(h)
        return (v * f(w.astype(np.complex128))) @ v.conj().T
    h2 = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    inv = E.map_eigvals("inv", h2, ctx=ctx)
    assert rel_err(inv, h2) < 1e-13  # its own inverse
    sq = E.map_eigvals("sqrt", h2, ctx=ctx)
    assert rel_err(sq @ sq, h2) < 1e-12
    rng = np.random.default_rng(5)
    for n in (3, 8, 17):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        w = np.concatenate([np.linspace(0.5, 2.0, n - n // 2), -np.linspace(0.5, 2.0, n // 2)])  # exact +-lambda pairs
        q, _ = np.linalg.qr(a)
        h = (q * w) @ q.conj().T
        h = (h + h.conj().T) / 2
        assert rel_err(E.map_eigvals("inv", h, ctx=ctx), np.linalg.inv(h)) < 1e-10
        assert rel_err(E.map_eigvals("sqrt", h, ctx=ctx), ref(np.sqrt, h)) < 1e-10
        # batch mixing a PSD and an indefinite matrix: only the indefinite one takes the shifted route
        p = a @ a.conj().T
        out = E.map_eigvals("inv", np.stack([p, h]), ctx=ctx)
        assert rel_err(out[0], np.linalg.inv(p)) < 1e-8 and rel_err(out[1], np.linalg.inv(h)) < 1e-10


CASES = [
    # name, graph, chi, bp iterations, edge, maxdim, cutoff
    ("grid3x3_chi3_notrunc", lambda: O.grid_graph((3, 3)), 3, 8, 5, None, None),
    ("grid3x3_chi3_maxdim2", lambda: O.grid_graph((3, 3)), 3, 8, 6, 2, None),
    ("grid3x3_chi4_cutoff", lambda: O.grid_graph((3, 3)), 4, 8, 3, None, 1e-3),
    ("grid3x2_ragged", lambda: O.grid_graph((3, 2)), [2, 3, 4, 2, 3, 2, 4], 8, 2, 3, 1e-6),
    ("cubic2_chi2", lambda: O.grid_graph((2, 2, 2)), 2, 8, 4, None, None),
    ("chain4_chi3", lambda: O.chain_graph(4), 3, 3, 1, 4, None),
]


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,mk,chi,iters,e,maxdim,cutoff", CASES, ids=[c[0] for c in CASES])
def test_apply2_matches_oracle(ctx, dtype, normalize, name, mk, chi, iters, e, maxdim, cutoff):
    g = mk()
    net, psi = make_pair(g, chi, dtype)
    msgs, bpc = bp_both(net, psi, ctx, iters)
    v1, v2 = g.edges[e]
    gate = O.random_unitary(4, seed=11, dtype=dtype).reshape(2, 2, 2, 2)
    ref, info = O.simple_update_bp(net, msgs, e, gate, maxdim=maxdim, cutoff=cutoff, normalize=normalize)
    got = {}
    out = E.apply(gate, bpc, (v1, v2), maxdim=maxdim, cutoff=cutoff, normalize=normalize,
                  callback=lambda **kw: got.update(kw))
    # test/test_apply.jl:62: the callback sees the truncation error; here it must also match the oracle's
    assert out.edge_dim(e) == info["newdim"]
    assert rel_err(got["singular_values"], info["svals"][:info["newdim"]]) < TOL
    assert abs(got["truncation_error"] - info["truncerr"]) < TOL
    new = [out.factor(v) for v in range(g.nv)]
    assert rel_err(pair_tensor(new, g, e), pair_tensor(ref.tensors, g, e)) < 1e-9
    for v in range(g.nv):
        if v not in (v1, v2):
            assert np.array_equal(new[v], net.tensors[v])
    # out-of-place: the input cache is untouched (apply.jl:106 copies)
    assert np.array_equal(bpc.factor(v1), net.tensors[v1])
    # downstream: BP on the updated network, messages away from the gated bond and <Z> agree
    seq = O.parallel_edge_sequence(g)
    m2 = O.reset_edge_messages(ref, msgs, e)
    m2, _, _ = O.bp_update(ref, m2, seq=seq, groups=O.synchronous_groups(seq), maxiter=4)
    out = E.update(out, maxiter=4, edge_sequence=[[x] for x in seq])
    for k, m in m2.items():
        if g.eid[k] != e:
            assert rel_err(out.message(k), m) < 1e-9
    ez = E.expect(out, "Z")
    for v in range(g.nv):
        assert abs(ez[v] - O.expect1(ref, m2, v, O.PAULI_Z)) < 1e-9


def graded_site(rng, shape, k, kappa, dtype):
    """Site tensor whose matricisation (every other bond) x (site, bond k) has singular values graded from 1 to 1 / kappa:
    the matrix simple_update_bp QR-factorises (src/apply.jl:70-76) is then ill conditioned."""
    nd = len(shape)
    cols = shape[0] * shape[1 + k]
    rows = int(np.prod(shape)) // cols
    cplx = np.dtype(dtype).kind == "c"

    def rnd(m, n):
        x = rng.standard_normal((m, n))
        return x + 1j * rng.standard_normal((m, n)) if cplx else x
    r = min(rows, cols)
    u, _ = np.linalg.qr(rnd(rows, r))
    v, _ = np.linalg.qr(rnd(cols, r))
    sv = np.logspace(0, -np.log10(kappa), r)
    mat = (u * sv) @ v.conj().T                                   # rows x (s, l)
    outer = [shape[1 + j] for j in range(nd - 1) if j != k]
    t = mat.reshape(outer + [shape[0], shape[1 + k]])             # [outer..., s, l]
    src = [1 + j for j in range(nd - 1) if j != k] + [0, 1 + k]
    return np.ascontiguousarray(np.transpose(t, np.argsort(src))).astype(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kappa", [1e4, 1e6, 1e8], ids=["k1e4", "k1e6", "k1e8"])
@pytest.mark.parametrize("name,mk,chi,e", [("grid3x3_chi3", lambda: O.grid_graph((3, 3)), 3, 5),
                                          ("grid4x4_chi16_tile", lambda: O.grid_graph((4, 4)), 16, 13)], ids=["chi3", "chi16"])
def test_apply2_ill_conditioned_sites(dtype, kappa, name, mk, chi, e):
    # The engine takes R from a Cholesky factorisation of the bond environment (Gram matrix: condition number squared)
    # where the reference QR-factorises the absorbed site tensor.  Sides whose Cholesky pivots span more than 1e4 repeat
    # the factorisation on A R1^+ (CholeskyQR2), thin sides (degree <= 2) invert (Y^H Y) ~ 1 instead of R R^H; a failed
    # pivot (kappa = 1e8: the Gram matrix is singular to working precision) falls back to the Jacobi eigen route.
    # Singular values, truncation error and kept dimension agree with the oracle's QR route to 1e-10 at every condition
    # number; the updated pair to 1e-10 up to kappa = 1e6 (measured 1e-13 / 8e-12) and to 1e-6 at 1e8 (measured 3e-8).
    g = mk()
    net, psi = make_pair(g, chi, dtype)
    v1, v2 = g.edges[e]
    rng = np.random.default_rng(17)
    for v in (v1, v2):
        t = graded_site(rng, net.tensors[v].shape, g.slot(v, e), kappa, dtype)
        net.tensors[v] = t
        psi.tensors[v] = t.copy()
    ctx = E.Context(0)
    msgs, bpc = bp_both(net, psi, ctx, 10)
    gate = O.random_unitary(4, seed=11, dtype=dtype).reshape(2, 2, 2, 2)
    for maxdim, cutoff in ((chi, None), (None, 1e-10)):
        ref, info = O.simple_update_bp(net, msgs, e, gate, maxdim=maxdim, cutoff=cutoff)
        got = {}
        out = E.apply(gate, bpc, (v1, v2), maxdim=maxdim, cutoff=cutoff, callback=lambda **kw: got.update(kw))
        sv = info["svals"]
        n = info["newdim"]
        if cutoff is not None:
            # a kept / dropped decision within rounding of the threshold may flip: compare where the spectrum has a gap
            w = sv ** 2 / np.sum(sv ** 2)
            tail = np.cumsum(w[::-1])[::-1]
            if np.any(np.abs(tail - cutoff) < 1e-3 * cutoff):
                continue
        assert out.edge_dim(e) == n, (out.edge_dim(e), n)
        sv_err = float(np.max(np.abs(got["singular_values"] - sv[:n])) / sv[0])
        te_err = abs(got["truncation_error"] - info["truncerr"])
        new = [out.factor(v) for v in range(g.nv)]
        pair_err = rel_err(pair_tensor(new, g, e), pair_tensor(ref.tensors, g, e))
        assert sv_err < TOL and te_err < TOL, (maxdim, cutoff, sv_err, te_err, pair_err)
        assert pair_err < (TOL if kappa <= 1e6 else 1e-6), (maxdim, cutoff, kappa, pair_err)
    if kappa <= 1e6:
        assert ctx.cholqr2_count() > 0, "the second Cholesky pass did not run on an ill-conditioned side"


@pytest.mark.parametrize("dtype", DTYPES)
def test_apply2_exact_on_tree(ctx, dtype):
    # BP environments are exact on a tree, so an untruncated simple update reproduces the exact gate
    g = O.random_tree_graph(6, seed=3)
    net, psi = make_pair(g, 2, dtype)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx, messages="default"))
    e = 2
    gate = O.random_unitary(4, seed=5, dtype=dtype).reshape(2, 2, 2, 2)
    out = E.apply(gate, bpc, g.edges[e])
    exact = O.exact_apply2(net, e, gate)
    got = O._state_vector(O.Network(g, [out.factor(v) for v in range(g.nv)], dtype))
    fid = abs(np.vdot(exact, got)) ** 2 / (np.vdot(exact, exact).real * np.vdot(got, got).real)
    assert abs(fid - 1.0) < 1e-12
    assert rel_err(got, exact) < 1e-10


@pytest.mark.parametrize("dtype", DTYPES)
def test_apply_layer_batched(ctx, dtype):
    # a whole vertex-disjoint colour layer in one batched call == the same gates one at a time
    g = O.grid_graph((4, 4))
    net, psi = make_pair(g, 3, dtype)
    msgs, bpc = bp_both(net, psi, ctx, 6)
    layer = O.edge_coloring(g)[0]
    gates = [O.random_unitary(4, seed=20 + i, dtype=dtype).reshape(2, 2, 2, 2) for i in range(len(layer))]
    batched = bpc.copy()
    info = E.apply_layer(gates, batched, [g.edges[e] for e in layer], maxdim=4, cutoff=1e-8)
    ref = net
    for gt, e in zip(gates, layer):
        ref, inf = O.simple_update_bp(ref, msgs, e, gt, maxdim=4, cutoff=1e-8)
        i = layer.index(e)
        assert info["newdim"][i] == inf["newdim"]
        assert abs(info["truncation_error"][i] - inf["truncerr"]) < TOL
        assert rel_err(info["singular_values"][i], inf["svals"][:inf["newdim"]]) < TOL
    new = [batched.factor(v) for v in range(g.nv)]
    for e in layer:
        assert rel_err(pair_tensor(new, g, e), pair_tensor(ref.tensors, g, e)) < 1e-9


def test_apply_errors_mirror_reference(ctx):
    g = O.grid_graph((2, 2))
    net, psi = make_pair(g, 2, np.complex128)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    gate = np.eye(4).reshape(2, 2, 2, 2)
    with pytest.raises(E.ITNError, match="must be neighbors"):
        E.apply(gate, bpc, (0, 3))  # src/apply.jl:127-129
    with pytest.raises(E.ITNError, match="more than 2 sites"):
        E.apply(gate, bpc, (0, 1, 2))  # src/apply.jl:143
    with pytest.raises(E.ITNError, match="vertex-disjoint"):
        E.apply_layer([gate, gate], bpc.copy(), [(0, 1), (1, 3)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_product_state_growth(ctx, dtype):
    # bond dimension 1 -> d^2-limited growth, the start of a TEBD run (config 3 regime)
    g = O.grid_graph((2, 3))
    tensors = []
    rng = np.random.default_rng(5)
    for v in range(g.nv):
        t = rng.standard_normal((2,) + (1,) * g.degree(v))
        tensors.append(t.astype(dtype))
    net = O.Network(g, tensors, dtype)
    psi = E.ITensorNetwork(E.NamedGraph(g.nv, g.edges), [t.copy() for t in tensors], dtype)
    msgs, bpc = bp_both(net, psi, ctx, 2)
    gate = O.random_unitary(4, seed=9, dtype=dtype).reshape(2, 2, 2, 2)
    ref, info = O.simple_update_bp(net, msgs, 0, gate, maxdim=None, cutoff=1e-14)
    got = {}
    out = E.apply(gate, bpc, g.edges[0], cutoff=1e-14, callback=lambda **kw: got.update(kw))
    assert out.edge_dim(0) == info["newdim"] == 2
    assert rel_err(got["singular_values"], info["svals"][:2]) < TOL
    new = [out.factor(v) for v in range(g.nv)]
    assert rel_err(pair_tensor(new, g, 0), pair_tensor(ref.tensors, g, 0)) < 1e-10


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("e,maxdim", [(17, 16), (18, 16), (17, 7), (18, 11), (8, 16), (4, 16)],
                         ids=["interior_h", "interior_v", "interior_h_trunc7", "interior_v_trunc11", "mixed", "mixed2"])
def test_apply2_dmma_tile_path(dtype, e, maxdim):
    # degree-4, chi = 16, d = 2 sites take the DMMA tile kernels (k_fast phase 1 + k_benv + k_rebuild);
    # the result must match the oracle AND the generic kernels.
    g = O.grid_graph((4, 4))
    net, psi = make_pair(g, 16, dtype)
    c = E.Context(0)
    msgs, bpc = bp_both(net, psi, c, 3)
    v1, v2 = g.edges[e]
    gate = O.random_unitary(4, seed=21, dtype=dtype).reshape(2, 2, 2, 2)
    ref, info = O.simple_update_bp(net, msgs, e, gate, maxdim=maxdim, cutoff=1e-13)
    got = {}
    out = E.apply(gate, bpc, (v1, v2), maxdim=maxdim, cutoff=1e-13, callback=lambda **kw: got.update(kw))
    assert out.edge_dim(e) == info["newdim"]
    assert rel_err(got["singular_values"], info["svals"][:info["newdim"]]) < TOL
    assert abs(got["truncation_error"] - info["truncerr"]) < TOL
    new = [out.factor(v) for v in (v1, v2)]
    tens = list(net.tensors)
    tens[v1], tens[v2] = new
    assert rel_err(pair_tensor(tens, g, e), pair_tensor(ref.tensors, g, e)) < 1e-9
    # second opinion: generic kernels only
    c2 = E.Context(0)
    c2.set_path(1)
    b2 = E.update(E.BeliefPropagationCache(psi, ctx=c2), maxiter=3, edge_sequence=E.parallel_edge_sequence(psi.graph))
    got2 = {}
    out2 = E.apply(gate, b2, (v1, v2), maxdim=maxdim, cutoff=1e-13, callback=lambda **kw: got2.update(kw))
    assert rel_err(got["singular_values"], got2["singular_values"]) < 1e-11
    tens2 = list(net.tensors)
    tens2[v1], tens2[v2] = out2.factor(v1), out2.factor(v2)
    assert rel_err(pair_tensor(tens, g, e), pair_tensor(tens2, g, e)) < 1e-9
    # BP keeps working on the updated network (tile copies are refreshed)
    if maxdim == 16:
        seq = O.parallel_edge_sequence(g)
        m2 = O.reset_edge_messages(ref, msgs, e)
        m2, _, _ = O.bp_update(ref, m2, seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
        out = E.update(out, maxiter=2, edge_sequence=[[x] for x in seq])
        for k, m in m2.items():
            if g.eid[k] != e:
                assert rel_err(out.message(k), m) < 1e-9


def test_prepared_layer_equals_gate_list(ctx):
    # prepare_layer packs a layer once (edge ids + gates in (esrc, edst) orientation); applying it equals the list form
    g = O.grid_graph((4, 3))
    net, psi = make_pair(g, 3, np.complex128)
    msgs, bpc = bp_both(net, psi, ctx, 4)
    layer = O.edge_coloring(g)[1]
    gate = O.random_unitary(4, seed=21).reshape(2, 2, 2, 2)
    pairs = [g.edges[e][::-1] if i % 2 else g.edges[e] for i, e in enumerate(layer)]  # both orientations
    a, b = bpc.copy(), bpc.copy()
    ia = E.apply_layer([gate] * len(layer), a, pairs, maxdim=3, cutoff=1e-12)
    ib = E.apply_layer(E.prepare_layer(b, [gate] * len(layer), pairs), b, maxdim=3, cutoff=1e-12)
    assert np.array_equal(ia["newdim"], ib["newdim"]) and np.array_equal(ia["truncation_error"], ib["truncation_error"])
    for v in range(g.nv):
        assert np.array_equal(a.factor(v), b.factor(v))
    assert a.edge_dim(layer[0]) == int(ia["newdim"][0])
