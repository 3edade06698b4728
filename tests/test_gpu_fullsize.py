"""Full-size checks of BASELINE.json configs 2, 3, 4, 5 against the oracle.

Config 2 (32x32, chi=8) and config 3 (heavy-hex, chi=32) are small enough for the oracle to run whole sweeps: messages,
<Z> and <ZZ> are compared at 1e-10 (north_star's requirement).  Configs 4 (64x64, chi=16) and 5 (16^3, chi=6) are
compared on a SAMPLE: the messages into sampled vertices are downloaded, one more production sweep runs on the device,
and every sampled outgoing message must equal the oracle's updated_message on the downloaded inputs (bench.py's
`sampled_parity`, the same check the bench line reports); size-independent properties (fixed-point residual, Hermitian
PSD messages, rescale => all region scalars 1) cover the rest."""
import numpy as np
import pytest

import itn_b200 as E

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def bench_psi(g, chi, dtype=np.complex128, d=2):
    """bench.py's synthetic state (one generator per vertex), so the tests check the bytes the bench times."""
    import torch

    import bench
    tensors, _host, _ = bench.make_psi(torch, g, chi, dtype, d, [True] * g.nv)
    return tensors, E.ITensorNetwork(g, tensors, dtype)


def test_config2_grid32_chi8_messages_and_observables_match_oracle():
    from oracle import itn_oracle as O
    go = O.grid_graph((32, 32))
    g = E.named_grid((32, 32))
    assert go.edges == g.edges
    tensors, psi = bench_psi(g, 8)
    net = O.Network(go, [np.ascontiguousarray(t) for t in tensors], np.complex128)
    ctx = E.Context(0)
    seq = O.parallel_edge_sequence(go)
    nsw = 10
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=nsw)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    E.update(bpc, maxiter=nsw, edge_sequence=[[e] for e in seq], inplace=True)
    worst = max(np.linalg.norm(bpc.message(k) - m) / np.linalg.norm(m) for k, m in msgs.items())
    assert worst < 1e-10, worst
    ez = E.expect(bpc, "Z")
    rng = np.random.default_rng(0)
    verts = list(rng.choice(g.nv, 64, replace=False)) + [0, 31, 1023]
    for v in verts:
        assert abs(ez[int(v)] - O.expect1(net, msgs, int(v), O.PAULI_Z)) < 1e-10, v
    eids = [int(i) for i in rng.choice(g.ne, 24, replace=False)]
    zz = E.expect2(bpc, eids, "Z", "Z")
    for e, got in zip(eids, zz):
        assert abs(got - O.expect2(net, msgs, e, O.PAULI_Z, O.PAULI_Z)) < 1e-10, e
    lz, lo = E.logscalar(bpc), O.logscalar(net, msgs)
    assert abs(lz - lo) < 1e-10 * abs(lo)
    # and to convergence: the reference's property tests at this size
    info = {}
    bpc = E.update(bpc, maxiter=60, tol=1e-14, edge_sequence=[[e] for e in seq], info=info)
    assert info["iterations"] < 60 and info["mean_diff"] <= 1e-14
    edges = [g.edges[i] for i in rng.choice(g.ne, 40, replace=False)]
    res = E.message_residuals(bpc, edges + [(v, u) for u, v in edges])
    assert np.max(res) < 1e-12  # fixed point (second order in the message error)
    for (u, v) in edges[:10]:
        m = bpc.message((u, v))
        assert np.linalg.norm(m - m.conj().T) < 1e-7 * np.linalg.norm(m)  # Hermitian at the fixed point
        assert np.min(np.linalg.eigvalsh((m + m.conj().T) / 2)) > -1e-9 * np.linalg.norm(m)  # positive semi-definite
    r = E.rescale(bpc)
    zv, ze = E.scalar_factors_quotient(r)
    assert np.allclose(zv, 1.0, atol=1e-10) and np.allclose(ze, 1.0, atol=1e-10)
    assert abs(E.scalar(r) - 1.0) < 1e-8


def _sampled_fullsize(dims, chi, nsample, max_contract_ms=None):
    import bench
    g = E.named_grid(dims)
    tensors, psi = bench_psi(g, chi)
    ctx = E.Context(0)
    seq = E.parallel_edge_sequence(g)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
    err, n = bench.sampled_parity(E, bpc, tensors, g, seq, [True] * g.nv, nsample=nsample)
    assert n >= nsample - 4 and err < 1e-10, (err, n)
    tm = bpc.last_timing()
    if max_contract_ms is not None:
        assert tm["contract_ms"] < max_contract_ms, tm
    # size-independent properties of the same state
    rng = np.random.default_rng(1)
    edges = [g.edges[i] for i in rng.choice(g.ne, 6, replace=False)]
    for (u, v) in edges:
        m = bpc.message((u, v))
        assert abs(np.linalg.norm(m) - 1.0) < 1e-12  # normalised (abstractbeliefpropagationcache.jl:234-237)
    lz = E.logscalar(bpc)
    assert np.isfinite(np.real(lz))
    bpc.close()


def test_config4_grid64_chi16_sampled_messages_match_oracle():
    # the DMMA sweep at full size against the ORACLE (not against another CUDA path)
    _sampled_fullsize((64, 64), 16, 24, max_contract_ms=200.0)  # the tile path ran (the generic kernels need ~350 ms per sweep)


def test_config5_cubic16_chi6_sampled_messages_match_oracle():
    _sampled_fullsize((16, 16, 16), 6, 20)


def test_config3_heavyhex_chi32_bp_and_gate_layer_match_oracle():
    # BASELINE config 3 at its full bond dimension: 127-qubit heavy-hex, chi = 32, ComplexF64.  Bond matrices of the
    # simple update are 128 x 128 (interior) down to 2 x 64 (leaves); the oracle still runs this size in seconds.
    from oracle import itn_oracle as O
    g = O.heavy_hex_eagle_graph()
    eg = E.heavy_hex_eagle()
    net = O.random_network(g, 32, dtype=np.complex128, seed=77)
    psi = E.ITensorNetwork(eg, [t.copy() for t in net.tensors], np.complex128)
    ctx = E.Context(0)
    seq = O.parallel_edge_sequence(g)
    sync = [[e] for e in seq]
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=3)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    E.update(bpc, maxiter=3, edge_sequence=sync, inplace=True)
    worst = max(np.linalg.norm(bpc.message(k) - m) / np.linalg.norm(m) for k, m in msgs.items())
    assert worst < 1e-10, worst
    layer = O.edge_coloring(g)[0]
    rng = np.random.default_rng(9)
    h = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    h = (h + h.conj().T) / 2
    w, v = np.linalg.eigh(h)
    gate = ((v * np.exp(-0.3j * w)) @ v.conj().T).reshape(2, 2, 2, 2)
    info = E.apply_layer([gate] * len(layer), bpc, [g.edges[e] for e in layer], maxdim=32, cutoff=1e-12)
    for i, e in enumerate(layer):
        new, inf = O.simple_update_bp(net, msgs, e, gate, maxdim=32, cutoff=1e-12)
        assert info["newdim"][i] == inf["newdim"], (e, info["newdim"][i], inf["newdim"])
        assert abs(info["truncation_error"][i] - inf["truncerr"]) < 1e-10
        sv = inf["svals"][:inf["newdim"]]
        assert np.max(np.abs(info["singular_values"][i] - sv)) < 1e-10 * sv[0]
        if i < 6:  # the updated pair (gauge invariant): A1' . A2' contracted over the new bond
            v1, v2 = g.edges[e]
            k1, k2 = g.slot(v1, e), g.slot(v2, e)
            pair_o = np.tensordot(new.tensors[v1], new.tensors[v2], axes=([1 + k1], [1 + k2]))
            pair_e = np.tensordot(bpc.factor(v1), bpc.factor(v2), axes=([1 + k1], [1 + k2]))
            assert np.linalg.norm(pair_e - pair_o) < 1e-9 * np.linalg.norm(pair_o)
