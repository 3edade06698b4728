"""Full-size checks (BASELINE.json configs 2 and 4) through size-independent properties: the oracle cannot run
these sizes in seconds, so the checks are the reference's own property tests (test_belief_propagation.jl:51-53,
90-91; test_normalize.jl:60-66) plus a device-side second opinion (DMMA kernels vs generic kernels)."""
import numpy as np
import pytest

import itn_b200 as E

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def torch_random_psi(g, chi, d=2, seed=1234):
    import torch
    gen = torch.Generator().manual_seed(seed)
    ts = []
    for v in range(g.nv):
        n = d * chi ** g.degree(v)
        flat = torch.randn(2 * n, generator=gen, dtype=torch.float64).numpy().view(np.complex128) * 2 ** -0.5
        ts.append(np.ndarray((d,) + (chi,) * g.degree(v), dtype=np.complex128, buffer=flat, order="F"))
    return E.ITensorNetwork(g, ts, np.complex128)


def test_config2_grid32_chi8_converges_and_observables_are_consistent():
    g = E.named_grid((32, 32))
    psi = torch_random_psi(g, 8)
    ctx = E.Context(0)
    info = {}
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=60, tol=1e-14, edge_sequence=E.parallel_edge_sequence(g),
                   info=info)
    assert info["iterations"] < 60 and info["mean_diff"] <= 1e-14
    rng = np.random.default_rng(0)
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
    ez = E.expect(bpc, "Z")
    vals = np.array([ez[v] for v in range(g.nv)])
    assert np.all(np.abs(vals.imag) < 1e-8) and np.all(np.abs(vals.real) <= 1 + 1e-10)
    rho = E.rdm2(bpc, [0, 100, 1000])
    for m in rho:
        assert abs(np.trace(m) - 1) < 1e-10 and np.min(np.linalg.eigvalsh((m + m.conj().T) / 2)) > -1e-8


def test_config4_grid64_chi16_dmma_sweep_equals_generic_updates():
    g = E.named_grid((64, 64))
    psi = torch_random_psi(g, 16)
    ctx = E.Context(0)
    seq = E.parallel_edge_sequence(g)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    del psi
    E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
    rng = np.random.default_rng(1)
    sample = [g.edges[i] for i in rng.choice(g.ne, 24, replace=False)]
    sample = sample + [(v, u) for u, v in sample[:8]]
    ctx.set_path(1)  # generic kernels for the single-message updates
    expected = {e: E.updated_message(bpc, e) for e in sample}
    ctx.set_path(0)
    info = {}
    E.update(bpc, maxiter=1, tol=0.0, edge_sequence=seq, inplace=True, info=info)  # one more DMMA sweep
    for e, m in expected.items():
        got = bpc.message(e)
        assert np.linalg.norm(got - m) < 1e-11 * np.linalg.norm(m), e
    assert 0.0 <= info["mean_diff"] < 1.0
    tm = bpc.last_timing()
    assert tm["contract_ms"] < 200.0  # the DMMA path ran (the generic kernels need ~350 ms per sweep)


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
