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
