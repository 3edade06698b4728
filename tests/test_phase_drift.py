"""Phase drift of BP messages on loopy graphs, and the engine's stabilisation.

update(::Algorithm"bp") (src/caches/abstractbeliefpropagationcache.jl:225-239, 272-329) normalises every new message by
its Frobenius norm only.  The update is multilinear in the z - 1 incoming messages, so a complex phase e^{i phi_e} on
message e propagates as phi_out = sum of phi_in: rounding-level phases grow by a factor ~ (z - 1) per sweep and reach
O(1) after a few dozen sweeps.  Scalars, expectation values and logZ (mod 2 pi i) do not notice, but the messages stop
being Hermitian, and `map_eigvals(...; ishermitian = true)` of the simple update then sees cos(phi) M: a BP-gauged TEBD
loop degrades after a few steps (measured on the engine before the fix: tools/dbg_tebd_cond.py, non-Hermiticity
5e-13 -> 4e-10 -> 4e-7 -> 5e-4 -> 0.6 over four Trotter steps of a 12 x 12 lattice).  The engine stores the Hermitian part
of every new message of a norm network (k_commit); these tests pin (a) that the drift is a property of the algorithm as
the reference states it (NumPy restatement, CPU) and (b) that the engine does not have it and still agrees with the
faithful restatement to 1e-10 while the latter's phases are small (GPU)."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O
from util import make_pair, rel_err


def non_hermiticity(msgs):
    return max(np.linalg.norm(m - m.conj().T) / np.linalg.norm(m) for m in msgs.values())


def test_reference_algorithm_drifts_and_the_hermitian_part_does_not():
    g = O.grid_graph((4, 4))
    net = O.random_network(g, 3, dtype=np.complex128, seed=5)
    seq = O.parallel_edge_sequence(g)
    grp = O.synchronous_groups(seq)
    msgs = O.identity_messages(net)
    stab = O.identity_messages(net)
    nh = []
    for _ in range(6):
        msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=grp, maxiter=10)
        stab, _, _ = O.bp_update(net, stab, seq=seq, groups=grp, maxiter=10, hermitize=True)
        nh.append(non_hermiticity(msgs))
    # exponential growth from rounding level to O(1) within 60 sweeps ...
    assert nh[0] < 1e-12 and nh[-1] > 0.1
    assert all(b > 50 * a for a, b in zip(nh[:4], nh[1:5]))
    # ... that gauge-invariant quantities do not see: logZ agrees up to a multiple of 2 pi i
    z0, z1 = O.logscalar(net, msgs), O.logscalar(net, stab)
    assert abs(z0.real - z1.real) < 1e-9 * abs(z1.real)
    k = (z0.imag - z1.imag) / (2 * np.pi)
    assert abs(k - round(k)) < 1e-6
    assert non_hermiticity(stab) < 1e-13
    # while the phases are still small the two iterations agree to rounding
    a, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=grp, maxiter=15)
    b, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=grp, maxiter=15, hermitize=True)
    assert max(rel_err(a[k_], b[k_]) for k_ in a) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("dims,chi", [((4, 4), 3), ((6, 6), 16), ((3, 3, 3), 2)])
def test_engine_messages_stay_hermitian_over_many_sweeps(dims, chi):
    ctx = E.Context(0)
    g = O.grid_graph(dims)
    net, psi = make_pair(g, chi, np.complex128, seed=5)
    seq = O.parallel_edge_sequence(g)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    # short run: parity with the faithful restatement (its phases are still at rounding level)
    ref, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=8)
    E.update(bpc, maxiter=8, edge_sequence=[[e] for e in seq], inplace=True)
    assert max(rel_err(bpc.message(k), m) for k, m in ref.items()) < 1e-10
    # long run: Hermitian to rounding, real logZ, and equal to the stabilised restatement
    E.update(bpc, maxiter=92, edge_sequence=[[e] for e in seq], inplace=True)
    got = {k: np.asarray(bpc.message(k)) for k in ref}
    assert non_hermiticity(got) < 1e-13
    z = E.logscalar(bpc)
    assert abs(z.imag) < 1e-9
    if np.prod(dims) <= 16:
        stab, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=100, hermitize=True)
        assert max(rel_err(got[k], m) for k, m in stab.items()) < 1e-10
        assert abs(z - O.logscalar(net, stab)) < 1e-9 * abs(z)


@pytest.mark.gpu
def test_bp_gauged_tebd_stays_well_conditioned():
    """Eight BP-gauged Trotter steps (unitary gate, chi 64 -> 16 truncation, two sweeps per colour layer) on a 10 x 10
    lattice: messages stay Hermitian and positive, the bond spectra stay where they started (before the fix the fifth
    step had sigma_16 / sigma_1 = 4e-20 and 503 of 528 messages with a non-positive eigenvalue)."""
    ctx = E.Context(0)
    g = O.grid_graph((10, 10))
    net, psi = make_pair(g, 16, np.complex128, seed=3)
    ng = E.NamedGraph(g.nv, g.edges)
    seq = E.parallel_edge_sequence(ng)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    E.update(bpc, maxiter=20, edge_sequence=seq, inplace=True)
    rng = np.random.default_rng(7)
    m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    w, v = np.linalg.eigh((m + m.conj().T) / 2)
    gate = ((v * np.exp(-0.05j * w)) @ v.conj().T).reshape(2, 2, 2, 2)
    lay = [([gate] * len(layer), [g.edges[e] for e in layer]) for layer in E.edge_coloring(ng)]
    first = None
    for step in range(8):
        info = E.tebd_step(bpc, lay, maxdim=16, cutoff=None, msg_mode=1, bp_maxiter=2, edge_sequence=seq)
        ratio = min(float(s[-1] / s[0]) for s in info["singular_values"])
        first = ratio if first is None else first
        assert ratio > 0.2 * first, (step, ratio, first)
    worst_h, worst_ev = 0.0, 1.0
    for (a, b) in g.edges:
        for e in ((a, b), (b, a)):
            mm = np.asarray(bpc.message(e))
            worst_h = max(worst_h, np.linalg.norm(mm - mm.conj().T) / np.linalg.norm(mm))
            ev = np.linalg.eigvalsh((mm + mm.conj().T) / 2)
            worst_ev = min(worst_ev, ev[0] / ev[-1])
    assert worst_h < 1e-12 and worst_ev > 1e-3
