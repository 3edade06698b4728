"""End-to-end gauged simple-update evolution (BASELINE config 3 shape at small bond dimension): kicked Ising on
the 127-qubit heavy-hex graph, one-site kicks + two-site ZZ gates applied colour layer by colour layer with BP
environments, BP re-run between layers.  The engine and the oracle run the same protocol; compared are the
gauge-invariant quantities: bond dimensions, truncation errors, <Z> on every qubit."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O

pytestmark = pytest.mark.gpu


def rx(theta):
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def rzz(theta):
    d = np.exp(-0.5j * theta * np.array([1, -1, -1, 1]))
    return np.diag(d).astype(np.complex128).reshape(2, 2, 2, 2)


def test_kicked_ising_heavy_hex():
    g = O.heavy_hex_eagle_graph()
    eg = E.heavy_hex_eagle()
    assert eg.edges == g.edges
    rng = np.random.default_rng(3)
    tensors = []
    for v in range(g.nv):  # random product state, bond dimension 1
        a = rng.standard_normal(2) + 1j * rng.standard_normal(2)
        tensors.append((a / np.linalg.norm(a)).reshape((2,) + (1,) * g.degree(v)))
    net = O.Network(g, [t.copy() for t in tensors], np.complex128)
    psi = E.ITensorNetwork(eg, [t.copy() for t in tensors], np.complex128)
    ctx = E.Context(0)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    msgs = O.identity_messages(net)
    seq = O.parallel_edge_sequence(g)
    sync = [[e] for e in seq]
    layers = O.edge_coloring(g)
    kick, zz = rx(0.7), rzz(-1.1)
    maxdim, cutoff, sweeps = 4, 1e-10, 6
    for step in range(2):
        for v in range(g.nv):
            net = O.apply1(net, v, kick)
        for v in range(g.nv):
            E.apply(kick, bpc, (v,), inplace=True)
        for layer in layers:
            msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=O.synchronous_groups(seq), maxiter=sweeps)
            E.update(bpc, maxiter=sweeps, edge_sequence=sync, inplace=True)
            info = E.apply_layer([zz] * len(layer), bpc, [g.edges[e] for e in layer], maxdim=maxdim, cutoff=cutoff)
            for i, e in enumerate(layer):
                net, inf = O.simple_update_bp(net, msgs, e, zz, maxdim=maxdim, cutoff=cutoff)
                msgs = O.reset_edge_messages(net, msgs, e)
                assert info["newdim"][i] == inf["newdim"], (step, e)
                assert abs(info["truncation_error"][i] - inf["truncerr"]) < 1e-9
    msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=O.synchronous_groups(seq), maxiter=10)
    E.update(bpc, maxiter=10, edge_sequence=sync, inplace=True)
    ez = E.expect(bpc, "Z")
    worst = max(abs(ez[v] - O.expect1(net, msgs, v, O.PAULI_Z)) for v in range(g.nv))
    assert worst < 1e-8, worst
    assert max(bpc.edge_dim(e) for e in range(g.ne)) == maxdim


def test_tebd_step_single_call_equals_host_loop():
    # itn_apply_layers (include/itn_b200.h): gate layers + BP sweeps in one library call = the same calls made one by one
    g = O.grid_graph((5, 4))
    eg = E.named_grid((5, 4))
    psi = E.random_tensornetwork(5, np.complex128, eg, link_space=3)
    ctx = E.Context(0)
    seq = [[e] for e in O.parallel_edge_sequence(g)]
    a = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=5, edge_sequence=seq)
    b = a.copy()
    gate = rzz(-0.9)
    kick = O.random_unitary(4, seed=5).reshape(2, 2, 2, 2)
    layers = []
    for li, layer in enumerate(O.edge_coloring(g)):
        layers.append(([gate if li % 2 == 0 else kick] * len(layer), [g.edges[e] for e in layer]))
    infos = []
    for gates, pairs in layers:
        infos.append(E.apply_layer(gates, a, pairs, maxdim=4, cutoff=1e-11))
        E.update(a, maxiter=3, edge_sequence=seq, inplace=True)
    info = {}
    res = E.tebd_step(b, layers, maxdim=4, cutoff=1e-11, bp_maxiter=3, edge_sequence=seq, info=info)
    assert info["bp_iterations"] == 3 * len(layers)
    off = 0
    for li, inf in enumerate(infos):
        n = len(inf["newdim"])
        assert np.array_equal(res["newdim"][off:off + n], inf["newdim"])
        assert np.array_equal(res["truncation_error"][off:off + n], inf["truncation_error"])
        for i in range(n):
            assert np.array_equal(res["singular_values"][off + i], inf["singular_values"][i])
        off += n
    for v in range(g.nv):
        assert np.array_equal(a.factor(v), b.factor(v))
    for (u, v) in g.edges:
        assert np.array_equal(a.message((u, v)), b.message((u, v)))


def test_fifty_trotter_steps_track_the_oracle():
    # 50 kicked-Ising steps on a 3 x 3 grid (200 gate layers, 600 truncations with cutoff = 1e-10 and maxdim = 4),
    # BP re-run before every layer: the engine's Gram / Cholesky route (CholeskyQR2 where the pivots ask for it) and the
    # oracle's QR route (src/apply.jl:70-93) must stay together gate by gate -- kept dimension equal, singular values and
    # truncation error to 1e-10 -- and end with the same <Z>.  Both sides keep the Hermitian part of the messages
    # (without it the reference's own phases drift after ~40 sweeps).
    g = O.grid_graph((3, 3))
    rng = np.random.default_rng(11)
    tensors = []
    for v in range(g.nv):
        a = rng.standard_normal(2) + 1j * rng.standard_normal(2)
        tensors.append((a / np.linalg.norm(a)).reshape((2,) + (1,) * g.degree(v)))
    net = O.Network(g, [t.copy() for t in tensors], np.complex128)
    psi = E.ITensorNetwork(E.NamedGraph(g.nv, g.edges), [t.copy() for t in tensors], np.complex128)
    ctx = E.Context(0)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    msgs = O.identity_messages(net)
    seq = O.parallel_edge_sequence(g)
    sync = [[e] for e in seq]
    layers = O.edge_coloring(g)
    kick, zz = rx(0.4), rzz(-0.6)
    maxdim, cutoff, sweeps = 4, 1e-10, 5
    worst_sv = worst_te = 0.0
    for step in range(50):
        for v in range(g.nv):
            net = O.apply1(net, v, kick)
            E.apply(kick, bpc, (v,), inplace=True)
        for layer in layers:
            msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=O.synchronous_groups(seq), maxiter=sweeps, hermitize=True)
            E.update(bpc, maxiter=sweeps, edge_sequence=sync, inplace=True)
            info = E.apply_layer([zz] * len(layer), bpc, [g.edges[e] for e in layer], maxdim=maxdim, cutoff=cutoff,
                                 normalize=True)
            for i, e in enumerate(layer):
                net, inf = O.simple_update_bp(net, msgs, e, zz, maxdim=maxdim, cutoff=cutoff, normalize=True)
                msgs = O.reset_edge_messages(net, msgs, e)
                n = inf["newdim"]
                assert info["newdim"][i] == n, (step, e, info["newdim"][i], n)
                sv = np.asarray(info["singular_values"][i])[:n]
                worst_sv = max(worst_sv, float(np.max(np.abs(sv - inf["svals"][:n])) / inf["svals"][0]))
                worst_te = max(worst_te, abs(info["truncation_error"][i] - inf["truncerr"]))
    assert worst_sv < 1e-10 and worst_te < 1e-10, (worst_sv, worst_te)
    msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=O.synchronous_groups(seq), maxiter=10, hermitize=True)
    E.update(bpc, maxiter=10, edge_sequence=sync, inplace=True)
    ez = E.expect(bpc, "Z")
    worst = max(abs(ez[v] - O.expect1(net, msgs, v, O.PAULI_Z)) for v in range(g.nv))
    assert worst < 1e-9, worst
    assert max(bpc.edge_dim(e) for e in range(g.ne)) == maxdim
