"""Analytic known-answer tests: states whose BP quantities have closed forms.  They pin the ORACLE (CPU, `not gpu`) and the
ENGINE (`gpu`) to numbers that come from neither -- the part of "parity unpinned" that can be pinned without a Julia runtime
(the reference holds no golden vectors):

  product states     bond dimension 1 (and zero-padded to chi > 1: rank-one messages) on loopy graphs: BP is exact,
                     <Z_v> = cos(2 theta_v), <Z_u Z_v> = product, log Z = sum_v log |psi_v|^2
  weighted GHZ       a |0..0> + b |1..1> on a tree as delta tensors (chi = 2): <Z_v> = (|a|^2 - |b|^2) / (|a|^2 + |b|^2),
                     <Z_u Z_v> = 1, Z = |a|^2 + |b|^2, two-site RDM = diag(|a|^2, 0, 0, |b|^2) / Z
  Schmidt spectra    exp(-i theta Z(x)Z) on |+>|+> = cos(theta) |++> - i sin(theta) |-->: singular values (cos, sin), so
                     maxdim / cutoff truncation has closed-form kept dimension and truncation error (src/apply.jl:81-88,
                     relative cutoff on squared singular values); CNOT (H (x) 1) |00>: (1, 1) / sqrt 2
  weighted W state   sum_i c_i |0..1_i..0> on a chain as an MPS of bond dimension 2: <Z_i> = 1 - 2 p_i, <Z_i Z_j> = 1 - 2 p_i - 2 p_j,
                     Z = sum |c|^2, the two-site RDM (off-diagonal coherence c_i conj(c_j) / Z), and the Schmidt coefficients
                     sqrt(sum_{i <= k} p_i), sqrt(sum_{i > k} p_i) across every bond (an identity gate returns them)
  cluster state      prod CZ |+>^n on a tree (chi = 2, signs): <Z> = <ZZ> = 0, Z = 1, edge RDMs 1/4 in the interior and
                     (1 + X_leaf Z_neighbour) / 4 at a leaf -- the stabiliser structure
"""
import math

import numpy as np
import pytest

from oracle import itn_oracle as O

Z = np.array([[1.0, 0.0], [0.0, -1.0]])
TOL = 1e-12


def product_network(g, thetas, phis, chi, dtype, scales):
    ts = []
    for v in range(g.nv):
        vec = scales[v] * np.array([math.cos(thetas[v]), np.exp(1j * phis[v]) * math.sin(thetas[v])])
        if np.dtype(dtype).kind != "c":
            vec = vec.real
        t = np.zeros((2,) + (chi,) * g.degree(v), dtype=dtype)
        t[(slice(None),) + (0,) * g.degree(v)] = vec
        ts.append(t)
    return O.Network(g, ts, dtype)


def ghz_network(g, a, b, dtype=np.complex128):
    ts = []
    for v in range(g.nv):
        t = np.zeros((2,) + (2,) * g.degree(v), dtype=dtype)
        t[(0,) * (1 + g.degree(v))] = a if v == 0 else 1.0
        t[(1,) * (1 + g.degree(v))] = b if v == 0 else 1.0
        ts.append(t)
    return O.Network(g, ts, dtype)


def zz_gate(theta):
    zz = np.kron(Z, Z)
    u = np.diag(np.exp(-1j * theta * np.diag(zz)))
    return u.reshape(2, 2, 2, 2)


def plus_chain(n, chi=1):
    g = O.chain_graph(n)
    ts = []
    for v in range(n):
        t = np.zeros((2,) + (chi,) * g.degree(v), dtype=np.complex128)
        t[(slice(None),) + (0,) * g.degree(v)] = 1 / math.sqrt(2)
        ts.append(t)
    return g, O.Network(g, ts, np.complex128)


# ---- the same checks for the oracle and the engine -------------------------------------------------------------------


class OracleSide:
    name = "oracle"

    def bp(self, net, maxiter):
        g = net.graph
        if g.is_tree():
            msgs, _, _ = O.bp_update(net, {}, seq=O.default_edge_sequence(g), maxiter=1)
        else:
            seq = O.parallel_edge_sequence(g)
            msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=maxiter)
        self.net, self.msgs = net, msgs

    def expect_z(self, v):
        return O.expect1(self.net, self.msgs, v, Z)

    def expect_zz(self, e):
        return O.expect2(self.net, self.msgs, e, Z, Z)

    def rdm2(self, e):
        return O.rdm2(self.net, self.msgs, e)

    def logz(self):
        return O.logscalar(self.net, self.msgs)

    def gate(self, e, gate, maxdim=None, cutoff=None):
        _, info = O.simple_update_bp(self.net, self.msgs, e, gate, maxdim=maxdim, cutoff=cutoff)
        return info["newdim"], info["truncerr"], np.asarray(info["svals"][:info["newdim"]])


class EngineSide:
    name = "engine"

    def __init__(self):
        import itn_b200 as E
        self.E = E
        self.ctx = E.Context(0)

    def bp(self, net, maxiter):
        E = self.E
        g = net.graph
        eg = E.NamedGraph(g.nv, g.edges)
        psi = E.ITensorNetwork(eg, [t.copy() for t in net.tensors], net.dtype)
        if g.is_tree():
            self.bpc = E.update(E.BeliefPropagationCache(psi, ctx=self.ctx, messages="default"))
        else:
            self.bpc = E.update(E.BeliefPropagationCache(psi, ctx=self.ctx), maxiter=maxiter, edge_sequence=E.parallel_edge_sequence(eg))
        self.g = g

    def expect_z(self, v):
        return self.E.expect(self.bpc, "Z", vertices=[v])[v]

    def expect_zz(self, e):
        return self.E.expect2(self.bpc, [e], "Z", "Z")[0]

    def rdm2(self, e):
        return self.E.rdm2(self.bpc, [e])[0]

    def logz(self):
        return self.E.logscalar(self.bpc)

    def gate(self, e, gate, maxdim=None, cutoff=None):
        got = {}
        self.E.apply(gate, self.bpc, self.g.edges[e], maxdim=maxdim, cutoff=cutoff, callback=lambda **kw: got.update(kw))
        sv = np.asarray(got["singular_values"])
        return len(sv), got["truncation_error"], sv


SIDES = [pytest.param(OracleSide, id="oracle"), pytest.param(EngineSide, id="engine", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("side", SIDES)
@pytest.mark.parametrize("dtype", [np.float64, np.complex128], ids=["f64", "c128"])
@pytest.mark.parametrize("chi", [1, 3])
def test_product_state_on_a_loopy_grid(side, dtype, chi):
    s = side()
    g = O.grid_graph((3, 4))
    rng = np.random.default_rng(3)
    th, ph, sc = rng.uniform(0, math.pi, g.nv), rng.uniform(0, 2 * math.pi, g.nv), rng.uniform(0.5, 2.0, g.nv)
    if np.dtype(dtype).kind != "c":
        ph = np.zeros(g.nv)  # real amplitudes
    s.bp(product_network(g, th, ph, chi, dtype, sc), maxiter=3)
    for v in range(g.nv):
        assert abs(s.expect_z(v) - math.cos(2 * th[v])) < TOL
    for e in (0, 5, len(g.edges) - 1):
        u, v = g.edges[e]
        assert abs(s.expect_zz(e) - math.cos(2 * th[u]) * math.cos(2 * th[v])) < TOL
    assert abs(s.logz() - float(np.sum(np.log(sc ** 2)))) < 1e-11


@pytest.mark.parametrize("side", SIDES)
def test_weighted_ghz_on_a_tree(side):
    s = side()
    g = O.random_tree_graph(8, seed=4)
    a, b = 0.8 * np.exp(0.3j), 0.35 * np.exp(-1.1j)
    s.bp(ghz_network(g, a, b), maxiter=1)
    zn = abs(a) ** 2 + abs(b) ** 2
    want = (abs(a) ** 2 - abs(b) ** 2) / zn
    for v in range(g.nv):
        assert abs(s.expect_z(v) - want) < TOL
    for e in range(len(g.edges)):
        assert abs(s.expect_zz(e) - 1.0) < TOL
    assert abs(np.exp(s.logz()) - zn) < TOL * zn
    rho = s.rdm2(2)
    assert np.allclose(rho, np.diag([abs(a) ** 2, 0, 0, abs(b) ** 2]) / zn, atol=TOL)


@pytest.mark.parametrize("side", SIDES)
@pytest.mark.parametrize("chi", [1, 2])
def test_schmidt_spectrum_of_a_zz_rotation(side, chi):
    theta = 0.1
    c2, s2 = math.cos(theta) ** 2, math.sin(theta) ** 2
    for n, e in ((2, 0), (4, 1)):
        # no truncation: singular values (cos, sin)
        s = side()
        g, net = plus_chain(n, chi)
        s.bp(net, 1)
        dim, terr, sv = s.gate(e, zz_gate(theta))
        assert dim == 2 * chi or dim == 2  # the reference keeps every singular value without a cutoff (zeros included)
        assert np.allclose(sorted(sv, reverse=True)[:2], [math.cos(theta), math.sin(theta)], atol=TOL) and abs(terr) < TOL
        # maxdim = 1: the discarded weight is sin^2 / (cos^2 + sin^2)
        s = side()
        s.bp(plus_chain(n, chi)[1], 1)
        dim, terr, sv = s.gate(e, zz_gate(theta), maxdim=1)
        assert dim == 1 and abs(terr - s2) < TOL and abs(sv[0] - math.sqrt(c2)) < TOL
        # relative cutoff above sin^2: truncates; below: keeps both
        s = side()
        s.bp(plus_chain(n, chi)[1], 1)
        dim, terr, sv = s.gate(e, zz_gate(theta), cutoff=2 * s2)
        assert dim == 1 and abs(terr - s2) < TOL
        s = side()
        s.bp(plus_chain(n, chi)[1], 1)
        dim, terr, sv = s.gate(e, zz_gate(theta), cutoff=0.5 * s2)
        assert dim == 2 and abs(terr) < TOL and np.allclose(sv, [math.cos(theta), math.sin(theta)], atol=TOL)


@pytest.mark.parametrize("side", SIDES)
def test_bell_pair_from_cnot_hadamard(side):
    h = np.array([[1, 1], [1, -1]]) / math.sqrt(2)
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=float)  # control = first site
    u = cnot @ np.kron(h, np.eye(2))                      # row index (s1, s2) with s1 slow
    gate = u.reshape(2, 2, 2, 2).astype(np.complex128)    # g[s1', s2', s1, s2]
    g = O.chain_graph(2)
    ts = [np.zeros((2, 1), dtype=np.complex128) for _ in range(2)]
    ts[0][0, 0] = ts[1][0, 0] = 1.0
    s = side()
    s.bp(O.Network(g, ts, np.complex128), 1)
    dim, terr, sv = s.gate(0, gate)
    assert dim == 2 and np.allclose(sv, [1 / math.sqrt(2)] * 2, atol=TOL) and abs(terr) < TOL
    s = side()
    s.bp(O.Network(g, ts, np.complex128), 1)
    dim, terr, sv = s.gate(0, gate, maxdim=1)
    assert dim == 1 and abs(terr - 0.5) < TOL


def w_chain(c):
    """sum_i c_i |0..1_i..0> on an open chain as an MPS of bond dimension 2 (state "no excitation yet" / "one passed"):
    A^0 = 1, A^1 = c_i |0><1|, boundary vectors <0| and |1>."""
    n = len(c)
    g = O.chain_graph(n)
    a0 = np.eye(2, dtype=np.complex128)
    ts = []
    for v in range(n):
        a1 = np.zeros((2, 2), dtype=np.complex128)
        a1[0, 1] = c[v]
        a = np.stack([a0, a1])                     # [s, left, right]
        if v == 0:
            t = a[:, 0, :]                         # bond to vertex 1 only
        elif v == n - 1:
            t = a[:, :, 1]
        else:
            t = a                                  # inc[v] = [edge to v - 1, edge to v + 1]
        ts.append(np.ascontiguousarray(t))
    return g, O.Network(g, ts, np.complex128)


@pytest.mark.parametrize("side", SIDES)
def test_weighted_w_state_on_a_chain(side):
    # BP is exact on a chain.  With p_i = |c_i|^2 / sum |c|^2:  <Z_i> = 1 - 2 p_i,  <Z_i Z_j> = 1 - 2 p_i - 2 p_j,
    # Z = sum |c|^2, the two-site RDM has the block (c_i, c_j)^H (c_i, c_j) / Z on {|10>, |01>}, and across the bond
    # (k, k + 1) the Schmidt coefficients are sqrt(sum_{i <= k} p_i), sqrt(sum_{i > k} p_i): an identity gate returns them.
    rng = np.random.default_rng(21)
    n = 7
    c = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    g, net = w_chain(c)
    p = np.abs(c) ** 2 / np.sum(np.abs(c) ** 2)
    s = side()
    s.bp(net, 1)
    for v in range(n):
        assert abs(s.expect_z(v) - (1 - 2 * p[v])) < TOL
    for e in range(n - 1):
        assert abs(s.expect_zz(e) - (1 - 2 * p[e] - 2 * p[e + 1])) < TOL
    assert abs(np.exp(s.logz()) - np.sum(np.abs(c) ** 2)) < 1e-11 * np.sum(np.abs(c) ** 2)
    e = 3
    rho = np.asarray(s.rdm2(e)).reshape(4, 4)
    rho = rho / np.trace(rho)
    want = np.zeros((4, 4), dtype=np.complex128)   # index s_e + 2 s_{e+1} (first site fastest, as rdm2 returns it)
    want[0, 0] = 1 - p[e] - p[e + 1]
    amp = np.array([c[e], c[e + 1]]) / np.sqrt(np.sum(np.abs(c) ** 2))   # index 1 = |1_e 0_{e+1}> carries c_e, index 2 c_{e+1}
    want[1:3, 1:3] = np.outer(amp, amp.conj())
    assert np.allclose(rho, want, atol=1e-11)
    ident = np.eye(4, dtype=np.complex128).reshape(2, 2, 2, 2)
    for k in (0, 2, 5):
        s = side()
        s.bp(w_chain(c)[1], 1)
        dim, terr, sv = s.gate(k, ident)
        sv = np.sort(np.asarray(sv))[::-1]
        wl, wr = np.sum(p[:k + 1]), np.sum(p[k + 1:])
        lo, hi = sorted([wl, wr])
        assert abs(sv[1] / sv[0] - math.sqrt(lo / hi)) < 1e-11 and abs(terr) < TOL
        assert np.all(sv[2:] < 1e-7 * sv[0])   # the gate bond carries two Schmidt states, whatever dimension is kept


def cluster_network(g):
    """Graph (cluster) state prod_{(u,v)} CZ_uv |+>^n as a tensor network of bond dimension 2: on every edge the lower
    vertex copies its spin onto the bond, the higher one applies (-1)^(s a):  T_v[s, a..] = 2^-1/2 prod_out delta(a, s)
    prod_in (-1)^(s a)."""
    ts = []
    for v in range(g.nv):
        t = np.zeros((2,) + (2,) * g.degree(v), dtype=np.complex128)
        for idx in np.ndindex(*t.shape):
            s, amp = idx[0], 1 / math.sqrt(2)
            for k, e in enumerate(g.inc[v]):
                a = idx[1 + k]
                if min(g.edges[e]) == v:
                    amp *= 1.0 if a == s else 0.0
                else:
                    amp *= -1.0 if (s and a) else 1.0
            t[idx] = amp
        ts.append(t)
    return O.Network(g, ts, np.complex128)


@pytest.mark.parametrize("side", SIDES)
def test_cluster_state_on_a_tree(side):
    # Stabilisers K_v = X_v prod_{w ~ v} Z_w.  On a tree BP is exact: <Z_v> = 0 and <Z_u Z_v> = 0 everywhere, the state is
    # normalised (Z = 1), the RDM of an edge whose endpoints both have further neighbours is 1/4, and the RDM of a leaf l
    # with its neighbour v is (1 + X_l Z_v) / 4 (the only stabiliser product supported on the pair).
    g = O.comb_tree_graph(3, 3)
    s = side()
    s.bp(cluster_network(g), 1)
    for v in range(g.nv):
        assert abs(s.expect_z(v)) < TOL
    for e in range(len(g.edges)):
        assert abs(s.expect_zz(e)) < TOL
    assert abs(s.logz()) < 1e-11
    x = np.array([[0.0, 1.0], [1.0, 0.0]])
    for e, (u, v) in enumerate(g.edges):
        rho = np.asarray(s.rdm2(e)).reshape(4, 4)       # index s_u + 2 s_v
        du, dv = g.degree(u), g.degree(v)
        if du > 1 and dv > 1:
            want = np.eye(4) / 4
        elif du == 1 and dv > 1:
            want = (np.eye(4) + np.kron(Z, x)) / 4      # kron(op_v, op_u): s_u is the fast index
        elif dv == 1 and du > 1:
            want = (np.eye(4) + np.kron(x, Z)) / 4
        else:
            continue
        assert np.allclose(rho, want, atol=1e-11), (e, u, v)
