"""CPU oracle for the ITensorNetworks.jl belief-propagation / simple-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product path (``libitn_b200.so``) never does.

PARITY UNPINNED against numeric outputs of the Julia reference: the reference is pure
Julia whose arithmetic lives in un-vendored packages (ITensors/NDTensors, NamedGraphs,
TensorOperations; ``Project.toml:9-62``, no Manifest), there is no Julia runtime in this
image, and the reference's tests hold no golden vectors.  What *is*
pinned: every property the reference's own tests assert for this path
(``tests/test_oracle.py`` restates ``test/test_belief_propagation.jl:18-99``,
``test/test_expect.jl:12-39``, ``test/test_normalize.jl:16-66``,
``test/test_map_eigvals.jl:7-34``, ``test/test_apply.jl:12-66``) plus brute-force exact
contraction on small networks.

This is a dense NumPy restatement (float64 / complex128) of
  src/caches/abstractbeliefpropagationcache.jl:32-36,99-107,225-239,272-329,349-408
  src/caches/beliefpropagationcache.jl:100-139
  src/formnetworks/quadraticformnetwork.jl:96-124, src/initialize_cache.jl:14-29
  src/edge_sequences.jl:32-51
  src/expect.jl:5-19, src/normalize.jl:13-34,63-80, src/apply.jl:9-146
  src/inner.jl:100-171, src/formnetworks/bilinearformnetwork.jl:23-94 (bilinear forms <phi|psi>, <phi|A|psi>)
  test/utils.jl:23-38 (input generator)
with the index conventions of SURVEY.md Appendix A:
  site tensor  A_v[s, a_1..a_z]   (s = physical index, a_k = bond to the k-th incident edge)
  message      M_{u->v}[a, a']    (a = ket-side bond index, a' = bra-side copy)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------
# graphs
# ----------------------------------------------------------------------------------------


@dataclass
class Graph:
    """Undirected simple graph on vertices 0..nv-1 with a fixed edge list (u < v not required)."""

    nv: int
    edges: list  # list[(u, v)]
    coords: list | None = None

    def __post_init__(self):
        self.edges = [(int(u), int(v)) for u, v in self.edges]
        self.inc = [[] for _ in range(self.nv)]  # incident edge ids, ascending
        for e, (u, v) in enumerate(self.edges):
            self.inc[u].append(e)
            self.inc[v].append(e)
        self.eid = {}
        for e, (u, v) in enumerate(self.edges):
            self.eid[(u, v)] = e
            self.eid[(v, u)] = e

    @property
    def ne(self):
        return len(self.edges)

    def other(self, e, v):
        u, w = self.edges[e]
        return w if u == v else u

    def neighbors(self, v):
        return [self.other(e, v) for e in self.inc[v]]

    def degree(self, v):
        return len(self.inc[v])

    def slot(self, v, e):
        """Position of edge e among v's bond axes (tensor axis = 1 + slot)."""
        return self.inc[v].index(e)

    def is_tree(self):
        return self.ne == self.nv - 1 and len(self.components()) == 1

    def is_forest(self):
        return self.ne == self.nv - len(self.components())

    def components(self):
        seen = [False] * self.nv
        comps = []
        for r in range(self.nv):
            if seen[r]:
                continue
            stack, comp = [r], []
            seen[r] = True
            while stack:
                x = stack.pop()
                comp.append(x)
                for y in self.neighbors(x):
                    if not seen[y]:
                        seen[y] = True
                        stack.append(y)
            comps.append(sorted(comp))
        return comps


def grid_graph(dims):
    """named_grid(dims) analogue; vertices in row-major order (last coordinate fastest)."""
    dims = tuple(int(x) for x in dims)
    nv = int(np.prod(dims))
    idx = np.arange(nv).reshape(dims)
    edges = []
    coords = [tuple(int(c) for c in np.unravel_index(i, dims)) for i in range(nv)]
    for i in range(nv):
        c = coords[i]
        for ax in range(len(dims)):
            if c[ax] + 1 < dims[ax]:
                c2 = list(c)
                c2[ax] += 1
                edges.append((i, int(idx[tuple(c2)])))
    return Graph(nv, edges, coords)


def chain_graph(n):
    return Graph(n, [(i, i + 1) for i in range(n - 1)])


def comb_tree_graph(nx, ny):
    """named_comb_tree((nx, ny)): a backbone of nx vertices, each with a tooth of ny vertices."""
    idx = lambda i, j: i * ny + j
    edges = [(idx(i, 0), idx(i + 1, 0)) for i in range(nx - 1)]
    for i in range(nx):
        edges += [(idx(i, j), idx(i, j + 1)) for j in range(ny - 1)]
    return Graph(nx * ny, edges)


def random_tree_graph(n, seed=0):
    rng = np.random.default_rng(seed)
    return Graph(n, [(int(rng.integers(0, i)), i) for i in range(1, n)])


def heavy_hex_eagle_graph():
    """127-qubit heavy-hex (IBM Eagle) coupling graph: 7 rows joined by 4 bridge qubits each; 144 edges."""
    rows = [list(range(0, 14)), list(range(18, 33)), list(range(37, 52)), list(range(56, 71)),
            list(range(75, 90)), list(range(94, 109)), list(range(113, 127))]
    edges = []
    for r in rows:
        edges += [(r[i], r[i + 1]) for i in range(len(r) - 1)]
    bridges = {14: (0, 18), 15: (4, 22), 16: (8, 26), 17: (12, 30),
               33: (20, 39), 34: (24, 43), 35: (28, 47), 36: (32, 51),
               52: (37, 56), 53: (41, 60), 54: (45, 64), 55: (49, 68),
               71: (58, 77), 72: (62, 81), 73: (66, 85), 74: (70, 89),
               90: (75, 94), 91: (79, 98), 92: (83, 102), 93: (87, 106),
               109: (96, 114), 110: (100, 118), 111: (104, 122), 112: (108, 126)}
    for b, (u, w) in bridges.items():
        edges += [(u, b), (b, w)]
    g = Graph(127, edges)
    assert g.ne == 144
    return g


def edge_coloring(g: Graph):
    """Greedy proper edge colouring: list of lists of edge ids, each list vertex-disjoint."""
    colors = []
    used = [set() for _ in range(g.nv)]
    col_of = {}
    for e, (u, v) in enumerate(g.edges):
        c = 0
        while c in used[u] or c in used[v]:
            c += 1
        used[u].add(c)
        used[v].add(c)
        col_of[e] = c
        while len(colors) <= c:
            colors.append([])
        colors[c].append(e)
    return colors


# ----------------------------------------------------------------------------------------
# edge schedules  (src/edge_sequences.jl:32-51)
# ----------------------------------------------------------------------------------------


def forest_cover(g: Graph):
    """Cover the edge set with spanning forests, greedily (NamedGraphs.forest_cover analogue).

    The exact edge order of the un-vendored NamedGraphs routine is unpinned; the
    host always passes the explicit sequence to the engine, so engine and oracle agree.
    """
    remaining = set(range(g.ne))
    forests = []
    while remaining:
        seen = [False] * g.nv
        forest = []
        for r in range(g.nv):
            if seen[r]:
                continue
            seen[r] = True
            queue = [r]
            while queue:
                x = queue.pop(0)
                for e in g.inc[x]:
                    if e not in remaining:
                        continue
                    y = g.other(e, x)
                    if not seen[y]:
                        seen[y] = True
                        forest.append(e)
                        queue.append(y)
        remaining -= set(forest)
        forests.append(forest)
    return forests


def _post_order_dfs_edges(nv, tree_edges, edges, root):
    adj = {}
    for e in tree_edges:
        u, v = edges[e]
        adj.setdefault(u, []).append(v)
        adj.setdefault(v, []).append(u)
    out = []

    def rec(x, parent):
        for y in adj.get(x, []):
            if y != parent:
                rec(y, x)
                out.append((y, x))  # child -> parent, emitted after the child's subtree

    import sys
    sys.setrecursionlimit(max(10000, 4 * nv))
    rec(root, -1)
    return out


def default_edge_sequence(g: Graph):
    """Forest-cover schedule: every directed edge exactly once (src/edge_sequences.jl:32-47)."""
    seq = []
    for forest in forest_cover(g):
        # connected components of the forest
        verts = sorted({x for e in forest for x in g.edges[e]})
        sub = Graph(g.nv, [g.edges[e] for e in forest])
        for comp in sub.components():
            if len(comp) < 2:
                continue
            comp_set = set(comp)
            tree_edges = [e for e in forest if g.edges[e][0] in comp_set]
            te = _post_order_dfs_edges(g.nv, tree_edges, g.edges, comp[0])
            seq += te + [(b, a) for (a, b) in reversed(te)]
    return seq


def parallel_edge_sequence(g: Graph):
    """src/edge_sequences.jl:49-51: all edges then all reversed edges, one group each."""
    return list(g.edges) + [(v, u) for (u, v) in g.edges]


# ----------------------------------------------------------------------------------------
# network container + generator  (test/utils.jl:23-38)
# ----------------------------------------------------------------------------------------


@dataclass
class Network:
    graph: Graph
    tensors: list  # tensors[v].shape == (d_v, chi_{inc[v][0]}, chi_{inc[v][1]}, ...)
    dtype: type = np.complex128
    # bra layer of a BilinearFormNetwork <phi|psi> (formnetworks/bilinearformnetwork.jl:23-42): bra[v] = phi_v as given
    # (conjugated on use, `dag`); None = QuadraticFormNetwork, bra = ket
    bra: list | None = None

    def copy(self):
        return Network(self.graph, [t.copy() for t in self.tensors], self.dtype,
                       None if self.bra is None else [None if t is None else t.copy() for t in self.bra])

    def bra_tensor(self, v):
        if self.bra is None or self.bra[v] is None:
            return self.tensors[v]
        return self.bra[v]

    def edge_dim(self, e):
        u, _ = self.graph.edges[e]
        return self.tensors[u].shape[1 + self.graph.slot(u, e)]


def random_network(g: Graph, chi, d=2, dtype=np.complex128, seed=1234):
    """iid N(0,1) (real) / CN(0,1) (complex, Julia randn(ComplexF64) convention) site tensors."""
    rng = np.random.default_rng(seed)
    chis = chi if isinstance(chi, (list, tuple, np.ndarray)) else [chi] * g.ne
    tensors = []
    for v in range(g.nv):
        shape = (d,) + tuple(int(chis[e]) for e in g.inc[v])
        if np.dtype(dtype).kind == "c":
            t = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / math.sqrt(2.0)
        else:
            t = rng.standard_normal(shape)
        tensors.append(np.ascontiguousarray(t.astype(dtype)))
    return Network(g, tensors, dtype)


def identity_messages(net: Network):
    """identity_messages (quadraticformnetwork.jl:96-124): delta on both directions, un-normalised."""
    msgs = {}
    for e, (u, v) in enumerate(net.graph.edges):
        chi = net.edge_dim(e)
        msgs[(u, v)] = np.eye(chi, dtype=net.dtype)
        msgs[(v, u)] = np.eye(chi, dtype=net.dtype)
    return msgs


# ----------------------------------------------------------------------------------------
# message update  (abstractbeliefpropagationcache.jl:225-239)
# ----------------------------------------------------------------------------------------


def _absorb(t, axis, m):
    """out[.., a', ..] = sum_a t[.., a, ..] * m[a, a']  (mode product on `axis`)."""
    return np.moveaxis(np.tensordot(t, m, axes=([axis], [0])), -1, axis)


def absorbed_ket(net, msgs, v, skip_edges=()):
    """Ket tensor of v with every incoming message absorbed except those on `skip_edges`."""
    g = net.graph
    b = net.tensors[v]
    for k, e in enumerate(g.inc[v]):
        if e in skip_edges:
            continue
        u = g.other(e, v)
        b = _absorb(b, 1 + k, msgs[(u, v)])
    return b


def updated_message(net, msgs, v, w, normalize=True, hermitize=False):
    """M_{v->w}[l, l'] = sum A_v[s,..a..,l] conj(A_v[s,..a'..,l']) prod_{u != w} M_{u->v}[a,a'].

    hermitize (NOT in the reference; the engine's stabilisation, csrc/itn_generic.cu k_commit): keep the Hermitian part
    of the new message of a norm network.  The update is multilinear in the incoming messages, so the complex phases of
    the messages add up along the graph and grow like (z - 1)^sweeps from rounding level; the Frobenius normalisation of
    the reference does not remove them (tests/test_phase_drift.py measures it on this restatement)."""
    g = net.graph
    e = g.eid[(v, w)]
    k = g.slot(v, e)
    a = net.bra_tensor(v)
    b = absorbed_ket(net, msgs, v, skip_edges=(e,))
    axes = [i for i in range(a.ndim) if i != 1 + k]
    m = np.tensordot(b, a.conj(), axes=(axes, axes))
    if hermitize and net.bra is None:
        m = 0.5 * (m + m.conj().T)
    if normalize:
        n = np.linalg.norm(m)
        if n != 0:
            m = m / n
    return m


def updated_message_local(a, incoming, k, normalize=True):
    """The same update (abstractbeliefpropagationcache.jl:225-239) stated on one vertex alone: `a` = site tensor
    [s, a_1..a_z], `incoming[j]` = message into the vertex on bond slot j (slot k is ignored), result = message leaving
    on slot k.  Used where only a sample of a large network is brought to the host (bench.py, full-size tests)."""
    b = a
    for j in range(a.ndim - 1):
        if j != k:
            b = _absorb(b, 1 + j, incoming[j])
    axes = [i for i in range(a.ndim) if i != 1 + k]
    m = np.tensordot(b, a.conj(), axes=(axes, axes))
    if normalize:
        n = np.linalg.norm(m)
        if n != 0:
            m = m / n
    return m


def message_diff(a, b):
    """1 - |<a^, b^>|^2, first argument conjugated (abstractbeliefpropagationcache.jl:32-36)."""
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    f = abs(np.vdot(a / na, b / nb)) ** 2
    return 1.0 - f


def bp_update(net, msgs, seq=None, groups=None, maxiter=1, tol=None, normalize=True,
              return_history=False, hermitize=False):
    """update(::Algorithm"bp") (abstractbeliefpropagationcache.jl:272-329).

    seq    : list of directed edges (v, w)
    groups : None -> sequential Gauss-Seidel over seq (:272-287); otherwise a list of index
             ranges (start, stop) into seq: every group is computed from the pre-sweep messages
             and all results are written back at the end of the sweep (:294-308, intended
             semantics; within a group updates are sequential on a scratch copy).
    Returns (msgs, iterations_done, last_mean_diff).
    """
    g = net.graph
    if seq is None:
        seq = default_edge_sequence(g)
    msgs = dict(msgs)
    iters = 0
    mean_diff = float("nan")
    history = []
    for _ in range(maxiter):
        diff = 0.0
        if groups is None:
            for (v, w) in seq:
                new = updated_message(net, msgs, v, w, normalize, hermitize)
                if tol is not None:
                    diff += message_diff(new, msgs[(v, w)])
                msgs[(v, w)] = new
        else:
            new_msgs = {}
            for (lo, hi) in groups:
                scratch = dict(msgs)
                for (v, w) in seq[lo:hi]:
                    new = updated_message(net, scratch, v, w, normalize, hermitize)
                    if tol is not None:
                        diff += message_diff(new, scratch[(v, w)])
                    scratch[(v, w)] = new
                    new_msgs[(v, w)] = new
            msgs.update(new_msgs)
        iters += 1
        if return_history:
            history.append({k: m.copy() for k, m in msgs.items()})
        if tol is not None:
            # update divides by length(edge_sequence): the number of groups in the grouped form (:319-321)
            mean_diff = diff / (len(seq) if groups is None else len(groups))
            if mean_diff <= tol:
                break
    if return_history:
        return msgs, iters, mean_diff, history
    return msgs, iters, mean_diff


def synchronous_groups(seq):
    return [(i, i + 1) for i in range(len(seq))]


# ----------------------------------------------------------------------------------------
# region scalars, logscalar  (beliefpropagationcache.jl:107-119, abstract :83-97,:397-412)
# ----------------------------------------------------------------------------------------


def vertex_scalar(net, msgs, v):
    b = absorbed_ket(net, msgs, v)
    return np.vdot(net.bra_tensor(v), b)  # sum conj(A) * B


def edge_scalar(net, msgs, e):
    u, v = net.graph.edges[e]
    return np.sum(msgs[(u, v)] * msgs[(v, u)])  # no conjugate


def region_scalars(net, msgs):
    zv = np.array([vertex_scalar(net, msgs, v) for v in range(net.graph.nv)])
    ze = np.array([edge_scalar(net, msgs, e) for e in range(net.graph.ne)])
    return zv, ze


def logscalar(net, msgs):
    zv, ze = region_scalars(net, msgs)
    if np.any(zv.real < 0):
        zv = zv.astype(np.complex128)
    if np.any(ze.real < 0):
        ze = ze.astype(np.complex128)
    if np.any(ze == 0):
        return -np.inf
    with np.errstate(divide="ignore"):
        return np.sum(np.log(zv)) - np.sum(np.log(ze))


def scalar(net, msgs):
    return np.exp(logscalar(net, msgs))


# ----------------------------------------------------------------------------------------
# rescale / normalize  (beliefpropagationcache.jl:121-139, abstract :349-395, normalize.jl:63-80)
# ----------------------------------------------------------------------------------------


def _isreal(x):
    return np.imag(x) == 0


def rescale_messages(net, msgs):
    msgs = dict(msgs)
    for e, (u, v) in enumerate(net.graph.edges):
        me = msgs[(u, v)] / np.linalg.norm(msgs[(u, v)])
        mer = msgs[(v, u)] / np.linalg.norm(msgs[(v, u)])
        n = np.sum(me * mer)
        if _isreal(n):
            s = np.sign(np.real(n))
            me = me * s
            n = n * s
        sf = 1.0 / np.sqrt(n)
        msgs[(u, v)] = (sf * me).astype(net.dtype)
        msgs[(v, u)] = (sf * mer).astype(net.dtype)
    return msgs


def rescale_partitions(net, msgs, verts=None):
    """verts = all ket and bra vertices (normalize.jl:75-76): each ket/bra tensor / its norm,
    then Z_v^(-1/2) spread over ket and bra.  The engine keeps bra == conj(ket), so the ket
    weight is taken as |Z_v|^(-1/2) (identical when Z_v is real positive, the <psi|psi> case).
    `verts` (abstractbeliefpropagationcache.jl:349-379): only these sites (ket and bra of each) are touched;
    partitions without a listed vertex are skipped (:365)."""
    net = net.copy()
    vs = range(net.graph.nv) if verts is None else sorted(set(int(v) for v in verts))
    for v in vs:
        net.tensors[v] = net.tensors[v] / np.linalg.norm(net.tensors[v])
    for v in vs:
        zv = vertex_scalar(net, msgs, v)
        net.tensors[v] = net.tensors[v] * (abs(zv) ** -0.5)
    return net


def rescale(net, msgs, verts=None):
    msgs = rescale_messages(net, msgs)
    net = rescale_partitions(net, msgs, verts)
    return net, msgs


# ----------------------------------------------------------------------------------------
# expectation values  (expect.jl:5-19; two-site RDM idiom test_belief_propagation.jl:64-91)
# ----------------------------------------------------------------------------------------


def rdm1(net, msgs, v):
    """rho[s, s'] = sum A[s,a..] conj(A[s',a'..]) prod M[a,a'] (un-normalised)."""
    b = absorbed_ket(net, msgs, v)
    a = net.tensors[v]
    axes = list(range(1, a.ndim))
    return np.tensordot(b, a.conj(), axes=(axes, axes))


def expect1(net, msgs, v, op):
    """<O_v> = sum_{s,s'} O[s', s] rho[s, s'] / tr(rho);  O indexed [s_out(bra side), s_in(ket side)]."""
    rho = rdm1(net, msgs, v)
    return np.sum(op.T * rho) / np.trace(rho)


def bond_env(net, msgs, v, e):
    """E[s, l, s', l'] = sum_{outer} B[s, a'.., l] conj(A[s', a'.., l'])  with all messages but e absorbed."""
    g = net.graph
    k = g.slot(v, e)
    a = net.tensors[v]
    b = absorbed_ket(net, msgs, v, skip_edges=(e,))
    axes = [i for i in range(1, a.ndim) if i != 1 + k]
    return np.tensordot(b, a.conj(), axes=(axes, axes))  # [s, l, s', l']


def rdm2(net, msgs, e):
    """rho[(s_u, s_v), (s_u', s_v')] for the edge e = (u, v), normalised to unit trace; s_u fastest."""
    u, v = net.graph.edges[e]
    eu = bond_env(net, msgs, u, e)
    ev = bond_env(net, msgs, v, e)
    rho = np.einsum("alcm,blem->abce", eu, ev)  # [s_u, s_v, s_u', s_v']
    du, dv = rho.shape[0], rho.shape[1]
    rho = rho.transpose(1, 0, 3, 2).reshape(du * dv, du * dv)  # row = s_u + du*s_v
    return rho / np.trace(rho)


def expect2(net, msgs, e, op_u, op_v):
    u, v = net.graph.edges[e]
    eu = bond_env(net, msgs, u, e)
    ev = bond_env(net, msgs, v, e)
    rho = np.einsum("alcm,blem->abce", eu, ev)
    num = np.einsum("abce,ca,eb->", rho, op_u, op_v)
    den = np.einsum("abab->", rho)
    return num / den


# ----------------------------------------------------------------------------------------
# truncation + map_eigvals  (apply.jl:9-25; NDTensors truncate! semantics, SURVEY A.7 - unpinned)
# ----------------------------------------------------------------------------------------


def truncate_spectrum(p, maxdim=None, cutoff=None, mindim=1):
    """Weights p sorted descending. Returns (kept n, truncerr). Relative cutoff on the discarded sum."""
    p = np.array(p, dtype=np.float64).copy()
    origm = len(p)
    for n in range(origm - 1, -1, -1):  # zero out negative weights at the tail
        if p[n] >= 0:
            break
        p[n] = 0.0
    if origm == 1:
        return 1, 0.0
    maxdim = origm if maxdim is None else min(int(maxdim), origm)
    n = origm
    truncerr = 0.0
    while n > maxdim:
        truncerr += p[n - 1]
        n -= 1
    scale = float(np.sum(p))
    if scale == 0.0:
        scale = 1.0
    if cutoff is not None:
        while n > mindim and truncerr + p[n - 1] <= cutoff * scale:
            truncerr += p[n - 1]
            n -= 1
    truncerr /= scale
    return max(n, 1), truncerr


def map_eigvals(f, m, cutoff=None):
    """f applied to the eigenvalues of Hermitian m (apply.jl:21-25).  Exactly diagonal input
    short-circuits (map_diag, :22); otherwise eigenvalues are sorted by decreasing magnitude
    and truncated with the relative `cutoff` before f is applied (pseudo-inverse semantics)."""
    if np.count_nonzero(m - np.diag(np.diagonal(m))) == 0:
        return np.diag(f(np.diagonal(m).astype(m.dtype))).astype(m.dtype)
    w, u = np.linalg.eigh((m + m.conj().T) / 2)
    order = np.argsort(-np.abs(w), kind="stable")
    w, u = w[order], u[:, order]
    n, _ = truncate_spectrum(w, cutoff=cutoff) if cutoff is not None else (len(w), 0.0)
    w, u = w[:n], u[:, :n]
    return (u * f(w.astype(m.dtype))) @ u.conj().T


# ----------------------------------------------------------------------------------------
# gates  (apply.jl:33-146)
# ----------------------------------------------------------------------------------------


def apply1(net, v, gate, normalize=False):
    """A'[s', ..] = sum_s gate[s', s] A[s, ..]  (apply.jl:108-116)."""
    net = net.copy()
    t = np.tensordot(gate, net.tensors[v], axes=([1], [0]))
    if normalize:
        t = t / np.linalg.norm(t)
    net.tensors[v] = t.astype(net.dtype)
    return net


def simple_update_bp(net, msgs, e, gate, maxdim=None, cutoff=None, normalize=False):
    """Two-site gate on the edge e=(v1, v2) with BP (product) environments (apply.jl:33-95).

    gate[s1', s2', s1, s2].  Returns (new_net, info) with info = dict(svals, truncerr, newdim).
    """
    g = net.graph
    v1, v2 = g.edges[e]
    eps = np.finfo(np.float64).eps
    eig_cutoff = 10 * eps

    def side(v):
        a = net.tensors[v]
        k = g.slot(v, e)
        outer = [(j, f) for j, f in enumerate(g.inc[v]) if f != e]
        sq, isq = {}, {}
        for j, f in outer:
            env = msgs[(g.other(f, v), v)]
            sq[j] = map_eigvals(np.sqrt, env, cutoff=eig_cutoff)
            isq[j] = map_eigvals(lambda x: 1.0 / np.sqrt(x), env, cutoff=eig_cutoff)
        at = a
        for j, _ in outer:
            at = _absorb(at, 1 + j, sq[j])
        # matrix with rows = outer bonds, cols = (s, shared bond)
        perm = [1 + j for j, _ in outer] + [0, 1 + k]
        atp = at.transpose(perm)
        outer_shape = atp.shape[:-2]
        d, chi = atp.shape[-2], atp.shape[-1]
        mat = atp.reshape(int(np.prod(outer_shape, dtype=np.int64)), d * chi)
        q, r = np.linalg.qr(mat, mode="reduced")
        rk = r.shape[0]
        return dict(v=v, k=k, outer=outer, isq=isq, q=q, r=r.reshape(rk, d, chi),
                    outer_shape=outer_shape, d=d)

    s1, s2 = side(v1), side(v2)
    theta = np.einsum("asl,btl->asbt", s1["r"], s2["r"])  # [r1, s1, r2, s2]
    theta = np.einsum("xyst,asbt->axby", gate, theta)  # [r1, s1', r2, s2']
    r1, d1, r2, d2 = theta.shape
    mat = theta.reshape(r1 * d1, r2 * d2)
    u, sv, vh = np.linalg.svd(mat, full_matrices=False)
    n, truncerr = truncate_spectrum(sv ** 2, maxdim=maxdim, cutoff=cutoff)
    u, sv, vh = u[:, :n], sv[:n], vh[:n, :]
    rs = np.sqrt(sv)
    new_r1 = (u * rs).reshape(r1, d1, n)  # [r1, s1, l]
    new_r2 = (rs[:, None] * vh).reshape(n, r2, d2).transpose(1, 2, 0)  # [r2, s2, l]

    new = net.copy()
    for sd, new_r in ((s1, new_r1), (s2, new_r2)):
        v, k, outer = sd["v"], sd["k"], sd["outer"]
        q = sd["q"].reshape(sd["outer_shape"] + (sd["q"].shape[1],))  # [outer..., r]
        for pos, (j, _) in enumerate(outer):
            q = _absorb(q, pos, sd["isq"][j].conj().T)  # Q'[a] = sum_a' Q[a'] conj(W[a, a'])
        t = np.tensordot(q, new_r, axes=([q.ndim - 1], [0]))  # [outer..., s, l]
        # back to [s, bonds in incident order]
        nb = len(g.inc[v])
        src_axes = {1 + j: pos for pos, (j, _) in enumerate(outer)}
        src_axes[0] = len(outer)
        src_axes[1 + k] = len(outer) + 1
        t = t.transpose([src_axes[i] for i in range(nb + 1)])
        if normalize:
            t = t / np.linalg.norm(t)
        new.tensors[v] = np.ascontiguousarray(t.astype(net.dtype))
    return new, dict(svals=sv, truncerr=truncerr, newdim=n)


# ----------------------------------------------------------------------------------------
# tree gauge  (abstractitensornetwork.jl:376-420) and apply(...; ortho = true)  (apply.jl:109-111, 130-132)
# ----------------------------------------------------------------------------------------


def qr_edge(net, u, v):
    """qr!(tn, u => v) (abstractitensornetwork.jl:376-385): Q of the tensor at u over (all its other indices) -> bond
    replaces it, R is multiplied into the tensor at v.  In place on a copy; returns the new network."""
    g = net.graph
    e = g.eid[(u, v)]
    new = net.copy()
    a = net.tensors[u]
    k = 1 + g.slot(u, e)
    am = np.moveaxis(a, k, -1)
    q, r = np.linalg.qr(am.reshape(-1, am.shape[-1]), mode="reduced")
    chi = am.shape[-1]
    if q.shape[1] < chi:  # fewer rows than bond states: pad with zero columns so that the bond keeps its extent
        q = np.concatenate([q, np.zeros((q.shape[0], chi - q.shape[1]), dtype=q.dtype)], axis=1)
        r = np.concatenate([r, np.zeros((chi - r.shape[0], chi), dtype=r.dtype)], axis=0)
    new.tensors[u] = np.ascontiguousarray(np.moveaxis(q.reshape(am.shape), -1, k).astype(net.dtype))
    kb = 1 + g.slot(v, e)
    b = np.moveaxis(net.tensors[v], kb, 0)
    b2 = np.tensordot(r, b, axes=([1], [0]))
    new.tensors[v] = np.ascontiguousarray(np.moveaxis(b2, 0, kb).astype(net.dtype))
    return new


def gauge_walk(net, edges):
    """gauge_walk(tn, edges) (abstractitensornetwork.jl:387-393)."""
    for (u, v) in edges:
        net = qr_edge(net, u, v)
    return net


def spanning_tree_edges(g: Graph):
    """Edge ids of a breadth-first spanning forest (the tree itself when g is a tree): the `steiner_tree` of all vertices
    in edge_sequence_between_regions (abstractitensornetwork.jl:399-405); on a loopy graph which spanning tree the
    un-vendored routine picks is unpinned, and the host passes the explicit sequence to the engine."""
    return forest_cover(g)[0] if g.ne else []


def tree_gauge_sequence(g: Graph, region):
    """Edges (child, parent) that move the gauge from everywhere to `region` (a vertex or a list of vertices):
    post-order DFS edges of the spanning tree rooted at region[0], without the edges inside the region
    (edge_sequence_between_regions, abstractitensornetwork.jl:399-405)."""
    region = [region] if np.isscalar(region) else list(region)
    if set(region) == set(range(g.nv)):
        return []
    te = _post_order_dfs_edges(g.nv, spanning_tree_edges(g), g.edges, region[0])
    rs = set(region)
    return [(a, b) for (a, b) in te if not (a in rs and b in rs)]


def tree_orthogonalize(net, region):
    """tree_orthogonalize(psi, region) = tree_gauge(psi, region) (abstractitensornetwork.jl:407-420)."""
    return gauge_walk(net, tree_gauge_sequence(net.graph, region))


def apply2_ortho(net, e, gate, maxdim=None, cutoff=None, normalize=False):
    """apply(o, psi; ortho = true) on the edge e with the default `envs = ITensor[]` (apply.jl:97-139): gauge the tree
    towards the first gate vertex, then simple_update_bp with no environments (identity messages)."""
    v1, _ = net.graph.edges[e]
    net = tree_orthogonalize(net, v1)
    return simple_update_bp(net, identity_messages(net), e, gate, maxdim=maxdim, cutoff=cutoff, normalize=normalize)


def reset_edge_messages(net, msgs, e):
    """Engine convention after a gate changed a bond: both directed messages on e restart from identity."""
    msgs = dict(msgs)
    u, v = net.graph.edges[e]
    chi = net.edge_dim(e)
    msgs[(u, v)] = np.eye(chi, dtype=net.dtype)
    msgs[(v, u)] = np.eye(chi, dtype=net.dtype)
    return msgs


# ----------------------------------------------------------------------------------------
# brute-force exact contraction (small networks only) - the known-answer side of the property tests
# ----------------------------------------------------------------------------------------


def _state_vector(net):
    """Full wavefunction psi[s_0, ..., s_{nv-1}] by sequential contraction (tiny networks only)."""
    g = net.graph
    import string
    letters = iter(string.ascii_letters)
    site = [next(letters) for _ in range(g.nv)]
    bond = [next(letters) for _ in range(g.ne)]
    subs = []
    for v in range(g.nv):
        subs.append(site[v] + "".join(bond[e] for e in g.inc[v]))
    expr = ",".join(subs) + "->" + "".join(site)
    return np.einsum(expr, *net.tensors, optimize="greedy")


def bilinear_network(phi: Network, psi: Network):
    """inner_network(phi, psi) (src/inner.jl:139-152 -> BilinearFormNetwork with the identity operator layer,
    formnetworks/bilinearformnetwork.jl:23-42,74-94): ket = psi, bra = dag(phi).  Bond dimensions may differ: the smaller
    tensor is zero-padded, which changes no contraction."""
    g = psi.graph
    dtype = np.result_type(phi.dtype, psi.dtype)
    kets, bras = [], []
    for v in range(g.nv):
        a, b = psi.tensors[v].astype(dtype), phi.tensors[v].astype(dtype)
        shape = tuple(max(x, y) for x, y in zip(a.shape, b.shape))
        pa = np.zeros(shape, dtype=dtype)
        pa[tuple(slice(0, n) for n in a.shape)] = a
        pb = np.zeros(shape, dtype=dtype)
        pb[tuple(slice(0, n) for n in b.shape)] = b
        kets.append(pa)
        bras.append(pb)
    return Network(g, kets, dtype, bras)


def apply_operator_network(op: Network, psi: Network):
    """A|psi> for an operator network A_v[s', s, b_1..b_z] on the same graph: the operator and ket layers of
    inner_network(phi, A, psi) (src/inner.jl:154-171) contracted site by site, bonds fused as a_k + chi_k * b_k."""
    g = psi.graph
    out = []
    for v in range(g.nv):
        a, t = op.tensors[v], psi.tensors[v]
        z = t.ndim - 1
        r = np.tensordot(a, t, axes=([1], [0]))  # [s', b_1..b_z, a_1..a_z]
        r = np.transpose(r, [0] + [i for k in range(z) for i in (1 + k, 1 + z + k)])  # [s', b_1, a_1, b_2, a_2, ...]
        # C-order reshape of each (b_k, a_k) pair: a_k fastest
        out.append(np.ascontiguousarray(r.reshape([r.shape[0]] + [r.shape[1 + 2 * k] * r.shape[2 + 2 * k] for k in range(z)])))
    return Network(g, out, np.result_type(op.dtype, psi.dtype))


def exact_inner_operator(phi: Network, op: Network, psi: Network):
    """<phi|A|psi> by brute force (inner(phi, A, psi; alg = "exact"), src/inner.jl:77-98); tiny networks only."""
    import string
    g = psi.graph
    letters = iter(string.ascii_letters)
    so = [next(letters) for _ in range(g.nv)]
    si = [next(letters) for _ in range(g.nv)]
    b = [next(letters) for _ in range(g.ne)]
    subs = [so[v] + si[v] + "".join(b[e] for e in g.inc[v]) for v in range(g.nv)]
    full = np.einsum(",".join(subs) + "->" + "".join(so) + "".join(si), *op.tensors, optimize="greedy")
    apsi = np.tensordot(full, _state_vector(psi), axes=(list(range(g.nv, 2 * g.nv)), list(range(g.nv))))
    return np.vdot(_state_vector(phi), apsi)


def partition_network(net: Network, groups):
    """Super-site network of `net` for a partition of its vertices into `groups` (partitioned_vertices of
    BeliefPropagationCache, src/caches/beliefpropagationcache.jl:20-35; column grouping in test/test_expect.jl:22-39).
    One einsum per partition: internal edges are summed, the site indices of the members fuse into one site index
    (first member fastest) and the edges to each neighbouring partition fuse into one bond (ascending edge id, first
    fastest).  Returns (Network on the quotient graph, group_of)."""
    import string
    g = net.graph
    group_of = {int(v): gi for gi, grp in enumerate(groups) for v in grp}
    bundles = {}
    for e, (u, v) in enumerate(g.edges):
        a, b = group_of[u], group_of[v]
        if a != b:
            bundles.setdefault((min(a, b), max(a, b)), []).append(e)
    qedges = sorted(bundles)
    qg = Graph(len(groups), qedges)
    tensors = []
    for gi, grp in enumerate(groups):
        letters = iter(string.ascii_letters)
        site = {v: next(letters) for v in grp}
        bond = {}
        for v in grp:
            for e in g.inc[v]:
                if e not in bond:
                    bond[e] = next(letters)
        subs = [site[v] + "".join(bond[e] for e in g.inc[v]) for v in grp]
        out = "".join(site[v] for v in grp)
        shape = [int(np.prod([net.tensors[v].shape[0] for v in grp]))]
        for qe in qg.inc[gi]:
            es = bundles[qedges[qe]]
            out += "".join(bond[e] for e in es)
            shape.append(int(np.prod([net.edge_dim(e) for e in es])))
        t = np.einsum(",".join(subs) + "->" + out, *[net.tensors[v] for v in grp])
        tensors.append(np.ascontiguousarray(t.reshape(shape, order="F")))
    return Network(qg, tensors, net.dtype), group_of


def lift_operator(site_dims, pos, o):
    """Operator on member `pos` of a partition as an operator on the fused site index (first member fastest)."""
    out = np.ones((1, 1), dtype=np.asarray(o).dtype)
    for q, d in enumerate(site_dims):
        out = np.kron(np.asarray(o) if q == pos else np.eye(d), out)
    return out


def exact_inner(phi: Network, psi: Network):
    """<phi|psi> by brute force (inner(phi, psi; alg = "exact"), src/inner.jl:60-75)."""
    return np.vdot(_state_vector(phi), _state_vector(psi))


def exact_norm_sqr(net):
    psi = _state_vector(net)
    return np.vdot(psi, psi)


def exact_expect1(net, v, op):
    psi = _state_vector(net)
    opsi = np.moveaxis(np.tensordot(op, psi, axes=([1], [v])), 0, v)
    return np.vdot(psi, opsi) / np.vdot(psi, psi)


def exact_rdm2(net, e):
    u, v = net.graph.edges[e]
    psi = _state_vector(net)
    nv = psi.ndim
    rest = [i for i in range(nv) if i not in (u, v)]
    rho = np.tensordot(psi, psi.conj(), axes=(rest, rest))  # [s_a, s_b, s_a', s_b'] a<b in axis order
    if u > v:
        rho = rho.transpose(1, 0, 3, 2)
    du, dv = rho.shape[0], rho.shape[1]
    rho = rho.transpose(1, 0, 3, 2).reshape(du * dv, du * dv)
    return rho / np.trace(rho)


def exact_apply2(net, e, gate):
    """psi' = gate applied on (v1, v2) to the full state (tiny networks)."""
    v1, v2 = net.graph.edges[e]
    psi = _state_vector(net)
    out = np.tensordot(gate, psi, axes=([2, 3], [v1, v2]))  # [s1', s2', rest...]
    return np.moveaxis(out, [0, 1], [v1, v2])


PAULI_Z = np.array([[1.0, 0.0], [0.0, -1.0]])
PAULI_X = np.array([[0.0, 1.0], [1.0, 0.0]])


def random_unitary(n, seed=0, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "c":
        m = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    else:
        m = rng.standard_normal((n, n))
    q, r = np.linalg.qr(m)
    return (q * (np.diagonal(r) / np.abs(np.diagonal(r)))).astype(dtype)
