"""Shared helpers: the same seeded inputs go to the oracle (NumPy) and the engine (CUDA through the C ABI)."""
import numpy as np

import itn_b200 as E
from oracle import itn_oracle as O


def make_pair(graph_o, chi, dtype, seed=1234, d=2):
    """Oracle network + engine-side host network holding identical bytes."""
    net = O.random_network(graph_o, chi, d=d, dtype=dtype, seed=seed)
    g = E.NamedGraph(graph_o.nv, graph_o.edges)
    psi = E.ITensorNetwork(g, [t.copy() for t in net.tensors], dtype)
    return net, psi


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = max(np.linalg.norm(b), 1e-300)
    return np.linalg.norm(a - b) / den


def assert_messages_close(bpc, msgs, tol=1e-10):
    worst = 0.0
    for k, m in msgs.items():
        worst = max(worst, rel_err(bpc.message(k), m))
    assert worst < tol, f"message mismatch {worst:.3e}"
    return worst


def exact_site_environment(net, v):
    """Exact environment of site v in <psi|psi>: every other ket and bra tensor contracted, the bonds of v left open as
    (ket, bra) pairs in incident-edge order (the `environment(qf, state_vertices(qf, [v]); alg = "exact")` of
    test/test_forms.jl:62-63, for a single-site partition)."""
    import string
    g = net.graph
    letters = iter(string.ascii_letters)
    site = [next(letters) for _ in range(g.nv)]
    kb = [next(letters) for _ in range(g.ne)]
    bb = [next(letters) for _ in range(g.ne)]
    ops, subs = [], []
    for u in range(g.nv):
        if u == v:
            continue
        ops += [net.tensors[u], net.tensors[u].conj()]
        subs += [site[u] + "".join(kb[e] for e in g.inc[u]), site[u] + "".join(bb[e] for e in g.inc[u])]
    out = "".join(kb[e] + bb[e] for e in g.inc[v])
    return np.einsum(",".join(subs) + "->" + out, *ops)
