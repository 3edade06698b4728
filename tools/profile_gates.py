"""Runs one simple-update Trotter step (one two-site gate per edge, colour layer by colour layer) on a square-lattice
PEPS so that `ncu --metrics gpu__time_duration.sum` can list the launches of a gate layer; prints per-layer wall time.

    python tools/profile_gates.py [L] [chi] [nlayers]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np  # noqa: E402

import itn_b200 as E  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nlayers = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = 2
g = E.named_grid((L, L))
psi = E.random_tensornetwork(1234, np.complex128, g, link_space=chi, d=d)
ctx = E.Context(0)
bpc = E.BeliefPropagationCache(psi, ctx=ctx)
seq = E.parallel_edge_sequence(g)
E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
rng = np.random.default_rng(7)
m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
h = (m + m.conj().T) / 2
w, v = np.linalg.eigh(h)
gate = ((v * np.exp(-0.05 * w)) @ v.conj().T).astype(np.complex128).reshape(d, d, d, d)
layers = E.edge_coloring(g)
ctx.sync()
l0 = ctx.launch_count()
print(f"setup launches: {l0}", flush=True)
for rep in range(2):
    work = bpc.copy()
    ctx.sync()
    for li, layer in enumerate(layers[:nlayers]):
        t0 = time.perf_counter()
        info = E.apply_layer([gate] * len(layer), work, [g.edges[e] for e in layer], maxdim=chi, cutoff=1e-12)
        ctx.sync()
        dt = time.perf_counter() - t0
        print(f"rep {rep} layer {li}: {len(layer)} gates, {1e3 * dt:.2f} ms, {len(layer) / dt:.0f} gates/s, "
              f"launches so far {ctx.launch_count() - l0}", flush=True)
    work.close()
