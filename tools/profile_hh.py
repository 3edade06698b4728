"""BASELINE config 3 at full size: 127-qubit heavy-hex PEPS, chi = 32, ComplexF64.  Times synchronous BP sweeps and the
simple-update colour layers of a Trotter step (maxdim 32), printing per-call wall times; run it under
`ncu --metrics gpu__time_duration.sum` for the launch list or with ITN_TRACE=1 for the host phases of itn_apply2.

    python tools/profile_hh.py [chi] [reps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np  # noqa: E402

import itn_b200 as E  # noqa: E402

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = E.heavy_hex_eagle()
psi = E.random_tensornetwork(77, np.complex128, g, link_space=chi, d=2)
ctx = E.Context(0)
bpc = E.BeliefPropagationCache(psi, ctx=ctx)
seq = E.parallel_edge_sequence(g)
E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
ctx.sync()
t0 = time.perf_counter()
n_sweeps = 20
E.update(bpc, maxiter=n_sweeps, edge_sequence=seq, inplace=True)
ctx.sync()
dt = (time.perf_counter() - t0) / n_sweeps
print(f"heavy-hex chi={chi}: {1e3 * dt:.3f} ms per synchronous sweep, {2 * g.ne / dt:.3e} message updates/s "
      f"(device ms of the last update: {bpc.last_timing()})", flush=True)
rng = np.random.default_rng(9)
h = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
h = (h + h.conj().T) / 2
w, v = np.linalg.eigh(h)
gate = ((v * np.exp(-0.3j * w)) @ v.conj().T).reshape(2, 2, 2, 2)
layers = E.edge_coloring(g)
for rep in range(reps):
    work = bpc.copy()
    ctx.sync()
    tot, ng = 0.0, 0
    for li, layer in enumerate(layers):
        t0 = time.perf_counter()
        info = E.apply_layer([gate] * len(layer), work, [g.edges[e] for e in layer], maxdim=chi, cutoff=1e-12)
        ctx.sync()
        dt = time.perf_counter() - t0
        tot += dt
        ng += len(layer)
        print(f"rep {rep} layer {li}: {len(layer)} gates, {1e3 * dt:.2f} ms, max new dim {int(max(info['newdim']))}", flush=True)
    print(f"rep {rep}: Trotter step {1e3 * tot:.2f} ms, {ng / tot:.0f} gates/s", flush=True)
    work.close()
