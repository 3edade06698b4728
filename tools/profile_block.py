"""Synchronous BP sweeps on one of the lattices that run on the block path (csrc/itn_block.cu), for ncu.

    python tools/profile_block.py cubic 8 6 [sweeps]      # 8^3 cubic lattice, chi = 6 (BASELINE config 5's vertex types)
    python tools/profile_block.py grid 32 8               # config 2
    python tools/profile_block.py heavyhex 0 32           # config 3

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_block.py ...
    ncu --set full --clock-control none --import-source on -k regex:k_block -c 12 -o rep python tools/profile_block.py ...
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np  # noqa: E402

import itn_b200 as E  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "cubic"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
chi = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
g = E.heavy_hex_eagle() if kind == "heavyhex" else E.named_grid((n, n, n) if kind == "cubic" else (n, n))
psi = E.random_tensornetwork(7, np.complex128, g, link_space=chi, d=2)
ctx = E.Context(0)
bpc = E.BeliefPropagationCache(psi, ctx=ctx)
seq = E.parallel_edge_sequence(g)
E.update(bpc, maxiter=2, edge_sequence=seq, inplace=True)
ctx.sync()
t0 = time.perf_counter()
E.update(bpc, maxiter=sweeps, edge_sequence=seq, inplace=True)
ctx.sync()
dt = (time.perf_counter() - t0) / sweeps
c = 8.0
fl = sum(g.degree(v) * c * g.degree(v) * 2 * float(chi) ** (g.degree(v) + 1) for v in range(g.nv))
print(f"{kind} n={n} chi={chi}: {1e3 * dt:.3f} ms per sweep, {fl / dt / 1e12:.2f} algorithmic TFLOP/s, paths {ctx.path_counts()}, "
      f"device {bpc.last_timing()}")
