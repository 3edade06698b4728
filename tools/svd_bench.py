"""Times the batched Jacobi SVD of the gate path (2048 bond matrices of 64 x 64 ComplexF64 = one colour layer of the
64 x 64 chi = 16 lattice) on both kernels and checks them against LAPACK.

    python tools/svd_bench.py [batch] [m] [n]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np  # noqa: E402

import itn_b200 as E  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ctx = E.Context(0)
rng = np.random.default_rng(3)
for dtype in (np.complex128, np.float64):
    a = rng.standard_normal((b, m, n))
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.standard_normal((b, m, n))
    a = a.astype(dtype)
    ref = np.linalg.svd(a[:16], compute_uv=False)
    for variant in (0, 3, 2, 1):
        ts = []
        for rep in range(4):
            sig, _, ms = E.svd_batch(a, variant=variant, ctx=ctx)
            ts.append(ms)
        err = np.max(np.abs(sig[:16, :ref.shape[1]] - ref)) / np.max(ref)
        print(f"{np.dtype(dtype).name} batch {b} {m}x{n} variant {variant}: ms {['%.3f' % t for t in ts]}  "
              f"max rel err vs LAPACK {err:.2e}", flush=True)
