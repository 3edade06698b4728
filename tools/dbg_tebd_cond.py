"""Conditioning of the BP messages and bond spectra along a BP-gauged TEBD run (one GPU): why do environments get flagged
rank deficient after a few steps?  usage: dbg_tebd_cond.py [n] [steps] [k_bp] [msg_mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np, torch
import itn_b200 as E
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k_bp = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g = E.named_grid((n, n))
ctx = E.Context(0)
tensors, host, _ = bench.make_psi(torch, g, 16, np.complex128, 2, [True] * g.nv)
psi = E.ITensorNetwork(g, tensors, np.complex128)
seq = E.parallel_edge_sequence(g)
bpc = E.BeliefPropagationCache(psi, ctx=ctx)
E.update(bpc, maxiter=20, edge_sequence=seq, inplace=True)
rng = np.random.default_rng(7)
m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)); h = (m + m.conj().T) / 2
w, v = np.linalg.eigh(h)
gate = ((v * np.exp(-0.05j * w)) @ v.conj().T).reshape(2, 2, 2, 2)
layers = E.edge_coloring(g)
lay = [([gate] * len(layer), [g.edges[e] for e in layer]) for layer in layers]

def report(tag):
    ratios, herm = [], []
    for (u, w_) in g.edges:
        for e in ((u, w_), (w_, u)):
            mm = np.asarray(bpc.message(e))
            herm.append(np.linalg.norm(mm - mm.conj().T) / np.linalg.norm(mm))
            ev = np.linalg.eigvalsh((mm + mm.conj().T) / 2)
            ratios.append(ev[0] / ev[-1])
    norms = [np.linalg.norm(np.asarray(bpc.factor(v_))) for v_ in range(0, g.nv, max(1, g.nv // 16))]
    ratios = np.array(ratios)
    print(f"{tag}: message lambda_min/lambda_max  min {ratios.min():.3e}  median {np.median(ratios):.3e}  "
          f"#(<1e-12) {int((ratios < 1e-12).sum())} of {len(ratios)};  non-hermiticity max {max(herm):.2e};  "
          f"tensor norms {min(norms):.3e} .. {max(norms):.3e}", flush=True)

report("start")
for s in range(steps):
    info = E.tebd_step(bpc, lay, maxdim=16, cutoff=None, msg_mode=mode, bp_maxiter=k_bp, edge_sequence=seq)
    sv = [np.asarray(x) for x in info["singular_values"]]
    r = np.array([x[-1] / x[0] for x in sv])
    print(f"step {s}: sigma_16/sigma_1 min {r.min():.3e} median {np.median(r):.3e}; terr max {info['truncation_error'].max():.3e}; paths {ctx.path_counts()}", flush=True)
    report(f"after step {s}")
