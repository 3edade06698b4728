"""Times the phases of the end-to-end call (upload, plan + relayout, sweep, download) on the bench workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np, torch
import itn_b200 as E

dims, chi, d = (64, 64), 16, 2
g = E.named_grid(dims)
sizes = [d * chi ** g.degree(v) for v in range(g.nv)]
offs = np.concatenate([[0], np.cumsum(sizes)])
host = torch.empty(int(offs[-1]) * 2, dtype=torch.float64).pin_memory()
torch.randn(host.shape, out=host)
hnp = host.numpy()
tensors = [np.ndarray((d,) + (chi,) * g.degree(v), dtype=np.complex128, buffer=hnp[int(offs[v]) * 2:int(offs[v + 1]) * 2].view(np.complex128), order="F") for v in range(g.nv)]
psi = E.ITensorNetwork(g, tensors, np.complex128)
ctx = E.Context(0)
seq = E.parallel_edge_sequence(g)
out_host = torch.empty(2 * g.ne * chi * chi * 2, dtype=torch.float64).pin_memory()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = E.BeliefPropagationCache(psi, ctx=ctx)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    E.update(c, maxiter=1, edge_sequence=seq, inplace=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    E.update(c, maxiter=1, edge_sequence=seq, inplace=True)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    c.messages_into(out_host.numpy())
    t5 = time.perf_counter()
    c.close(); torch.cuda.synchronize(); t6 = time.perf_counter()
    print(f"rep {rep}: construct(host) {t1-t0:.3f}s +drain {t2-t1:.3f}s | first update (plan+relayout+sweep) {t3-t2:.3f}s | second update {t4-t3:.3f}s | download {t5-t4:.3f}s | close {t6-t5:.3f}s")

# the deferred path bench.py's e2e arm uses: the constructor registers the host tensors, update() streams them
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = E.BeliefPropagationCache(psi, ctx=ctx, defer_upload=True)
    t1 = time.perf_counter()
    E.update(c, maxiter=1, edge_sequence=seq, inplace=True)
    t2 = time.perf_counter()
    c.messages_into(out_host.numpy())
    t3 = time.perf_counter()
    c.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"deferred rep {rep}: construct {t1-t0:.3f}s | update (upload + sweep) {t2-t1:.3f}s | download {t3-t2:.3f}s | close {t4-t3:.3f}s | total {t4-t0:.3f}s")
