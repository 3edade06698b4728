"""Times the pieces of a BP-gauged Trotter step on a partitioned 64x64 chi=16 lattice (torchrun, N ranks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np, torch, torch.distributed as dist
import itn_b200 as E
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = E.named_grid((n, n))
ctx = E.Context(local)
if world > 1:
    E.init_distributed(ctx, rank, world)
owner = E.partition_vertices(g, world) if world > 1 else None
mine = [owner is None or owner[v] == rank for v in range(g.nv)]
tensors, host, _ = bench.make_psi(torch, g, 16, np.complex128, 2, mine)
psi = E.ITensorNetwork(g, tensors, np.complex128)
seq = E.parallel_edge_sequence(g)
bpc = E.BeliefPropagationCache(psi, ctx=ctx, owner=owner)
E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
rng = np.random.default_rng(7)
m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)); h = (m + m.conj().T) / 2
w, v = np.linalg.eigh(h)
gate = ((v * np.exp(-0.05 * w)) @ v.conj().T).reshape(2, 2, 2, 2)
layers = E.edge_coloring(g)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for mode in (0, 1):
    for rep in range(nrep):
        out = []
        for layer in layers:
            sync(); t0 = time.perf_counter()
            E.apply_layer([gate] * len(layer), bpc, [g.edges[e] for e in layer], maxdim=16, cutoff=None, msg_mode=mode)
            sync(); t1 = time.perf_counter()
            E.update(bpc, maxiter=2, edge_sequence=seq, inplace=True)
            sync(); t2 = time.perf_counter()
            out.append((round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1), 1), ctx.path_counts()))
        if rank == 0:
            print(f"msg_mode={mode} rep={rep}: (gate ms, 2 sweeps ms, path counts) per layer: {out}", flush=True)
lay = [([gate] * len(layer), [g.edges[e] for e in layer]) for layer in layers]
for use_copy in (False, True):
    work = bpc.copy() if use_copy else bpc
    for rep in range(nrep):
        sync(); t0 = time.perf_counter()
        info = E.tebd_step(work, lay, maxdim=16, cutoff=None, msg_mode=1, bp_maxiter=2, edge_sequence=seq)
        t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
        if rank == 0:
            print("   newdim min/max", int(info["newdim"].min()), int(info["newdim"].max()), "terr max", float(info["truncation_error"].max()), flush=True)
        print(f"rank {rank} copy={use_copy} tebd_step rep {rep}: call {1e3 * (t1 - t0):.1f} ms, +sync {1e3 * (t2 - t1):.1f} ms, paths {ctx.path_counts()}", flush=True)
if world > 1:
    dist.destroy_process_group()
