"""torchrun entry: colour layers of a simple-update Trotter step on a graph-partitioned square-lattice PEPS, one process
per GPU; prints per-layer wall times (max over ranks) and, with ITN_TRACE=1, the host phases of itn_apply2 on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/profile_gates_dist.py [L] [chi] [reps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    import itn_b200 as E
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    E.init_distributed(ctx, rank, world)
    g = E.named_grid((L, L))
    owner = E.partition_vertices(g, world)
    psi = E.random_tensornetwork(1234, np.complex128, g, link_space=chi, d=2)
    bpc = E.BeliefPropagationCache(psi, ctx=ctx, owner=owner)
    del psi
    seq = E.parallel_edge_sequence(g)
    E.update(bpc, maxiter=3, edge_sequence=seq, inplace=True)
    rng = np.random.default_rng(7)
    m = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    h = (m + m.conj().T) / 2
    w, v = np.linalg.eigh(h)
    gate = ((v * np.exp(-0.05 * w)) @ v.conj().T).astype(np.complex128).reshape(2, 2, 2, 2)
    layers = E.edge_coloring(g)
    work = bpc.copy()
    for rep in range(reps):
        for li, layer in enumerate(layers):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            E.apply_layer([gate] * len(layer), work, [g.edges[e] for e in layer], maxdim=chi, cutoff=None)
            t_call = time.perf_counter() - t0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt, t_call], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"rep {rep} layer {li}: {len(layer)} gates on {world} GPUs, {1e3 * float(t[0]):.2f} ms "
                      f"(call returned after {1e3 * float(t[1]):.2f} ms), {len(layer) / float(t[0]):.0f} gates/s", flush=True)
    work.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
