"""Partitioned BP and gate layers on N GPUs vs the single-process oracle (and observables).

One process per GPU (the layout bench.py's contract launches):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Single process, one host thread per GPU (itn_ctx_create_group = ncclCommInitAll, SURVEY.md 8b "so Julia needs no MPI"):
    python tests/dist_gpu_check.py --threads 2
Prints 'DIST_OK' when every rank's stored messages, the all-reduced convergence measure, the region scalars, <Z>,
cut-edge RDMs and the gate layers match the oracle to 1e-10."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))


def run_checks(E, O, ctx, rank, world):
    """The checks of one rank; every rank (process or thread) runs the same sequence of collective calls."""
    worst = 0.0
    cases = (((6, 4), 3, np.complex128), ((8, 8), 16, np.complex128), ((4, 4), 2, np.float64), ((4, 4, 4), 2, np.complex128))
    for dims, chi, dtype in cases:
        g = O.grid_graph(dims)
        eg = E.named_grid(dims)
        owner = E.partition_vertices(eg, world)
        net = O.random_network(g, chi, dtype=dtype, seed=1234)
        psi = E.ITensorNetwork(eg, [t.copy() for t in net.tensors], dtype)
        seq = O.parallel_edge_sequence(g)
        iters = 4
        ref, _, ref_diff = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq),
                                       maxiter=iters, tol=0.0)
        bpc = E.BeliefPropagationCache(psi, ctx=ctx, owner=owner)
        info = {}
        E.update(bpc, maxiter=iters, tol=0.0, edge_sequence=[[e] for e in seq], inplace=True, info=info)
        for (u, v), m in ref.items():
            if owner[u] == rank or owner[v] == rank:
                worst = max(worst, np.linalg.norm(bpc.message((u, v)) - m) / np.linalg.norm(m))
        worst = max(worst, abs(info["mean_diff"] - ref_diff))
        zv, ze = E.scalar_factors_quotient(bpc)
        zv_o, ze_o = O.region_scalars(net, ref)
        worst = max(worst, np.linalg.norm(zv - zv_o) / np.linalg.norm(zv_o), np.linalg.norm(ze - ze_o) / np.linalg.norm(ze_o))
        worst = max(worst, abs(E.logscalar(bpc) - O.logscalar(net, ref)) * 1e-1)
        ez = E.expect(bpc, "Z")
        worst = max(worst, max(abs(ez[v] - O.expect1(net, ref, v, O.PAULI_Z)) for v in range(g.nv)))
        # two-site RDMs on interior and cut edges (test_belief_propagation.jl:64-91): every rank gets every matrix
        pick = [e for e, (u, v) in enumerate(g.edges) if owner[u] != owner[v]][:3] + [0, g.ne - 1]
        for m, e in zip(E.rdm2(bpc, pick), pick):
            worst = max(worst, np.linalg.norm(m - O.rdm2(net, ref, e)))
        r = E.rescale(bpc)
        zv2, ze2 = E.scalar_factors_quotient(r)
        worst = max(worst, float(np.max(np.abs(zv2 - 1))), float(np.max(np.abs(ze2 - 1))))
        # one-site gates on EVERY vertex, called identically on every rank (non-owners skip, src/apply.jl:108-116)
        kick = O.random_unitary(2, seed=5, dtype=dtype)
        for v in range(g.nv):
            net = O.apply1(net, v, kick)
            E.apply(kick, bpc, (v,), inplace=True)
        ref, _, _ = O.bp_update(net, ref, seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
        E.update(bpc, maxiter=2, edge_sequence=[[e] for e in seq], inplace=True)
        ez = E.expect(bpc, "Z")
        worst = max(worst, max(abs(ez[v] - O.expect1(net, ref, v, O.PAULI_Z)) for v in range(g.nv)))
        # simple-update gate layers on the partitioned network (src/apply.jl:33-95): interior edges and edges that cross
        # the cut (the guest rank ships its bond environment, the owner returns the T factor); then BP again and <Z>
        gate = O.random_unitary(4, seed=11, dtype=dtype).reshape(2, 2, 2, 2)
        gate = (gate + 0.5 * np.eye(4, dtype=dtype).reshape(2, 2, 2, 2)).astype(dtype)  # non-unitary: truncation error > 0
        maxdim = max(2, chi - 1) if chi <= 3 else chi
        msgs = ref
        for layer in O.edge_coloring(g):
            info2 = E.apply_layer([gate] * len(layer), bpc, [g.edges[e] for e in layer], maxdim=maxdim, cutoff=1e-13)
            for i, e in enumerate(layer):
                net, inf = O.simple_update_bp(net, msgs, e, gate, maxdim=maxdim, cutoff=1e-13)
                msgs = O.reset_edge_messages(net, msgs, e)
                assert info2["newdim"][i] == inf["newdim"], (dims, e, info2["newdim"][i], inf["newdim"])
                worst = max(worst, abs(info2["truncation_error"][i] - inf["truncerr"]))
                sv = inf["svals"][:inf["newdim"]]
                worst = max(worst, float(np.max(np.abs(info2["singular_values"][i] - sv)) / sv[0]))
            msgs, _, _ = O.bp_update(net, msgs, seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
            E.update(bpc, maxiter=2, edge_sequence=[[e] for e in seq], inplace=True)
        ez = E.expect(bpc, "Z")
        worst = max(worst, max(abs(ez[v] - O.expect1(net, msgs, v, O.PAULI_Z)) for v in range(g.nv)))
    return worst


def main_threads(n):
    import threading

    import itn_b200 as E
    from oracle import itn_oracle as O
    ctxs = E.Context.group(list(range(n)))
    worst = [None] * n
    errs = [None] * n

    def work(i):
        try:
            worst[i] = run_checks(E, O, ctxs[i], i, n)
        except BaseException as ex:  # a rank that dies would leave its peers waiting in a collective
            errs[i] = ex
            os._exit(3)

    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    w = max(worst)
    print(f"single process, {n} threads: worst error over ranks: {w:.3e}")
    print("DIST_OK" if w < 1e-10 else "DIST_FAIL")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--threads":
        return main_threads(int(sys.argv[2]))
    import torch
    import torch.distributed as dist

    import itn_b200 as E
    from oracle import itn_oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.Context(local)
    E.init_distributed(ctx, rank, world)
    worst = run_checks(E, O, ctx, rank, world)
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"worst error over ranks: {float(t[0]):.3e}")
        print("DIST_OK" if float(t[0]) < 1e-10 else "DIST_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
