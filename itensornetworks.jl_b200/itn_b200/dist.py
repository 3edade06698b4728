"""Host side of the multi-GPU path: graph partition, halo plan, NCCL bootstrap.

No reference counterpart (the reference is single-process).  One process per GPU; every
rank builds the same graph, stores the site tensors of the vertices it owns, and runs the synchronous sweep
on them; `libitn_b200` exchanges the messages that cross a cut once per sweep (csrc/itn_dist.cu).  The plan
below is the host-side statement of what the library does, used by the CPU (gloo) tests and by callers that
want to know the traffic.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


def partition_vertices(graph, nparts):
    """owner[v] for a strip partition: grids are cut along their first axis (contiguous row blocks),
    anything else into contiguous vertex-id blocks.  NVSwitch gives every pair of GPUs the same bandwidth,
    so only the cut size matters, not which ranks are neighbours."""
    nparts = int(nparts)
    if nparts <= 1:
        return [0] * graph.nv
    names = getattr(graph, "names", None)
    if names and isinstance(names[0], tuple) and len(names[0]) >= 1:
        n0 = max(c[0] for c in names) + 1
        if n0 >= nparts:
            return [min(nparts - 1, (c[0] * nparts) // n0) for c in names]
    return [min(nparts - 1, (v * nparts) // graph.nv) for v in range(graph.nv)]


def directed_id(graph, u, v):
    """2e for esrc -> edst, 2e + 1 for the reverse direction (the library's message numbering)."""
    e = graph.eid[(u, v)]
    return 2 * e + (0 if graph.edges[e] == (u, v) else 1)


def halo_plan(graph, owner, rank, edges):
    """Messages of one sweep that cross a cut, as seen from `rank`:
    {peer: {"send": [(u, v), ...], "recv": [(u, v), ...]}} with both lists ordered by directed id, so that the
    k-th message sent by one side is the k-th message received by the other."""
    plan = {}
    seen = set()
    for (u, v) in sorted(edges, key=lambda e: directed_id(graph, *e)):
        if (u, v) in seen:
            continue
        seen.add((u, v))
        ou, ov = owner[u], owner[v]
        if ou == ov:
            continue
        if ou == rank:
            plan.setdefault(ov, {"send": [], "recv": []})["send"].append((u, v))
        if ov == rank:
            plan.setdefault(ou, {"send": [], "recv": []})["recv"].append((u, v))
    return plan


def halo_bytes_per_sweep(graph, owner, edge_dims, itemsize):
    """Bytes every rank sends per synchronous sweep (for the traffic table in DESIGN.md)."""
    nparts = max(owner) + 1
    out = [0] * nparts
    for e, (u, v) in enumerate(graph.edges):
        if owner[u] != owner[v]:
            b = int(edge_dims[e]) ** 2 * itemsize
            out[owner[u]] += b
            out[owner[v]] += b
    return out


def init_distributed(ctx, rank, world):
    """Create the library's NCCL communicator: rank 0 draws the id, torch.distributed (any backend) broadcasts
    its 128 bytes, every rank calls itn_ctx_init_dist."""
    import torch
    import torch.distributed as dist
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib().itn_nccl_unique_id(buf))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    buf2 = (C.c_ubyte * 128).from_buffer_copy(raw)
    check(lib().itn_ctx_init_dist(ctx.h, int(rank), int(world), buf2))
    ctx.rank, ctx.world = rank, world
