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


def _factorizations(n, k):
    """All ordered k-tuples of positive integers with product n."""
    if k == 1:
        return [(n,)]
    out = []
    for f in range(1, n + 1):
        if n % f == 0:
            out += [(f,) + rest for rest in _factorizations(n // f, k - 1)]
    return out


def cut_edges(graph, owner):
    """(total number of edges crossing a cut, largest number of cut edges at one rank)."""
    per = {}
    tot = 0
    for (u, v) in graph.edges:
        if owner[u] != owner[v]:
            tot += 1
            per[owner[u]] = per.get(owner[u], 0) + 1
            per[owner[v]] = per.get(owner[v], 0) + 1
    return tot, max(per.values()) if per else 0


def _brick_owner(names, shape, parts):
    own = []
    for c in names:
        r = 0
        for ax in range(len(shape)):
            r = r * parts[ax] + min(parts[ax] - 1, (c[ax] * parts[ax]) // shape[ax])
        own.append(r)
    return own


def partition_vertices(graph, nparts, kind="auto"):
    """owner[v] of a graph partition into `nparts` ranks.
      "strips"  grids are cut along their first axis (contiguous row blocks / planes),
      "bricks"  grids are cut into a p_0 x p_1 x ... array of near-cubic blocks (the factorisation of nparts with the
                fewest cut edges, ties broken by the busiest rank's cut),
      "auto"    whichever of the two has fewer cut edges at the busiest rank (the halo exchange and the cut-gate traffic
                of a rank scale with its own cut; NVSwitch gives every pair of GPUs the same bandwidth, so it does not
                matter which ranks are neighbours).
    Graphs without grid coordinates are cut into contiguous vertex-id blocks."""
    nparts = int(nparts)
    if nparts <= 1:
        return [0] * graph.nv
    names = getattr(graph, "names", None)
    if names and isinstance(names[0], tuple) and len(names[0]) >= 1:
        nd = len(names[0])
        shape = [max(c[ax] for c in names) + 1 for ax in range(nd)]
        strips = None
        if shape[0] >= nparts:
            strips = _brick_owner(names, shape, (nparts,) + (1,) * (nd - 1))
        if kind == "strips" and strips is not None:
            return strips
        best = None
        for parts in _factorizations(nparts, nd):
            if any(parts[ax] > shape[ax] for ax in range(nd)):
                continue
            own = _brick_owner(names, shape, parts)
            tot, mx = cut_edges(graph, own)
            key = (tot, mx) if kind == "bricks" else (mx, tot)
            if best is None or key < best[0]:
                best = (key, own)
        if best is not None:
            return best[1]
        if strips is not None:
            return strips
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


def gate_exchange_plan(graph, owner, rank, pairs, edge_dims, sdims, planes=2):
    """What itn_apply2 exchanges for a vertex-disjoint layer of two-site gates on a partitioned network, as seen from
    `rank` (csrc/itn_linalg.cu): site-level work runs where the site lives, edge-level work on the rank that owns esrc
    (the "owner"); for an edge crossing a cut the other rank (the "guest") sends its bond environment C (n x n,
    n = d chi) and receives its T factor (n x d cand, cand = largest possible new bond dimension).
    Returns {peer: {"send_C": [(gate index, doubles)], "recv_C": [...], "send_T": [...], "recv_T": [...]}}; every list is
    in gate order, which is the order of the segments inside the NCCL buffers on both sides."""
    plan = {}
    for i, (a, b) in enumerate(pairs):
        e = graph.eid[(a, b)]
        u, v = graph.edges[e]  # engine orientation (esrc, edst)
        ou, ov = owner[u], owner[v]
        if ou == ov or rank not in (ou, ov):
            continue
        chi = int(edge_dims[e])

        def outer(x):
            n = 1
            for f in graph.inc[x]:
                if f != e:
                    n *= int(edge_dims[f])
            return n
        r = [min(outer(x), sdims[x] * chi) for x in (u, v)]
        cand = min(r[0] * sdims[u], r[1] * sdims[v])
        nn = sdims[v] * chi  # the guest always holds edst
        csize = planes * nn * nn
        tsize = planes * nn * sdims[v] * cand
        if rank == ou:  # owner: receives C, sends T
            pl = plan.setdefault(ov, {"send_C": [], "recv_C": [], "send_T": [], "recv_T": []})
            pl["recv_C"].append((i, csize))
            pl["send_T"].append((i, tsize))
        else:           # guest
            pl = plan.setdefault(ou, {"send_C": [], "recv_C": [], "send_T": [], "recv_T": []})
            pl["send_C"].append((i, csize))
            pl["recv_T"].append((i, tsize))
    return plan


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
