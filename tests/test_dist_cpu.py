"""world_size-2 gloo test of the multi-GPU host logic on CPU (no GPU, no NCCL).

Each rank owns one strip of the lattice, updates the messages leaving its vertices with the oracle's
updated_message (the checker standing in for the CUDA kernels, which need a GPU), and exchanges the
messages that cross the cut following itn_b200.halo_plan -- the same plan csrc/itn_dist.cu builds.  After k
synchronous sweeps every rank must hold exactly the messages of the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dims, chi, iters, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
        import torch
        import torch.distributed as dist

        import itn_b200 as E
        from oracle import itn_oracle as O
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        g = O.grid_graph(dims)
        eg = E.named_grid(dims)
        owner = E.partition_vertices(eg, world)
        net = O.random_network(g, chi, dtype=np.complex128, seed=1234)
        seq = [e for grp in E.parallel_edge_sequence(eg) for e in grp]
        plan = E.halo_plan(eg, owner, rank, seq)
        # a rank stores the messages on every edge touching one of its vertices
        msgs = {k: m for k, m in O.identity_messages(net).items() if owner[k[0]] == rank or owner[k[1]] == rank}
        local_diff = 0.0
        for _ in range(iters):
            new = {}
            local_diff = 0.0
            for (v, w) in seq:
                if owner[v] != rank:
                    continue
                new[(v, w)] = O.updated_message(net, msgs, v, w)
                local_diff += O.message_diff(new[(v, w)], msgs[(v, w)])
            msgs.update(new)
            reqs, bufs = [], []
            for peer, pl in sorted(plan.items()):
                if pl["send"]:
                    sb = torch.from_numpy(np.concatenate([msgs[k].ravel(order="F").view(np.float64) for k in pl["send"]]))
                    reqs.append(dist.isend(sb, peer))
                if pl["recv"]:
                    rb = torch.empty(sum(2 * msgs[k].size for k in pl["recv"]), dtype=torch.float64)
                    reqs.append(dist.irecv(rb, peer))
                    bufs.append((pl["recv"], rb))
            for r in reqs:
                r.wait()
            for keys, rb in bufs:
                off, a = 0, rb.numpy()
                for k in keys:
                    n = msgs[k].size
                    msgs[k] = a[off:off + 2 * n].view(np.complex128).reshape(msgs[k].shape, order="F").copy()
                    off += 2 * n
            t = torch.tensor([local_diff], dtype=torch.float64)
            dist.all_reduce(t)
            mean_diff = float(t[0]) / len(seq)
        ref, _, ref_diff = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq),
                                       maxiter=iters, tol=0.0)
        worst = max(np.linalg.norm(msgs[k] - ref[k]) for k in msgs)
        q.put((rank, worst, abs(mean_diff - ref_diff), len(msgs), sorted(plan)))
        dist.destroy_process_group()
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex), None, None, None))


def test_partitioned_sweeps_match_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    dims, chi, iters = (4, 3), 2, 3
    ps = [ctx.Process(target=_worker, args=(r, 2, port, dims, chi, iters, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, worst, ddiff, nstored, peers in res:
        assert not isinstance(worst, str), worst
        assert worst == 0.0, f"rank {rank}: partitioned sweep differs from the single-process oracle by {worst}"
        assert ddiff < 1e-15
        assert peers == [1 - rank]
        assert 0 < nstored < 2 * 17  # fewer messages than the full lattice (17 edges)


def test_halo_plan_is_symmetric():
    sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
    import itn_b200 as E
    for dims, world in (((8, 8), 4), ((6, 5), 3), ((4, 4, 4), 2), ((64, 64), 8)):
        g = E.named_grid(dims)
        owner = E.partition_vertices(g, world)
        assert sorted(set(owner)) == list(range(world))
        seq = [e for grp in E.parallel_edge_sequence(g) for e in grp]
        plans = [E.halo_plan(g, owner, r, seq) for r in range(world)]
        for r in range(world):
            for peer, pl in plans[r].items():
                assert pl["send"] == plans[peer][r]["recv"]
                assert pl["recv"] == plans[peer][r]["send"]
        ncut = sum(1 for (u, v) in g.edges if owner[u] != owner[v])
        assert sum(len(pl["send"]) for p in plans for pl in p.values()) == 2 * ncut
    g64 = E.named_grid((64, 64))
    b = E.halo_bytes_per_sweep(g64, E.partition_vertices(g64, 8, kind="strips"), [16] * 8064, 16)
    assert b[0] == 64 * 256 * 16 and b[3] == 2 * 64 * 256 * 16  # 256 KiB per cut and direction
    # bricks: 2 x 4 blocks cut 256 edges in total (strips: 448) and at most 80 at one rank (strips: 128)
    assert E.cut_edges(g64, E.partition_vertices(g64, 8, kind="strips")) == (448, 128)
    assert E.cut_edges(g64, E.partition_vertices(g64, 8, kind="bricks")) == (256, 80)
    g3 = E.named_grid((16, 16, 16))
    assert E.cut_edges(g3, E.partition_vertices(g3, 8, kind="strips")) == (1792, 512)
    assert E.cut_edges(g3, E.partition_vertices(g3, 8)) == (768, 192)  # auto = 2 x 2 x 2 bricks


def test_gate_exchange_plan_is_symmetric():
    # cut-edge two-site gates (itn_apply2 on a partitioned network): what one rank sends is what its peer expects
    sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
    import itn_b200 as E
    for dims, world, chi in (((8, 8), 4, 16), ((6, 5), 3, 4), ((4, 4, 4), 2, 2), ((64, 64), 8, 16)):
        g = E.named_grid(dims)
        owner = E.partition_vertices(g, world)
        sd = [2] * g.nv
        ed = [chi] * g.ne
        for layer in E.edge_coloring(g):
            pairs = [g.edges[e] for e in layer]
            plans = [E.gate_exchange_plan(g, owner, r, pairs, ed, sd) for r in range(world)]
            ncut = sum(1 for (u, v) in pairs if owner[u] != owner[v])
            assert sum(len(pl["send_C"]) for p in plans for pl in p.values()) == ncut
            for r in range(world):
                for peer, pl in plans[r].items():
                    assert pl["send_C"] == plans[peer][r]["recv_C"] and pl["recv_C"] == plans[peer][r]["send_C"]
                    assert pl["send_T"] == plans[peer][r]["recv_T"] and pl["recv_T"] == plans[peer][r]["send_T"]
    # 64x64 chi=16 d=2 ComplexF64: C is 32 x 32 (16 KiB), T is 32 x (2 * 64) (64 KiB) per cut gate
    g = E.named_grid((64, 64))
    owner = E.partition_vertices(g, 8)
    lay = max(E.edge_coloring(g), key=lambda l: sum(owner[g.edges[e][0]] != owner[g.edges[e][1]] for e in l))
    pl = E.gate_exchange_plan(g, owner, 3, [g.edges[e] for e in lay], [16] * g.ne, [2] * g.nv)
    sizes = {c for p in pl.values() for (_, c) in p["send_C"] + p["recv_C"]}
    assert sizes == {2 * 32 * 32}


def _gate_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "itensornetworks.jl_b200"))
        import torch
        import torch.distributed as dist

        import itn_b200 as E
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        g = E.named_grid((6, 4))
        owner = E.partition_vertices(g, world)
        ok = True
        for layer in E.edge_coloring(g):
            pairs = [g.edges[e] for e in layer]
            plan = E.gate_exchange_plan(g, owner, rank, pairs, [3] * g.ne, [2] * g.nv)
            for what_s, what_r in (("send_C", "recv_C"), ("send_T", "recv_T")):
                reqs, bufs = [], []
                for peer, pl in sorted(plan.items()):
                    if pl[what_s]:  # segment of gate i is filled with the marker 1000 * i + sender rank
                        sb = torch.cat([torch.full((n,), 1000.0 * i + rank, dtype=torch.float64) for i, n in pl[what_s]])
                        reqs.append(dist.isend(sb, peer))
                    if pl[what_r]:
                        rb = torch.empty(sum(n for _, n in pl[what_r]), dtype=torch.float64)
                        reqs.append(dist.irecv(rb, peer))
                        bufs.append((peer, pl[what_r], rb))
                for r in reqs:
                    r.wait()
                for peer, segs, rb in bufs:
                    off = 0
                    for i, n in segs:
                        ok = ok and bool((rb[off:off + n] == 1000.0 * i + peer).all())
                        off += n
        q.put((rank, ok))
        dist.destroy_process_group()
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))


def test_gate_exchange_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, ok in res:
        assert ok is True, (rank, ok)


def test_inner_network_host_layout_matches_oracle():
    """inner_network (host mirror): zero-padding of unequal bond dimensions (src/inner.jl:139-171) agrees with the oracle's
    restatement; no device call is involved.  The operator layer of <phi|A|psi> is contracted on the device
    (tests/test_gpu_inner.py::test_operator_layer_is_contracted_on_the_device)."""
    import numpy as np

    import itn_b200 as E
    from oracle import itn_oracle as O
    g = O.random_tree_graph(7, seed=3)
    phi = O.random_network(g, [2, 3, 2, 1, 2, 3], dtype=np.complex128, seed=1)
    psi = O.random_network(g, [3, 2, 3, 2, 2, 1], dtype=np.complex128, seed=2)
    rng = np.random.default_rng(0)
    ops = [rng.standard_normal((2, 2) + (2,) * len(g.inc[v])) + 0j for v in range(g.nv)]
    eg = E.NamedGraph(g.nv, g.edges)
    ket, bra = E.inner_network(E.ITensorNetwork(eg, phi.tensors), E.ITensorNetwork(eg, psi.tensors))
    ref2 = O.bilinear_network(phi, psi)
    for v in range(g.nv):
        assert ket.tensors[v].shape == bra.tensors[v].shape == ref2.tensors[v].shape
        assert np.array_equal(ket.tensors[v], ref2.tensors[v]) and np.array_equal(bra.tensors[v], ref2.bra[v])
    ref = O.bilinear_network(phi, O.apply_operator_network(O.Network(g, ops, np.complex128), psi))
    # BP on the tree (oracle) reproduces the brute-force <phi|A|psi>
    msgs, _, _ = O.bp_update(ref, {}, seq=O.default_edge_sequence(g), maxiter=1)
    exact = O.exact_inner_operator(phi, O.Network(g, ops, np.complex128), psi)
    assert abs(O.scalar(ref, msgs) - exact) < 1e-10 * abs(exact)
