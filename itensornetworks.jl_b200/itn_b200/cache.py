"""Host mirror of the reference's BP-cache API, forwarding to libitn_b200 through the C ABI.

Mirrors (same names, argument meaning and error behaviour):
  BeliefPropagationCache            src/caches/beliefpropagationcache.jl:13-35
  update / update_message / updated_message / message(s) / set_message
                                    src/caches/abstractbeliefpropagationcache.jl:173-337
  environment                       src/caches/beliefpropagationcache.jl:100-105
  region_scalar / vertex_scalars / edge_scalars / logscalar / scalar
                                    beliefpropagationcache.jl:107-119, abstract :83-97,:397-412
  rescale / normalize               abstract :349-395, src/normalize.jl:13-80
  expect                            src/expect.jl:5-109
  apply                             src/apply.jl:97-160
The reference's mutators are out-of-place (every one returns a fresh cache); the mirror keeps that
default and offers `inplace=True` where copying device state would dominate.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ITNError, check, i32, lib
from .graphs import default_edge_sequence, tree_gauge_sequence
from .network import ITensorNetwork

_DTYPE_CODE = {np.dtype(np.float64): 0, np.dtype(np.complex128): 1}
# The engine computes in Float64 / ComplexF64.  The reference's tests also run Float32 / ComplexF32 networks and pin that
# the element type is preserved (test_belief_propagation.jl:54, test_normalize.jl:59): single-precision inputs are widened
# at the boundary and everything handed back (messages, factors, scalars, observables) is narrowed to the input's type.
_COMPUTE = {np.dtype(np.float32): np.dtype(np.float64), np.dtype(np.complex64): np.dtype(np.complex128),
            np.dtype(np.float64): np.dtype(np.float64), np.dtype(np.complex128): np.dtype(np.complex128)}


class Context:
    """One per process and GPU (itn_ctx)."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib().itn_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.world = 0, 1  # set by itn_b200.init_distributed

    @classmethod
    def group(cls, devices):
        """Single-process multi-GPU: one context per listed device, rank i of n = position in the list, communicators from
        one ncclCommInitAll (itn_ctx_create_group).  Collective calls (update, apply on cut edges, rdm2 on cut edges, ...)
        must be issued by one host thread per context."""
        devices = [int(d) for d in devices]
        arr, pd = i32(devices)
        hs = (C.c_void_p * len(devices))()
        check(lib().itn_ctx_create_group(len(devices), pd, hs))
        out = []
        for i, d in enumerate(devices):
            c = cls.__new__(cls)
            c.h = C.c_void_p(hs[i])
            c.device = d
            c.rank, c.world = i, len(devices)
            out.append(c)
        return out

    def sync(self):
        check(lib().itn_ctx_sync(self.h))

    def launch_count(self):
        n = C.c_int64()
        check(lib().itn_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def path_counts(self):
        """Message updates computed so far by (tile path, block path, shape-generic kernels)."""
        out = (C.c_int64 * 3)()
        check(lib().itn_ctx_path_counts(self.h, out))
        return tuple(int(x) for x in out)

    def cholqr2_count(self):
        """Gate sides whose R factor took the second Cholesky pass (ill-conditioned sites, csrc/itn_linalg.cu)."""
        n = C.c_int64()
        check(lib().itn_ctx_cholqr2_count(self.h, C.byref(n)))
        return n.value

    def set_path(self, mode):
        """0 = auto (DMMA tile path where it applies), 1 = shape-generic DMMA kernels only, 2 = FMA kernels only."""
        check(lib().itn_ctx_set_path(self.h, int(mode)))

    def close(self):
        if getattr(self, "h", None):
            lib().itn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class BeliefPropagationCache:
    """BP cache of <psi|psi> with the default one-site partition; state lives on the device."""

    def __init__(self, psi: ITensorNetwork = None, ctx: Context = None, messages="identity", owner=None, dist=None,
                 defer_upload=False, bra: ITensorNetwork = None, partitioned_vertices=None, _handle=None, _like=None):
        if _handle is not None:  # clone
            self.ctx, self.graph, self.dtype, self.h = _like.ctx, _like.graph, _like.dtype, _handle
            self.eltype = _like.eltype
            self._host_refs = None
            self.sdims = list(_like.sdims)
            self.owner, self.rank = _like.owner, _like.rank
            self.bilinear = getattr(_like, "bilinear", False)
            self._edims = None if _like._edims is None else list(_like._edims)
            self.partition = _like.partition
            return
        self.ctx = ctx or default_context()
        # partitioned_vertices (src/caches/beliefpropagationcache.jl:20-35): several sites per partition.  The sites of a
        # partition are merged into one super-site (device contractions, partitions.py); everything below then sees the
        # super-site network, self.partition maps original vertices to (partition, position).
        self.partition = None
        if partitioned_vertices is not None and any(len(grp) > 1 for grp in partitioned_vertices):
            from .partitions import partitioned_network
            assert bra is None and owner is None, "multi-site partitions: single GPU, quadratic form"
            psi, self.partition = partitioned_network(psi, partitioned_vertices, self.ctx)
        self.graph = psi.graph
        self.eltype = np.dtype(psi.dtype)      # what the caller sees
        self.dtype = _COMPUTE[self.eltype]     # what the device computes in
        g = self.graph
        self.sdims = [t.shape[0] for t in psi.tensors]
        # bond dimensions from the tensor shapes, one pass over the vertices (psi.edge_dim(e) per edge costs a list search each)
        ed = [0] * g.ne
        for inc_v, t in zip(g.inc, psi.tensors):
            shp = t.shape
            for k, e in enumerate(inc_v):
                ed[e] = shp[1 + k]
        self._edims = ed
        # multi-GPU: owner[v] = rank that stores vertex v; dist = (rank, nranks) of this process
        self.owner = None if owner is None else [int(x) for x in owner]
        self.rank = self.ctx.rank if dist is None else int(dist[0])
        a0, p0 = i32([u for u, _ in g.edges])
        a1, p1 = i32([v for _, v in g.edges])
        a2, p2 = i32(self._edims)
        a3, p3 = i32(self.sdims)
        h = C.c_void_p()
        a4, p4 = i32(self.owner) if self.owner is not None else (None, None)
        check(lib().itn_net_create(self.ctx.h, _DTYPE_CODE[self.dtype], g.nv, g.ne, p0, p1, p2, p3, p4, C.byref(h)))
        self.h = h
        # all site tensors in one call (pipelined copy + import).  defer_upload=True only registers the host arrays
        # (kept alive in self._host_refs, and they must not be modified) and lets the first synchronous update()
        # overlap the copy with its sweep (ITN_HOST_DEFERRED, include/itn_b200.h).
        mine = [v for v in range(g.nv) if self.owner is None or self.owner[v] == self.rank]
        self.set_factors(mine, [psi.tensors[v] for v in mine], defer=defer_upload)
        # BilinearFormNetwork <bra|psi> (src/formnetworks/bilinearformnetwork.jl:23-42): explicit bra layer, same shapes
        self.bilinear = bra is not None
        if bra is not None:
            for v in mine:
                self.set_bra_factor(v, bra.tensors[v])
        # initialize_cache (src/initialize_cache.jl:14-29): identity messages on loopy graphs, none on trees
        if messages == "identity" or (messages == "default" and not g.is_tree()):
            check(lib().itn_msg_set_identity(self.h))
        elif isinstance(messages, dict):
            for (u, v), m in messages.items():
                self.set_message((u, v), m)

    # -- lifetime ---------------------------------------------------------------------------
    def copy(self):
        h = C.c_void_p()
        check(lib().itn_net_clone(self.h, C.byref(h)))
        return BeliefPropagationCache(_handle=h, _like=self)

    def close(self):
        if getattr(self, "h", None):
            lib().itn_net_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().itn_sync(self.h))
        self._host_refs = None

    # -- factors ------------------------------------------------------------------------------
    def edge_dim(self, e):
        # host copy of the library's edge dims (they change only through the gate calls below, which return them)
        ed = self._edims
        if ed is None:
            ed = self._edims = [0] * self.graph.ne
            d = C.c_int32()
            for f in range(self.graph.ne):
                check(lib().itn_net_edge_dim(self.h, f, C.byref(d)))
                ed[f] = d.value
        return ed[int(e)]

    def _note_newdims(self, eids, newdim):
        if self._edims is not None:
            for e, k in zip(eids, newdim):
                self._edims[e] = int(k)

    def _shape(self, v):
        return (self.sdims[v],) + tuple(self.edge_dim(e) for e in self.graph.inc[v])

    def set_factor(self, v, t):
        """bpc[v] = t (ket; the bra is dag(prime(ket)) implicitly, test_belief_propagation.jl:41-45)."""
        t = np.asarray(t, dtype=self.dtype)
        if t.shape != self._shape(v):
            raise ITNError(2, f"tensor of vertex {v} has shape {t.shape}, expected {self._shape(v)}")
        f = np.asfortranarray(t)  # column-major bytes with axes [site, bonds...]
        check(lib().itn_net_set_tensor(self.h, int(v), f.ctypes.data_as(C.c_void_p), f.ndim, None))

    def set_bra_factor(self, v, t):
        """bra tensor phi_v of a bilinear form (passed un-conjugated; the engine applies `dag`)."""
        t = np.asfortranarray(np.asarray(t, dtype=self.dtype))
        assert t.shape == self._shape(v), f"bra tensor {v} has shape {t.shape}, expected {self._shape(v)}"
        check(lib().itn_net_set_bra_tensor(self.h, v, t.ctypes.data_as(C.c_void_p), t.ndim, None))
        self.bilinear = True

    def set_factors(self, verts, tensors, defer=False):
        """itn_net_set_tensors: many site tensors in one pipelined upload (axes [site, bonds...])."""
        hosts, addrs = [], []
        dt = self.dtype
        from_buffer, addressof = C.c_char.from_buffer, C.addressof
        if self._edims is None:
            self.edge_dim(0) if self.graph.ne else None  # fills the host copy of the bond dimensions
        ed, sd, inc = self._edims, self.sdims, self.graph.inc
        for v, t in zip(verts, tensors):
            # fast path: an F-ordered array of the device dtype is passed as it is (no per-tensor conversion calls)
            if not (type(t) is np.ndarray and t.dtype == dt and t.flags.f_contiguous):
                t = np.asfortranarray(np.asarray(t, dtype=dt))
            if t.shape != (sd[v], *[ed[e] for e in inc[v]]):
                raise ITNError(2, f"tensor of vertex {v} has shape {t.shape}, expected {self._shape(v)}")
            hosts.append(t)
            try:  # address of the first element; ndarray.ctypes costs 2 us per tensor (8 ms for a 64 x 64 lattice)
                addrs.append(addressof(from_buffer(t.T)))
            except (TypeError, ValueError, BufferError):  # read-only or empty buffer
                addrs.append(t.ctypes.data)
        n = len(hosts)
        _, pv = i32(list(verts))
        ptrs = (C.c_void_p * max(n, 1))(*addrs)
        check(lib().itn_net_set_tensors(self.h, n, pv, ptrs, None, None, 1 if defer else 0))
        self._host_refs = hosts if defer else None

    def factor(self, v):
        shape = self._shape(v)
        out = np.empty(shape, dtype=self.dtype, order="F")
        check(lib().itn_net_get_tensor(self.h, int(v), out.ctypes.data_as(C.c_void_p), len(shape), None))
        return np.ascontiguousarray(out).astype(self.eltype, copy=False)

    def tensornetwork(self):
        return ITensorNetwork(self.graph, [self.factor(v) for v in range(self.graph.nv)], self.eltype)

    # -- messages -----------------------------------------------------------------------------
    def message(self, edge):
        u, v = edge
        chi = self.edge_dim(self.graph.eid[(u, v)])
        out = np.empty((chi, chi), dtype=self.dtype, order="F")
        check(lib().itn_msg_get(self.h, int(u), int(v), out.ctypes.data_as(C.c_void_p)))
        return np.ascontiguousarray(out).astype(self.eltype, copy=False)

    def set_message(self, edge, m):
        u, v = edge
        f = np.asfortranarray(np.asarray(m, dtype=self.dtype))
        check(lib().itn_msg_set(self.h, int(u), int(v), f.ctypes.data_as(C.c_void_p)))

    def last_timing(self):
        """Device time of the last update() in ms (CUDA events on the library stream) and the share of the
        contraction kernels."""
        a, b = C.c_double(), C.c_double()
        check(lib().itn_bp_last_timing(self.h, C.byref(a), C.byref(b)))
        return {"total_ms": a.value, "contract_ms": b.value}

    def messages_into(self, out):
        """Download every locally stored message into one host array (directed id order 2e, 2e+1), interleaved complex."""
        check(lib().itn_msg_get_all(self.h, out.ctypes.data_as(C.c_void_p), out.nbytes))

    def messages(self):
        out = {}
        for (u, v) in self.graph.edges:
            out[(u, v)] = self.message((u, v))
            out[(v, u)] = self.message((v, u))
        return out


# ---------------------------------------------------------------------------------------------
# update
# ---------------------------------------------------------------------------------------------


def default_bp_maxiter(bpc):
    """default_bp_maxiter (abstractbeliefpropagationcache.jl:38-42): 1 on trees, otherwise unspecified."""
    return 1 if bpc.graph.is_tree() else None


class EdgeSequence:
    """An edge sequence in the engine's wire format (int32 source / destination arrays, group offsets), built once by
    prepare_sequence: a driver that calls update() with the same schedule every step does not pay the Python-side
    marshalling of 16128 edges (4 ms on the 64 x 64 lattice) per call.  Mirrors prepare_layer for gate layers."""

    def __init__(self, edge_sequence):
        seq = list(edge_sequence)
        self.grouped = len(seq) > 0 and isinstance(seq[0], list)
        if self.grouped:
            flat = [e for grp in seq for e in grp]
            self._ptr, self.gp = i32(np.cumsum([0] + [len(grp) for grp in seq]))
            self.ng = len(seq)
        else:
            flat, self._ptr, self.gp, self.ng = seq, None, None, 0
        self.n = len(flat)
        self._src, self.ps = i32([u for u, _ in flat])
        self._dst, self.pd = i32([v for _, v in flat])


def prepare_sequence(edge_sequence):
    return edge_sequence if isinstance(edge_sequence, EdgeSequence) else EdgeSequence(edge_sequence)


def update(bpc, maxiter="default", tol=None, edge_sequence=None, normalize=True, inplace=False, info=None):
    """update(bpc; alg="bp", maxiter, tol, edge_sequence, message_update_alg=(; normalize)).

    edge_sequence: list of directed edges (sequential sweep), list of edge lists (grouped / parallel sweep), or the
    EdgeSequence made from either by prepare_sequence."""
    if maxiter == "default":
        maxiter = default_bp_maxiter(bpc)
    if maxiter is None:
        raise ITNError(1, "You need to specify a number of iterations for BP!")
    if edge_sequence is None:
        edge_sequence = default_edge_sequence(bpc.graph)
    es = prepare_sequence(edge_sequence)
    out = bpc if inplace else bpc.copy()
    iters, diff = C.c_int32(), C.c_double()
    check(lib().itn_bp_update(out.h, es.ps, es.pd, es.n, es.gp, es.ng, int(maxiter), -1.0 if tol is None else float(tol),
                              1 if normalize else 0, C.byref(iters), C.byref(diff)))
    out._host_refs = None  # deferred host tensors have been consumed
    if info is not None:
        info["iterations"] = iters.value
        info["mean_diff"] = diff.value
    return out


def updated_message(bpc, edge, normalize=True):
    u, v = edge
    chi = bpc.edge_dim(bpc.graph.eid[(u, v)])
    out = np.empty((chi, chi), dtype=bpc.dtype, order="F")
    check(lib().itn_updated_message(bpc.h, int(u), int(v), 1 if normalize else 0, out.ctypes.data_as(C.c_void_p)))
    return np.ascontiguousarray(out).astype(bpc.eltype, copy=False)


def update_message(bpc, edge, normalize=True):
    out = bpc.copy()
    out.set_message(edge, updated_message(bpc, edge, normalize))
    return out


def message(bpc, edge):
    return bpc.message(edge)


def message_diff(a, b):
    """message_diff (abstractbeliefpropagationcache.jl:32-36) for two host messages."""
    a = np.asarray(a)
    b = np.asarray(b)
    return 1.0 - abs(np.vdot(a / np.linalg.norm(a), b / np.linalg.norm(b))) ** 2


def message_residuals(bpc, edges=None):
    """message_diff(updated_message(bpc, e), message(bpc, e)) for every directed edge, on the device."""
    if edges is None:
        edges = list(bpc.graph.edges) + [(v, u) for u, v in bpc.graph.edges]
    a, ps = i32([u for u, _ in edges])
    b, pd = i32([v for _, v in edges])
    out = np.zeros(len(edges), dtype=np.float64)
    check(lib().itn_message_residuals(bpc.h, ps, pd, len(edges), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def environment(bpc, verts):
    """environment(bpc, verts) for whole-site vertex sets: the incoming messages on the boundary
    of `verts` (edges inside the set are excluded).  Returns [((u, v), M_{u->v}), ...]."""
    vs = set(int(v) for v in verts)
    if bpc.partition is not None:  # messages into the partitions that contain `verts` (fused bond indices)
        vs = set(bpc.partition.group_of[v] for v in vs)
    out = []
    for v in sorted(vs):
        for e in bpc.graph.inc[v]:
            u = bpc.graph.other(e, v)
            if u not in vs:
                out.append(((u, v), bpc.message((u, v))))
    return out


# ---------------------------------------------------------------------------------------------
# scalars, rescale, normalize
# ---------------------------------------------------------------------------------------------


def scalar_factors_quotient(bpc):
    zv = np.empty(bpc.graph.nv, dtype=bpc.dtype)
    ze = np.empty(max(bpc.graph.ne, 1), dtype=bpc.dtype)
    check(lib().itn_region_scalars(bpc.h, zv.ctypes.data_as(C.c_void_p), ze.ctypes.data_as(C.c_void_p)))
    return zv.astype(bpc.eltype, copy=False), ze[: bpc.graph.ne].astype(bpc.eltype, copy=False)


def vertex_scalars(bpc):
    return scalar_factors_quotient(bpc)[0]


def edge_scalars(bpc):
    return scalar_factors_quotient(bpc)[1]


def region_scalar(bpc, region):
    if isinstance(region, tuple):
        return edge_scalars(bpc)[bpc.graph.eid[region]]
    return vertex_scalars(bpc)[region]


def logscalar(bpc):
    out = (C.c_double * 2)()
    check(lib().itn_logscalar(bpc.h, out))
    return complex(out[0], out[1]) if out[1] != 0.0 else out[0]


def scalar(bpc):
    return np.exp(logscalar(bpc))


def rescale(bpc, inplace=False, verts=None):
    """rescale(bpc; verts) (abstractbeliefpropagationcache.jl:391-395): rescale_messages over every edge, then
    rescale_partitions restricted to `verts` (None = every vertex; a vertex stands for its ket and bra)."""
    out = bpc if inplace else bpc.copy()
    if verts is None:
        check(lib().itn_rescale(out.h))
    else:
        vs = [int(v) for v in verts]
        if bpc.partition is not None:
            vs = sorted(set(bpc.partition.group_of[v] for v in vs))
        _, pv = i32(vs)
        check(lib().itn_rescale_verts(out.h, pv, len(vs)))
    return out


def _cache_for(psi, cache, update_cache, cache_update_kwargs, ctx=None, cache_construction_kwargs=None):
    """The `cache!` / `update_cache` / `cache_update_kwargs` protocol (src/expect.jl:21-41)."""
    if isinstance(psi, BeliefPropagationCache):
        cache, psi = psi, None
        update_cache = False if update_cache is None else update_cache
    if cache is None:
        cache = BeliefPropagationCache(psi, ctx=ctx, messages="default", **(cache_construction_kwargs or {}))
        update_cache = True if update_cache is None else update_cache
    elif update_cache is None:
        update_cache = False
    if update_cache:
        cache = update(cache, inplace=True, **(cache_update_kwargs or {}))
    return cache


def normalize(psi, alg="bp", cache=None, update_cache=None, cache_update_kwargs=None, ctx=None):
    """normalize(psi; alg="bp", cache!, update_cache, cache_update_kwargs) (src/normalize.jl:63-80)."""
    assert alg == "bp", "only alg=\"bp\" runs on the engine"
    cache = _cache_for(psi, cache, update_cache, cache_update_kwargs, ctx)
    check(lib().itn_rescale(cache.h))
    return cache.tensornetwork()


def _pad_to(t, shape, dtype):
    out = np.zeros(shape, dtype=dtype)
    out[tuple(slice(0, n) for n in t.shape)] = t
    return out


def inner_network(phi: ITensorNetwork, psi: ITensorNetwork, operator: ITensorNetwork = None, ctx=None):
    """inner_network(phi, psi) / inner_network(phi, A, psi) (src/inner.jl:139-171): returns (ket, bra) host networks of
    equal shapes for BeliefPropagationCache(ket, bra=bra).  The operator layer (a list of arrays A_v[s', s, b_1..b_z] with
    the bond order of the state) and the ket layer of a vertex form one partition of the three-layer network
    (src/formnetworks/bilinearformnetwork.jl:23-42): A_v is contracted into the ket on the DEVICE (itn_tensordot, the
    same in-partition `contract` the multi-site partitions use) and the pair of bonds (a_k, b_k) becomes the fused bond
    a_k + chi_k b_k of the partition; the host only permutes and reshapes axes.  Differing bond dimensions are zero-padded."""
    g = psi.graph
    ops = None if operator is None else [np.asarray(a) for a in getattr(operator, "tensors", operator)]
    dtype = np.result_type(phi.dtype, psi.dtype, *([a.dtype for a in ops] if ops else []))
    if ops is not None:
        from .partitions import tensordot as device_tensordot
        ctx = ctx or default_context()
    kets = []
    for v in range(g.nv):
        t = psi.tensors[v]
        if ops is not None:
            a = ops[v]
            z = t.ndim - 1
            r = device_tensordot(a, t, [1], [0], ctx)  # [s', b_1..b_z, a_1..a_z], contracted by the engine
            r = np.transpose(r, [0] + [i for k in range(z) for i in (1 + k, 1 + z + k)])  # [s', b_1, a_1, b_2, a_2, ..]
            t = r.reshape([r.shape[0]] + [r.shape[1 + 2 * k] * r.shape[2 + 2 * k] for k in range(z)])  # C order: a fastest
        kets.append(t)
    ks, bs = [], []
    for v in range(g.nv):
        shape = tuple(max(x, y) for x, y in zip(kets[v].shape, phi.tensors[v].shape))
        ks.append(_pad_to(kets[v], shape, dtype))
        bs.append(_pad_to(phi.tensors[v], shape, dtype))
    return ITensorNetwork(g, ks, dtype), ITensorNetwork(g, bs, dtype)


def loginner(phi, psi, operator=None, alg="bp", cache=None, update_cache=None, cache_update_kwargs=None, ctx=None,
             messages="default_bilinear"):
    """loginner(phi, psi; alg = "bp") / loginner(phi, A, psi; alg = "bp") (src/inner.jl:100-137) = logscalar of the BP
    cache of the bilinear form network (src/contract.jl:41-58).  As in the reference the cache of a bilinear form has no
    default messages (initialize_cache fallback, src/initialize_cache.jl:10-12): on trees the forest-cover sequence
    creates them; on loopy graphs pass `messages="identity"` or a dict, and `maxiter` in cache_update_kwargs."""
    assert alg == "bp", "only alg=\"bp\" runs on the engine"
    if cache is None:
        ket, bra = inner_network(phi, psi, operator, ctx=ctx)
        cache = BeliefPropagationCache(ket, ctx=ctx, bra=bra, messages=None if messages == "default_bilinear" else messages)
        update_cache = True if update_cache is None else update_cache
    elif update_cache is None:
        update_cache = False
    if update_cache:
        cache = update(cache, inplace=True, **(cache_update_kwargs or {}))
    return logscalar(cache)


def inner(phi, psi, operator=None, **kwargs):
    """inner(phi, psi; alg = "bp") / inner(phi, A, psi; alg = "bp") (src/inner.jl:139-171): exp(loginner)."""
    return np.exp(loginner(phi, psi, operator, **kwargs))


def norm_sqr(psi, cache=None, update_cache=None, cache_update_kwargs=None, ctx=None):
    cache = _cache_for(psi, cache, update_cache, cache_update_kwargs, ctx)
    return scalar(cache)


# ---------------------------------------------------------------------------------------------
# observables
# ---------------------------------------------------------------------------------------------

_OPS = {
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Sz": 0.5 * np.array([[1, 0], [0, -1]], dtype=np.complex128),
    "Sx": 0.5 * np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Id": np.eye(2, dtype=np.complex128),
}


def op(name, dtype=np.complex128):
    m = _OPS[name]
    if np.dtype(dtype).kind != "c":
        assert np.all(m.imag == 0), f"operator {name} is not real"
        return m.real.astype(dtype)
    return m.astype(dtype)


def expect(psi, operator, vertices=None, alg="bp", cache=None, update_cache=None, cache_update_kwargs=None, ctx=None,
           cache_construction_kwargs=None):
    """expect(psi, op, vertices; alg="bp", cache!, update_cache, cache_update_kwargs, cache_construction_kwargs)
    (src/expect.jl:58-109); cache_construction_kwargs = {"partitioned_vertices": [[v, ...], ...]} groups several sites
    per BP partition (test/test_expect.jl:22-39).

    `operator` is a name ("Sz", "Z", ...) or a d x d matrix O[s_out, s_in]. Returns {vertex: value}."""
    assert alg == "bp", "only alg=\"bp\" runs on the engine"
    cache = _cache_for(psi, cache, update_cache, cache_update_kwargs, ctx, cache_construction_kwargs)
    o = op(operator, cache.dtype) if isinstance(operator, str) else np.asarray(operator, dtype=cache.dtype)
    if cache.partition is not None:
        # the operator acts on one factor of the fused site index of its partition
        pm = cache.partition
        if vertices is None:
            vertices = sorted(pm.group_of)
        vertices = [int(v) for v in vertices]
        packed = [np.asfortranarray(pm.lift_operator(v, o).astype(cache.dtype)).ravel(order="F") for v in vertices]
        ops = np.ascontiguousarray(np.concatenate(packed)) if packed else np.zeros(0, dtype=cache.dtype)
        out = np.empty(len(vertices), dtype=cache.dtype)
        _, pv = i32([pm.group_of[v] for v in vertices])
        check(lib().itn_expect1(cache.h, pv, len(vertices), ops.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        out = out.astype(cache.eltype, copy=False)
        return {v: out[i] for i, v in enumerate(vertices)}
    if vertices is None:
        vertices = list(range(cache.graph.nv))
    ops = np.stack([np.asfortranarray(o).ravel(order="F")] * len(vertices)) if len(vertices) else np.zeros((0, 4))
    ops = np.ascontiguousarray(ops, dtype=cache.dtype)
    out = np.empty(len(vertices), dtype=cache.dtype)
    _, pv = i32(vertices)
    check(lib().itn_expect1(cache.h, pv, len(vertices), ops.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    out = out.astype(cache.eltype, copy=False)
    return {int(v): out[i] for i, v in enumerate(vertices)}


def rdm2(bpc, edges):
    """Two-site reduced density matrices from the BP environment (test_belief_propagation.jl:64-91)."""
    eids = [bpc.graph.eid[tuple(e)] if not np.isscalar(e) else int(e) for e in edges]
    ds = [bpc.sdims[bpc.graph.edges[e][0]] * bpc.sdims[bpc.graph.edges[e][1]] for e in eids]
    tot = sum(d * d for d in ds)
    out = np.empty(tot, dtype=bpc.dtype)
    _, pe = i32(eids)
    check(lib().itn_rdm2(bpc.h, pe, len(eids), out.ctypes.data_as(C.c_void_p)))
    res, off = [], 0
    for d in ds:
        res.append(out[off:off + d * d].reshape(d, d, order="F").astype(bpc.eltype))
        off += d * d
    return res


def expect2(bpc, edges, op_u, op_v):
    """<O_u O_v> on edges (u, v) = tr(rho_uv (O_u (x) O_v)); row index of rho = s_u + d_u * s_v."""
    ou = op(op_u, bpc.dtype) if isinstance(op_u, str) else np.asarray(op_u)
    ov = op(op_v, bpc.dtype) if isinstance(op_v, str) else np.asarray(op_v)
    kron = np.kron(ov, ou)
    return [np.trace(r @ kron) for r in rdm2(bpc, edges)]


# ---------------------------------------------------------------------------------------------
# gates
# ---------------------------------------------------------------------------------------------


def gauge_walk(bpc, edges, inplace=False):
    """gauge_walk(tn, edges) (src/abstractitensornetwork.jl:387-393): qr!(tn, u => v) for every (u, v) of `edges`, on the
    device (itn_gauge_walk).  The messages of the cache are left as they are (they belong to the old gauge)."""
    out = bpc if inplace else bpc.copy()
    edges = [(int(u), int(v)) for u, v in edges]
    for u, v in edges:
        if not bpc.graph.has_edge(u, v):
            raise ITNError(1, "Edge not in graph.")
    a_s, ps = i32([u for u, _ in edges])
    a_d, pd = i32([v for _, v in edges])
    check(lib().itn_gauge_walk(out.h, ps, pd, len(edges)))
    del a_s, a_d
    return out


def tree_gauge(bpc, region, inplace=False):
    """tree_gauge(psi, region) (src/abstractitensornetwork.jl:407-418): move the gauge of the whole network towards
    `region` (a vertex or a list of vertices), treating the network as the tree spanned by a spanning tree."""
    return gauge_walk(bpc, tree_gauge_sequence(bpc.graph, region), inplace=inplace)


tree_orthogonalize = tree_gauge  # src/abstractitensornetwork.jl:420


def apply(gate, bpc, verts, maxdim=None, cutoff=None, normalize=False, callback=None, inplace=False, msg_mode=0,
          ortho=False):
    """apply(o, psi; envs, maxdim, cutoff, normalize, ortho, callback) (src/apply.jl:97-146) on a BP cache:
    the product environment is the cache's current messages.  `verts` = (v,) or (v1, v2).  ortho=True first gauges the
    network towards verts[0] (tree_orthogonalize, :109-111 and :130-132); as in the reference the environments are whatever
    the caller holds (the default `envs = ITensor[]` of the reference is a cache with identity messages)."""
    verts = tuple(int(v) for v in verts)
    if bpc.partition is not None:
        raise ITNError(1, "`apply` requires a product environment (`envs` with no shared edges); the cache groups several "
                          "sites per partition. Contract `envs` to product form before calling.")
    if len(verts) == 2 and not bpc.graph.has_edge(*verts):
        raise ITNError(1, "Vertices where the gates are being applied must be neighbors for now.")
    out = bpc if inplace else bpc.copy()
    if ortho and len(verts) in (1, 2):
        tree_gauge(out, verts[0], inplace=True)
    if len(verts) == 1:
        g = np.asfortranarray(np.asarray(gate, dtype=bpc.dtype))
        _, pv = i32(verts)
        check(lib().itn_apply1(out.h, pv, 1, g.ctypes.data_as(C.c_void_p), 1 if normalize else 0))
        return out
    if len(verts) == 2:
        if not bpc.graph.has_edge(*verts):
            raise ITNError(1, "Vertices where the gates are being applied must be neighbors for now.")
        info = apply_layer([gate], out, [verts], maxdim=maxdim, cutoff=cutoff, normalize=normalize, msg_mode=msg_mode)
        if callback is not None:
            callback(singular_values=info["singular_values"][0], truncation_error=info["truncation_error"][0])
        return out
    if len(verts) < 1:
        raise ITNError(1, "Gate being applied does not share indices with tensor network.")
    raise ITNError(1, "Gates with more than 2 sites is not supported yet.")


def apply_layer(gates, bpc, pairs=None, maxdim=None, cutoff=None, normalize=False, msg_mode=0):
    """A vertex-disjoint layer of two-site gates in one batched call (in place).
    gates[i][s1', s2', s1, s2] acts on pairs[i] = (v1, v2)."""
    if isinstance(gates, GateLayer):  # packed once by prepare_layer (a Trotter driver applies the same layers every step)
        eids, packed, pe = gates.eids, gates.packed, gates.pe
    else:
        eids, packed = _pack_gates(bpc, gates, pairs)
        packed = np.concatenate(packed) if eids else np.zeros(0, dtype=bpc.dtype)
        _, pe = i32(eids)
    n = len(eids)
    dmax = max(bpc.sdims) if bpc.sdims else 1
    stride = max([dmax * dmax * bpc.edge_dim(e) for e in eids] + [1])
    newdim = np.zeros(n, dtype=np.int32)
    terr = np.zeros(n, dtype=np.float64)
    sv = np.zeros((n, stride), dtype=np.float64)
    check(lib().itn_apply2(bpc.h, pe, n, packed.ctypes.data_as(C.c_void_p), 0 if maxdim is None else int(maxdim),
                           -1.0 if cutoff is None else float(cutoff), 1 if normalize else 0, int(msg_mode),
                           newdim.ctypes.data_as(C.POINTER(C.c_int32)), terr.ctypes.data_as(C.POINTER(C.c_double)),
                           sv.ctypes.data_as(C.POINTER(C.c_double)), stride))
    bpc._note_newdims(eids, newdim)
    return {"newdim": newdim, "truncation_error": terr, "singular_values": _SvalRows(sv, newdim)}


class GateLayer:
    """A vertex-disjoint layer of two-site gates in the engine's wire format (edge ids + gates packed in (esrc, edst)
    orientation), built once by prepare_layer and passed to apply_layer in place of the gate list."""

    def __init__(self, eids, packed):
        self.eids = list(eids)
        self.packed = np.ascontiguousarray(packed)
        self._ids, self.pe = i32(self.eids)


def prepare_layer(bpc, gates, pairs):
    eids, packed = _pack_gates(bpc, gates, pairs)
    return GateLayer(eids, np.concatenate(packed) if eids else np.zeros(0, dtype=bpc.dtype))


class _SvalRows:
    """singular_values[i] = the kept singular values of gate i (a view into the padded result rows, sliced on access:
    a colour layer of the 64 x 64 lattice has 2048 gates and most callers read a few of them)."""

    def __init__(self, sv, newdim):
        self.sv, self.newdim = sv, newdim

    def __len__(self):
        return len(self.newdim)

    def __getitem__(self, i):
        return self.sv[i, :self.newdim[i]]

    def __iter__(self):
        return (self.sv[i, :k] for i, k in enumerate(self.newdim))


def _pack_gates(bpc, gates, pairs):
    g = bpc.graph
    eids, packed, memo = [], [], {}
    for gate, (v1, v2) in zip(gates, pairs):
        e = g.eid.get((v1, v2))
        if e is None:
            raise ITNError(1, "Vertices where the gates are being applied must be neighbors for now.")
        d1, d2 = bpc.sdims[v1], bpc.sdims[v2]
        flip = g.edges[e] != (v1, v2)
        key = (id(gate), d1, d2, flip)
        if key not in memo:
            gt = np.asarray(gate, dtype=bpc.dtype).reshape(d1, d2, d1, d2)
            if flip:
                gt = gt.transpose(1, 0, 3, 2)
            memo[key] = np.asfortranarray(gt).ravel(order="F")
        eids.append(e)
        packed.append(memo[key])
    return eids, packed


def tebd_step(bpc, layers, maxdim=None, cutoff=None, normalize=False, msg_mode=0, bp_maxiter=0, bp_tol=None,
              edge_sequence=None, info=None):
    """One Trotter step in ONE library call (itn_apply_layers): `layers` is a list of (gates, pairs) colour layers, each a
    vertex-disjoint batch of two-site gates; after every layer `bp_maxiter` BP sweeps over `edge_sequence` (list of
    single-edge groups = synchronous sweep, list of edges = sequential) refresh the environments of the next layer.
    Mirrors the host loop `psi = apply(o, psi; envs...); bpc = update(bpc; cache_update_kwargs...)` of a TEBD driver
    (src/apply.jl:97-160 with the cache protocol of src/expect.jl:21-41).  In place; returns per-gate results."""
    all_e, all_g, ptr = [], [], [0]
    for gates, pairs in layers:
        e, g = _pack_gates(bpc, gates, pairs)
        all_e += e
        all_g += g
        ptr.append(len(all_e))
    n = len(all_e)
    packed = np.ascontiguousarray(np.concatenate(all_g)) if n else np.zeros(0, dtype=bpc.dtype)
    dmax = max(bpc.sdims) if bpc.sdims else 1
    stride = max([dmax * dmax * bpc.edge_dim(e) for e in all_e] + [1])
    if maxdim is not None:
        stride = max(stride, dmax * dmax * int(maxdim))  # bonds may grow between the layers of the step
    newdim = np.zeros(n, dtype=np.int32)
    terr = np.zeros(n, dtype=np.float64)
    sv = np.zeros((n, stride), dtype=np.float64)
    es = prepare_sequence([] if edge_sequence is None else edge_sequence)  # keeps its index arrays alive during the call
    ps, pd, gp, grouped = es.ps, es.pd, es.gp, es.grouped
    a_e, pe = i32(all_e)
    a_l, pl = i32(ptr)
    iters = C.c_int32()
    check(lib().itn_apply_layers(bpc.h, len(layers), pl, pe, packed.ctypes.data_as(C.c_void_p),
                                 0 if maxdim is None else int(maxdim), -1.0 if cutoff is None else float(cutoff),
                                 1 if normalize else 0, int(msg_mode), ps, pd, es.n, gp,
                                 es.ng if grouped else 0, int(bp_maxiter), -1.0 if bp_tol is None else float(bp_tol),
                                 1, newdim.ctypes.data_as(C.POINTER(C.c_int32)), terr.ctypes.data_as(C.POINTER(C.c_double)),
                                 sv.ctypes.data_as(C.POINTER(C.c_double)), stride, C.byref(iters)))
    del a_e, a_l
    bpc._host_refs = None
    bpc._note_newdims(all_e, newdim)  # layers in order: the last gate on an edge wins
    if info is not None:
        info["bp_iterations"] = iters.value
    return {"newdim": newdim, "truncation_error": terr, "layer_ptr": ptr, "singular_values": _SvalRows(sv, newdim)}


def map_eigvals(f, mats, cutoff=None, ctx=None):
    """map_eigvals(f, A, ...; ishermitian=true, cutoff) (src/apply.jl:21-25) for a batch of Hermitian
    matrices; f in {"sqrt", "invsqrt", "inv"}."""
    ctx = ctx or default_context()
    mats = np.asarray(mats)
    dtype = np.dtype(np.complex128 if mats.dtype.kind == "c" else np.float64)
    single = mats.ndim == 2
    m = mats[None] if single else mats
    n, chi = m.shape[0], m.shape[1]
    inp = np.ascontiguousarray(np.stack([np.asfortranarray(x.astype(dtype)).ravel(order="F") for x in m]))
    out = np.empty_like(inp)
    fn = {"sqrt": 0, "invsqrt": 1, "inv": 2}[f]
    check(lib().itn_map_eigvals(ctx.h, _DTYPE_CODE[dtype], fn, chi, n, inp.ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p), -1.0 if cutoff is None else float(cutoff)))
    res = np.stack([o.reshape(chi, chi, order="F") for o in out])
    return res[0] if single else res


def svd_batch(mats, variant=0, want_us=False, ctx=None):
    """Singular values of a batch of matrices through the engine's batched Jacobi kernels (the `factorize_svd` step of
    simple_update_bp, src/apply.jl:81-88).  Returns (sigma[batch, n] descending, U*Sigma or None, device milliseconds);
    variant 1 forces the shape-generic kernel."""
    ctx = ctx or default_context()
    mats = np.asarray(mats)
    dtype = np.dtype(np.complex128 if mats.dtype.kind == "c" else np.float64)
    b, m, n = mats.shape
    inp = np.ascontiguousarray(np.stack([np.asfortranarray(x.astype(dtype)).ravel(order="F") for x in mats]))
    sig = np.zeros((b, n), dtype=np.float64)
    us = np.empty_like(inp) if want_us else None
    ms = C.c_double(0.0)
    check(lib().itn_svd_batch(ctx.h, _DTYPE_CODE[dtype], m, n, b, inp.ctypes.data_as(C.c_void_p),
                              sig.ctypes.data_as(C.POINTER(C.c_double)),
                              us.ctypes.data_as(C.c_void_p) if want_us else None, int(variant), C.byref(ms)))
    if want_us:
        us = np.stack([u.reshape(m, n, order="F") for u in us])
    return sig, us, ms.value
