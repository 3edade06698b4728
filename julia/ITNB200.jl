# ITNB200.jl -- Julia shim over libitn_b200.so: a drop-in `AbstractBeliefPropagationCache` whose site tensors and
# messages live on a B200.
#
# STATUS: NOT EXECUTED.  There is no Julia runtime in the build image or on the GPU boxes, so this file has never been
# loaded.  What is checked (tests/test_abi.py): every `ccall` names an exported symbol with the argument count and the
# pointer / scalar kinds of its C prototype, every method of the abstract-cache interface
# (src/caches/abstractbeliefpropagationcache.jl:44-69) has a B200 method below, and the subtype declaration carries both
# type parameters.  The same C ABI is exercised end to end by the Python ctypes mirror (itensornetworks.jl_b200/itn_b200/),
# which is what the parity tests run.
#
# How the reference reaches this code.  Its front-ends take an algorithm tag and a cache reference:
#     expect(psi, ops; alg, cache!, update_cache, cache_update_kwargs, cache_construction_kwargs)      src/expect.jl:21-41
#     normalize(psi; alg, cache!, ...), rescale(alg, tn; ...)                                          src/normalize.jl:13-80
#     environment(tn, verts; alg, cache!, ...)                                                         src/environment.jl:6-49
#     logscalar(tn; alg, cache!, ...), scalar                                                          src/contract.jl:41-62
# and call `initialize_cache(Algorithm(alg), tn; ...)`, `update(cache; ...)`, `environment(cache, verts)`,
# `rescale(cache; verts)`, `logscalar(cache)` on whatever comes back.  With this module loaded,
#     expect(psi, ops; alg = "bp_b200")        normalize(psi; alg = "bp_b200")        logscalar(norm_sqr_network(psi); alg = "bp_b200")
# build a `B200BeliefPropagationCache` (initialize_cache(::Algorithm"bp_b200", ...)) and every later call lands in the
# methods below, i.e. in `ccall`s; `cache! = Ref(bpc)` hands an existing device cache to any of them.
module ITNB200

using Adapt: Adapt
using DataGraphs: DataGraphs
using Dictionaries: Dictionary, set!
using Graphs: Graphs, dst, edges, is_tree, src, vertices
using ITensors: ITensors, ITensor, Index, array, commonind, commoninds, dag, dim, inds, itensor, noprime, prime
using ITensors.NDTensors: @Algorithm_str, Algorithm
using ITensorNetworks: ITensorNetworks, AbstractBeliefPropagationCache, AbstractITensorNetwork, BeliefPropagationCache,
  ITensorNetwork, QuadraticFormNetwork, bra_vertex, default_edge_sequence, default_partitioned_vertices, ket_network,
  ket_vertex, ket_vertices, operator_vertex, original_state_vertex, siteinds, tensornetwork
using NamedGraphs: NamedGraphs, NamedEdge
using NamedGraphs.PartitionedGraphs: PartitionedGraph, PartitionedGraphs, QuotientEdge, QuotientVertex,
  boundary_quotientedges, quotient_graph, quotientedges, quotientvertices, unpartitioned_graph

const LIB = get(ENV, "ITN_B200_LIB", joinpath(@__DIR__, "..", "itensornetworks.jl_b200", "lib", "libitn_b200.so"))

# src/apply.jl:120-128 and friends raise ErrorException; keep that type so `@test_throws ErrorException` holds.
function check(status::Cint)
  status == 0 && return nothing
  msg = unsafe_string(ccall((:itn_last_error, LIB), Cstring, ()))
  return error(msg)
end

mutable struct Context
  h::Ptr{Cvoid}
  function Context(device::Integer=0; stream::Ptr{Cvoid}=C_NULL)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:itn_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), device, stream, out))
    ctx = new(out[])
    finalizer(c -> ccall((:itn_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    return ctx
  end
  function Context(h::Ptr{Cvoid})   # adopts a handle made by itn_ctx_create_group
    ctx = new(h)
    finalizer(c -> ccall((:itn_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    return ctx
  end
end
const DEFAULT_CONTEXT = Ref{Union{Nothing,Context}}(nothing)
default_context() = (isnothing(DEFAULT_CONTEXT[]) && (DEFAULT_CONTEXT[] = Context()); DEFAULT_CONTEXT[])

dtype_code(::Type{Float64}) = Cint(0)
dtype_code(::Type{ComplexF64}) = Cint(1)
compute_type(::Type{T}) where {T<:Real} = Float64          # Float32 inputs are widened at the boundary
compute_type(::Type{T}) where {T<:Complex} = ComplexF64

"""
BP cache of the norm network <psi|psi> (a `QuadraticFormNetwork`) with the default one-site partition
{(v, "ket"), (v, "bra"), (v, "operator")} (src/formnetworks/abstractformnetwork.jl:89-91).  The ket tensors and the
messages live on the device behind `h`; the bra layer (dag(prime(ket))) and the identity operator layer are implicit
there.  `qf` is the host-side form network: index structure always, tensor DATA only after `sync_host!` (vertices whose
device tensor is newer are listed in `stale`).  Mirrors BeliefPropagationCache (src/caches/beliefpropagationcache.jl:13-35).
"""
mutable struct B200BeliefPropagationCache{V,PV,QF<:QuadraticFormNetwork,PTN<:PartitionedGraph} <:
               AbstractBeliefPropagationCache{V,PV}
  h::Ptr{Cvoid}
  ctx::Context
  qf::QF
  ptn::PTN
  verts::Vector               # original state vertices, id = position - 1
  vid::Dict{Any,Int}
  eds::Vector                 # undirected edges of the ket network, id = position - 1
  eid::Dict{Any,Int}          # (u, v) and (v, u) -> id
  elt::Type                   # element type the caller sees
  celt::Type                  # Float64 or ComplexF64: what the device computes in
  stale::Set{Any}             # original vertices whose host tensors (ket and bra) are older than the device's
end

function finalize_cache!(bpc::B200BeliefPropagationCache)
  finalizer(b -> ccall((:itn_net_destroy, LIB), Cint, (Ptr{Cvoid},), b.h), bpc)
  return bpc
end

ket_tensor(bpc::B200BeliefPropagationCache, v) = tensornetwork(bpc.qf)[ket_vertex(bpc.qf, v)]
link_index(bpc::B200BeliefPropagationCache, e) = commonind(ket_tensor(bpc, src(e)), ket_tensor(bpc, dst(e)))
site_index(bpc::B200BeliefPropagationCache, v) =
  commonind(ket_tensor(bpc, v), tensornetwork(bpc.qf)[operator_vertex(bpc.qf, v)])

# axis_edge[i] = edge id carried by axis i of the ITensor's storage, or -1 for the site index (include/itn_b200.h)
function axis_edges(bpc::B200BeliefPropagationCache, v, t::ITensor)
  s = site_index(bpc, v)
  out = Int32[]
  for i in inds(t)
    if noprime(i) == noprime(s)
      push!(out, -1)
    else
      k = findfirst(e -> (src(e) == v || dst(e) == v) && noprime(link_index(bpc, e)) == noprime(i), bpc.eds)
      isnothing(k) && error("index $i of the tensor at $v is neither its site index nor a link of the network")
      push!(out, k - 1)
    end
  end
  return out
end

function B200BeliefPropagationCache(qf::QuadraticFormNetwork;
  partitioned_vertices=default_partitioned_vertices(qf), messages=nothing, ctx::Context=default_context())
  ptn = PartitionedGraph(qf, partitioned_vertices)
  all(pv -> length(pv) == 3 && length(unique(first.(pv))) == 1, collect(partitioned_vertices)) ||
    error("libitn_b200 runs the default one-site partition {ket, bra, operator} per vertex; multi-site partitions " *
          "are built by the host mirror (itn_b200.partitions) and are not wired into the Julia shim")
  ket = ket_network(qf)
  verts = collect(vertices(ket))
  vid = Dict{Any,Int}(v => i - 1 for (i, v) in enumerate(verts))
  eds = collect(edges(ket))
  eid = Dict{Any,Int}()
  for (k, e) in enumerate(eds)
    eid[(src(e), dst(e))] = k - 1
    eid[(dst(e), src(e))] = k - 1
  end
  elt = promote_type(map(v -> eltype(ket[v]), verts)...)
  celt = compute_type(elt)
  esrc = Int32[vid[src(e)] for e in eds]
  edst = Int32[vid[dst(e)] for e in eds]
  edim = Int32[dim(commonind(ket[src(e)], ket[dst(e)])) for e in eds]
  sdim = Int32[dim(only(siteinds(ket, v))) for v in verts]
  out = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:itn_net_create, LIB), Cint,
    (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
    ctx.h, dtype_code(celt), length(verts), length(eds), esrc, edst, edim, sdim, C_NULL, out))
  V = eltype(collect(vertices(qf)))
  PV = eltype(collect(keys(partitioned_vertices)))
  bpc = B200BeliefPropagationCache{V,PV,typeof(qf),typeof(ptn)}(out[], ctx, qf, ptn, verts, vid, eds, eid, elt, celt, Set{Any}())
  finalize_cache!(bpc)
  # all site tensors in one pipelined upload (itn_net_set_tensors): column-major `array(t)`, axis_edge maps the ITensor
  # index order to edges so nothing is permuted on the host
  arrs = [Array{celt}(array(ket[v])) for v in verts]
  axes = [axis_edges(bpc, v, ket[v]) for v in verts]
  ids = Int32[vid[v] for v in verts]
  nds = Int32[ndims(a) for a in arrs]
  ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
  flat = reduce(vcat, axes; init=Int32[])
  GC.@preserve arrs check(ccall((:itn_net_set_tensors, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Int32}, Cint),
    bpc.h, length(verts), ids, ptrs, nds, flat, 0))
  # initialize_cache (src/initialize_cache.jl:14-29): identity messages on loopy quotient graphs, none on trees
  if isnothing(messages)
    is_tree(quotient_graph(ptn)) ||
      check(ccall((:itn_msg_set_identity, LIB), Cint, (Ptr{Cvoid},), bpc.h))
  else
    ITensorNetworks.set_messages!(bpc, messages)
  end
  return bpc
end
B200BeliefPropagationCache(psi::ITensorNetwork; kwargs...) = B200BeliefPropagationCache(QuadraticFormNetwork(psi); kwargs...)

# ---- algorithm tag: initialize_cache / update ------------------------------------------------------------------------

function ITensorNetworks.initialize_cache(alg::Algorithm"bp_b200", fn::QuadraticFormNetwork; kwargs...)
  return B200BeliefPropagationCache(fn; kwargs...)
end
# environment(alg, tn, verts; ...) partitions first (src/environment.jl:41-49) and hands over the PartitionedGraph
function ITensorNetworks.initialize_cache(alg::Algorithm"bp_b200", ptn::PartitionedGraph; kwargs...)
  fn = unpartitioned_graph(ptn)
  fn isa QuadraticFormNetwork || error("alg = \"bp_b200\" contracts norm networks (QuadraticFormNetwork)")
  return B200BeliefPropagationCache(fn; partitioned_vertices=PartitionedGraphs.partitioned_vertices(ptn), kwargs...)
end
ITensorNetworks.initialize_cache(alg::Algorithm"bp_b200", psi::ITensorNetwork; kwargs...) =
  B200BeliefPropagationCache(QuadraticFormNetwork(psi); kwargs...)

ITensorNetworks.default_update_alg(::B200BeliefPropagationCache) = "bp_b200"
ITensorNetworks.default_message_update_alg(::B200BeliefPropagationCache) = "contract"
ITensorNetworks.default_bp_maxiter(bpc::B200BeliefPropagationCache) =
  ITensorNetworks.default_bp_maxiter(quotient_graph(bpc.ptn))
ITensorNetworks.default_bp_maxiter(::Algorithm, bpc::B200BeliefPropagationCache) = ITensorNetworks.default_bp_maxiter(bpc)
ITensorNetworks.default_bp_edge_sequence(bpc::B200BeliefPropagationCache) = default_edge_sequence(bpc.ptn)
ITensorNetworks.default_edge_sequence(::Algorithm, bpc::B200BeliefPropagationCache) = default_edge_sequence(bpc.ptn)

function ITensorNetworks.set_default_kwargs(alg::Algorithm"bp_b200", bpc::B200BeliefPropagationCache)
  verbose = get(alg.kwargs, :verbose, false)
  maxiter = get(alg.kwargs, :maxiter, ITensorNetworks.default_bp_maxiter(bpc))
  edge_sequence = get(alg.kwargs, :edge_sequence, ITensorNetworks.default_bp_edge_sequence(bpc))
  tol = get(alg.kwargs, :tol, nothing)
  message_update_alg = ITensorNetworks.set_default_kwargs(get(alg.kwargs, :message_update_alg, Algorithm("contract")))
  return Algorithm("bp_b200"; verbose, maxiter, edge_sequence, tol, message_update_alg)
end

vertex_of(pv::QuotientVertex) = parent(pv)       # the default partition is keyed by the original state vertex
endpoints(pe::QuotientEdge) = (vertex_of(src(pe)), vertex_of(dst(pe)))

# update(::Algorithm"bp", bpc) (abstractbeliefpropagationcache.jl:313-329) in ONE library call: the sweeps, message_diff
# and the tol test run on the device.  A vector of QuotientEdges is the sequential schedule, a vector of vectors the
# grouped one (:294-308).
function ITensorNetworks.update(alg::Algorithm"bp_b200", bpc::B200BeliefPropagationCache)
  isnothing(alg.kwargs.maxiter) && error("You need to specify a number of iterations for BP!")
  seq = alg.kwargs.edge_sequence
  grouped = !isempty(seq) && first(seq) isa AbstractVector
  flat = grouped ? reduce(vcat, seq) : seq
  s = Int32[bpc.vid[endpoints(e)[1]] for e in flat]
  d = Int32[bpc.vid[endpoints(e)[2]] for e in flat]
  gp = grouped ? Int32[0; cumsum(length.(seq))] : Int32[0]
  out = copy(bpc)
  iters = Ref{Int32}(0)
  diff = Ref{Float64}(NaN)
  tol = alg.kwargs.tol
  normalize = alg.kwargs.message_update_alg.kwargs.normalize
  GC.@preserve s d gp check(ccall((:itn_bp_update, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Cint, Cdouble, Cint, Ptr{Int32}, Ptr{Cdouble}),
    out.h, s, d, length(flat), grouped ? pointer(gp) : Ptr{Int32}(C_NULL), grouped ? length(seq) : 0,
    alg.kwargs.maxiter, isnothing(tol) ? -1.0 : tol, normalize, iters, diff))
  alg.kwargs.verbose && !isnothing(tol) && diff[] <= tol &&
    println("BP converged to desired precision after $(iters[]) iterations.")
  return out
end

# ---- required interface (abstractbeliefpropagationcache.jl:44-69) ---------------------------------------------------------

ITensorNetworks.partitioned_tensornetwork(bpc::B200BeliefPropagationCache) = (sync_host!(bpc); bpc.ptn)
ITensorNetworks.partitions(bpc::B200BeliefPropagationCache) = quotientvertices(bpc.ptn)
PartitionedGraphs.quotientedges(bpc::B200BeliefPropagationCache) = quotientedges(bpc.ptn)
PartitionedGraphs.partitioned_vertices(bpc::B200BeliefPropagationCache) = PartitionedGraphs.partitioned_vertices(bpc.ptn)
PartitionedGraphs.quotient_graph(bpc::B200BeliefPropagationCache) = quotient_graph(bpc.ptn)
Graphs.vertices(bpc::B200BeliefPropagationCache) = vertices(bpc.ptn)

function Base.copy(bpc::B200BeliefPropagationCache{V,PV,QF,PTN}) where {V,PV,QF,PTN}   # beliefpropagationcache.jl:43-47
  out = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:itn_net_clone, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), bpc.h, out))
  qf = copy(bpc.qf)
  ptn = PartitionedGraph(qf, PartitionedGraphs.partitioned_vertices(bpc.ptn))
  c = B200BeliefPropagationCache{V,PV,QF,typeof(ptn)}(out[], bpc.ctx, qf, ptn, bpc.verts, bpc.vid, bpc.eds, bpc.eid,
    bpc.elt, bpc.celt, copy(bpc.stale))
  return finalize_cache!(c)
end

# device -> host for the tensors a gate or a rescale changed (bra = dag(prime(ket)), quadraticformnetwork.jl:126-137)
function sync_host!(bpc::B200BeliefPropagationCache)
  for v in collect(bpc.stale)
    old = ket_tensor(bpc, v)
    is = inds(old)
    ax = axis_edges(bpc, v, old)
    dims = Int[]
    for (i, a) in zip(is, ax)
      if a < 0
        push!(dims, dim(i))
      else
        d = Ref{Int32}(0)
        check(ccall((:itn_net_edge_dim, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, a, d))
        push!(dims, d[])
      end
    end
    buf = Array{bpc.celt}(undef, dims...)
    GC.@preserve buf ax check(ccall((:itn_net_get_tensor, LIB), Cint,
      (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, bpc.vid[v], buf, ndims(buf), ax))
    # bond dimensions may have changed under a gate: fresh link indices of the new size are made by update_host_links!
    new_is = map((i, n) -> dim(i) == n ? i : Index(n; tags=ITensors.tags(i)), is, dims)
    ITensorNetworks.update(bpc.qf, v, itensor(Array{bpc.elt}(buf), new_is...))
  end
  empty!(bpc.stale)
  return bpc
end

# messages as ITensors on (dag(prime(link)), link), the convention of identity_messages (quadraticformnetwork.jl:117):
# device matrix M[a, a'] with a the ket link and a' its primed (bra) copy
function message_itensor(bpc::B200BeliefPropagationCache, u, v)
  l = link_index(bpc, bpc.eds[bpc.eid[(u, v)] + 1])
  m = Matrix{bpc.celt}(undef, dim(l), dim(l))
  GC.@preserve m check(ccall((:itn_msg_get, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[u], bpc.vid[v], m))
  return itensor(Matrix{bpc.elt}(m), l, dag(prime(l)))
end

function ITensorNetworks.message(bpc::B200BeliefPropagationCache, pe::QuotientEdge)
  u, v = endpoints(pe)
  return ITensor[message_itensor(bpc, u, v)]
end
function ITensorNetworks.messages(bpc::B200BeliefPropagationCache)
  out = Dictionary{QuotientEdge,Vector{ITensor}}()
  for pe in quotientedges(bpc.ptn), e in (pe, reverse(pe))
    set!(out, e, ITensorNetworks.message(bpc, e))
  end
  return out
end
ITensorNetworks.messages(bpc::B200BeliefPropagationCache, es) = map(e -> ITensorNetworks.message(bpc, e), es)

function ITensorNetworks.set_message!(bpc::B200BeliefPropagationCache, pe::QuotientEdge, message)
  u, v = endpoints(pe)
  l = link_index(bpc, bpc.eds[bpc.eid[(u, v)] + 1])
  t = length(message) == 1 ? only(message) : ITensors.contract(message)
  a = Matrix{bpc.celt}(array(t, l, dag(prime(l))))
  GC.@preserve a check(ccall((:itn_msg_set, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[u], bpc.vid[v], a))
  return bpc
end
function ITensorNetworks.set_messages!(bpc::B200BeliefPropagationCache, quotientedges_messages)
  for pe in eachindex(quotientedges_messages)
    ITensorNetworks.set_message!(bpc, pe, quotientedges_messages[pe])
  end
  return bpc
end
ITensorNetworks.delete_messages!(bpc::B200BeliefPropagationCache, pes::Vector{<:QuotientEdge}) =
  error("libitn_b200 keeps a message on every directed edge once it is set; build a new cache to drop messages")

# updated_message(::Algorithm"contract", bpc, edge) (:225-239) on the device, without storing it
function ITensorNetworks.updated_message(alg::Algorithm"contract", bpc::B200BeliefPropagationCache, pe::QuotientEdge)
  u, v = endpoints(pe)
  l = link_index(bpc, bpc.eds[bpc.eid[(u, v)] + 1])
  m = Matrix{bpc.celt}(undef, dim(l), dim(l))
  GC.@preserve m check(ccall((:itn_updated_message, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[u], bpc.vid[v], alg.kwargs.normalize, m))
  return ITensor[itensor(Matrix{bpc.elt}(m), l, dag(prime(l)))]
end

# message_diff(updated_message(bpc, e), message(bpc, e)) for a list of directed edges (:32-36), on the device
function message_residuals(bpc::B200BeliefPropagationCache, pes)
  s = Int32[bpc.vid[endpoints(e)[1]] for e in pes]
  d = Int32[bpc.vid[endpoints(e)[2]] for e in pes]
  out = zeros(Float64, length(pes))
  GC.@preserve s d out check(ccall((:itn_message_residuals, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Cdouble}), bpc.h, s, d, length(pes), out))
  return out
end

# environment(bpc, verts) (beliefpropagationcache.jl:100-105): messages into the partitions of `verts` + the tensors of
# those partitions that are not in `verts` (ket / bra / operator tensors from the host form network)
function ITensorNetworks.environment(bpc::B200BeliefPropagationCache, verts::Vector; kwargs...)
  sync_host!(bpc)
  pvs = unique(QuotientVertex(original_state_vertex(bpc.qf, v)) for v in verts)
  bpes = boundary_quotientedges(bpc.ptn, pvs; dir=:in)
  ms = reduce(vcat, ITensorNetworks.messages(bpc, bpes); init=ITensor[])
  inside = reduce(vcat, [collect(vertices(bpc.ptn, pv)) for pv in pvs]; init=[])
  central = ITensor[tensornetwork(bpc.qf)[v] for v in setdiff(inside, verts)]
  return vcat(ms, central)
end

# region_scalar (beliefpropagationcache.jl:107-119), vertex_scalars / edge_scalars (abstract :83-97): one call for all
function region_scalars(bpc::B200BeliefPropagationCache)
  zv = Vector{bpc.celt}(undef, length(bpc.verts))
  ze = Vector{bpc.celt}(undef, max(length(bpc.eds), 1))
  GC.@preserve zv ze check(ccall((:itn_region_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), bpc.h, zv, ze))
  return zv, ze
end
function ITensorNetworks.region_scalar(bpc::B200BeliefPropagationCache, pv::QuotientVertex; kwargs...)
  return bpc.elt(region_scalars(bpc)[1][bpc.vid[vertex_of(pv)] + 1])
end
function ITensorNetworks.region_scalar(bpc::B200BeliefPropagationCache, pe::QuotientEdge; kwargs...)
  return bpc.elt(region_scalars(bpc)[2][bpc.eid[endpoints(pe)] + 1])
end
function ITensorNetworks.vertex_scalars(bpc::B200BeliefPropagationCache, pvs=ITensorNetworks.partitions(bpc); kwargs...)
  zv = region_scalars(bpc)[1]
  return map(pv -> bpc.elt(zv[bpc.vid[vertex_of(pv)] + 1]), pvs)
end
function ITensorNetworks.edge_scalars(bpc::B200BeliefPropagationCache, pes=quotientedges(bpc); kwargs...)
  ze = region_scalars(bpc)[2]
  return map(pe -> bpc.elt(ze[bpc.eid[endpoints(pe)] + 1]), pes)
end

function ITensorNetworks.logscalar(bpc::B200BeliefPropagationCache)      # abstract :397-408
  out = zeros(Float64, 2)
  check(ccall((:itn_logscalar, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), bpc.h, out))
  return out[2] == 0 ? out[1] : complex(out[1], out[2])
end

# rescale (abstract :349-395; beliefpropagationcache.jl:121-139).  `verts` are vertices of the form network: a site is
# rescaled when its ket (equivalently its bra) is listed, which is how normalize passes them (normalize.jl:75-76).
function ITensorNetworks.rescale_messages(bpc::B200BeliefPropagationCache, pes)
  out = copy(bpc)
  check(ccall((:itn_rescale_verts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), out.h, Int32[], 0))
  return out
end
ITensorNetworks.rescale_messages(bpc::B200BeliefPropagationCache) = ITensorNetworks.rescale_messages(bpc, quotientedges(bpc))
function ITensorNetworks.rescale(bpc::B200BeliefPropagationCache; verts=nothing, kwargs...)
  out = copy(bpc)
  if isnothing(verts)
    check(ccall((:itn_rescale, LIB), Cint, (Ptr{Cvoid},), out.h))
    union!(out.stale, out.verts)
  else
    sites = unique(original_state_vertex(bpc.qf, v) for v in verts)
    ids = Int32[bpc.vid[v] for v in sites]
    GC.@preserve ids check(ccall((:itn_rescale_verts, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), out.h, ids, length(ids)))
    union!(out.stale, sites)
  end
  return out
end
function ITensorNetworks.rescale_partitions(bpc::B200BeliefPropagationCache, partitions;
  verts=reduce(vcat, [collect(vertices(bpc.ptn, pv)) for pv in partitions]; init=[]))
  # the library rescales messages and partitions together; rescaling messages that are already rescaled is the identity
  return ITensorNetworks.rescale(bpc; verts)
end

# bpc[v] = t / update_factor(s) (abstract :158-171; beliefpropagationcache.jl:88-91).  Writing a ket tensor uploads it;
# the bra is implicit on the device, so a bra write must be the dual of the ket (it is kept on the host side only).
function DataGraphs.set_vertex_data!(bpc::B200BeliefPropagationCache, value::ITensor, vertex)
  sync_host!(bpc)
  DataGraphs.set_vertex_data!(bpc.qf, value, vertex)
  if vertex == ket_vertex(bpc.qf, original_state_vertex(bpc.qf, vertex))
    v = original_state_vertex(bpc.qf, vertex)
    a = Array{bpc.celt}(array(value))
    ax = axis_edges(bpc, v, value)
    GC.@preserve a ax check(ccall((:itn_net_set_tensor, LIB), Cint,
      (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, bpc.vid[v], a, ndims(a), ax))
  end
  return bpc
end
function ITensorNetworks.update_factors(bpc::B200BeliefPropagationCache, factors)
  out = copy(bpc)
  for vertex in eachindex(factors)
    out[vertex] = factors[vertex]
  end
  return out
end
function ITensorNetworks.update_factor(bpc::B200BeliefPropagationCache, vertex, factor)
  out = copy(bpc)
  out[vertex] = factor
  return out
end
ITensorNetworks.tensornetwork(bpc::B200BeliefPropagationCache) = (sync_host!(bpc); bpc.qf)

# ---- Adapt (abstract :109-137): where the storage lives ---------------------------------------------------------------------
# The device IS the storage of this cache, so adapting its messages / factors to a device array type is the identity;
# adapting to a host array type hands back a plain BeliefPropagationCache with everything downloaded, and
# `adapt(B200Device(ctx), bpc::BeliefPropagationCache)` moves a host cache (tensors and messages) onto the GPU.
struct B200Device
  ctx::Context
end
B200Device() = B200Device(default_context())

ITensorNetworks.map_messages(f, bpc::B200BeliefPropagationCache, pes=nothing) = bpc
ITensorNetworks.map_factors(f, bpc::B200BeliefPropagationCache, vs=nothing) = bpc
Adapt.adapt_structure(to::B200Device, bpc::B200BeliefPropagationCache) = bpc
function Adapt.adapt_structure(to::B200Device, bpc::BeliefPropagationCache)
  fn = tensornetwork(bpc)
  fn isa QuadraticFormNetwork || error("B200Device holds norm networks (QuadraticFormNetwork)")
  return B200BeliefPropagationCache(fn; partitioned_vertices=PartitionedGraphs.partitioned_vertices(bpc),
    messages=ITensorNetworks.messages(bpc), ctx=to.ctx)
end
function Adapt.adapt_structure(to::Type{<:Array}, bpc::B200BeliefPropagationCache)
  sync_host!(bpc)
  return BeliefPropagationCache(PartitionedGraph(copy(bpc.qf), PartitionedGraphs.partitioned_vertices(bpc.ptn));
    messages=ITensorNetworks.messages(bpc))
end

# ---- observables ----------------------------------------------------------------------------------------------------------

# expect(alg, psi, ops; cache!, ...) (src/expect.jl:21-41): the generic method works through `environment` above (one
# download per operator); this specialisation evaluates every single-site operator in one call (itn_expect1).
function ITensorNetworks.expect(alg::Algorithm"bp_b200", ψ::AbstractITensorNetwork, ops;
  (cache!)=nothing, update_cache=isnothing(cache!), cache_update_kwargs=(;), cache_construction_kwargs=(;), kwargs...)
  if isnothing(cache!)
    cache! = Ref(ITensorNetworks.initialize_cache(alg, QuadraticFormNetwork(ψ); cache_construction_kwargs...))
  end
  if update_cache
    cache![] = ITensorNetworks.update(cache![]; cache_update_kwargs...)
  end
  bpc = cache![]
  ids = Int32[bpc.vid[only(o.sites)] for o in ops]
  mats = reduce(vcat, [vec(Matrix{bpc.celt}(array(ITensors.op(o.which_op, site_index(bpc, only(o.sites))),
      prime(site_index(bpc, only(o.sites))), site_index(bpc, only(o.sites))))) for o in ops]; init=bpc.celt[])
  out = Vector{bpc.celt}(undef, length(ops))
  GC.@preserve ids mats out check(ccall((:itn_expect1, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), bpc.h, ids, length(ids), mats, out))
  return map(bpc.elt, out)
end

# two-site reduced density matrices (test/test_belief_propagation.jl:64-91) on a list of edges: (d_u d_v) x (d_u d_v)
# matrices with row index s_u + d_u s_v, unit trace
function rdm2(bpc::B200BeliefPropagationCache, es)
  ids = Int32[bpc.eid[(src(e), dst(e))] for e in es]
  D = [dim(site_index(bpc, src(bpc.eds[i + 1]))) * dim(site_index(bpc, dst(bpc.eds[i + 1]))) for i in ids]
  out = Vector{bpc.celt}(undef, sum(abs2, D))
  GC.@preserve ids out check(ccall((:itn_rdm2, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}),
    bpc.h, ids, length(ids), out))
  offs = cumsum([0; abs2.(D)])
  return [reshape(out[offs[k]+1:offs[k+1]], D[k], D[k]) for k in eachindex(D)]
end

# ---- gates -----------------------------------------------------------------------------------------------------------------
# apply(o, psi; envs, maxdim, cutoff, normalize, callback) (src/apply.jl:97-146) with the BP product environment taken
# from the cache: `envs = B200Environment(bpc)` replaces `envs = environment(bpc, ...)`, the cache is updated in place
# (tensors, bond dimension, messages on the gated edge) and the gated state is returned as in the reference.
struct B200Environment
  bpc::B200BeliefPropagationCache
  msg_mode::Int          # message stored on the gated edge: 0 = identity, 1 = diag(singular values)
end
B200Environment(bpc::B200BeliefPropagationCache) = B200Environment(bpc, 1)

gate_sites(bpc::B200BeliefPropagationCache, o::ITensor) =
  [v for v in bpc.verts if !isempty(commoninds(o, ITensor(site_index(bpc, v))))]

function ITensors.apply(o::ITensor, ψ::AbstractITensorNetwork; envs::B200Environment, normalize=false, ortho=false,
  callback=Returns(nothing), maxdim=nothing, cutoff=nothing, apply_kwargs...)
  bpc = envs.bpc
  vs = gate_sites(bpc, o)
  # ortho = true: tree_orthogonalize(psi, v1) before the gate (src/apply.jl:109-111, 130-132), on the device
  ortho && length(vs) in (1, 2) && tree_orthogonalize!(bpc, vs[1])
  if length(vs) == 1
    v = only(vs)
    s = site_index(bpc, v)
    g = Matrix{bpc.celt}(array(o, prime(s), s))
    ids = Int32[bpc.vid[v]]
    GC.@preserve g ids check(ccall((:itn_apply1, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Cint),
      bpc.h, ids, 1, g, normalize))
    push!(bpc.stale, v)
  elseif length(vs) == 2
    haskey(bpc.eid, (vs[1], vs[2])) || error("Vertices where the gates are being applied must be neighbors for now.")
    k = bpc.eid[(vs[1], vs[2])]
    e = bpc.eds[k + 1]
    s1, s2 = site_index(bpc, src(e)), site_index(bpc, dst(e))
    g = Array{bpc.celt,4}(array(o, prime(s1), prime(s2), s1, s2))      # g[s1', s2', s1, s2], (1, 2) = (esrc, edst)
    newdim = Ref{Int32}(0)
    terr = Ref{Float64}(0.0)
    stride = dim(s1) * dim(s2) * dim(link_index(bpc, e))
    sv = zeros(Float64, stride)
    ids = Int32[k]
    GC.@preserve g ids sv check(ccall((:itn_apply2, LIB), Cint,
      (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Cint, Cdouble, Cint, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
      bpc.h, ids, 1, g, isnothing(maxdim) ? 0 : maxdim, isnothing(cutoff) ? -1.0 : cutoff, normalize, envs.msg_mode,
      newdim, terr, sv, stride))
    callback(; singular_values=sv[1:newdim[]], truncation_error=terr[])
    push!(bpc.stale, src(e), dst(e))
  elseif length(vs) < 1
    error("Gate being applied does not share indices with tensor network.")
  else
    error("Gates with more than 2 sites is not supported yet.")
  end
  return ket_network(ITensorNetworks.tensornetwork(bpc))
end

# gauge_walk(tn, edges) (src/abstractitensornetwork.jl:387-393) on the device tensors: qr!(tn, src => dst) edge by edge
# (include/itn_b200.h: itn_gauge_walk).  The messages of the cache keep their values (they belong to the old gauge).
function gauge_walk!(bpc::B200BeliefPropagationCache, es)
  s = Int32[bpc.vid[src(e)] for e in es]
  d = Int32[bpc.vid[dst(e)] for e in es]
  GC.@preserve s d check(ccall((:itn_gauge_walk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint),
    bpc.h, s, d, length(es)))
  foreach(e -> push!(bpc.stale, src(e), dst(e)), es)
  return bpc
end

# tree_gauge / tree_orthogonalize (src/abstractitensornetwork.jl:395-420): the walk is planned by the reference's own
# edge_sequence_between_regions on the ket graph, the QR steps run on the device
function tree_orthogonalize!(bpc::B200BeliefPropagationCache, region)
  region = region isa AbstractVector ? region : [region]
  g = NamedGraphs.NamedGraph(bpc.verts)                       # graph of the ket network: original vertices, its bonds
  foreach(e -> Graphs.add_edge!(g, src(e) => dst(e)), bpc.eds)
  es = ITensorNetworks.edge_sequence_between_regions(g, collect(vertices(g)), region)
  return gauge_walk!(bpc, es)
end

# Single-process multi-GPU (include/itn_b200.h: itn_ctx_create_group = ncclCommInitAll): one Context per device, rank i
# of n = position in `devices`.  Collective calls (update, apply on cut edges, ...) are issued by one task per context,
# e.g. `Threads.@spawn` per GPU.
function context_group(devices::Vector{<:Integer})
  devs = Int32.(devices)
  hs = Vector{Ptr{Cvoid}}(undef, length(devs))
  GC.@preserve devs hs check(ccall((:itn_ctx_create_group, LIB), Cint, (Cint, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
    length(devs), devs, hs))
  return [Context(h) for h in hs]
end

function context_rank(ctx::Context)
  r = Ref{Int32}(0)
  n = Ref{Int32}(1)
  check(ccall((:itn_ctx_rank, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), ctx.h, r, n))
  return Int(r[]), Int(n[])
end

# One Trotter step in one call (include/itn_b200.h: itn_apply_layers): `layers` is a vector of vectors of two-site gate
# ITensors; after every layer `bp_maxiter` synchronous BP sweeps refresh the environments (north_star's
# `apply(gate, psi; cache_update_kwargs, maxdim, cutoff)` loop without returning to Julia between its parts).
function tebd_step!(bpc::B200BeliefPropagationCache, layers; maxdim=nothing, cutoff=nothing, normalize=false, msg_mode=1,
  bp_maxiter=0, bp_tol=nothing)
  eids = Int32[]
  ptr = Int32[0]
  packed = bpc.celt[]
  for layer in layers
    for o in layer
      vs = gate_sites(bpc, o)
      (length(vs) == 2 && haskey(bpc.eid, (vs[1], vs[2]))) ||
        error("Vertices where the gates are being applied must be neighbors for now.")
      k = bpc.eid[(vs[1], vs[2])]
      e = bpc.eds[k + 1]
      s1, s2 = site_index(bpc, src(e)), site_index(bpc, dst(e))
      push!(eids, k)
      append!(packed, vec(Array{bpc.celt,4}(array(o, prime(s1), prime(s2), s1, s2))))
      push!(bpc.stale, src(e), dst(e))
    end
    push!(ptr, length(eids))
  end
  seq = reduce(vcat, [[(src(e), dst(e)), (dst(e), src(e))] for e in bpc.eds])   # every directed edge, its own group
  s = Int32[bpc.vid[a] for (a, _) in seq]
  d = Int32[bpc.vid[b] for (_, b) in seq]
  gp = Int32.(0:length(seq))
  n = length(eids)
  stride = 4 * maximum(dim(link_index(bpc, e)) for e in bpc.eds) * (isnothing(maxdim) ? 1 : 1) + (isnothing(maxdim) ? 0 : 4 * maxdim)
  newdim = zeros(Int32, n)
  terr = zeros(Float64, n)
  sv = zeros(Float64, stride * n)
  iters = Ref{Int32}(0)
  GC.@preserve eids ptr packed s d gp newdim terr sv check(ccall((:itn_apply_layers, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Cint, Cdouble, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Cint,
     Ptr{Int32}, Cint, Cint, Cdouble, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Int32}),
    bpc.h, length(layers), ptr, eids, packed, isnothing(maxdim) ? 0 : maxdim, isnothing(cutoff) ? -1.0 : cutoff,
    normalize, msg_mode, s, d, length(seq), gp, length(seq), bp_maxiter, isnothing(bp_tol) ? -1.0 : bp_tol, 1,
    newdim, terr, sv, stride, iters))
  return (; newdim, truncation_error=terr, singular_values=[sv[(i-1)*stride+1:(i-1)*stride+newdim[i]] for i in 1:n],
    bp_iterations=iters[])
end

# map_eigvals(f, A, ...; ishermitian = true, cutoff) (src/apply.jl:21-25) for one Hermitian matrix; f in (:sqrt, :invsqrt, :inv)
function map_eigvals_matrix(ctx::Context, f::Symbol, m::AbstractMatrix; cutoff=nothing)
  elt = compute_type(eltype(m))
  a = Matrix{elt}(m)
  out = similar(a)
  fn = Dict(:sqrt => 0, :invsqrt => 1, :inv => 2)[f]
  GC.@preserve a out check(ccall((:itn_map_eigvals, LIB), Cint,
    (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
    ctx.h, dtype_code(elt), fn, size(a, 1), 1, a, out, isnothing(cutoff) ? -1.0 : cutoff))
  return out
end

# ---- bilinear forms: inner(phi, psi; alg = "bp") / loginner (src/inner.jl:100-171) ------------------------------------------
# BilinearFormNetwork(phi, psi) with an explicit bra layer.  phi must live on the same graph with the same link dimensions
# as psi (pad the smaller tensor with zeros otherwise).  The link indices of phi[v] are matched to those of psi[v] edge by
# edge; phi is passed un-conjugated, the engine applies `dag`.
function set_bra_factor!(bpc::B200BeliefPropagationCache, phi::ITensorNetwork, v)
  t = phi[v]
  s = only(siteinds(phi, v))
  ax = Int32[i == s ? -1 : findfirst(e -> (src(e) == v || dst(e) == v) &&
                                      i == commonind(phi[src(e)], phi[dst(e)]), bpc.eds) - 1 for i in inds(t)]
  a = Array{bpc.celt}(array(t))
  GC.@preserve a ax check(ccall((:itn_net_set_bra_tensor, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, bpc.vid[v], a, ndims(a), ax))
  return bpc
end

function loginner_bp(phi::ITensorNetwork, psi::ITensorNetwork; ctx::Context=default_context(), update_kwargs...)
  # the fallback initialize_cache of the reference gives a bilinear form no default messages (src/initialize_cache.jl:10-12):
  # trees need none, loopy graphs get identity messages here and `maxiter` from the caller
  bpc = B200BeliefPropagationCache(psi; ctx)
  for v in bpc.verts
    set_bra_factor!(bpc, phi, v)
  end
  return ITensorNetworks.logscalar(ITensorNetworks.update(bpc; update_kwargs...))
end
inner_bp(phi, psi; kwargs...) = exp(loginner_bp(phi, psi; kwargs...))

end # module
