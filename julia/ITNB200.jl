# ITNB200.jl -- Julia shim over libitn_b200.so (UNTESTED: there is no Julia runtime in the build image or on the
# GPU boxes; the same ABI is exercised by the Python ctypes mirror in itensornetworks.jl_b200/itn_b200/).
#
# `B200BeliefPropagationCache <: AbstractBeliefPropagationCache` holds an `itn_net*` and forwards the interface
# of src/caches/abstractbeliefpropagationcache.jl:44-69 to the C ABI declared in include/itn_b200.h.  Host code
# (graphs, index bookkeeping, OpSum -> gate arrays, edge sequences) stays in ITensorNetworks.jl.
module ITNB200

using ITensors: ITensors, ITensor, Index, array, dim, inds, itensor, commonind, noncommoninds
using ITensorNetworks: ITensorNetworks, AbstractBeliefPropagationCache, ITensorNetwork, default_edge_sequence,
  edges, vertices, src, dst, siteinds, linkinds
using NamedGraphs: NamedGraphs, NamedEdge

const LIB = get(ENV, "ITN_B200_LIB", joinpath(@__DIR__, "..", "itensornetworks.jl_b200", "lib", "libitn_b200.so"))

struct ITNError <: Exception
  code::Cint
  msg::String
end
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
end

"""
BP cache of <psi|psi> with the default one-site partition whose tensors and messages live on a B200.
Mirrors BeliefPropagationCache (src/caches/beliefpropagationcache.jl:13-35).
"""
mutable struct B200BeliefPropagationCache{V} <: AbstractBeliefPropagationCache{V}
  h::Ptr{Cvoid}
  ctx::Context
  psi::ITensorNetwork{V}          # host copy of the index structure (tensors are authoritative on the device)
  verts::Vector{V}                # vertex <-> 0-based id
  vid::Dict{V,Int}
  eds::Vector{NamedEdge{V}}       # undirected edge list, id = position - 1
  elt::Type
end

dtype_code(::Type{Float64}) = Cint(0)
dtype_code(::Type{ComplexF64}) = Cint(1)

function B200BeliefPropagationCache(psi::ITensorNetwork{V}; ctx::Context=Context(), messages=:default) where {V}
  verts = collect(vertices(psi))
  vid = Dict(v => i - 1 for (i, v) in enumerate(verts))
  eds = collect(edges(psi))
  elt = promote_type(map(v -> eltype(psi[v]), verts)...)
  esrc = Int32[vid[src(e)] for e in eds]
  edst = Int32[vid[dst(e)] for e in eds]
  edim = Int32[dim(commonind(psi[src(e)], psi[dst(e)])) for e in eds]
  sdim = Int32[dim(only(siteinds(psi, v))) for v in verts]
  out = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:itn_net_create, LIB), Cint,
    (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
    ctx.h, dtype_code(elt), length(verts), length(eds), esrc, edst, edim, sdim, C_NULL, out))
  bpc = B200BeliefPropagationCache{V}(out[], ctx, psi, verts, vid, eds, elt)
  finalizer(b -> ccall((:itn_net_destroy, LIB), Cint, (Ptr{Cvoid},), b.h), bpc)
  # all site tensors in one pipelined upload (itn_net_set_tensors): column-major `array(t)`, axis_edge maps the ITensor
  # index order to edges so nothing is permuted on the host.  flags = 1 (ITN_HOST_DEFERRED) would only register the
  # arrays and let the first synchronous update() overlap the copy with its sweep; the arrays must then stay rooted
  # (keep `arrs` in the cache object) until that call returns.
  arrs = [array(psi[v]) for v in verts]
  axes = [axis_edges(bpc, v, psi[v]) for v in verts]
  ids = Int32[vid[v] for v in verts]
  nds = Int32[ndims(a) for a in arrs]
  ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
  flat = reduce(vcat, axes; init=Int32[])
  GC.@preserve arrs check(ccall((:itn_net_set_tensors, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Int32}, Cint),
    bpc.h, length(verts), ids, ptrs, nds, flat, 0))
  # initialize_cache (src/initialize_cache.jl:14-29): identity messages on loopy graphs only
  if messages === :identity || (messages === :default && !NamedGraphs.is_tree(psi))
    check(ccall((:itn_msg_set_identity, LIB), Cint, (Ptr{Cvoid},), bpc.h))
  end
  return bpc
end

# axis_edge[i] = edge id carried by axis i of the ITensor's storage, or -1 for the site index
function axis_edges(bpc::B200BeliefPropagationCache, v, t::ITensor)
  s = only(siteinds(bpc.psi, v))
  return Int32[i == s ? -1 : findfirst(e -> (src(e) == v || dst(e) == v) &&
                                       i == commonind(bpc.psi[src(e)], bpc.psi[dst(e)]), bpc.eds) - 1
               for i in inds(t)]
end

function set_factor!(bpc::B200BeliefPropagationCache, v, t::ITensor)   # bpc[v] = t (beliefpropagationcache.jl:88-91)
  a = array(t)                                   # column-major, axes in inds(t) order
  ax = axis_edges(bpc, v, t)
  GC.@preserve a ax check(ccall((:itn_net_set_tensor, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, bpc.vid[v], a, ndims(a), ax))
  return bpc
end

Base.copy(bpc::B200BeliefPropagationCache{V}) where {V} = begin   # beliefpropagationcache.jl:43-47
  out = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:itn_net_clone, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), bpc.h, out))
  c = B200BeliefPropagationCache{V}(out[], bpc.ctx, bpc.psi, bpc.verts, bpc.vid, bpc.eds, bpc.elt)
  finalizer(b -> ccall((:itn_net_destroy, LIB), Cint, (Ptr{Cvoid},), b.h), c)
  c
end

# update(bpc; maxiter, tol, edge_sequence, message_update_alg = (; normalize))   (abstract...cache.jl:313-337)
function ITensorNetworks.update(bpc::B200BeliefPropagationCache; maxiter=nothing, tol=nothing,
  edge_sequence=default_edge_sequence(bpc.psi), normalize=true, kwargs...)
  isnothing(maxiter) && NamedGraphs.is_tree(bpc.psi) && (maxiter = 1)
  isnothing(maxiter) && error("You need to specify a number of iterations for BP!")
  grouped = !isempty(edge_sequence) && first(edge_sequence) isa AbstractVector
  flat = grouped ? reduce(vcat, edge_sequence) : edge_sequence
  s = Int32[bpc.vid[src(e)] for e in flat]
  d = Int32[bpc.vid[dst(e)] for e in flat]
  gp = grouped ? Int32[0; cumsum(length.(edge_sequence))] : Int32[]
  out = copy(bpc)
  iters = Ref{Int32}(0); diff = Ref{Float64}(NaN)
  GC.@preserve s d gp check(ccall((:itn_bp_update, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Cint, Cdouble, Cint, Ptr{Int32}, Ptr{Cdouble}),
    out.h, s, d, length(flat), grouped ? pointer(gp) : C_NULL, grouped ? length(edge_sequence) : 0,
    maxiter, isnothing(tol) ? -1.0 : tol, normalize, iters, diff))
  return out
end

# message(bpc, edge) (abstract...cache.jl:173-175) as an ITensor on (prime(link)', link)-style indices is built by
# the caller from this matrix M[a, a'] (a: ket link, a': bra link).
function message_matrix(bpc::B200BeliefPropagationCache, e)
  chi = Ref{Int32}(0)
  eid = findfirst(x -> (src(x), dst(x)) == (src(e), dst(e)) || (src(x), dst(x)) == (dst(e), src(e)), bpc.eds) - 1
  check(ccall((:itn_net_edge_dim, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, eid, chi))
  m = Matrix{bpc.elt}(undef, chi[], chi[])
  GC.@preserve m check(ccall((:itn_msg_get, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[src(e)], bpc.vid[dst(e)], m))
  return m
end

function ITensorNetworks.logscalar(bpc::B200BeliefPropagationCache)      # abstract...cache.jl:397-408
  out = zeros(Float64, 2)
  check(ccall((:itn_logscalar, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), bpc.h, out))
  return out[2] == 0 ? out[1] : complex(out[1], out[2])
end

function ITensorNetworks.rescale(bpc::B200BeliefPropagationCache; kwargs...)   # abstract...cache.jl:391-395
  out = copy(bpc)
  check(ccall((:itn_rescale, LIB), Cint, (Ptr{Cvoid},), out.h))
  return out
end

# expect(psiIpsi, op) (src/expect.jl:5-19) for a list of vertices and one d x d operator matrix O[s_out, s_in]
function expect_matrix(bpc::B200BeliefPropagationCache, O::AbstractMatrix, verts=bpc.verts)
  ids = Int32[bpc.vid[v] for v in verts]
  ops = repeat(vec(Matrix{bpc.elt}(O)), length(verts))
  out = Vector{bpc.elt}(undef, length(verts))
  GC.@preserve ids ops out check(ccall((:itn_expect1, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), bpc.h, ids, length(ids), ops, out))
  return Dict(zip(verts, out))
end

# apply(o, psi; envs = environment(bpc, ...), maxdim, cutoff, normalize, callback) (src/apply.jl:97-146) with the
# environments taken from the cache's own messages; `gate` is the d1 x d2 x d1 x d2 array g[s1', s2', s1, s2].
function apply_gate!(bpc::B200BeliefPropagationCache, gate::AbstractArray, v1, v2; maxdim=nothing, cutoff=nothing,
  normalize=false, callback=Returns(nothing), msg_mode=0)
  eid = findfirst(x -> Set((src(x), dst(x))) == Set((v1, v2)), bpc.eds)
  isnothing(eid) && error("Vertices where the gates are being applied must be neighbors for now.")
  e = bpc.eds[eid]
  g = Array{bpc.elt,4}(gate)
  src(e) == v1 || (g = permutedims(g, (2, 1, 4, 3)))
  newdim = Ref{Int32}(0); terr = Ref{Float64}(0.0); sv = zeros(Float64, 512)
  ids = Int32[eid - 1]
  GC.@preserve g ids sv check(ccall((:itn_apply2, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Cint, Cdouble, Cint, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
    bpc.h, ids, 1, g, isnothing(maxdim) ? 0 : maxdim, isnothing(cutoff) ? -1.0 : cutoff, normalize, msg_mode,
    newdim, terr, sv, length(sv)))
  callback(; singular_values=sv[1:newdim[]], truncation_error=terr[])
  return bpc
end

# set_message!(bpc, edge, M) (abstract...cache.jl:197-200): M[a, a'] on the directed edge src(e) -> dst(e)
function set_message_matrix!(bpc::B200BeliefPropagationCache, e, m::AbstractMatrix)
  a = Matrix{bpc.elt}(m)
  GC.@preserve a check(ccall((:itn_msg_set, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[src(e)], bpc.vid[dst(e)], a))
  return bpc
end

# updated_message(bpc, edge; normalize) without storing it (abstract...cache.jl:225-239, test_belief_propagation.jl:51)
function updated_message_matrix(bpc::B200BeliefPropagationCache, e; normalize=true)
  chi = size(message_matrix(bpc, e), 1)
  m = Matrix{bpc.elt}(undef, chi, chi)
  GC.@preserve m check(ccall((:itn_updated_message, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}),
    bpc.h, bpc.vid[src(e)], bpc.vid[dst(e)], normalize, m))
  return m
end

# message_diff(updated_message(bpc, e), message(bpc, e)) for a list of directed edges (abstract...cache.jl:32-36)
function message_residuals(bpc::B200BeliefPropagationCache, es)
  s = Int32[bpc.vid[src(e)] for e in es]
  d = Int32[bpc.vid[dst(e)] for e in es]
  out = zeros(Float64, length(es))
  GC.@preserve s d out check(ccall((:itn_message_residuals, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Cint, Ptr{Cdouble}), bpc.h, s, d, length(es), out))
  return out
end

# vertex_scalars / edge_scalars (abstract...cache.jl:83-97; region_scalar beliefpropagationcache.jl:107-119)
function region_scalars(bpc::B200BeliefPropagationCache)
  zv = Vector{bpc.elt}(undef, length(bpc.verts))
  ze = Vector{bpc.elt}(undef, length(bpc.eds))
  GC.@preserve zv ze check(ccall((:itn_region_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), bpc.h, zv, ze))
  return Dict(zip(bpc.verts, zv)), Dict(zip(bpc.eds, ze))
end

# two-site reduced density matrices (test_belief_propagation.jl:64-91) on a list of edges: (d_u d_v) x (d_u d_v), unit trace
function rdm2(bpc::B200BeliefPropagationCache, es)
  ids = Int32[findfirst(x -> Set((src(x), dst(x))) == Set((src(e), dst(e))), bpc.eds) - 1 for e in es]
  D = [dim(only(siteinds(bpc.psi, src(bpc.eds[i + 1])))) * dim(only(siteinds(bpc.psi, dst(bpc.eds[i + 1])))) for i in ids]
  out = Vector{bpc.elt}(undef, sum(abs2, D))
  GC.@preserve ids out check(ccall((:itn_rdm2, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}),
    bpc.h, ids, length(ids), out))
  offs = cumsum([0; abs2.(D)])
  return [reshape(out[offs[k]+1:offs[k+1]], D[k], D[k]) for k in eachindex(D)]
end

# one-site apply (src/apply.jl:108-116): gate[s', s] on vertex v, in place
function apply_gate!(bpc::B200BeliefPropagationCache, gate::AbstractMatrix, v; normalize=false)
  g = Matrix{bpc.elt}(gate)
  ids = Int32[bpc.vid[v]]
  GC.@preserve g ids check(ccall((:itn_apply1, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Cvoid}, Cint),
    bpc.h, ids, 1, g, normalize))
  return bpc
end

# map_eigvals(f, A, ...; ishermitian = true, cutoff) (src/apply.jl:21-25) for one Hermitian matrix; f in (:sqrt, :invsqrt, :inv)
function map_eigvals_matrix(ctx::Context, f::Symbol, m::AbstractMatrix; cutoff=nothing)
  elt = eltype(m) <: Complex ? ComplexF64 : Float64
  a = Matrix{elt}(m)
  out = similar(a)
  fn = Dict(:sqrt => 0, :invsqrt => 1, :inv => 2)[f]
  GC.@preserve a out check(ccall((:itn_map_eigvals, LIB), Cint,
    (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
    ctx.h, dtype_code(elt), fn, size(a, 1), 1, a, out, isnothing(cutoff) ? -1.0 : cutoff))
  return out
end

# One Trotter step in one call (include/itn_b200.h: itn_apply_layers): `layers` is a vector of vectors of
# (gate::Array{T,4}, v1, v2); after every layer `bp_maxiter` synchronous BP sweeps refresh the environments.
function tebd_step!(bpc::B200BeliefPropagationCache, layers; maxdim=nothing, cutoff=nothing, normalize=false, msg_mode=0,
  bp_maxiter=0, bp_tol=nothing)
  eids = Int32[]; ptr = Int32[0]; packed = bpc.elt[]
  for layer in layers
    for (gate, v1, v2) in layer
      eid = findfirst(x -> Set((src(x), dst(x))) == Set((v1, v2)), bpc.eds)
      isnothing(eid) && error("Vertices where the gates are being applied must be neighbors for now.")
      g = Array{bpc.elt,4}(gate)
      src(bpc.eds[eid]) == v1 || (g = permutedims(g, (2, 1, 4, 3)))
      push!(eids, eid - 1); append!(packed, vec(g))
    end
    push!(ptr, length(eids))
  end
  seq = reduce(vcat, [[e, reverse(e)] for e in bpc.eds])          # every directed edge, each its own group
  s = Int32[bpc.vid[src(e)] for e in seq]; d = Int32[bpc.vid[dst(e)] for e in seq]
  gp = Int32.(0:length(seq))
  n = length(eids); stride = 256
  newdim = zeros(Int32, n); terr = zeros(Float64, n); sv = zeros(Float64, stride * n); iters = Ref{Int32}(0)
  GC.@preserve eids ptr packed s d gp newdim terr sv check(ccall((:itn_apply_layers, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Cint, Cdouble, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Cint,
     Ptr{Int32}, Cint, Cint, Cdouble, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Int32}),
    bpc.h, length(layers), ptr, eids, packed, isnothing(maxdim) ? 0 : maxdim, isnothing(cutoff) ? -1.0 : cutoff,
    normalize, msg_mode, s, d, length(seq), gp, length(seq), bp_maxiter, isnothing(bp_tol) ? -1.0 : bp_tol, 1,
    newdim, terr, sv, stride, iters))
  return (; newdim, truncation_error=terr, singular_values=[sv[(i-1)*stride+1:(i-1)*stride+newdim[i]] for i in 1:n],
    bp_iterations=iters[])
end

# inner(phi, psi; alg = "bp") / loginner (src/inner.jl:100-171): BilinearFormNetwork(phi, psi) with an explicit bra layer.
# phi must live on the same graph with the same link dimensions as psi (pad the smaller tensor with zeros otherwise;
# for inner(phi, A, psi) contract A[v] into psi[v] first and fuse the link pairs with combiners).  The link indices of
# phi[v] are matched to those of psi[v] edge by edge; phi is passed un-conjugated, the engine applies `dag`.
function set_bra_factor!(bpc::B200BeliefPropagationCache, phi::ITensorNetwork, v)
  t = phi[v]
  s = only(siteinds(phi, v))
  ax = Int32[i == s ? -1 : findfirst(e -> (src(e) == v || dst(e) == v) &&
                                      i == commonind(phi[src(e)], phi[dst(e)]), bpc.eds) - 1 for i in inds(t)]
  a = Array{bpc.elt}(array(t))
  GC.@preserve a ax check(ccall((:itn_net_set_bra_tensor, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Int32}), bpc.h, bpc.vid[v], a, ndims(a), ax))
  return bpc
end

function loginner_bp(phi::ITensorNetwork, psi::ITensorNetwork; ctx::Context=Context(), messages=:none, update_kwargs...)
  # the fallback initialize_cache of the reference gives a bilinear form no default messages (src/initialize_cache.jl:10-12)
  bpc = B200BeliefPropagationCache(psi; ctx, messages)
  for v in bpc.verts
    set_bra_factor!(bpc, phi, v)
  end
  return ITensorNetworks.logscalar(ITensorNetworks.update(bpc; update_kwargs...))
end
inner_bp(phi, psi; kwargs...) = exp(loginner_bp(phi, psi; kwargs...))

end # module
