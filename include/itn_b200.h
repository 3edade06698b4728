/* libitn_b200 — C ABI of the B200-native belief-propagation / simple-update engine.
 *
 * The reference (ITensorNetworks.jl) has no FFI: its extension mechanism is Julia multiple
 * dispatch on `AbstractBeliefPropagationCache` (src/caches/abstractbeliefpropagationcache.jl:44-69)
 * plus the `cache!` / `cache_update_kwargs` protocol (src/expect.jl:21-41, src/normalize.jl:13-34,
 * src/environment.jl:21-39, src/contract.jl:41-58).  A Julia `B200BeliefPropagationCache <:
 * AbstractBeliefPropagationCache` holds an `itn_net*` and forwards those methods to the entry points
 * below through `ccall`.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - status codes: 0 = OK, >0 = error; `itn_last_error()` has the text (thread local).  No C++
 *     exception or longjmp crosses this boundary.
 *   - the library owns every device allocation; host pointers are borrowed for the duration of one
 *     call only.  Every call that writes through a host pointer returns after the data is complete.
 *   - dtype 0 = Float64, 1 = ComplexF64 (host layout: interleaved re,im as in Julia/NumPy).
 *   - vertices are 0..nv-1, undirected edges 0..ne-1 with endpoints (esrc[e], edst[e]).
 *     Directed message "u -> v" is addressed by its endpoints.
 *   - host tensors are column-major (Julia order).  A site tensor's axes are described by
 *     `axis_edge[i]` = edge id carried by axis i, or -1 for the physical (site) axis.
 *   - messages are chi x chi column-major matrices M[a, a'] with a = ket-side bond index and
 *     a' = bra-side (primed) copy.
 *   - a handle is used by one host thread at a time; handles of DIFFERENT contexts may be used concurrently, handles
 *     that share a context share its stream and block cache and must be serialised by the caller.
 */
#ifndef ITN_B200_H
#define ITN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct itn_ctx itn_ctx;
typedef struct itn_net itn_net;

enum {
  ITN_OK = 0,
  ITN_EINVAL = 1,       /* bad argument (reference: `error(...)` -> ErrorException) */
  ITN_ESHAPE = 2,       /* shape / dimension mismatch */
  ITN_ECUDA = 3,        /* CUDA runtime failure (incl. "no device": there is no CPU fallback) */
  ITN_ENCCL = 4,        /* NCCL failure */
  ITN_ENOMEM = 5,       /* device allocation failed */
  ITN_EUNSUPPORTED = 6  /* valid in the reference but outside this engine's scope */
};

enum { ITN_F64 = 0, ITN_C128 = 1 };

/* ---- library / context ------------------------------------------------------------------- */

const char* itn_last_error(void);
int itn_version(void);

/* One context per process and GPU.  `stream` may be an existing cudaStream_t (e.g. the host
 * framework's current stream, so that the host can bracket calls with its own CUDA events) or
 * NULL to let the library create one. */
int itn_ctx_create(int device, void* stream, itn_ctx** out);
int itn_ctx_destroy(itn_ctx* ctx);
int itn_ctx_sync(itn_ctx* ctx);

/* Multi-GPU (one process per GPU).  Rank 0 obtains an id, the host broadcasts its 128 bytes by
 * any means (MPI.jl, Distributed, torch.distributed), every rank calls itn_ctx_init_dist.  The
 * library then exchanges boundary messages itself over NCCL/NVLink; no reference counterpart
 * (the reference is single-process). */
int itn_nccl_unique_id(void* out_128_bytes);
int itn_ctx_init_dist(itn_ctx* ctx, int rank, int nranks, const void* id_128_bytes);
/* Single-process multi-GPU: n contexts on n distinct devices with communicators from one
 * ncclCommInitAll; out_n[i] is rank i of n.  Collective entry points may wait for their peers on the host: drive every
 * context from its own host thread (one Julia task / Python thread per GPU), as one would drive one process per GPU. */
int itn_ctx_create_group(int n, const int32_t* devices, itn_ctx** out_n);
int itn_ctx_rank(const itn_ctx* ctx, int32_t* rank, int32_t* nranks);

/* ---- network = partitioned <psi|psi> tensor network + messages ----------------------------
 * Replaces BeliefPropagationCache(ptn; messages) (src/caches/beliefpropagationcache.jl:13-35) over
 * QuadraticFormNetwork(psi) with the default one-site partition
 * (src/formnetworks/abstractformnetwork.jl:89-91): only the ket is stored, bra = conj(ket) and the
 * identity operator layer are implicit.  `owner[v]` = rank owning vertex v (NULL: all on rank 0). */
int itn_net_create(itn_ctx* ctx, int dtype, int nv, int ne, const int32_t* esrc, const int32_t* edst,
                   const int32_t* edim, const int32_t* sdim, const int32_t* owner, itn_net** out);
int itn_net_clone(const itn_net* net, itn_net** out); /* Base.copy (beliefpropagationcache.jl:43-47) */
int itn_net_destroy(itn_net* net);                    /* callable from a Julia finalizer */
int itn_sync(itn_net* net);

int itn_net_edge_dim(const itn_net* net, int e, int32_t* out);
int itn_net_tensor_size(const itn_net* net, int v, int64_t* out_elems);

/* bpc[v] = t / update_factor (abstractbeliefpropagationcache.jl:158-171, beliefpropagationcache.jl:88-91).
 * Bond extents must match the current edge dims. */
int itn_net_set_tensor(itn_net* net, int v, const void* host, int nd, const int32_t* axis_edge);
int itn_net_get_tensor(const itn_net* net, int v, void* host, int nd, const int32_t* axis_edge);

/* Bra layer of a bilinear form <phi|psi>: BilinearFormNetwork(phi, psi) with the identity operator layer
 * (src/formnetworks/bilinearformnetwork.jl:23-42, inner_network src/inner.jl:139-171), the network behind
 * inner(phi, psi; alg = "bp") and loginner (src/inner.jl:100-171).  phi_v is passed as given (NOT conjugated: the engine
 * applies `dag`), with the axis description of itn_net_set_tensor and the same extents as the ket; hosts holding
 * different bond dimensions zero-pad the smaller tensor.  Vertices without a bra tensor keep bra = conj(ket).
 * While a bra layer is set, messages are general (not Hermitian) matrices M[a_ket, a_bra] and only itn_bp_update,
 * itn_updated_message, itn_message_residuals, itn_region_scalars, itn_logscalar, the message accessors and
 * itn_net_clone are defined; observables, rescale and gates return ITN_EUNSUPPORTED (as in the reference, where
 * expect / apply / normalize take a QuadraticFormNetwork).  <phi|A|psi> (src/inner.jl:154-171) enters as
 * <phi|(A psi)> with the operator layer contracted into the ket site by site on the host (fused bond indices).
 * itn_net_clear_bra returns to the quadratic form <psi|psi>. */
int itn_net_set_bra_tensor(itn_net* net, int v, const void* host, int nd, const int32_t* axis_edge);
int itn_net_clear_bra(itn_net* net);

/* The constructor's data path, BeliefPropagationCache(ptn) over the whole network
 * (src/caches/beliefpropagationcache.jl:20-35): n site tensors in one call.  hosts[i] is the column-major
 * tensor of vertex verts[i]; nd / axis_edge (nullable = canonical [site, bonds...]) are the per-tensor axis
 * counts and the concatenated axis descriptions of itn_net_set_tensor.  The copy is pipelined: a second stream
 * moves chunk c + 1 over PCIe while the main stream de-interleaves and permutes chunk c.
 *   flags = 0                  returns when every host buffer has been consumed.
 *   flags = ITN_HOST_DEFERRED  only registers the buffers; they must stay valid and unchanged until the next
 *       call on this handle that consumes tensors returns (itn_bp_update, itn_sync, any observable or gate).
 *       A synchronous itn_bp_update then runs its first sweep vertex chunk by vertex chunk behind the copy,
 *       so the upload of psi overlaps the DMMA kernels instead of preceding them. */
enum { ITN_HOST_DEFERRED = 1 };
int itn_net_set_tensors(itn_net* net, int n, const int32_t* verts, const void* const* hosts, const int32_t* nd,
                        const int32_t* axis_edge, int flags);

/* identity_messages (src/formnetworks/quadraticformnetwork.jl:96-124): delta on every directed edge. */
int itn_msg_set_identity(itn_net* net);
/* set_message! / message (abstractbeliefpropagationcache.jl:173-200). */
int itn_msg_set(itn_net* net, int src, int dst, const void* host_chi2);
int itn_msg_get(const itn_net* net, int src, int dst, void* host_chi2);
/* messages(bpc) in one call: every message stored on this rank, back to back in directed-id order
 * (2e = esrc->edst, 2e+1 = edst->esrc), each chi x chi column-major; `bytes` = capacity of `host`
 * (may be pinned memory: one device->host copy).  Unset / remote messages are skipped. */
int itn_msg_get_all(const itn_net* net, void* host, int64_t bytes);

/* ---- belief propagation ------------------------------------------------------------------ */

/* update(::Algorithm"bp", bpc) (abstractbeliefpropagationcache.jl:313-329) with
 * update_iteration sequential (:272-287; group_ptr == NULL: Gauss-Seidel over the list, executed as
 * dependency wavefronts with identical arithmetic) or grouped/synchronous (:294-308; group_ptr =
 * ngroups+1 offsets into the list; every group reads the pre-sweep messages, results are written
 * back at the end of the sweep.  Single-edge groups are the synchronous sweep that the batched DMMA kernels and the
 * multi-GPU path run; a group of several edges is a sequential pass over its edges that starts from the pre-sweep
 * messages (single GPU, per-message kernels), and the diff of a sweep is divided by the number of groups as in :319-321).
 * updated_message(::Algorithm"contract") (:225-239) and message_diff (:32-36) run on the device.
 * tol < 0: `tol = nothing` (no diff computed).  Outputs may be NULL. */
int itn_bp_update(itn_net* net, const int32_t* seq_src, const int32_t* seq_dst, int nseq,
                  const int32_t* group_ptr, int ngroups, int maxiter, double tol, int normalize,
                  int32_t* iters, double* last_mean_diff);

/* updated_message(bpc, edge) without storing it (test_belief_propagation.jl:51). */
int itn_updated_message(itn_net* net, int src, int dst, int normalize, void* host_chi2);
/* message_diff(updated_message(bpc, e), message(bpc, e)) for the n given directed edges. */
int itn_message_residuals(itn_net* net, const int32_t* src, const int32_t* dst, int n, double* out_n);

/* ---- scalars, rescale, observables --------------------------------------------------------- */

/* vertex_scalars / edge_scalars (abstract :83-97; region_scalar beliefpropagationcache.jl:107-119).
 * z_v: nv scalars, z_e: ne scalars, in the network dtype. */
int itn_region_scalars(itn_net* net, void* z_v, void* z_e);
/* logscalar(bpc) (abstract :397-408); out = {re, im}; re = -Inf if an edge scalar is zero. */
int itn_logscalar(itn_net* net, double out_re_im[2]);
/* rescale(bpc) = rescale_messages + rescale_partitions over all ket/bra vertices
 * (beliefpropagationcache.jl:121-139, abstract :349-395; normalize.jl:63-80). */
int itn_rescale(itn_net* net);
/* rescale(bpc; verts) / rescale_partitions(bpc, partitions; verts) (abstract :349-395): every message pair is rescaled
 * (rescale_messages always runs over all quotient edges), but only the site tensors of the n listed vertices (ket and
 * bra of each, the k = 2 case of :363-372) are divided so that their region scalar becomes 1; the other vertices keep
 * their tensors.  n = 0 rescales the messages only. */
int itn_rescale_verts(itn_net* net, const int32_t* verts, int n);

/* expect(psiIpsi, op) (src/expect.jl:5-19) for n vertices; ops: n matrices d x d, column-major
 * O[s_out, s_in]; out: n scalars (network dtype). */
int itn_expect1(itn_net* net, const int32_t* verts, int n, const void* ops, void* out_n);
/* two-site RDM idiom (test_belief_propagation.jl:64-91) on n edges: out = n matrices
 * (d_u d_v) x (d_u d_v), column-major, row index = s_u + d_u * s_v, unit trace. */
int itn_rdm2(itn_net* net, const int32_t* eids, int n, void* out);

/* ---- gates ---------------------------------------------------------------------------------- */

/* one-site apply (src/apply.jl:108-116): A'[s',..] = sum_s gate[s', s] A[s,..]. */
int itn_apply1(itn_net* net, const int32_t* verts, int n, const void* gates_n_dxd, int normalize);

/* two-site apply with BP product environments = simple_update_bp (src/apply.jl:33-95,117-139) on a
 * vertex-disjoint batch of n edges.  gates: n arrays g[s1', s2', s1, s2] column-major with
 * (1, 2) = (esrc, edst) of the edge.  maxdim <= 0: none; cutoff < 0: `cutoff = nothing`.
 * The environments are the cache's current messages.  Outputs (nullable): new bond dimension,
 * truncation error (callback's truncation_error, apply.jl:89) and singular values (svals_stride
 * doubles per edge, zero padded).  msg_mode: what to store as both messages on a gated edge —
 * 0 = identity, 1 = diag(singular values) (the Vidal-gauge fixed point of the updated pair). */
int itn_apply2(itn_net* net, const int32_t* eids, int n, const void* gates, int maxdim,
               double cutoff, int normalize, int msg_mode, int32_t* newdim_n, double* truncerr_n,
               double* svals, int svals_stride);

/* The gauged simple-update loop as one call (`apply(o, psi; cache_update_kwargs, maxdim, cutoff)` of
 * north_star): for each colour layer l = 0 .. nlayers-1, itn_apply2 on eids[layer_ptr[l] .. layer_ptr[l+1]) (gates
 * concatenated in the same order, outputs indexed like eids), followed - when bp_maxiter > 0 - by itn_bp_update with
 * the given sequence (same meaning of seq / group_ptr / tol / normalize as itn_bp_update), so that the next layer sees
 * refreshed environments.  This is the host loop `for layer: psi = apply(gates, psi; envs); bpc = update(bpc)` of a
 * TEBD driver without returning to the host language between its parts.  bp_iters_total (nullable): sweeps run. */
int itn_apply_layers(itn_net* net, int nlayers, const int32_t* layer_ptr, const int32_t* eids, const void* gates,
                     int maxdim, double cutoff, int normalize, int msg_mode, const int32_t* seq_src,
                     const int32_t* seq_dst, int nseq, const int32_t* group_ptr, int ngroups, int bp_maxiter,
                     double bp_tol, int bp_normalize, int32_t* newdim_n, double* truncerr_n, double* svals,
                     int svals_stride, int32_t* bp_iters_total);

/* gauge_walk(tn, edges) (src/abstractitensornetwork.jl:387-393): qr!(tn, src[i] => dst[i]) for i = 0 .. n-1, the loop
 * behind tree_gauge / tree_orthogonalize (:395-420), which apply(o, psi; ortho = true) runs towards the first gate vertex
 * before the update (src/apply.jl:109-111, 130-132).  After step i the tensor of src[i] is an isometry from (site, other
 * bonds) to the bond and the square factor has been multiplied into dst[i]; the state is unchanged.  The factor is the
 * Hermitian one (G^(1/2) of the Gram matrix, twice: orthonormal to eps), not Householder's triangular R: the gauge differs
 * from the reference's by a unitary on the bond, every gauge-invariant quantity agrees.  Messages are left untouched (they
 * belong to the old gauge: re-run itn_bp_update or reset them).  Single GPU. */
int itn_gauge_walk(itn_net* net, const int32_t* src, const int32_t* dst, int n);

/* map_eigvals(f, A, ...; ishermitian = true, cutoff) (src/apply.jl:21-25) for a batch of n
 * Hermitian chi x chi host matrices; fn: 0 = sqrt, 1 = inv o sqrt, 2 = inv.  cutoff < 0: none. */
int itn_map_eigvals(itn_ctx* ctx, int dtype, int fn, int chi, int n, const void* host_in,
                    void* host_out, double cutoff);

/* ---- generalised partitions ------------------------------------------------------------------ */

/* Contraction of two host tensors over `npairs` axis pairs (numpy.tensordot convention: the result carries the free axes
 * of a, then the free axes of b), column-major, on the device.  This is the `contract` of the tensors inside one
 * multi-site partition (src/caches/abstractbeliefpropagationcache.jl:232-233 for partitioned_vertices with several sites
 * per partition, test/test_expect.jl:22-39): the host merges the site tensors of a partition into one super-site tensor
 * (fused site and bond indices) with it and hands the resulting network to itn_net_create; BP, scalars and expect then
 * run on the existing kernels.  Small tensors, set-up work. */
int itn_tensordot(itn_ctx* ctx, int dtype, const void* a_host, int nda, const int32_t* dims_a, const void* b_host, int ndb,
                  const int32_t* dims_b, int npairs, const int32_t* axes_a, const int32_t* axes_b, void* out_host);

/* ---- instrumentation ------------------------------------------------------------------------ */

/* Singular values (sorted descending, n per matrix) and optionally U*Sigma (columns in the kernel's internal order, not
 * sorted) of `batch` host matrices m x n, column-major, through the batched one-sided Jacobi kernels that
 * simple_update_bp's `factorize_svd` (src/apply.jl:81-88) runs on.  variant: 0 = the kernel the gate path selects,
 * 1 = the shape-generic kernel, 2 = the round-robin m, n <= 64 kernel (second opinions).  device_ms (nullable): CUDA-event time of the decomposition alone. */
int itn_svd_batch(itn_ctx* ctx, int dtype, int m, int n, int batch, const void* host_in, double* host_sigma,
                  void* host_us, int variant, double* device_ms);

/* The block geometry, operation lists and shared-memory fibre tables the block path (csrc/itn_block.cu) uses for a vertex
 * with site dimension d, degree z and bond extents chi[0..z) in a bucket of `nverts` vertices, as a flat int32 stream
 * (layout: csrc/itn_block.cu).  Host only: needs no device.  *nout = number of entries (0: the signature has no plan and
 * runs on the shape-generic kernels); `out` is filled when cap >= *nout.  tests/block_emulator.py replays it in NumPy. */
int itn_block_plan_export(int dtype, int d, int z, const int32_t* chi, int nverts, int32_t* out, int cap, int32_t* nout);

/* Number of kernels this library has launched on ctx since creation (bench.py "gpu_launches"). */
int itn_ctx_launch_count(const itn_ctx* ctx, int64_t* out);
/* Message updates computed so far on ctx by {the degree-4 chi=16 tile path (csrc/itn_fast.cu), the block path
 * (csrc/itn_block.cu), the shape-generic per-message / per-vertex kernels (csrc/itn_generic.cu)}: which kernels ran. */
int itn_ctx_path_counts(const itn_ctx* ctx, int64_t* out3);
/* Gate sides (itn_apply2 / itn_apply_layers on ctx so far) whose R factor was refined by a second Cholesky pass on
 * A R1^+ (CholeskyQR2): the sides where the reference's QR (src/apply.jl:70-76) sees an ill-conditioned matrix. */
int itn_ctx_cholqr2_count(const itn_ctx* ctx, int64_t* out);
/* Select the message-update implementation: 0 = auto (DMMA tile path where a bucket qualifies, the shape-generic
 * DMMA kernels otherwise), 1 = shape-generic DMMA kernels only, 2 = plain FMA kernels only (the parity tests use
 * 1 and 2 as second and third opinions on the device). */
int itn_ctx_set_path(itn_ctx* ctx, int mode);
/* Device time of the last itn_bp_update in milliseconds (CUDA events on the context stream) and
 * the share spent in the dominant contraction kernels. */
int itn_bp_last_timing(const itn_net* net, double* total_ms, double* contract_ms);

#ifdef __cplusplus
}
#endif
#endif /* ITN_B200_H */
