// Internal data model of libitn_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "itn_b200.h"

#define ITN_MAX_MODES 10  // site + up to 9 bonds

struct ItnError : std::runtime_error {
  int code;
  ItnError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void itn_set_error(const std::string& s);

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw ItnError(_e == cudaErrorMemoryAllocation ? ITN_ENOMEM : ITN_ECUDA,                   \
                     std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
  } while (0)

#define ITN_REQUIRE(cond, code, msg)          \
  do {                                        \
    if (!(cond)) throw ItnError((code), (msg)); \
  } while (0)

struct itn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;  // host -> device staging copies that overlap kernels on `stream` (created on demand)
  int rank = 0, nranks = 1;
  void* nccl = nullptr;      // ncclComm_t
  void* nccl_lib = nullptr;  // dlopen handle
  int64_t launches = 0;
  int64_t cholqr2_sides = 0;         // gate sides whose R factor took the second pass (CholeskyQR2, itn_apply2)
  int64_t path_msgs[3] = {0, 0, 0};  // messages computed by the tile path, the block path and the shape-generic kernels
  int nets_alive = 0;          // handles created on this context and not yet destroyed
  bool destroy_pending = false;  // itn_ctx_destroy was called while networks were alive (finalizer order)
  int path_mode = 0;
  int sm_count = 148;
  size_t ws_budget = (size_t)6 << 30;  // scratch budget for the generic path (bytes)
  // Large blocks (slabs of site tensors, tile-major copies, staging) are recycled here instead of going back to the
  // driver's pool: every block is used on `stream` only (or on copy_stream between events that order it against
  // `stream`), so a freed block can be handed to later work on the same stream without synchronising, and the
  // multi-GB allocations of a gate layer or a fresh cache never wait for the driver to map new physical memory.
  std::multimap<size_t, void*> big_free;        // size -> block
  std::unordered_map<void*, size_t> big_live;   // block -> size
  size_t big_cached = 0;                        // bytes parked in big_free
  int* pinned_flags = nullptr;                  // page-locked landing zone of small device flags (grown on demand)
  size_t pinned_flags_n = 0;
};

// Planar storage: re plane [0, n), im plane [n, 2n) (complex only).
// Several tensors created by one batched call (a gate layer) share one device allocation.
struct DevSlab {
  void* base = nullptr;
  int refs = 0;
};
struct DevTensor {
  double* p = nullptr;
  int64_t n = 0;
  DevSlab* slab = nullptr;  // null: p is its own allocation
};

// One step of a mode-product chain: out[l, b, r] = sum_a in[l, a, r] * m(a, b)
struct ModeStep {
  long long L, R;
  int K, N;          // K = summed extent, N = output extent
  const double* m;   // planar matrix, K x N column-major (or N x K if trans)
  long long mplane;  // offset of the imaginary plane of m
  int trans, conj;
};

// "Contract vertex v leaving a set of modes open" job (generic path).
struct VJob {
  const double* a;   // source tensor, planar, canonical order [site, bonds...]
  double* ap;        // permuted copy: closed modes first, open modes last (== a if identity)
  double* w[2];      // ping-pong scratch (planar, n each)
  const double* b;   // bra tensor when the network is a bilinear form <phi|psi> (null: bra = ket)
  double* bp;        // permuted copy of b (== b if identity), the A' of the close below when b is set
  double* out;       // staged result, planar No x No:  out[o + No*o'] = sum_x B[x,o] conj(A'[x,o'])
  long long n;       // elements of the tensor
  long long X;       // product of closed extents
  int No;            // product of open extents
  int nm;            // number of modes
  int dims[ITN_MAX_MODES];
  long long pstride[ITN_MAX_MODES];  // stride (in ap) of source mode i
  int identity_perm;
  int nsteps;
  ModeStep steps[ITN_MAX_MODES];
};

// Axis description of a host tensor (any axis order, interleaved complex) relative to the canonical device layout.
struct Marshal {
  int nd;
  int dims[ITN_MAX_MODES];           // host axis extents
  long long cstride[ITN_MAX_MODES];  // canonical stride of host axis i
};

// A site tensor whose host buffer has been registered but not copied yet (itn_net_set_tensors, ITN_HOST_DEFERRED).
struct PendingUpload {
  int v;
  const void* host;
  Marshal m;
};

struct CommitJob {
  const double* staged;  // planar No x No
  double* dest;          // planar message (may be scratch)
  const double* old;     // planar message to diff against (may be null)
  int n2;                // No*No
};

struct itn_net {
  itn_ctx* ctx = nullptr;
  int dtype = 0;
  bool cplx = false;
  int nv = 0, ne = 0;
  std::vector<int> esrc, edst, edim, sdim, owner;
  std::vector<std::vector<int>> inc;  // incident edge ids, ascending
  std::unordered_map<uint64_t, int> dmap;  // (src,dst) -> directed id (2e: esrc->edst, 2e+1: reverse)
  std::vector<DevTensor> T;  // per vertex
  // bra layer of a BilinearFormNetwork <phi|psi> (src/formnetworks/bilinearformnetwork.jl:23-42): phi_v as given
  // (conjugated on use); empty / null entries mean bra = ket (QuadraticFormNetwork)
  std::vector<DevTensor> Tb;
  int nbra = 0;
  bool has_bra() const { return nbra > 0; }
  const double* bra(int v) const { return (nbra > 0 && Tb[v].p) ? Tb[v].p : T[v].p; }
  std::vector<DevTensor> M;  // per directed edge
  uint64_t topo_version = 0;  // bumped whenever a tensor pointer / bond dim changes
  std::vector<uint64_t> tver;  // per vertex: bumped whenever the contents (or storage) of its site tensor change
  // Lazy canonical copies: after a gate layer on the tile path the new site tensor of a vertex may exist in the tile-major
  // layouts only (k_rebuild skipped the canonical write; storage is allocated).  canon_stale[v] marks it; every reader of
  // T[v] outside the tile kernels goes through itn_canon_ensure* first (itn_flush_pending does it for whole entry points).
  std::vector<char> canon_stale;
  int n_canon_stale = 0;
  void touch(int v) {  // the canonical tensor of v was (re)written
    if (tver.size() != (size_t)nv) tver.assign(nv, 0);
    tver[v]++;
    topo_version++;
    if (!canon_stale.empty() && canon_stale[v]) {
      canon_stale[v] = 0;
      --n_canon_stale;
    }
  }
  double last_total_ms = 0, last_contract_ms = 0;
  std::vector<PendingUpload> pending;  // deferred host tensors: device storage exists, contents arrive with the next consumer
  void* fast = nullptr;  // fast-path cache (owned by itn_fast.cu)
  void* block = nullptr; // block-path cache (owned by itn_block.cu)
  void* dist = nullptr;  // halo exchange plan and buffers (owned by itn_dist.cu)

  int planes() const { return cplx ? 2 : 1; }
  int other(int e, int v) const { return esrc[e] == v ? edst[e] : esrc[e]; }
  int slot(int v, int e) const {
    for (size_t i = 0; i < inc[v].size(); ++i)
      if (inc[v][i] == e) return (int)i;
    return -1;
  }
  // message flowing INTO v along edge e
  int msg_into(int v, int e) const { return edst[e] == v ? 2 * e : 2 * e + 1; }
  int did(int s, int d) const {
    auto it = dmap.find(((uint64_t)(uint32_t)s << 32) | (uint32_t)d);
    return it == dmap.end() ? -1 : it->second;
  }
  long long tensor_elems(int v) const {
    long long n = sdim[v];
    for (int e : inc[v]) n *= edim[e];
    return n;
  }
};

// copies every deferred host tensor (itn_net_set_tensors with ITN_HOST_DEFERRED) to the device; no-op when none is pending
void itn_flush_pending(itn_net* net);  // deferred uploads AND lazy canonical copies: T[v] is valid for every local vertex
void itn_flush_uploads(itn_net* net);  // deferred uploads only (callers that handle canon_stale themselves)
// materialise the canonical copy of vertices whose truth lives in the tile-major layouts (itn_fast.cu)
void itn_canon_ensure_all(itn_net* net);
void itn_canon_ensure(itn_net* net, int v);
void itn_canon_ensure_outside_sweep(itn_net* net);  // every stale vertex that is not part of the planned tile sweep

// ---- device memory helpers (stream ordered) ----
void* itn_dev_alloc(itn_ctx* ctx, size_t bytes);
void itn_dev_free(itn_ctx* ctx, void* p);
// releases the storage of t (its own allocation, or one reference on its slab) and clears it
void itn_tensor_free(itn_ctx* ctx, DevTensor& t);

struct DevBuf {  // RAII scratch buffer
  itn_ctx* ctx;
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf(itn_ctx* c, size_t b) : ctx(c), bytes(b) { p = b ? itn_dev_alloc(c, b) : nullptr; }
  ~DevBuf() {
    if (p) itn_dev_free(ctx, p);
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  template <class T>
  T* as() const { return (T*)p; }
};

// ---- generic vertex-contraction engine (itn_generic.cu) ----
struct JobSpec {
  int v;
  uint32_t open_mask;  // bit i set: mode i (0 = site, 1+k = k-th incident edge) stays open
  double* out;         // device, planar No x No
  // optional: mats[k] != nullptr replaces the incoming message on bond slot k (planar chi x chi, device)
  const double* const* mats = nullptr;
  // true: no message is absorbed on any closed bond (plain Gram matrix of the site tensor over the closed modes)
  bool no_messages = false;
  // optional: a tensor of the same shape that stands in for the site tensor of v (device, canonical planar)
  const double* tensor = nullptr;
};
// Runs all specs in [lo, hi) batches bounded by the workspace budget. Results in spec.out.
void itn_run_vertex_jobs(itn_net* net, const std::vector<JobSpec>& specs);
// All outgoing messages of a vertex in one go (synchronous sweeps), partial absorptions shared between the outputs.
struct SweepSpec {
  int v;
  double* out[ITN_MAX_MODES];  // staged (un-normalised) message leaving along bond slot k, planar chi_k x chi_k
};
bool itn_vertex_sweep_ok(const itn_net* net, int v);
void itn_run_vertex_sweeps(itn_net* net, const std::vector<SweepSpec>& specs);
void itn_run_commit(itn_net* net, const std::vector<CommitJob>& jobs, int normalize, double* d_diffs);
void itn_run_commit_dev(itn_net* net, const CommitJob* d_jobs, size_t n, int normalize, double* d_diffs);
int itn_open_extent(const itn_net* net, int v, uint32_t open_mask);

struct ModeProdSpec {
  const double* src;  // planar, canonical order
  long long n;        // elements
  int nm;
  int dims[ITN_MAX_MODES];
  int nsteps;
  int mode[ITN_MAX_MODES];          // which mode each step acts on
  const double* mat[ITN_MAX_MODES];  // planar dims[mode] x dims[mode]: out[.., b, ..] = sum_a in[.., a, ..] mat[a + K b]
  int trans[ITN_MAX_MODES];          // non-zero: use mat[b + K a] instead
  double *w0, *w1;                   // ping-pong outputs, n * planes doubles each
};
void itn_run_modeprods(itn_ctx* ctx, bool cplx, const std::vector<ModeProdSpec>& specs, std::vector<const double*>& result);

// ---- fast path (itn_fast.cu) ----
// Plans a synchronous sweep: handled[i] = 1 for the message jobs (directed id dids[i], source vertex
// srcv[i]) that the DMMA kernels compute; returns their number.  The sweep writes the un-normalised
// new messages of those jobs to staged[i].
// first / nfirst (optional): vertices flagged in `first` take the leading sweep positions; *nfirst = how many did.
int itn_fast_bp_plan(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv, std::vector<char>& handled,
                     const std::vector<char>* first = nullptr, int* nfirst = nullptr);
void itn_fast_bp_sweep(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv,
                       const std::vector<char>& handled, double* const* staged);
void itn_fast_release(itn_net* net);
// The same sweep in pieces, so that the first sweep can overlap the host -> device upload of the site tensors:
// begin uploads the pointer tables, range runs the three phases for sweep positions [lo, hi), end reduces the
// per-CTA partials into staged[].  itn_fast_bp_sweep = begin + range(0, n) + end.
void itn_fast_bp_sweep_begin(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv,
                             const std::vector<char>& handled, double* const* staged);
void itn_fast_bp_sweep_range(itn_net* net, int lo, int hi);
void itn_fast_bp_sweep_end(itn_net* net);
// Vertices of the planned sweep in sweep-position order, and whether position r is bucket slot r for every r.
const std::vector<int>& itn_fast_sweep_vertices(itn_net* net, bool* contiguous);
// (Re)build the tile-major copies of bucket slots [lo, hi) from the canonical tensors (which must be complete).
void itn_fast_relayout_range(itn_net* net, int lo, int hi);
// simple update on the tile path (degree 4, all bonds 16, d = 2): bond environments and the rebuild A . T
struct FastBenvJob {
  int v, slot;                 // vertex and bond slot of the gate bond
  const double* const* mats;   // [4] messages to absorb per slot (nullptr: the network's own), device planar
  double* C;                   // out: planar n x n, n = 16 d, C[(s + d l) + n (s' + d l')]
};
struct FastRebuildJob {
  int v, slot, chi_new;
  const double* T;             // planar n x (d chi_new)
  double* out;                 // new tensor, canonical planar
  bool lazy = false;           // chi_new == 16: the canonical copy may be left unwritten (itn_net::canon_stale)
};
bool itn_fast_gate_site_ok(itn_net* net, int v);
void itn_fast_bond_envs(itn_net* net, const std::vector<FastBenvJob>& jobs);
void itn_fast_rebuild(itn_net* net, const std::vector<FastRebuildJob>& jobs);
// after the new tensors are committed: marks the tile-major copies that itn_fast_rebuild wrote directly as current
void itn_fast_commit_direct(itn_net* net);

// ---- block path (itn_block.cu): synchronous sweeps of any degree 2..8 / bond extent <= 32 on the canonical layout ----
// plan: handled[i] = 2 for the message jobs the block kernels take (jobs with handled[i] != 0 are left alone); returns
// their number.  begin prepares one itn_bp_update call (device tables, scratch tensors; staged[i] receives the
// un-normalised new message of job i), run executes one sweep, end releases the per-call buffers.
int itn_block_bp_plan(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv, std::vector<char>& handled);
void itn_block_bp_begin(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv,
                        const std::vector<char>& handled, double* const* staged);
void itn_block_bp_run(itn_net* net, bool all_side);  // all_side: the tile path shares the sweep (every bucket on a side stream)
void itn_block_bp_join(itn_net* net);               // the context stream waits for the side streams of the last run
void itn_block_bp_end(itn_net* net);
void itn_block_release(itn_net* net);

// ---- multi-GPU (itn_dist.cu) ----
bool itn_is_local(const itn_net* net, int v);
// Sends every message listed in `dids` whose destination vertex lives on another rank to that rank and
// receives the matching ones (ordered by directed id on both sides); one grouped NCCL send/recv per peer.
void itn_dist_exchange(itn_net* net, const std::vector<int>& dids);
// The same exchange in two halves, overlapped with the interior of the sweep: prepare (plan, side stream) once per call;
// begin after the messages that leave this rank are committed (pack + send / recv on the side stream); end after the last
// local read of the pre-sweep boundary messages is enqueued (the context stream waits for the transfer and unpacks).
void itn_dist_exchange_prepare(itn_net* net, const std::vector<int>& dids);
void itn_dist_exchange_begin(itn_net* net);
void itn_dist_exchange_end(itn_net* net);
void itn_dist_allreduce_sum(itn_ctx* ctx, double* dev, int n);
// One grouped point-to-point exchange on the context stream: per peer, sn doubles out of sbuf and rn doubles into rbuf.
struct P2PSeg {
  int rank;
  const double* sbuf;
  size_t sn;
  double* rbuf;
  size_t rn;
};
void itn_dist_p2p(itn_ctx* ctx, const std::vector<P2PSeg>& segs);
void itn_dist_release(itn_net* net);

// ---- small linear algebra (itn_linalg.cu) ----
// Batched Hermitian Jacobi eigen-decomposition based matrix function, planar matrices on device.
void itn_dev_map_eigvals(itn_ctx* ctx, bool cplx, int fn, int chi, int n, const double* const* d_in_ptrs,
                         double* const* d_out_ptrs, double cutoff);

template <class T>
static inline T* itn_upload(itn_ctx* ctx, const std::vector<T>& h, DevBuf& buf) {
  if (h.empty()) return nullptr;
  cudaError_t e = cudaMemcpyAsync(buf.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) throw ItnError(ITN_ECUDA, std::string("upload: ") + cudaGetErrorString(e));
  // the host vector may die before the copy executes: pageable copies are staged synchronously by the
  // runtime, so this is safe for std::vector sources.
  return (T*)buf.p;
}

// ITN_TRACE=1: host-side phase times of an entry point on stderr (where the host keeps the GPU waiting)
struct HostTrace {
  bool on;
  const char* what;
  std::chrono::steady_clock::time_point last;
  std::string line;
  explicit HostTrace(const char* w) : on(getenv("ITN_TRACE") != nullptr), what(w), last(std::chrono::steady_clock::now()) {}
  void mark(const char* phase) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, " %s %.2f", phase, std::chrono::duration<double, std::milli>(now - last).count());
    line += buf;
    last = now;
  }
  ~HostTrace() {
    if (on) fprintf(stderr, "[itn trace] %s host ms:%s\n", what, line.c_str());
  }
};

#define ITN_LAUNCH_CHECK(ctx)                                                        \
  do {                                                                               \
    (ctx)->launches++;                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      throw ItnError(ITN_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)
