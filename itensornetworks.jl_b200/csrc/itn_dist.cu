// Multi-GPU plumbing: one process per GPU, graph-partitioned network, boundary messages over NCCL.
//
// The reference is single-process (SURVEY.md sections 2 and 5: no communication layer exists), so this file
// has no reference counterpart.  A synchronous sweep (abstractbeliefpropagationcache.jl:294-308) needs, for
// every vertex, only the messages flowing into it; each rank therefore owns the site tensors of its
// vertices, computes the messages leaving them, and once per sweep ships the messages that cross a cut
// to the rank owning their destination: pack (one kernel) -> grouped ncclSend/ncclRecv per peer -> unpack.
// The convergence test adds one ncclAllReduce of a single double.
//
// NCCL is loaded with dlopen (the host process - torch here, Julia in production - usually has it loaded
// already), so the library has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <map>

#include "itn_internal.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) throw ItnError(ITN_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define LOAD(sym)                                                                                   \
  a.sym = (decltype(a.sym))dlsym(a.lib, "nccl" #sym);                                               \
  if (!a.sym) throw ItnError(ITN_ENCCL, "libnccl is missing symbol nccl" #sym);
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommInitAll) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd)
  LOAD(AllReduce) LOAD(GetErrorString)
#undef LOAD
  return a;
}

#define NCCL_CHECK(expr)                                                                             \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) throw ItnError(ITN_ENCCL, std::string(#expr) + ": " + api().GetErrorString(_r)); \
  } while (0)

struct HaloJob {
  double* msg;    // planar message
  long long off;  // offset (doubles) in the packed buffer
  int n;          // doubles (all planes)
};
__global__ void k_halo_pack(const HaloJob* __restrict__ jobs, double* __restrict__ buf) {
  const HaloJob J = jobs[blockIdx.x];
  for (int i = threadIdx.x; i < J.n; i += blockDim.x) buf[J.off + i] = J.msg[i];
}
__global__ void k_halo_unpack(const HaloJob* __restrict__ jobs, const double* __restrict__ buf) {
  const HaloJob J = jobs[blockIdx.x];
  for (int i = threadIdx.x; i < J.n; i += blockDim.x) J.msg[i] = buf[J.off + i];
}

struct Peer {
  int rank;
  long long send_off, send_n, recv_off, recv_n;  // doubles
};
struct DistPlan {
  std::vector<int> dids;  // the sweep this plan was built for
  uint64_t topo_version = ~0ull;
  std::vector<Peer> peers;
  int nsend = 0, nrecv = 0;
  HaloJob *d_send = nullptr, *d_recv = nullptr;
  double *sendbuf = nullptr, *recvbuf = nullptr;
  // split exchange (itn_dist_exchange_begin / _end): pack + send / recv run on a side stream while the context stream
  // computes the part of the sweep that no other rank waits for
  cudaStream_t comm = nullptr;
  cudaEvent_t ready = nullptr, done = nullptr;
  bool in_flight = false;
};

void free_plan(itn_net* net, DistPlan* p) {
  itn_ctx* ctx = net->ctx;
  itn_dev_free(ctx, p->d_send);
  itn_dev_free(ctx, p->d_recv);
  itn_dev_free(ctx, p->sendbuf);
  itn_dev_free(ctx, p->recvbuf);
  p->d_send = p->d_recv = nullptr;
  p->sendbuf = p->recvbuf = nullptr;
  p->peers.clear();
}

}  // namespace

bool itn_is_local(const itn_net* net, int v) { return net->ctx->nranks == 1 || net->owner[v] == net->ctx->rank; }

void itn_dist_release(itn_net* net) {
  if (!net->dist) return;
  DistPlan* p = (DistPlan*)net->dist;
  free_plan(net, p);
  if (p->comm) cudaStreamDestroy(p->comm);
  if (p->ready) cudaEventDestroy(p->ready);
  if (p->done) cudaEventDestroy(p->done);
  delete p;
  net->dist = nullptr;
}

void itn_dist_allreduce_sum(itn_ctx* ctx, double* dev, int n) {
  if (ctx->nranks == 1) return;
  ITN_REQUIRE(ctx->nccl, ITN_ENCCL, "context is not initialised for multi-GPU use (itn_ctx_init_dist)");
  NCCL_CHECK(api().AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
}

void itn_dist_p2p(itn_ctx* ctx, const std::vector<P2PSeg>& segs) {
  if (ctx->nranks == 1) return;
  ITN_REQUIRE(ctx->nccl, ITN_ENCCL, "context is not initialised for multi-GPU use (itn_ctx_init_dist)");
  bool any = false;
  for (const P2PSeg& sg : segs) any = any || sg.sn || sg.rn;
  if (!any) return;
  NCCL_CHECK(api().GroupStart());
  for (const P2PSeg& sg : segs) {
    if (sg.sn) NCCL_CHECK(api().Send(sg.sbuf, sg.sn, ncclDouble, sg.rank, (ncclComm_t)ctx->nccl, ctx->stream));
    if (sg.rn) NCCL_CHECK(api().Recv(sg.rbuf, sg.rn, ncclDouble, sg.rank, (ncclComm_t)ctx->nccl, ctx->stream));
  }
  NCCL_CHECK(api().GroupEnd());
}

static DistPlan* ensure_plan(itn_net* net, const std::vector<int>& dids) {
  itn_ctx* ctx = net->ctx;
  ITN_REQUIRE(ctx->nccl, ITN_ENCCL, "context is not initialised for multi-GPU use (itn_ctx_init_dist)");
  DistPlan* p = (DistPlan*)net->dist;
  if (!p) net->dist = p = new DistPlan();
  if (p->topo_version != net->topo_version || p->dids != dids) {
    free_plan(net, p);
    p->dids = dids;
    p->topo_version = net->topo_version;
    // messages u -> v of this sweep with owner[u] != owner[v]; both sides order them by directed id
    std::vector<int> sorted = dids;
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    std::map<int, std::vector<int>> send, recv;  // peer -> directed ids
    for (int did : sorted) {
      const int e = did / 2;
      const int u = (did & 1) ? net->edst[e] : net->esrc[e];
      const int v = (did & 1) ? net->esrc[e] : net->edst[e];
      const int ou = net->owner[u], ov = net->owner[v];
      if (ou == ov) continue;
      if (ou == ctx->rank) send[ov].push_back(did);
      if (ov == ctx->rank) recv[ou].push_back(did);
    }
    std::vector<HaloJob> sj, rj;
    long long soff = 0, roff = 0;
    std::map<int, Peer> peers;
    const int P = net->planes();
    for (auto& kv : send) {
      Peer& pr = peers[kv.first];
      pr.rank = kv.first;
      pr.send_off = soff;
      for (int did : kv.second) {
        ITN_REQUIRE(net->M[did].p, ITN_EINVAL, "boundary message is not allocated");
        const int nd = (int)(net->M[did].n * P);
        sj.push_back({net->M[did].p, soff, nd});
        soff += nd;
      }
      pr.send_n = soff - pr.send_off;
    }
    for (auto& kv : recv) {
      Peer& pr = peers[kv.first];
      pr.rank = kv.first;
      pr.recv_off = roff;
      for (int did : kv.second) {
        ITN_REQUIRE(net->M[did].p, ITN_EINVAL, "boundary message is not allocated");
        const int nd = (int)(net->M[did].n * P);
        rj.push_back({net->M[did].p, roff, nd});
        roff += nd;
      }
      pr.recv_n = roff - pr.recv_off;
    }
    for (auto& kv : peers) p->peers.push_back(kv.second);
    p->nsend = (int)sj.size();
    p->nrecv = (int)rj.size();
    if (p->nsend) {
      p->d_send = (HaloJob*)itn_dev_alloc(ctx, sj.size() * sizeof(HaloJob));
      p->sendbuf = (double*)itn_dev_alloc(ctx, (size_t)soff * sizeof(double));
      CUDA_CHECK(cudaMemcpyAsync(p->d_send, sj.data(), sj.size() * sizeof(HaloJob), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (p->nrecv) {
      p->d_recv = (HaloJob*)itn_dev_alloc(ctx, rj.size() * sizeof(HaloJob));
      p->recvbuf = (double*)itn_dev_alloc(ctx, (size_t)roff * sizeof(double));
      CUDA_CHECK(cudaMemcpyAsync(p->d_recv, rj.data(), rj.size() * sizeof(HaloJob), cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // sj / rj are about to go out of scope
  }
  return p;
}

// pack -> grouped send / recv on stream `st`
static void exchange_on(itn_net* net, DistPlan* p, cudaStream_t st) {
  itn_ctx* ctx = net->ctx;
  if (p->nsend) {
    k_halo_pack<<<p->nsend, 128, 0, st>>>(p->d_send, p->sendbuf);
    ITN_LAUNCH_CHECK(ctx);
  }
  NCCL_CHECK(api().GroupStart());
  for (const Peer& pr : p->peers) {
    if (pr.send_n) NCCL_CHECK(api().Send(p->sendbuf + pr.send_off, (size_t)pr.send_n, ncclDouble, pr.rank, (ncclComm_t)ctx->nccl, st));
    if (pr.recv_n) NCCL_CHECK(api().Recv(p->recvbuf + pr.recv_off, (size_t)pr.recv_n, ncclDouble, pr.rank, (ncclComm_t)ctx->nccl, st));
  }
  NCCL_CHECK(api().GroupEnd());
}

void itn_dist_exchange(itn_net* net, const std::vector<int>& dids) {
  itn_ctx* ctx = net->ctx;
  if (ctx->nranks == 1) return;
  DistPlan* p = ensure_plan(net, dids);
  exchange_on(net, p, ctx->stream);
  if (p->nrecv) {
    k_halo_unpack<<<p->nrecv, 128, 0, ctx->stream>>>(p->d_recv, p->recvbuf);
    ITN_LAUNCH_CHECK(ctx);
  }
}

void itn_dist_exchange_prepare(itn_net* net, const std::vector<int>& dids) {
  if (net->ctx->nranks == 1) return;
  DistPlan* p = ensure_plan(net, dids);
  if (!p->comm) {
    // highest priority: the few CTAs of pack / send / recv are dispatched ahead of the thousands of queued sweep CTAs
    int lo_prio = 0, hi_prio = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CUDA_CHECK(cudaStreamCreateWithPriority(&p->comm, cudaStreamNonBlocking, hi_prio));
    CUDA_CHECK(cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
  }
}

// Every message this rank sends has been committed on the context stream: pack and send / receive on the side stream.
void itn_dist_exchange_begin(itn_net* net) {
  itn_ctx* ctx = net->ctx;
  if (ctx->nranks == 1) return;
  DistPlan* p = (DistPlan*)net->dist;
  ITN_REQUIRE(p && p->comm, ITN_EINVAL, "halo exchange is not prepared");
  CUDA_CHECK(cudaEventRecord(p->ready, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(p->comm, p->ready, 0));
  exchange_on(net, p, p->comm);
  CUDA_CHECK(cudaEventRecord(p->done, p->comm));
  p->in_flight = true;
}

// Every local read of the pre-sweep boundary messages has been enqueued on the context stream: wait for the transfer and
// unpack the received messages there.
void itn_dist_exchange_end(itn_net* net) {
  itn_ctx* ctx = net->ctx;
  if (ctx->nranks == 1) return;
  DistPlan* p = (DistPlan*)net->dist;
  if (!p || !p->in_flight) return;
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, p->done, 0));
  if (p->nrecv) {
    k_halo_unpack<<<p->nrecv, 128, 0, ctx->stream>>>(p->d_recv, p->recvbuf);
    ITN_LAUNCH_CHECK(ctx);
  }
  p->in_flight = false;
}

#define API_BEGIN try {
#define API_END                              \
  }                                          \
  catch (const ItnError& e) {                \
    itn_set_error(e.what());                 \
    return e.code;                           \
  }                                          \
  catch (const std::exception& e) {          \
    itn_set_error(e.what());                 \
    return ITN_EINVAL;                       \
  }                                          \
  return ITN_OK;

extern "C" int itn_nccl_unique_id(void* out_128_bytes) {
  API_BEGIN
  ITN_REQUIRE(out_128_bytes, ITN_EINVAL, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  NCCL_CHECK(api().GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  API_END
}

extern "C" int itn_ctx_init_dist(itn_ctx* ctx, int rank, int nranks, const void* id_128_bytes) {
  API_BEGIN
  ITN_REQUIRE(ctx && id_128_bytes, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, ITN_EINVAL, "bad rank / nranks");
  ITN_REQUIRE(!ctx->nccl, ITN_EINVAL, "context is already initialised for multi-GPU use");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_128_bytes, sizeof(id));
  ncclComm_t comm;
  NCCL_CHECK(api().CommInitRank(&comm, nranks, id, rank));
  ctx->nccl = comm;
  ctx->nccl_lib = api().lib;
  ctx->rank = rank;
  ctx->nranks = nranks;
  API_END
}

// Single-process multi-GPU ("so Julia needs no MPI"): n contexts, one per listed device, whose
// communicators come from one ncclCommInitAll call; context i is rank i of n.  Every entry point that involves an
// exchange (itn_bp_update, itn_apply2, itn_rdm2, ...) is collective and may wait on the host for its peers, so the
// caller drives each context from its own host thread (Threads.@spawn per GPU in Julia, a Python thread per GPU in
// tests/test_gpu_dist.py) exactly as it would drive one process per GPU.
extern "C" int itn_ctx_create_group(int n, const int32_t* devices, itn_ctx** out_n) {
  API_BEGIN
  ITN_REQUIRE(n >= 1 && devices && out_n, ITN_EINVAL, "bad arguments");
  std::vector<int> devs(devices, devices + n);
  {
    std::vector<int> sorted = devs;
    std::sort(sorted.begin(), sorted.end());
    ITN_REQUIRE(std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end(), ITN_EINVAL,
                "itn_ctx_create_group: every context needs its own device");
  }
  std::vector<itn_ctx*> ctxs(n, nullptr);
  std::vector<ncclComm_t> comms(n, nullptr);
  try {
    for (int i = 0; i < n; ++i) {
      const int st = itn_ctx_create(devs[i], nullptr, &ctxs[i]);
      if (st != ITN_OK) throw ItnError(st, itn_last_error());
    }
    if (n > 1) NCCL_CHECK(api().CommInitAll(comms.data(), n, devs.data()));
  } catch (...) {
    for (itn_ctx* c : ctxs)
      if (c) itn_ctx_destroy(c);
    throw;
  }
  for (int i = 0; i < n; ++i) {
    if (n > 1) {
      ctxs[i]->nccl = comms[i];
      ctxs[i]->nccl_lib = api().lib;
    }
    ctxs[i]->rank = i;
    ctxs[i]->nranks = n;
    out_n[i] = ctxs[i];
  }
  API_END
}

extern "C" int itn_ctx_rank(const itn_ctx* ctx, int32_t* rank, int32_t* nranks) {
  API_BEGIN
  ITN_REQUIRE(ctx, ITN_EINVAL, "NULL argument");
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  API_END
}
