// C ABI of libitn_b200: handle lifetime, host<->device marshalling, BP driver, scalars, observables.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>

#include "itn_internal.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void itn_set_error(const std::string& s) { g_err = s; }
extern "C" const char* itn_last_error(void) { return g_err.c_str(); }
extern "C" int itn_version(void) { return 100; }

#define API_BEGIN try {
#define API_END                                  \
  }                                              \
  catch (const ItnError& e) {                    \
    itn_set_error(e.what());                     \
    return e.code;                               \
  }                                              \
  catch (const std::bad_alloc&) {                \
    itn_set_error("host allocation failed");     \
    return ITN_ENOMEM;                           \
  }                                              \
  catch (const std::exception& e) {              \
    itn_set_error(e.what());                     \
    return ITN_EINVAL;                           \
  }                                              \
  return ITN_OK;

namespace {
constexpr size_t kBigBlock = (size_t)32 << 20;    // blocks of at least 32 MiB are recycled by the context
constexpr size_t kBigRound = (size_t)64 << 20;    // their sizes are rounded up to 64 MiB so that similar requests match
constexpr size_t kBigCacheMax = (size_t)96 << 30; // parked bytes above this are returned to the driver

void big_trim(itn_ctx* ctx, size_t keep) {
  while (ctx->big_cached > keep && !ctx->big_free.empty()) {
    auto it = std::prev(ctx->big_free.end());
    cudaFreeAsync(it->second, ctx->stream);
    ctx->big_cached -= it->first;
    ctx->big_free.erase(it);
  }
}
}  // namespace

void* itn_dev_alloc(itn_ctx* ctx, size_t bytes) {
  void* p = nullptr;
  if (bytes >= kBigBlock) {
    const size_t want = (bytes + kBigRound - 1) / kBigRound * kBigRound;
    auto it = ctx->big_free.lower_bound(want);
    if (it != ctx->big_free.end() && it->first <= want + want / 4) {
      p = it->second;
      ctx->big_live[p] = it->first;
      ctx->big_cached -= it->first;
      ctx->big_free.erase(it);
      return p;
    }
    static const bool trace = getenv("ITN_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    bool retried = false;
    cudaError_t e = cudaMallocAsync(&p, want, ctx->stream);
    if (e != cudaSuccess) {  // give the parked blocks back and retry once
      cudaGetLastError();
      big_trim(ctx, 0);
      cudaStreamSynchronize(ctx->stream);
      e = cudaMallocAsync(&p, want, ctx->stream);
      retried = true;
    }
    if (trace) {
      size_t live = 0;
      for (auto& kv : ctx->big_live) live += kv.second;
      fprintf(stderr, "[itn trace] big block miss: %zu MiB in %.1f ms%s (parked %zu MiB in %zu blocks, live %zu MiB)\n", want >> 20,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
              retried ? " after trimming the cache" : "", ctx->big_cached >> 20, ctx->big_free.size(), live >> 20);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw ItnError(ITN_ENOMEM, std::string("device allocation of ") + std::to_string(bytes) +
                                     " bytes failed: " + cudaGetErrorString(e));
    }
    ctx->big_live[p] = want;
    return p;
  }
  cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    big_trim(ctx, 0);
    cudaStreamSynchronize(ctx->stream);
    e = cudaMallocAsync(&p, bytes ? bytes : 16, ctx->stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw ItnError(ITN_ENOMEM, std::string("device allocation of ") + std::to_string(bytes) +
                                   " bytes failed: " + cudaGetErrorString(e));
  }
  return p;
}
void itn_dev_free(itn_ctx* ctx, void* p) {
  if (!p) return;
  auto it = ctx->big_live.find(p);
  if (it != ctx->big_live.end()) {
    ctx->big_free.emplace(it->second, p);
    ctx->big_cached += it->second;
    ctx->big_live.erase(it);
    if (ctx->big_cached > kBigCacheMax) big_trim(ctx, kBigCacheMax * 3 / 4);
    return;
  }
  cudaFreeAsync(p, ctx->stream);
}

void itn_tensor_free(itn_ctx* ctx, DevTensor& t) {
  if (t.slab) {
    if (--t.slab->refs == 0) {
      itn_dev_free(ctx, t.slab->base);
      delete t.slab;
    }
  } else if (t.p) {
    itn_dev_free(ctx, t.p);
  }
  t.p = nullptr;
  t.n = 0;
  t.slab = nullptr;
}

static void set_device(const itn_ctx* ctx) { CUDA_CHECK(cudaSetDevice(ctx->device)); }

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
namespace {

// host (interleaved complex, arbitrary axis order) -> device planar canonical
template <bool C>
__global__ void k_import(const double* __restrict__ in, double* __restrict__ out, long long n, Marshal m) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i, off = 0;
    for (int a = 0; a < m.nd; ++a) {
      int d = m.dims[a];
      long long q = r / d;
      off += (r - q * d) * m.cstride[a];
      r = q;
    }
    if (C) {
      out[off] = in[2 * i];
      out[n + off] = in[2 * i + 1];
    } else {
      out[off] = in[i];
    }
  }
}
// the same for a batch of tensors staged back to back: block (job, y)
struct ImportJob {
  const double* src;  // staged host bytes
  double* dst;        // canonical planar tensor
  long long n;
  Marshal m;
};
template <bool C>
__global__ void __launch_bounds__(256) k_import_batch(const ImportJob* __restrict__ jobs) {
  const ImportJob& J = jobs[blockIdx.x];
  const long long n = J.n;
  const double* __restrict__ in = J.src;
  double* __restrict__ out = J.dst;
  bool ident = true;
  {
    long long st = 1;
    for (int a = 0; a < J.m.nd; ++a) {
      ident = ident && J.m.cstride[a] == st;
      st *= J.m.dims[a];
    }
  }
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
    long long off = i;
    if (!ident) {
      long long r = i;
      off = 0;
      for (int a = 0; a < J.m.nd; ++a) {
        int d = J.m.dims[a];
        long long q = r / d;
        off += (r - q * d) * J.m.cstride[a];
        r = q;
      }
    }
    if (C) {
      const double2 z = reinterpret_cast<const double2*>(in)[i];
      out[off] = z.x;
      out[n + off] = z.y;
    } else {
      out[off] = in[i];
    }
  }
}
template <bool C>
__global__ void k_export(const double* __restrict__ in, double* __restrict__ out, long long n, Marshal m) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i, off = 0;
    for (int a = 0; a < m.nd; ++a) {
      int d = m.dims[a];
      long long q = r / d;
      off += (r - q * d) * m.cstride[a];
      r = q;
    }
    if (C) {
      out[2 * i] = in[off];
      out[2 * i + 1] = in[n + off];
    } else {
      out[i] = in[off];
    }
  }
}

struct CopyJob {
  const double* src;
  double* dst;
  long long n;  // doubles
};
__global__ void __launch_bounds__(256) k_copy_batch(const CopyJob* __restrict__ jobs) {
  const CopyJob J = jobs[blockIdx.x];
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < J.n; i += (long long)gridDim.y * blockDim.x)
    J.dst[i] = J.src[i];
}

struct IdJob {
  double* p;
  int chi;
};
template <bool C>
__global__ void k_identity(const IdJob* __restrict__ jobs) {
  IdJob J = jobs[blockIdx.x];
  int n2 = J.chi * J.chi;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    J.p[i] = (i % J.chi == i / J.chi) ? 1.0 : 0.0;
    if (C) J.p[n2 + i] = 0.0;
  }
}

__global__ void k_sum_fixed(const double* __restrict__ x, int n, double* __restrict__ out) {
  // deterministic single-block sum
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

struct EdgePair {
  double* m1;  // planar
  double* m2;
  int n2;
};

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Z_e = sum m1[l,l'] * m2[l,l'] (no conjugate)   beliefpropagationcache.jl:115-119
template <bool C>
__global__ void k_edge_scalar(const EdgePair* __restrict__ ep, double* __restrict__ out /*2 per edge*/) {
  EdgePair E = ep[blockIdx.x];
  if (!E.m1) return;
  double sr = 0, si = 0;
  for (int i = threadIdx.x; i < E.n2; i += 32) {
    double ar = E.m1[i], br = E.m2[i];
    if (C) {
      double ai = E.m1[E.n2 + i], bi = E.m2[E.n2 + i];
      sr += ar * br - ai * bi;
      si += ar * bi + ai * br;
    } else {
      sr += ar * br;
    }
  }
  sr = wsum(sr);
  si = wsum(si);
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = sr;
    out[2 * blockIdx.x + 1] = si;
  }
}

// rescale_messages (beliefpropagationcache.jl:121-139), one warp per edge
template <bool C>
__global__ void k_rescale_msgs(const EdgePair* __restrict__ ep) {
  EdgePair E = ep[blockIdx.x];
  double s1 = 0, s2 = 0, nr = 0, ni = 0;
  for (int i = threadIdx.x; i < E.n2; i += 32) {
    double ar = E.m1[i], br = E.m2[i];
    double ai = C ? E.m1[E.n2 + i] : 0.0, bi = C ? E.m2[E.n2 + i] : 0.0;
    s1 += ar * ar + ai * ai;
    s2 += br * br + bi * bi;
    nr += ar * br - ai * bi;
    ni += ar * bi + ai * br;
  }
  s1 = wsum(s1); s2 = wsum(s2); nr = wsum(nr); ni = wsum(ni);
  const double n1 = sqrt(s1), n2 = sqrt(s2);
  nr /= (n1 * n2);
  ni /= (n1 * n2);
  double sgn = 1.0;
  if (ni == 0.0) {  // isreal(n): me[1] *= sign(n); n *= sign(n)
    sgn = (nr > 0.0) ? 1.0 : ((nr < 0.0) ? -1.0 : 0.0);
    nr *= sgn;
  }
  // sf = inv(sqrt(n)), principal branch
  double mod = sqrt(nr * nr + ni * ni);
  double sq_r, sq_i;
  if (ni == 0.0) {
    sq_r = sqrt(nr);
    sq_i = 0.0;
  } else {
    sq_r = sqrt(0.5 * (mod + nr));
    sq_i = copysign(sqrt(0.5 * (mod - nr)), ni);
  }
  double den = sq_r * sq_r + sq_i * sq_i;
  double fr = sq_r / den, fi = -sq_i / den;
  const double a1 = sgn / n1, a2 = 1.0 / n2;
  for (int i = threadIdx.x; i < E.n2; i += 32) {
    double ar = E.m1[i] * a1, br = E.m2[i] * a2;
    if (C) {
      double ai = E.m1[E.n2 + i] * a1, bi = E.m2[E.n2 + i] * a2;
      E.m1[i] = fr * ar - fi * ai;
      E.m1[E.n2 + i] = fr * ai + fi * ar;
      E.m2[i] = fr * br - fi * bi;
      E.m2[E.n2 + i] = fr * bi + fi * br;
    } else {
      E.m1[i] = fr * ar;
      E.m2[i] = fr * br;
    }
  }
}

struct ScaleJob {
  double* p;
  long long n;      // total doubles (both planes)
  const double* z;  // {re, im} of Z_v, or null
  double s;         // explicit factor when z is null
};
// A *= |Z_v|^(-1/2)   (rescale_partitions, abstractbeliefpropagationcache.jl:349-379 with k = 2)
__global__ void k_scale(const ScaleJob* __restrict__ jobs) {
  ScaleJob J = jobs[blockIdx.x];
  double f = J.s;
  if (J.z) f = 1.0 / sqrt(sqrt(J.z[0] * J.z[0] + J.z[1] * J.z[1]));
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < J.n; i += (long long)gridDim.y * blockDim.x)
    J.p[i] *= f;
}

struct ExpectJob {
  const double* rho;  // planar d x d, rho[s + d*s']
  const double* op;   // planar d x d, O[s_out + d*s_in]
  int d;
};
template <bool C>
__global__ void k_expect_finalize(const ExpectJob* __restrict__ jobs, int n, double* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  ExpectJob J = jobs[j];
  if (!J.rho) {
    out[2 * j] = out[2 * j + 1] = 0.0;
    return;
  }
  const int d = J.d, d2 = d * d;
  double nr = 0, ni = 0, tr = 0, ti = 0;
  for (int s = 0; s < d; ++s) {
    tr += J.rho[s + d * s];
    if (C) ti += J.rho[d2 + s + d * s];
    for (int sp = 0; sp < d; ++sp) {
      double orr = J.op[sp + d * s], rr = J.rho[s + d * sp];
      if (C) {
        double oi = J.op[d2 + sp + d * s], ri = J.rho[d2 + s + d * sp];
        nr += orr * rr - oi * ri;
        ni += orr * ri + oi * rr;
      } else {
        nr += orr * rr;
      }
    }
  }
  double den = tr * tr + ti * ti;
  out[2 * j] = (nr * tr + ni * ti) / den;
  out[2 * j + 1] = (ni * tr - nr * ti) / den;
}

struct Rdm2Job {
  const double* eu;  // planar (du*chi)^2 : E[(s + du*l) + Nu*(s' + du*l')]
  const double* ev;
  int du, dv, chi;
};
template <bool C>
__global__ void __launch_bounds__(256) k_rdm2_finalize(const Rdm2Job* __restrict__ jobs, double* __restrict__ out, int out_stride) {
  __shared__ double res[2 * 256];
  Rdm2Job J = jobs[blockIdx.x];
  const int du = J.du, dv = J.dv, chi = J.chi;
  const int Nu = du * chi, Nv = dv * chi, D = du * dv, nout = D * D;
  const long long pu = (long long)Nu * Nu, pv = (long long)Nv * Nv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < nout; o += nw) {
    int row = o % D, col = o / D;
    int su = row % du, sv = row / du, sup = col % du, svp = col / du;
    double sr = 0, si = 0;
    for (int t = lane; t < chi * chi; t += 32) {
      int l = t % chi, lp = t / chi;
      long long iu = (su + du * l) + (long long)Nu * (sup + du * lp);
      long long iv = (sv + dv * l) + (long long)Nv * (svp + dv * lp);
      double ar = J.eu[iu], br = J.ev[iv];
      if (C) {
        double ai = J.eu[pu + iu], bi = J.ev[pv + iv];
        sr += ar * br - ai * bi;
        si += ar * bi + ai * br;
      } else {
        sr += ar * br;
      }
    }
    sr = wsum(sr);
    si = wsum(si);
    if (lane == 0) {
      res[2 * o] = sr;
      res[2 * o + 1] = si;
    }
  }
  __syncthreads();
  double tr = 0, ti = 0;
  for (int s = 0; s < D; ++s) {
    tr += res[2 * (s + D * s)];
    ti += res[2 * (s + D * s) + 1];
  }
  double den = tr * tr + ti * ti;
  double* o_ = out + (long long)blockIdx.x * out_stride;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    double r = res[2 * o], i = res[2 * o + 1];
    o_[2 * o] = (r * tr + i * ti) / den;
    o_[2 * o + 1] = (i * tr - r * ti) / den;
  }
}

struct Apply1Job {
  double* t;        // planar tensor, site index fastest
  long long n;
  const double* g;  // planar d x d gate, g[s' + d*s]
  int d;
  int normalize;
};
template <bool C>
__global__ void __launch_bounds__(256) k_apply1(const Apply1Job* __restrict__ jobs, double* __restrict__ sumsq) {
  Apply1Job J = jobs[blockIdx.x];
  const int d = J.d, d2 = d * d;
  const long long cols = J.n / d;
  double ss = 0.0;
  for (long long c = (long long)blockIdx.y * blockDim.x + threadIdx.x; c < cols; c += (long long)gridDim.y * blockDim.x) {
    double xr[8], xi[8];
    for (int s = 0; s < d; ++s) {
      xr[s] = J.t[c * d + s];
      xi[s] = C ? J.t[J.n + c * d + s] : 0.0;
    }
    for (int sp = 0; sp < d; ++sp) {
      double yr = 0, yi = 0;
      for (int s = 0; s < d; ++s) {
        double gr = J.g[sp + d * s];
        double gi = C ? J.g[d2 + sp + d * s] : 0.0;
        yr += gr * xr[s] - gi * xi[s];
        yi += gr * xi[s] + gi * xr[s];
      }
      J.t[c * d + sp] = yr;
      if (C) J.t[J.n + c * d + sp] = yi;
      ss += yr * yr + yi * yi;
    }
  }
  if (sumsq) {
    ss = wsum(ss);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sumsq[blockIdx.x], ss);
  }
}

struct NormJob {
  double* p;
  long long n;  // total doubles
};
// deterministic per-tensor Frobenius norm then scale (one block per tensor)
__global__ void __launch_bounds__(256) k_normalize(const NormJob* __restrict__ jobs) {
  __shared__ double sh[8];
  __shared__ double tot;
  NormJob J = jobs[blockIdx.x];
  double s = 0.0;
  for (long long i = threadIdx.x; i < J.n; i += blockDim.x) s += J.p[i] * J.p[i];
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    tot = t;
  }
  __syncthreads();
  double f = 1.0 / sqrt(tot);
  for (long long i = threadIdx.x; i < J.n; i += blockDim.x) J.p[i] *= f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int itn_ctx_create(int device, void* stream, itn_ctx** out) {
  API_BEGIN
  ITN_REQUIRE(out != nullptr, ITN_EINVAL, "out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw ItnError(ITN_ECUDA, "no CUDA device available: libitn_b200 has no CPU fallback");
  }
  ITN_REQUIRE(device >= 0 && device < ndev, ITN_EINVAL, "device index out of range");
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  ITN_REQUIRE(prop.major >= 10, ITN_ECUDA,
              std::string("libitn_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                  std::to_string(prop.minor));
  itn_ctx* c = new itn_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  size_t freeb = 0, totalb = 0;
  CUDA_CHECK(cudaMemGetInfo(&freeb, &totalb));
  c->ws_budget = std::max<size_t>((size_t)1 << 30, freeb / 8);
  *out = c;
  API_END
}

static void ctx_destroy_now(itn_ctx* ctx);

extern "C" int itn_ctx_destroy(itn_ctx* ctx) {
  API_BEGIN
  if (!ctx) return ITN_OK;
  if (ctx->nets_alive > 0) {
    // finalizers run in any order (Julia GC, Python cycles): the context goes away with its last network
    ctx->destroy_pending = true;
    return ITN_OK;
  }
  ctx_destroy_now(ctx);
  API_END
}

static void ctx_destroy_now(itn_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->nccl && ctx->nccl_lib) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(ctx->nccl_lib, "ncclCommDestroy");
    if (f) f(ctx->nccl);
  }
  big_trim(ctx, 0);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->pinned_flags) cudaFreeHost(ctx->pinned_flags);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int itn_ctx_sync(itn_ctx* ctx) {
  API_BEGIN
  ITN_REQUIRE(ctx, ITN_EINVAL, "ctx is NULL");
  set_device(ctx);
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

extern "C" int itn_ctx_launch_count(const itn_ctx* ctx, int64_t* out) {
  API_BEGIN
  ITN_REQUIRE(ctx && out, ITN_EINVAL, "NULL argument");
  *out = ctx->launches;
  API_END
}

extern "C" int itn_ctx_set_path(itn_ctx* ctx, int mode) {
  API_BEGIN
  ITN_REQUIRE(ctx, ITN_EINVAL, "ctx is NULL");
  ITN_REQUIRE(mode >= 0 && mode <= 2, ITN_EINVAL, "mode must be 0 (auto), 1 (shape-generic DMMA kernels only) or 2 (FMA kernels only)");
  ctx->path_mode = mode;
  API_END
}

// ------------------------------------------------------------------------------------------------
// network
// ------------------------------------------------------------------------------------------------
// a rank stores the messages on every edge that touches one of its vertices
static bool msg_stored(const itn_net* net, int did) {
  const int e = did / 2;
  return itn_is_local(net, net->esrc[e]) || itn_is_local(net, net->edst[e]);
}

static void alloc_message(itn_net* net, int did) {
  int e = did / 2;
  long long n2 = (long long)net->edim[e] * net->edim[e];
  if (net->M[did].p && net->M[did].n == n2) return;
  if (net->M[did].p) itn_tensor_free(net->ctx, net->M[did]);
  net->M[did].p = (double*)itn_dev_alloc(net->ctx, (size_t)n2 * net->planes() * sizeof(double));
  net->M[did].n = n2;
}

// One shared, reference-counted allocation for a batch of tensors (a network has thousands of site tensors and tens
// of thousands of messages: one cudaMallocAsync instead of one per object).  sizes in elements per plane.
static void slab_alloc(itn_net* net, const std::vector<DevTensor*>& ts, const std::vector<long long>& elems) {
  if (ts.empty()) return;
  const int P = net->planes();
  auto align32 = [](size_t x) { return (x + 31) & ~(size_t)31; };
  size_t tot = 0;
  for (long long n : elems) tot += align32((size_t)n * P);
  DevSlab* slab = new DevSlab();
  try {
    slab->base = itn_dev_alloc(net->ctx, tot * sizeof(double));
  } catch (...) {
    delete slab;
    throw;
  }
  size_t off = 0;
  for (size_t i = 0; i < ts.size(); ++i) {
    if (ts[i]->p) itn_tensor_free(net->ctx, *ts[i]);
    ts[i]->p = (double*)slab->base + off;
    ts[i]->n = elems[i];
    ts[i]->slab = slab;
    slab->refs++;
    off += align32((size_t)elems[i] * P);
  }
}

static void set_identity_messages(itn_net* net, const std::vector<int>& dids) {
  if (dids.empty()) return;
  std::vector<DevTensor*> need;
  std::vector<long long> need_n;
  for (int did : dids) {
    if (!msg_stored(net, did)) continue;
    const long long n2 = (long long)net->edim[did / 2] * net->edim[did / 2];
    if (net->M[did].p && net->M[did].n == n2) continue;
    need.push_back(&net->M[did]);
    need_n.push_back(n2);
  }
  slab_alloc(net, need, need_n);
  std::vector<IdJob> jobs;
  for (int did : dids) {
    if (!msg_stored(net, did)) continue;
    jobs.push_back({net->M[did].p, net->edim[did / 2]});
  }
  if (jobs.empty()) return;
  DevBuf b(net->ctx, jobs.size() * sizeof(IdJob));
  const IdJob* d = itn_upload(net->ctx, jobs, b);
  if (net->cplx) k_identity<true><<<(unsigned)jobs.size(), 128, 0, net->ctx->stream>>>(d);
  else k_identity<false><<<(unsigned)jobs.size(), 128, 0, net->ctx->stream>>>(d);
  ITN_LAUNCH_CHECK(net->ctx);
}

extern "C" int itn_net_create(itn_ctx* ctx, int dtype, int nv, int ne, const int32_t* esrc, const int32_t* edst,
                              const int32_t* edim, const int32_t* sdim, const int32_t* owner, itn_net** out) {
  API_BEGIN
  ITN_REQUIRE(ctx && out, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(dtype == ITN_F64 || dtype == ITN_C128, ITN_EUNSUPPORTED, "dtype must be 0 (Float64) or 1 (ComplexF64)");
  ITN_REQUIRE(nv > 0 && ne >= 0, ITN_EINVAL, "nv must be > 0 and ne >= 0");
  ITN_REQUIRE(sdim && (ne == 0 || (esrc && edst && edim)), ITN_EINVAL, "NULL graph arrays");
  set_device(ctx);
  std::unique_ptr<itn_net> net(new itn_net());
  net->ctx = ctx;
  net->dtype = dtype;
  net->cplx = dtype == ITN_C128;
  net->nv = nv;
  net->ne = ne;
  net->esrc.assign(esrc, esrc + ne);
  net->edst.assign(edst, edst + ne);
  net->edim.assign(edim, edim + ne);
  net->sdim.assign(sdim, sdim + nv);
  net->owner.assign(nv, 0);
  if (owner) net->owner.assign(owner, owner + nv);
  for (int v = 0; v < nv; ++v)
    ITN_REQUIRE(net->owner[v] >= 0 && net->owner[v] < ctx->nranks, ITN_EINVAL,
                "owner[v] must be a rank of the context (call itn_ctx_init_dist before itn_net_create)");
  net->inc.assign(nv, {});
  // site dimensions above 8 are super-sites of multi-site partitions (fused site indices): BP, scalars and expect only
  for (int v = 0; v < nv; ++v) ITN_REQUIRE(sdim[v] >= 1 && sdim[v] <= 64, ITN_EUNSUPPORTED, "site dimension must be in 1..64");
  for (int e = 0; e < ne; ++e) {
    int s = esrc[e], d = edst[e];
    ITN_REQUIRE(s >= 0 && s < nv && d >= 0 && d < nv && s != d, ITN_EINVAL, "edge endpoint out of range or self loop");
    ITN_REQUIRE(edim[e] >= 1, ITN_ESHAPE, "bond dimension must be >= 1");
    ITN_REQUIRE(net->did(s, d) < 0 && net->did(d, s) < 0, ITN_EUNSUPPORTED, "multi-edges are not supported");
    net->dmap[((uint64_t)(uint32_t)s << 32) | (uint32_t)d] = 2 * e;
    net->dmap[((uint64_t)(uint32_t)d << 32) | (uint32_t)s] = 2 * e + 1;
    net->inc[s].push_back(e);
    net->inc[d].push_back(e);
  }
  for (int v = 0; v < nv; ++v)
    ITN_REQUIRE((int)net->inc[v].size() + 1 <= ITN_MAX_MODES, ITN_EUNSUPPORTED, "vertex degree above 9 is not supported");
  net->T.assign(nv, DevTensor());
  net->M.assign(2 * (size_t)ne, DevTensor());
  net->tver.assign(nv, 0);
  ctx->nets_alive++;
  *out = net.release();
  API_END
}

static void free_net_storage(itn_net* net) {
  for (auto& t : net->T)
    if (t.p) itn_tensor_free(net->ctx, t);
  for (auto& t : net->Tb)
    if (t.p) itn_tensor_free(net->ctx, t);
  net->nbra = 0;
  for (auto& m : net->M)
    if (m.p) itn_tensor_free(net->ctx, m);
  itn_fast_release(net);
  itn_block_release(net);
  itn_dist_release(net);
}

extern "C" int itn_net_destroy(itn_net* net) {
  API_BEGIN
  if (!net) return ITN_OK;
  cudaSetDevice(net->ctx->device);
  net->pending.clear();
  free_net_storage(net);
  itn_ctx* ctx = net->ctx;
  delete net;
  if (--ctx->nets_alive == 0 && ctx->destroy_pending) ctx_destroy_now(ctx);
  API_END
}

extern "C" int itn_net_clone(const itn_net* src, itn_net** out) {
  API_BEGIN
  ITN_REQUIRE(src && out, ITN_EINVAL, "NULL argument");
  set_device(src->ctx);
  itn_flush_pending(const_cast<itn_net*>(src));
  std::unique_ptr<itn_net> net(new itn_net(*src));
  net->fast = nullptr;
  net->block = nullptr;
  net->dist = nullptr;
  for (auto& t : net->T) t.p = nullptr, t.slab = nullptr;
  for (auto& t : net->Tb) t.p = nullptr, t.slab = nullptr;
  for (auto& m : net->M) m.p = nullptr, m.slab = nullptr;
  // two shared allocations (site tensors, messages) and one batched copy kernel instead of ~20k cudaMallocAsync + memcpy
  const int P = src->planes();
  std::vector<DevTensor*> ts, ms;
  std::vector<long long> tn, mn;
  std::vector<CopyJob> cj;
  for (int v = 0; v < src->nv; ++v)
    if (src->T[v].p) {
      ts.push_back(&net->T[v]);
      tn.push_back(src->T[v].n);
    }
  for (size_t d = 0; d < src->M.size(); ++d)
    if (src->M[d].p) {
      ms.push_back(&net->M[d]);
      mn.push_back(src->M[d].n);
    }
  for (size_t v = 0; v < src->Tb.size(); ++v)
    if (src->Tb[v].p) {
      ts.push_back(&net->Tb[v]);
      tn.push_back(src->Tb[v].n);
    }
  slab_alloc(net.get(), ts, tn);
  slab_alloc(net.get(), ms, mn);
  long long maxn = 0;
  for (size_t v = 0; v < src->Tb.size(); ++v)
    if (src->Tb[v].p) {
      cj.push_back({src->Tb[v].p, net->Tb[v].p, src->Tb[v].n * P});
      maxn = std::max<long long>(maxn, src->Tb[v].n * P);
    }
  for (int v = 0; v < src->nv; ++v)
    if (src->T[v].p) {
      cj.push_back({src->T[v].p, net->T[v].p, src->T[v].n * P});
      maxn = std::max<long long>(maxn, src->T[v].n * P);
    }
  for (size_t d = 0; d < src->M.size(); ++d)
    if (src->M[d].p) {
      cj.push_back({src->M[d].p, net->M[d].p, src->M[d].n * P});
      maxn = std::max<long long>(maxn, src->M[d].n * P);
    }
  if (!cj.empty()) {
    DevBuf jb(src->ctx, cj.size() * sizeof(CopyJob));
    const CopyJob* dj = itn_upload(src->ctx, cj, jb);
    unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn + 4095) / 4096, 64));
    k_copy_batch<<<dim3((unsigned)cj.size(), gy), 256, 0, src->ctx->stream>>>(dj);
    ITN_LAUNCH_CHECK(src->ctx);
  }
  src->ctx->nets_alive++;
  *out = net.release();
  API_END
}

extern "C" int itn_sync(itn_net* net) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  set_device(net->ctx);
  itn_flush_pending(net);
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  API_END
}

extern "C" int itn_net_edge_dim(const itn_net* net, int e, int32_t* out) {
  API_BEGIN
  ITN_REQUIRE(net && out, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(e >= 0 && e < net->ne, ITN_EINVAL, "edge id out of range");
  *out = net->edim[e];
  API_END
}

extern "C" int itn_net_tensor_size(const itn_net* net, int v, int64_t* out) {
  API_BEGIN
  ITN_REQUIRE(net && out, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
  *out = net->tensor_elems(v);
  API_END
}

static Marshal make_marshal(const itn_net* net, int v, int nd, const int32_t* axis_edge) {
  const int z = (int)net->inc[v].size();
  ITN_REQUIRE(nd == z + 1, ITN_ESHAPE,
              "tensor of vertex " + std::to_string(v) + " must have " + std::to_string(z + 1) + " axes");
  // canonical strides
  std::vector<long long> cs(z + 1);
  long long s = 1;
  cs[0] = 1;
  s = net->sdim[v];
  for (int k = 0; k < z; ++k) {
    cs[k + 1] = s;
    s *= net->edim[net->inc[v][k]];
  }
  Marshal m;
  m.nd = nd;
  std::vector<bool> seen(z + 1, false);
  for (int a = 0; a < nd; ++a) {
    int mode;
    if (!axis_edge) {
      mode = a;
    } else if (axis_edge[a] < 0) {
      mode = 0;
    } else {
      int k = net->slot(v, axis_edge[a]);
      ITN_REQUIRE(k >= 0, ITN_EINVAL, "axis_edge names an edge that is not incident to the vertex");
      mode = k + 1;
    }
    ITN_REQUIRE(!seen[mode], ITN_EINVAL, "axis_edge repeats an axis");
    seen[mode] = true;
    m.dims[a] = mode == 0 ? net->sdim[v] : net->edim[net->inc[v][mode - 1]];
    m.cstride[a] = cs[mode];
  }
  return m;
}

extern "C" int itn_net_set_tensor(itn_net* net, int v, const void* host, int nd, const int32_t* axis_edge) {
  API_BEGIN
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
  if (!itn_is_local(net, v)) return ITN_OK;  // another rank stores this vertex
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  Marshal m = make_marshal(net, v, nd, axis_edge);
  for (size_t i = 0; i < net->pending.size(); ++i)
    if (net->pending[i].v == v) {
      net->pending.erase(net->pending.begin() + i);
      break;
    }
  const long long n = net->tensor_elems(v);
  const int P = net->planes();
  if (!net->T[v].p || net->T[v].n != n) {
    if (net->T[v].p) itn_tensor_free(ctx, net->T[v]);
    net->T[v].p = (double*)itn_dev_alloc(ctx, (size_t)n * P * sizeof(double));
    net->T[v].n = n;
  }
  net->touch(v);
  DevBuf stage(ctx, (size_t)n * P * sizeof(double));
  CUDA_CHECK(cudaMemcpyAsync(stage.p, host, (size_t)n * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  unsigned g = (unsigned)std::min<long long>((n + 255) / 256, 4096);
  if (net->cplx) k_import<true><<<g, 256, 0, ctx->stream>>>(stage.as<double>(), net->T[v].p, n, m);
  else k_import<false><<<g, 256, 0, ctx->stream>>>(stage.as<double>(), net->T[v].p, n, m);
  ITN_LAUNCH_CHECK(ctx);
  API_END
}

// Bra layer of a bilinear form <phi|psi> (BilinearFormNetwork, src/formnetworks/bilinearformnetwork.jl:23-42; built by
// inner_network, src/inner.jl:139-171).  The quadratic form keeps bra = conj(ket) implicit; here phi_v is stored as given
// and conjugated by the closing kernels exactly where they conjugate the ket.  Same extents as the ket (a host that
// holds different bond dimensions zero-pads the smaller tensor, which changes no contraction).
extern "C" int itn_net_set_bra_tensor(itn_net* net, int v, const void* host, int nd, const int32_t* axis_edge) {
  API_BEGIN
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
  if (!itn_is_local(net, v)) return ITN_OK;
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  Marshal m = make_marshal(net, v, nd, axis_edge);
  const long long n = net->tensor_elems(v);
  const int P = net->planes();
  if (net->Tb.size() != (size_t)net->nv) net->Tb.resize(net->nv);
  DevTensor& t = net->Tb[v];
  if (!t.p || t.n != n) {
    if (t.p) itn_tensor_free(ctx, t);
    else net->nbra++;
    t.p = (double*)itn_dev_alloc(ctx, (size_t)n * P * sizeof(double));
    t.n = n;
  }
  net->touch(v);
  DevBuf stage(ctx, (size_t)n * P * sizeof(double));
  CUDA_CHECK(cudaMemcpyAsync(stage.p, host, (size_t)n * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  unsigned g = (unsigned)std::min<long long>((n + 255) / 256, 4096);
  if (net->cplx) k_import<true><<<g, 256, 0, ctx->stream>>>(stage.as<double>(), t.p, n, m);
  else k_import<false><<<g, 256, 0, ctx->stream>>>(stage.as<double>(), t.p, n, m);
  ITN_LAUNCH_CHECK(ctx);
  API_END
}

extern "C" int itn_net_clear_bra(itn_net* net) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  set_device(net->ctx);
  for (size_t v = 0; v < net->Tb.size(); ++v)
    if (net->Tb[v].p) {
      itn_tensor_free(net->ctx, net->Tb[v]);
      net->touch((int)v);
    }
  net->nbra = 0;
  API_END
}

// ------------------------------------------------------------------------------------------------
// pipelined host -> device upload of site tensors
// ------------------------------------------------------------------------------------------------
namespace {

cudaStream_t copy_stream(itn_ctx* ctx) {
  if (!ctx->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  return ctx->copy_stream;
}

// Streams items[cuts[c] .. cuts[c+1]) chunk by chunk: the copy stream fills one of two staging slots while the main
// stream imports (de-interleave + axis permutation) the other.  after_chunk(c) runs once the import of chunk c
// has been enqueued on the main stream, so whatever it launches there overlaps the copy of chunk c + 1.
// Host buffers are free again when the copy stream has drained (the caller synchronises it).
template <class F>
void upload_pipelined(itn_net* net, const std::vector<PendingUpload>& items, const std::vector<size_t>& cuts, F after_chunk) {
  itn_ctx* ctx = net->ctx;
  if (items.empty()) return;
  const int P = net->planes();
  cudaStream_t cs = copy_stream(ctx);
  std::vector<ImportJob> jobs(items.size());
  size_t slot_bytes = 0;
  long long maxn = 0;
  for (size_t c = 0; c + 1 < cuts.size(); ++c) {
    size_t off = 0;
    for (size_t i = cuts[c]; i < cuts[c + 1]; ++i) {
      const long long n = net->T[items[i].v].n;
      jobs[i].src = (const double*)off;  // byte offset inside the slot for now
      jobs[i].dst = net->T[items[i].v].p;
      jobs[i].n = n;
      jobs[i].m = items[i].m;
      off += (size_t)n * P * sizeof(double);
      maxn = std::max(maxn, n);
    }
    slot_bytes = std::max(slot_bytes, off);
  }
  DevBuf stage(ctx, 2 * slot_bytes), jb(ctx, jobs.size() * sizeof(ImportJob));
  for (size_t c = 0; c + 1 < cuts.size(); ++c)
    for (size_t i = cuts[c]; i < cuts[c + 1]; ++i)
      jobs[i].src = (const double*)(stage.as<char>() + (c & 1) * slot_bytes + (size_t)jobs[i].src);
  const ImportJob* dj = itn_upload(ctx, jobs, jb);
  cudaEvent_t ready, copied[2], imported[2];
  CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  for (int k = 0; k < 2; ++k) {
    CUDA_CHECK(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&imported[k], cudaEventDisableTiming));
  }
  auto cleanup = [&]() {
    cudaEventDestroy(ready);
    for (int k = 0; k < 2; ++k) {
      cudaEventDestroy(copied[k]);
      cudaEventDestroy(imported[k]);
    }
  };
  try {
    // the staging buffer and the destination tensors were allocated in main-stream order
    CUDA_CHECK(cudaEventRecord(ready, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(cs, ready, 0));
    unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn + 2047) / 2048, 64));
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
      const int k = (int)(c & 1);
      const size_t lo = cuts[c], hi = cuts[c + 1];
      if (hi == lo) {
        after_chunk(c);
        continue;
      }
      if (c >= 2) CUDA_CHECK(cudaStreamWaitEvent(cs, imported[k], 0));
      // tensors that follow each other in host memory (one big pinned buffer in vertex order is the common case) land
      // next to each other in the slot too: one copy per contiguous run instead of one per tensor
      for (size_t i = lo; i < hi;) {
        size_t bytes = (size_t)jobs[i].n * P * sizeof(double), j = i + 1;
        while (j < hi && (const char*)items[j].host == (const char*)items[i].host + bytes &&
               (const char*)jobs[j].src == (const char*)jobs[i].src + bytes)
          bytes += (size_t)jobs[j++].n * P * sizeof(double);
        CUDA_CHECK(cudaMemcpyAsync((void*)jobs[i].src, items[i].host, bytes, cudaMemcpyHostToDevice, cs));
        i = j;
      }
      CUDA_CHECK(cudaEventRecord(copied[k], cs));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, copied[k], 0));
      dim3 grid((unsigned)(hi - lo), gy);
      if (net->cplx) k_import_batch<true><<<grid, 256, 0, ctx->stream>>>(dj + lo);
      else k_import_batch<false><<<grid, 256, 0, ctx->stream>>>(dj + lo);
      ITN_LAUNCH_CHECK(ctx);
      CUDA_CHECK(cudaEventRecord(imported[k], ctx->stream));
      after_chunk(c);
    }
    CUDA_CHECK(cudaStreamSynchronize(cs));  // host buffers consumed
  } catch (...) {
    cudaStreamSynchronize(cs);
    cleanup();
    throw;
  }
  cleanup();
}

// chunk boundaries of roughly `target` bytes
std::vector<size_t> cuts_by_bytes(const itn_net* net, const std::vector<PendingUpload>& items, size_t lo, size_t hi,
                                  size_t target) {
  std::vector<size_t> cuts{lo};
  size_t acc = 0;
  for (size_t i = lo; i < hi; ++i) {
    acc += (size_t)net->T[items[i].v].n * net->planes() * sizeof(double);
    if (acc >= target) {
      cuts.push_back(i + 1);
      acc = 0;
    }
  }
  if (cuts.back() != hi) cuts.push_back(hi);
  return cuts;
}

constexpr size_t kUploadChunkBytes = (size_t)256 << 20;

}  // namespace

// Copies every registered-but-deferred host tensor to the device (all consumers other than the pipelined first BP sweep).
void itn_flush_pending(itn_net* net) {
  itn_canon_ensure_all(net);
  itn_flush_uploads(net);
}

void itn_flush_uploads(itn_net* net) {
  if (net->pending.empty()) return;
  std::vector<PendingUpload> items;
  items.swap(net->pending);
  upload_pipelined(net, items, cuts_by_bytes(net, items, 0, items.size(), kUploadChunkBytes), [](size_t) {});
  for (const PendingUpload& pu : items) net->touch(pu.v);  // tile-major copies planned while the tensors were pending are stale
}

/* BeliefPropagationCache(ptn) data path: every site tensor in one call (pipelined copy + import). */
extern "C" int itn_net_set_tensors(itn_net* net, int n, const int32_t* verts, const void* const* hosts, const int32_t* nd,
                                   const int32_t* axis_edge, int flags) {
  API_BEGIN
  ITN_REQUIRE(net && n >= 0 && (n == 0 || (verts && hosts)), ITN_EINVAL, "NULL argument");
  ITN_REQUIRE((flags & ~ITN_HOST_DEFERRED) == 0, ITN_EINVAL, "unknown flag");
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  std::vector<char> seen(net->nv, 0);
  size_t aoff = 0;
  std::vector<PendingUpload> items;
  for (int i = 0; i < n; ++i) {
    const int v = verts[i];
    ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
    ITN_REQUIRE(!seen[v], ITN_EINVAL, "vertex listed twice");
    ITN_REQUIRE(hosts[i] != nullptr, ITN_EINVAL, "NULL host tensor");
    seen[v] = 1;
    const int ndv = nd ? nd[i] : (int)net->inc[v].size() + 1;
    const int32_t* ax = axis_edge ? axis_edge + aoff : nullptr;
    aoff += ndv;
    if (!itn_is_local(net, v)) continue;  // another rank stores this vertex
    PendingUpload pu;
    pu.v = v;
    pu.host = hosts[i];
    pu.m = make_marshal(net, v, ndv, ax);
    items.push_back(pu);
  }
  // a newer registration replaces an older pending one
  if (!net->pending.empty()) {
    std::vector<PendingUpload> keep;
    for (const PendingUpload& pu : net->pending)
      if (!seen[pu.v]) keep.push_back(pu);
    net->pending.swap(keep);
  }
  {
    std::vector<DevTensor*> need;
    std::vector<long long> need_n;
    for (const PendingUpload& pu : items) {
      const int v = pu.v;
      const long long nel = net->tensor_elems(v);
      if (!net->T[v].p || net->T[v].n != nel) {
        need.push_back(&net->T[v]);
        need_n.push_back(nel);
      }
      net->touch(v);
    }
    slab_alloc(net, need, need_n);
  }
  net->topo_version++;
  if (flags & ITN_HOST_DEFERRED) {
    net->pending.insert(net->pending.end(), items.begin(), items.end());
  } else {
    upload_pipelined(net, items, cuts_by_bytes(net, items, 0, items.size(), kUploadChunkBytes), [](size_t) {});
  }
  API_END
}

extern "C" int itn_net_get_tensor(const itn_net* net_, int v, void* host, int nd, const int32_t* axis_edge) {
  API_BEGIN
  itn_net* net = const_cast<itn_net*>(net_);
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
  ITN_REQUIRE(net->T[v].p, ITN_EINVAL, "site tensor is not set");
  set_device(net->ctx);
  itn_flush_pending(net);
  itn_ctx* ctx = net->ctx;
  Marshal m = make_marshal(net, v, nd, axis_edge);
  const long long n = net->T[v].n;
  const int P = net->planes();
  DevBuf stage(ctx, (size_t)n * P * sizeof(double));
  unsigned g = (unsigned)std::min<long long>((n + 255) / 256, 4096);
  if (net->cplx) k_export<true><<<g, 256, 0, ctx->stream>>>(net->T[v].p, stage.as<double>(), n, m);
  else k_export<false><<<g, 256, 0, ctx->stream>>>(net->T[v].p, stage.as<double>(), n, m);
  ITN_LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(host, stage.p, (size_t)n * P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

extern "C" int itn_msg_set_identity(itn_net* net) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  set_device(net->ctx);
  std::vector<int> dids(2 * (size_t)net->ne);
  std::iota(dids.begin(), dids.end(), 0);
  set_identity_messages(net, dids);
  API_END
}

static Marshal msg_marshal(int chi) {
  Marshal m;
  m.nd = 2;
  m.dims[0] = m.dims[1] = chi;
  m.cstride[0] = 1;
  m.cstride[1] = chi;
  return m;
}

extern "C" int itn_msg_set(itn_net* net, int src, int dst, const void* host) {
  API_BEGIN
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  int did = net->did(src, dst);
  ITN_REQUIRE(did >= 0, ITN_EINVAL, "(src, dst) is not an edge of the network");
  if (!msg_stored(net, did)) return ITN_OK;  // neither endpoint lives on this rank
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  alloc_message(net, did);
  const long long n2 = net->M[did].n;
  const int P = net->planes();
  DevBuf stage(ctx, (size_t)n2 * P * sizeof(double));
  CUDA_CHECK(cudaMemcpyAsync(stage.p, host, (size_t)n2 * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  Marshal m = msg_marshal(net->edim[did / 2]);
  if (net->cplx) k_import<true><<<1, 256, 0, ctx->stream>>>(stage.as<double>(), net->M[did].p, n2, m);
  else k_import<false><<<1, 256, 0, ctx->stream>>>(stage.as<double>(), net->M[did].p, n2, m);
  ITN_LAUNCH_CHECK(ctx);
  API_END
}

static void download_planar(itn_net* net, const double* dev, long long n, void* host) {
  // planar device -> interleaved host
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  std::vector<double> tmp((size_t)n * P);
  CUDA_CHECK(cudaMemcpyAsync(tmp.data(), dev, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  double* h = (double*)host;
  if (net->cplx) {
    for (long long i = 0; i < n; ++i) {
      h[2 * i] = tmp[i];
      h[2 * i + 1] = tmp[n + i];
    }
  } else {
    memcpy(h, tmp.data(), (size_t)n * sizeof(double));
  }
}

extern "C" int itn_msg_get(const itn_net* net_, int src, int dst, void* host) {
  API_BEGIN
  itn_net* net = const_cast<itn_net*>(net_);
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  int did = net->did(src, dst);
  ITN_REQUIRE(did >= 0, ITN_EINVAL, "(src, dst) is not an edge of the network");
  ITN_REQUIRE(net->M[did].p, ITN_EINVAL, "message is not set");
  set_device(net->ctx);
  download_planar(net, net->M[did].p, net->M[did].n, host);
  API_END
}

namespace {
struct PackJob {
  const double* src;  // planar
  long long n;        // elements
  long long off;      // element offset in the packed output
};
// planar device arrays -> one packed buffer; interleave = 1: (re, im) pairs (host layout), 0: planar
template <bool C>
__global__ void k_pack(const PackJob* __restrict__ jobs, double* __restrict__ out, int interleave) {
  PackJob J = jobs[blockIdx.x];
  const int P = C ? 2 : 1;
  for (long long i = threadIdx.x; i < J.n; i += blockDim.x) {
    if (C) {
      if (interleave) {
        out[2 * (J.off + i)] = J.src[i];
        out[2 * (J.off + i) + 1] = J.src[J.n + i];
      } else {
        out[P * J.off + i] = J.src[i];
        out[P * J.off + J.n + i] = J.src[J.n + i];
      }
    } else {
      out[J.off + i] = J.src[i];
    }
  }
}
}  // namespace

extern "C" int itn_msg_get_all(const itn_net* net_, void* host, int64_t bytes) {
  API_BEGIN
  itn_net* net = const_cast<itn_net*>(net_);
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  std::vector<PackJob> jobs;
  long long off = 0;
  for (size_t d = 0; d < net->M.size(); ++d)
    if (net->M[d].p) {
      jobs.push_back({net->M[d].p, net->M[d].n, off});
      off += net->M[d].n;
    }
  const size_t need = (size_t)off * net->planes() * sizeof(double);
  ITN_REQUIRE((size_t)bytes >= need, ITN_ESHAPE, "host buffer too small for all messages");
  if (jobs.empty()) return ITN_OK;
  DevBuf jb(ctx, jobs.size() * sizeof(PackJob)), stage(ctx, need);
  const PackJob* dj = itn_upload(ctx, jobs, jb);
  if (net->cplx) k_pack<true><<<(unsigned)jobs.size(), 128, 0, ctx->stream>>>(dj, stage.as<double>(), 1);
  else k_pack<false><<<(unsigned)jobs.size(), 128, 0, ctx->stream>>>(dj, stage.as<double>(), 1);
  ITN_LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(host, stage.p, need, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

// ------------------------------------------------------------------------------------------------
// belief propagation driver
// ------------------------------------------------------------------------------------------------
namespace {

struct MsgJob {
  int did, v, k;  // directed id, source vertex, slot of the edge at v
};

std::vector<MsgJob> make_msg_jobs(itn_net* net, const int32_t* src, const int32_t* dst, int n) {
  std::vector<MsgJob> jobs(n);
  for (int i = 0; i < n; ++i) {
    int did = net->did(src[i], dst[i]);
    ITN_REQUIRE(did >= 0, ITN_EINVAL,
                "sequence entry " + std::to_string(i) + " (" + std::to_string(src[i]) + " -> " +
                    std::to_string(dst[i]) + ") is not an edge of the network");
    jobs[i] = {did, src[i], net->slot(src[i], did / 2)};
  }
  return jobs;
}

// staged outputs for a list of message jobs
struct Staged {
  DevBuf buf;
  std::vector<double*> ptr;
  Staged(itn_net* net, const std::vector<MsgJob>& jobs) : buf(net->ctx, total(net, jobs)) {
    size_t off = 0;
    ptr.resize(jobs.size());
    for (size_t i = 0; i < jobs.size(); ++i) {
      ptr[i] = (double*)((char*)buf.p + off);
      long long chi = net->edim[jobs[i].did / 2];
      off += (size_t)chi * chi * net->planes() * sizeof(double);
    }
  }
  static size_t total(itn_net* net, const std::vector<MsgJob>& jobs) {
    size_t t = 0;
    for (auto& j : jobs) {
      long long chi = net->edim[j.did / 2];
      t += (size_t)chi * chi * net->planes() * sizeof(double);
    }
    return t;
  }
};

void compute_messages(itn_net* net, const std::vector<MsgJob>& jobs, size_t lo, size_t hi,
                      const std::vector<double*>& staged) {
  // fast path first (synchronous full-bucket sweeps), generic otherwise
  std::vector<JobSpec> specs;
  specs.reserve(hi - lo);
  for (size_t i = lo; i < hi; ++i) specs.push_back({jobs[i].v, 1u << (jobs[i].k + 1), staged[i]});
  itn_run_vertex_jobs(net, specs);
}

// update_iteration(alg, bpc, edge_groups) with groups of several edges (abstractbeliefpropagationcache.jl:294-308,
// intended semantics): every group is a sequential (Gauss-Seidel) pass over its edges that starts from the PRE-SWEEP
// messages; the message of every edge of the group is collected, later groups override earlier ones, and everything is
// written back at the end of the sweep.  update() divides the summed diffs by the number of groups (:319-321).
// Every update gets its own result buffer, so inside a group only read-after-write dependencies order the work
// (dependency levels, one batched launch sequence per level).  Parity mode: per-message kernels, single GPU.
void bp_update_groups(itn_net* net, const std::vector<MsgJob>& jobs, const int32_t* group_ptr, int ngroups, int maxiter,
                      double tol, int normalize, int32_t* iters, double* last_mean_diff) {
  itn_ctx* ctx = net->ctx;
  ITN_REQUIRE(ctx->nranks == 1, ITN_EUNSUPPORTED,
              "groups of several edges run on a single GPU only (use single-edge groups on a partitioned network)");
  itn_flush_pending(net);
  const int nseq = (int)jobs.size();
  const bool want_diff = tol >= 0.0;
  const size_t nd = net->M.size();
  // plan: per job the buffer each incoming message is read from (-1: the pre-sweep message) and its level in the group
  std::vector<std::vector<int>> reads(nseq);
  std::vector<int> oldsrc(nseq, -1), level(nseq, 0);
  std::vector<int> final_job(nd, -1);
  {
    std::vector<int> overlay(nd, -1);
    for (int g = 0; g < ngroups; ++g) {
      std::vector<int> touched;
      for (int i = group_ptr[g]; i < group_ptr[g + 1]; ++i) {
        const MsgJob& J = jobs[i];
        int lv = 0;
        const int z = (int)net->inc[J.v].size();
        reads[i].assign(z, -1);
        for (int k = 0; k < z; ++k) {
          const int e = net->inc[J.v][k];
          if (e == J.did / 2) continue;
          const int m = net->msg_into(J.v, e);
          ITN_REQUIRE(overlay[m] >= 0 || net->M[m].p != nullptr, ITN_EINVAL,
                      "message into vertex " + std::to_string(J.v) + " on edge " + std::to_string(e) +
                          " does not exist when updating " + std::to_string(J.v) + " -> " +
                          std::to_string(net->other(J.did / 2, J.v)) + " (initialise messages or use the forest-cover sequence)");
          reads[i][k] = overlay[m];
          if (overlay[m] >= 0) lv = std::max(lv, level[overlay[m]] + 1);
        }
        if (want_diff) {
          ITN_REQUIRE(overlay[J.did] >= 0 || net->M[J.did].p != nullptr, ITN_EINVAL,
                      "tol requires an existing message on every edge of the sequence (tol = nothing on trees)");
          oldsrc[i] = overlay[J.did];
          if (overlay[J.did] >= 0) lv = std::max(lv, level[overlay[J.did]] + 1);
        }
        level[i] = lv;
        overlay[J.did] = i;
        final_job[J.did] = i;
        touched.push_back(J.did);
      }
      for (int d : touched) overlay[d] = -1;
    }
  }
  for (size_t d = 0; d < nd; ++d)
    if (final_job[d] >= 0 && !net->M[d].p) {
      alloc_message(net, (int)d);
      CUDA_CHECK(cudaMemsetAsync(net->M[d].p, 0, (size_t)net->M[d].n * net->planes() * sizeof(double), ctx->stream));
    }
  Staged raw(net, jobs), res(net, jobs);
  DevBuf diffs(ctx, (size_t)std::max(nseq, 1) * sizeof(double));
  DevBuf dsum(ctx, sizeof(double));
  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0));
  CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, ctx->stream));
  int done = 0;
  double mean = NAN;
  try {
    for (int it = 0; it < maxiter; ++it) {
      for (int g = 0; g < ngroups; ++g) {
        const int lo = group_ptr[g], hi = group_ptr[g + 1];
        int maxlv = -1;
        for (int i = lo; i < hi; ++i) maxlv = std::max(maxlv, level[i]);
        for (int lv = 0; lv <= maxlv; ++lv) {
          std::vector<JobSpec> specs;
          std::vector<std::vector<const double*>> mats;
          std::vector<CommitJob> cj;
          std::vector<int> idx;
          for (int i = lo; i < hi; ++i)
            if (level[i] == lv) idx.push_back(i);
          mats.resize(idx.size());
          for (size_t q = 0; q < idx.size(); ++q) {
            const int i = idx[q];
            mats[q].assign(reads[i].size(), nullptr);
            for (size_t k = 0; k < reads[i].size(); ++k)
              if (reads[i][k] >= 0) mats[q][k] = res.ptr[reads[i][k]];
          }
          for (size_t q = 0; q < idx.size(); ++q) {
            const int i = idx[q];
            specs.push_back({jobs[i].v, 1u << (jobs[i].k + 1), raw.ptr[i], mats[q].data()});
            const int chi = net->edim[jobs[i].did / 2];
            const double* old = !want_diff ? nullptr : (oldsrc[i] >= 0 ? res.ptr[oldsrc[i]] : net->M[jobs[i].did].p);
            cj.push_back({raw.ptr[i], res.ptr[i], old, chi * chi});
          }
          itn_run_vertex_jobs(net, specs);
          // diffs of this level land at consecutive slots; the sum below runs over all nseq slots in job order
          DevBuf ld(ctx, cj.size() * sizeof(double));
          itn_run_commit(net, cj, normalize, want_diff ? ld.as<double>() : nullptr);
          if (want_diff)
            for (size_t q = 0; q < idx.size(); ++q)
              CUDA_CHECK(cudaMemcpyAsync(diffs.as<double>() + idx[q], ld.as<double>() + q, sizeof(double),
                                         cudaMemcpyDeviceToDevice, ctx->stream));
        }
      }
      // write-back of the sweep: the last listed update of every directed edge
      std::vector<CommitJob> wb;
      for (size_t d = 0; d < nd; ++d)
        if (final_job[d] >= 0) {
          const int chi = net->edim[d / 2];
          wb.push_back({res.ptr[final_job[d]], net->M[d].p, nullptr, chi * chi});
        }
      itn_run_commit(net, wb, 0, nullptr);
      ++done;
      if (want_diff) {
        k_sum_fixed<<<1, 256, 0, ctx->stream>>>(diffs.as<double>(), nseq, dsum.as<double>());
        ITN_LAUNCH_CHECK(ctx);
        double sum = 0;
        CUDA_CHECK(cudaMemcpyAsync(&sum, dsum.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        mean = sum / ngroups;
        if (mean <= tol) break;
      }
    }
    CUDA_CHECK(cudaEventRecord(ev1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ev1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    net->last_total_ms = ms;
    net->last_contract_ms = ms;
  } catch (...) {
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    throw;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  if (iters) *iters = done;
  if (last_mean_diff) *last_mean_diff = mean;
}

}  // namespace

extern "C" int itn_bp_update(itn_net* net, const int32_t* seq_src, const int32_t* seq_dst, int nseq,
                             const int32_t* group_ptr, int ngroups, int maxiter, double tol, int normalize,
                             int32_t* iters, double* last_mean_diff) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  // update(::Algorithm"bp"): "You need to specify a number of iterations for BP!" (:315-317)
  ITN_REQUIRE(maxiter >= 0, ITN_EINVAL, "You need to specify a number of iterations for BP!");
  ITN_REQUIRE(nseq >= 0 && (nseq == 0 || (seq_src && seq_dst)), ITN_EINVAL, "bad edge sequence");
  set_device(net->ctx);
  itn_ctx* ctx = net->ctx;
  if (iters) *iters = 0;
  if (last_mean_diff) *last_mean_diff = NAN;
  if (nseq == 0 || maxiter == 0) return ITN_OK;
  const bool sync_mode = group_ptr != nullptr;
  if (!sync_mode) itn_flush_pending(net);  // only the synchronous sweep overlaps the upload (see below)
  bool multi_edge_groups = false;
  if (sync_mode) {
    ITN_REQUIRE(ngroups >= 1 && group_ptr[0] == 0 && group_ptr[ngroups] == nseq, ITN_EINVAL,
                "group_ptr must hold ngroups + 1 offsets from 0 to nseq");
    for (int i = 0; i < ngroups; ++i) {
      ITN_REQUIRE(group_ptr[i + 1] >= group_ptr[i], ITN_EINVAL, "group_ptr must be non-decreasing");
      multi_edge_groups = multi_edge_groups || group_ptr[i + 1] - group_ptr[i] != 1;
    }
  }
  HostTrace trace("bp_update");
  std::vector<MsgJob> jobs = make_msg_jobs(net, seq_src, seq_dst, nseq);
  trace.mark("jobs");
  if (multi_edge_groups) {
    bp_update_groups(net, jobs, group_ptr, ngroups, maxiter, tol, normalize, iters, last_mean_diff);
    return ITN_OK;
  }
  const bool want_diff = tol >= 0.0;
  // multi-GPU: every rank receives the same full sequence and keeps the updates whose source vertex it owns;
  // the messages crossing a cut are exchanged once per sweep (itn_dist.cu)
  const int nseq_global = nseq;
  std::vector<int> global_dids(nseq);
  for (int i = 0; i < nseq; ++i) global_dids[i] = jobs[i].did;
  if (ctx->nranks > 1) {
    ITN_REQUIRE(sync_mode, ITN_EUNSUPPORTED,
                "the sequential (Gauss-Seidel) schedule does not shard: use the grouped/parallel schedule on a partitioned network");
    std::vector<MsgJob> loc;
    for (auto& J : jobs)
      if (itn_is_local(net, J.v)) loc.push_back(J);
    jobs.swap(loc);
    nseq = (int)jobs.size();
  }

  // availability simulation + dependency levels
  std::vector<char> valid(net->M.size());
  for (size_t d = 0; d < net->M.size(); ++d) valid[d] = net->M[d].p != nullptr;
  std::vector<int> level(nseq, 0);
  {
    std::vector<int> lastw(net->M.size(), -1), lastr(net->M.size(), 0);
    for (int i = 0; i < nseq; ++i) {
      const MsgJob& J = jobs[i];
      int lv = 0;
      for (int e : net->inc[J.v]) {
        if (e == J.did / 2) continue;
        int m = net->msg_into(J.v, e);
        ITN_REQUIRE(valid[m], ITN_EINVAL,
                    "message into vertex " + std::to_string(J.v) + " on edge " + std::to_string(e) +
                        " does not exist when updating " + std::to_string(J.v) + " -> " +
                        std::to_string(net->other(J.did / 2, J.v)) + " (initialise messages or use the forest-cover sequence)");
        if (!sync_mode) lv = std::max(lv, lastw[m] + 1);
      }
      if (want_diff)
        ITN_REQUIRE(net->M[J.did].p != nullptr || (!sync_mode && lastw[J.did] >= 0) , ITN_EINVAL,
                    "tol requires an existing message on every edge of the sequence (tol = nothing on trees)");
      if (!sync_mode) {
        lv = std::max(lv, lastw[J.did] + 1);
        lv = std::max(lv, lastr[J.did]);
        for (int e : net->inc[J.v]) {
          if (e == J.did / 2) continue;
          int m = net->msg_into(J.v, e);
          lastr[m] = std::max(lastr[m], lv);
        }
        lastw[J.did] = lv;
        valid[J.did] = 1;
      }
      level[i] = lv;
    }
  }
  // messages that do not exist yet get storage (filled before first read by construction)
  for (auto& J : jobs) {
    if (!net->M[J.did].p) {
      alloc_message(net, J.did);
      CUDA_CHECK(cudaMemsetAsync(net->M[J.did].p, 0, (size_t)net->M[J.did].n * net->planes() * sizeof(double), ctx->stream));
    }
  }
  // stable order by level
  std::vector<int> order(nseq);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });
  std::vector<MsgJob> sjobs(nseq);
  for (int i = 0; i < nseq; ++i) sjobs[i] = jobs[order[i]];
  std::vector<size_t> lvl_ptr{0};
  for (int i = 1; i < nseq; ++i)
    if (level[order[i]] != level[order[i - 1]]) lvl_ptr.push_back(i);
  lvl_ptr.push_back(nseq);

  trace.mark("levels");
  Staged staged(net, sjobs);
  DevBuf diffs(ctx, (size_t)std::max(nseq, 1) * sizeof(double));
  DevBuf dsum(ctx, sizeof(double));
  std::vector<int> all_dids(nseq), all_src(nseq);
  for (int i = 0; i < nseq; ++i) {
    all_dids[i] = sjobs[i].did;
    all_src[i] = sjobs[i].v;
  }
  // synchronous sweeps: vertices that qualify go to the DMMA kernels, the rest to the generic kernels
  std::vector<char> handled;
  // multi-GPU: updates whose message leaves this rank, and the vertices that produce them (they are swept first, their
  // messages committed and sent while the interior of the sweep runs: itn_dist_exchange_begin / _end)
  const bool overlap_halo = sync_mode && ctx->nranks > 1 && getenv("ITN_NO_HALO_OVERLAP") == nullptr;
  std::vector<char> cut_job(nseq, 0), cut_vertex;
  if (overlap_halo) {
    cut_vertex.assign(net->nv, 0);
    for (int i = 0; i < nseq; ++i) {
      const int w = net->other(sjobs[i].did / 2, sjobs[i].v);
      if (net->owner[w] != ctx->rank) {
        cut_job[i] = 1;
        cut_vertex[sjobs[i].v] = 1;
      }
    }
  }
  int nfirst = 0;
  const int nfast = sync_mode ? itn_fast_bp_plan(net, all_dids, all_src, handled, overlap_halo ? &cut_vertex : nullptr, &nfirst) : 0;
  trace.mark("plan_tile");
  // vertices the tile path did not take: the block path (itn_block.cu) for every degree 2..8 / bond extent <= 32
  const int nblock = sync_mode ? itn_block_bp_plan(net, all_dids, all_src, handled) : 0;
  trace.mark("plan_block");
  // the rest of a synchronous sweep: vertices whose outgoing messages are all part of it go to the vertex-level DMMA
  // sweep (shared partial absorptions, itn_run_vertex_sweeps), whatever is left to the per-message kernels
  std::vector<JobSpec> slow_specs;
  std::vector<SweepSpec> vsweeps;
  if (sync_mode) {
    if (handled.size() != (size_t)nseq) handled.assign(nseq, 0);
    std::vector<int> cnt(net->nv, 0), first(net->nv, -1);
    for (int i = 0; i < nseq; ++i)
      if (!handled[i]) cnt[sjobs[i].v]++;
    std::vector<int> vs_index(net->nv, -1);
    for (int i = 0; i < nseq; ++i) {
      if (handled[i]) continue;
      const int v = sjobs[i].v;
      if (cnt[v] == (int)net->inc[v].size() && itn_vertex_sweep_ok(net, v) && net->T[v].p) {
        if (vs_index[v] < 0) {
          vs_index[v] = (int)vsweeps.size();
          SweepSpec sp;
          memset(&sp, 0, sizeof(sp));
          sp.v = v;
          vsweeps.push_back(sp);
        }
        vsweeps[vs_index[v]].out[sjobs[i].k] = staged.ptr[i];
      } else {
        slow_specs.push_back({v, 1u << (sjobs[i].k + 1), staged.ptr[i]});
      }
    }
    // a sequence that lists a directed edge twice leaves a slot of the vertex sweep unset: fall back for that vertex
    for (size_t q = 0; q < vsweeps.size();) {
      const int v = vsweeps[q].v;
      bool full = true;
      for (size_t k = 0; k < net->inc[v].size(); ++k) full = full && vsweeps[q].out[k] != nullptr;
      if (full) {
        ++q;
        continue;
      }
      for (int i = 0; i < nseq; ++i)
        if (!handled[i] && sjobs[i].v == v) slow_specs.push_back({v, 1u << (sjobs[i].k + 1), staged.ptr[i]});
      vsweeps.erase(vsweeps.begin() + q);
    }
  }
  // Deferred host tensors (itn_net_set_tensors, ITN_HOST_DEFERRED): in a synchronous sweep the outgoing messages of a
  // vertex depend on its own tensor and the pre-sweep messages only, so the first sweep runs vertex chunk by vertex
  // chunk behind the host -> device copy: copy(c + 1) overlaps import + relayout + DMMA phases of chunk c.
  bool pipelined_first = false;
  std::vector<PendingUpload> pl_items;
  std::vector<size_t> pl_cuts;
  std::vector<std::pair<int, int>> pl_ranges;  // sweep positions [lo, hi) completed by chunk c
  if (!net->pending.empty()) {
    bool contiguous = false;
    if (nfast > 0) {
      const std::vector<int>& sv = itn_fast_sweep_vertices(net, &contiguous);
      if (contiguous) {
        std::vector<int> pos(net->nv, -1);
        for (size_t r = 0; r < sv.size(); ++r) pos[sv[r]] = (int)r;
        std::vector<PendingUpload> others, bucket;
        for (const PendingUpload& pu : net->pending) (pos[pu.v] >= 0 ? bucket : others).push_back(pu);
        std::sort(bucket.begin(), bucket.end(), [&](const PendingUpload& a, const PendingUpload& b) { return pos[a.v] < pos[b.v]; });
        // chunk 0: tensors outside the bucket (generic kernels run after the loop); then the bucket in sweep order
        pl_items = others;
        pl_cuts = {0, others.size()};
        pl_ranges.push_back({0, 0});
        pl_items.insert(pl_items.end(), bucket.begin(), bucket.end());
        std::vector<size_t> bc = cuts_by_bytes(net, pl_items, others.size(), pl_items.size(), kUploadChunkBytes);
        int done_pos = 0;
        for (size_t c = 1; c < bc.size(); ++c) {
          pl_cuts.push_back(bc[c]);
          // sweep positions below the first still-pending bucket vertex are complete after this chunk
          const int upto = bc[c] < pl_items.size() ? pos[pl_items[bc[c]].v] : (int)sv.size();
          pl_ranges.push_back({done_pos, upto});
          done_pos = upto;
        }
        if (bucket.empty()) pl_ranges.back().second = (int)sv.size();
        pipelined_first = true;
        net->pending.clear();
      }
    }
    if (!pipelined_first) {
      itn_flush_uploads(net);
      if (nfast > 0) {  // rebuild the tile-major copies (the block path's share of `handled` is kept)
        std::vector<char> h1;
        itn_fast_bp_plan(net, all_dids, all_src, h1, overlap_halo ? &cut_vertex : nullptr, &nfirst);
      }
    }
  }
  // vertices whose site tensor exists tile-major only (after a gate layer) and that this call reads canonically
  if (net->n_canon_stale) {
    if (sync_mode && nfast > 0) itn_canon_ensure_outside_sweep(net);
    else itn_canon_ensure_all(net);
  }
  struct BlockCall {  // per-call buffers of the block path, released on every exit
    itn_net* net;
    bool on;
    ~BlockCall() {
      if (on) itn_block_bp_end(net);
    }
  } block_call{net, nblock > 0};
  trace.mark("plan_rest");
  if (nblock > 0) itn_block_bp_begin(net, all_dids, all_src, handled, staged.ptr.data());
  trace.mark("block_begin");
  if (overlap_halo) itn_dist_exchange_prepare(net, global_dids);
  // commit descriptors of a synchronous sweep: [cut jobs | the rest] (one list, diffs in the same order)
  std::vector<CommitJob> cj_sync;
  size_t n_cut_jobs = 0;
  if (sync_mode) {
    cj_sync.reserve(nseq);
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < nseq; ++i) {
        if ((cut_job[i] != 0) != (pass == 0)) continue;
        const int chi = net->edim[sjobs[i].did / 2];
        cj_sync.push_back({staged.ptr[i], net->M[sjobs[i].did].p, want_diff ? net->M[sjobs[i].did].p : nullptr, chi * chi});
      }
    for (int i = 0; i < nseq; ++i) n_cut_jobs += cut_job[i] ? 1 : 0;
  }
  // uploaded once per call (as are the pointer tables of the tile and block paths): nothing inside the sweep loop copies
  // from pageable memory, so the host runs ahead of the stream and the GPU never waits for the next sweep to be enqueued
  DevBuf cj_dev(ctx, std::max<size_t>(cj_sync.size(), 1) * sizeof(CommitJob));
  const CommitJob* d_cj = sync_mode ? itn_upload(ctx, cj_sync, cj_dev) : nullptr;
  // (the upload-pipelined first sweep uploads the tile path's tables itself; they stay valid for the later sweeps)
  if (sync_mode && nfast > 0 && !pipelined_first) itn_fast_bp_sweep_begin(net, all_dids, all_src, handled, staged.ptr.data());

  cudaEvent_t ev0, ev1;
  CUDA_CHECK(cudaEventCreate(&ev0));
  CUDA_CHECK(cudaEventCreate(&ev1));
  CUDA_CHECK(cudaEventRecord(ev0, ctx->stream));
  int done = 0;
  double mean = NAN;
  std::vector<cudaEvent_t> cev;  // event pairs around the contraction kernels
  try {
    bool halo_in_flight = false;
    for (int it = 0; it < maxiter; ++it) {
      for (size_t l = 0; l + 1 < lvl_ptr.size(); ++l) {
        const size_t lo = lvl_ptr[l], hi = lvl_ptr[l + 1];
        const bool timed = cev.size() < 4096;
        if (timed) {
          cudaEvent_t a, b;
          CUDA_CHECK(cudaEventCreate(&a));
          cev.push_back(a);
          CUDA_CHECK(cudaEventCreate(&b));
          cev.push_back(b);
          CUDA_CHECK(cudaEventRecord(a, ctx->stream));
        }
        if (nfast > 0 && pipelined_first && it == 0) {
          itn_fast_bp_sweep_begin(net, all_dids, all_src, handled, staged.ptr.data());
          int swept = 0;
          upload_pipelined(net, pl_items, pl_cuts, [&](size_t c) {
            const std::pair<int, int> r = pl_ranges[c];
            if (r.second > r.first) {
              itn_fast_relayout_range(net, r.first, r.second);
              itn_fast_bp_sweep_range(net, r.first, r.second);
              swept = r.second;
            }
          });
          const int ns = (int)itn_fast_sweep_vertices(net, nullptr).size();
          if (swept < ns) {  // bucket vertices that were already resident
            itn_fast_relayout_range(net, swept, ns);
            itn_fast_bp_sweep_range(net, swept, ns);
          }
          itn_fast_bp_sweep_end(net);
          if (nblock > 0) {
            itn_block_bp_run(net, false);
            itn_block_bp_join(net);
          }
          itn_run_vertex_sweeps(net, vsweeps);
          itn_run_vertex_jobs(net, slow_specs);
        } else if (sync_mode && overlap_halo) {
          // everything that produces a message for another rank first: block buckets (side streams), the cut-adjacent
          // part of the tile sweep, the generic kernels; commit and send those messages; then the interior
          if (nblock > 0) itn_block_bp_run(net, nfast > 0);
          if (nfast > 0) {
            itn_fast_bp_sweep_range(net, 0, nfirst);
          }
          if (nblock > 0) itn_block_bp_join(net);
          itn_run_vertex_sweeps(net, vsweeps);
          itn_run_vertex_jobs(net, slow_specs);
          itn_run_commit_dev(net, d_cj, n_cut_jobs, normalize, want_diff ? diffs.as<double>() : nullptr);
          itn_dist_exchange_begin(net);
          halo_in_flight = true;
          if (nfast > 0) {
            itn_fast_bp_sweep_range(net, nfirst, (int)itn_fast_sweep_vertices(net, nullptr).size());
            itn_fast_bp_sweep_end(net);
          }
        } else if (sync_mode) {
          // the block buckets (rim vertices of a lattice whose bulk is on the tile path) run on side streams behind the
          // tile sweep: they read the same pre-sweep messages and write their own staged outputs
          if (nblock > 0) itn_block_bp_run(net, nfast > 0);
          if (nfast > 0) {
            itn_fast_bp_sweep_range(net, 0, (int)itn_fast_sweep_vertices(net, nullptr).size());
            itn_fast_bp_sweep_end(net);
          }
          if (nblock > 0) itn_block_bp_join(net);
          itn_run_vertex_sweeps(net, vsweeps);
          itn_run_vertex_jobs(net, slow_specs);
        } else {
          compute_messages(net, sjobs, lo, hi, staged.ptr);
        }
        if (timed) CUDA_CHECK(cudaEventRecord(cev.back(), ctx->stream));
        if (halo_in_flight) {
          itn_run_commit_dev(net, d_cj + n_cut_jobs, cj_sync.size() - n_cut_jobs, normalize, want_diff ? diffs.as<double>() + n_cut_jobs : nullptr);
        } else if (sync_mode) {
          itn_run_commit_dev(net, d_cj, cj_sync.size(), normalize, want_diff ? diffs.as<double>() : nullptr);
        } else {
          std::vector<CommitJob> cj(hi - lo);
          for (size_t i = lo; i < hi; ++i) {
            int chi = net->edim[sjobs[i].did / 2];
            cj[i - lo] = {staged.ptr[i], net->M[sjobs[i].did].p, want_diff ? net->M[sjobs[i].did].p : nullptr, chi * chi};
          }
          itn_run_commit(net, cj, normalize, want_diff ? diffs.as<double>() + lo : nullptr);
        }
      }
      if (halo_in_flight) {
        itn_dist_exchange_end(net);
        halo_in_flight = false;
      } else if (ctx->nranks > 1) {
        itn_dist_exchange(net, global_dids);
      }
      ++done;
      if (sync_mode) {
        ctx->path_msgs[0] += nfast;
        ctx->path_msgs[1] += nblock;
        ctx->path_msgs[2] += nseq - nfast - nblock;
      } else {
        ctx->path_msgs[2] += nseq;
      }
      if (want_diff) {
        k_sum_fixed<<<1, 256, 0, ctx->stream>>>(diffs.as<double>(), nseq, dsum.as<double>());
        ITN_LAUNCH_CHECK(ctx);
        itn_dist_allreduce_sum(ctx, dsum.as<double>(), 1);
        double s = 0;
        CUDA_CHECK(cudaMemcpyAsync(&s, dsum.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        mean = s / nseq_global;
        if (mean <= tol) break;
      }
    }
    trace.mark("launch");
    CUDA_CHECK(cudaEventRecord(ev1, ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(ev1));
    trace.mark("wait");
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    net->last_total_ms = ms;
    double cms = 0;
    for (size_t i = 0; i + 1 < cev.size(); i += 2) {
      float t = 0;
      CUDA_CHECK(cudaEventElapsedTime(&t, cev[i], cev[i + 1]));
      cms += t;
    }
    net->last_contract_ms = cms;
  } catch (...) {
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    for (auto e : cev) cudaEventDestroy(e);
    throw;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  for (auto e : cev) cudaEventDestroy(e);
  if (iters) *iters = done;
  if (last_mean_diff) *last_mean_diff = mean;
  API_END
}

extern "C" int itn_ctx_path_counts(const itn_ctx* ctx, int64_t* out3) {
  API_BEGIN
  ITN_REQUIRE(ctx && out3, ITN_EINVAL, "NULL argument");
  for (int i = 0; i < 3; ++i) out3[i] = ctx->path_msgs[i];
  API_END
}

extern "C" int itn_ctx_cholqr2_count(const itn_ctx* ctx, int64_t* out) {
  API_BEGIN
  ITN_REQUIRE(ctx && out, ITN_EINVAL, "NULL argument");
  *out = ctx->cholqr2_sides;
  API_END
}

extern "C" int itn_bp_last_timing(const itn_net* net, double* total_ms, double* contract_ms) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  if (total_ms) *total_ms = net->last_total_ms;
  if (contract_ms) *contract_ms = net->last_contract_ms;
  API_END
}

static void updated_messages_to_scratch(itn_net* net, const std::vector<MsgJob>& jobs, int normalize,
                                        Staged& staged, Staged& dest, double* d_diffs) {
  for (auto& J : jobs)
    for (int e : net->inc[J.v])
      if (e != J.did / 2)
        ITN_REQUIRE(net->M[net->msg_into(J.v, e)].p, ITN_EINVAL, "an incoming message is not set");
  compute_messages(net, jobs, 0, jobs.size(), staged.ptr);
  std::vector<CommitJob> cj(jobs.size());
  for (size_t i = 0; i < jobs.size(); ++i) {
    int chi = net->edim[jobs[i].did / 2];
    cj[i] = {staged.ptr[i], dest.ptr[i], d_diffs ? net->M[jobs[i].did].p : nullptr, chi * chi};
  }
  itn_run_commit(net, cj, normalize, d_diffs);
}

extern "C" int itn_updated_message(itn_net* net, int src, int dst, int normalize, void* host) {
  API_BEGIN
  ITN_REQUIRE(net && host, ITN_EINVAL, "NULL argument");
  set_device(net->ctx);
  itn_flush_pending(net);
  int32_t s = src, d = dst;
  std::vector<MsgJob> jobs = make_msg_jobs(net, &s, &d, 1);
  Staged staged(net, jobs), dest(net, jobs);
  updated_messages_to_scratch(net, jobs, normalize, staged, dest, nullptr);
  long long chi = net->edim[jobs[0].did / 2];
  download_planar(net, dest.ptr[0], chi * chi, host);
  API_END
}

extern "C" int itn_message_residuals(itn_net* net, const int32_t* src, const int32_t* dst, int n, double* out) {
  API_BEGIN
  ITN_REQUIRE(net && out && (n == 0 || (src && dst)), ITN_EINVAL, "NULL argument");
  if (n == 0) return ITN_OK;
  set_device(net->ctx);
  itn_flush_pending(net);
  std::vector<MsgJob> jobs = make_msg_jobs(net, src, dst, n);
  for (auto& J : jobs) ITN_REQUIRE(net->M[J.did].p, ITN_EINVAL, "message is not set");
  Staged staged(net, jobs), dest(net, jobs);
  DevBuf diffs(net->ctx, (size_t)n * sizeof(double));
  updated_messages_to_scratch(net, jobs, 1, staged, dest, diffs.as<double>());
  CUDA_CHECK(cudaMemcpyAsync(out, diffs.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, net->ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  API_END
}

// ------------------------------------------------------------------------------------------------
// scalars / rescale
// ------------------------------------------------------------------------------------------------
static void require_all_set(itn_net* net) {
  for (int v = 0; v < net->nv; ++v)
    if (itn_is_local(net, v)) ITN_REQUIRE(net->T[v].p, ITN_EINVAL, "site tensor of vertex " + std::to_string(v) + " is not set");
  for (size_t d = 0; d < net->M.size(); ++d)
    if (msg_stored(net, (int)d)) ITN_REQUIRE(net->M[d].p, ITN_EINVAL, "a message is not set (run itn_bp_update first)");
}

// d_zv: 2 doubles per vertex {re, im}
static void vertex_scalars_dev(itn_net* net, double* d_zv) {
  std::vector<JobSpec> specs;
  // planar No=1 output: out[0] = re, out[1] = im for complex; real writes out[0] only
  CUDA_CHECK(cudaMemsetAsync(d_zv, 0, (size_t)net->nv * 2 * sizeof(double), net->ctx->stream));
  for (int v = 0; v < net->nv; ++v)
    if (itn_is_local(net, v)) specs.push_back({v, 0u, d_zv + 2 * v});
  itn_run_vertex_jobs(net, specs);
}

static void edge_scalars_dev(itn_net* net, double* d_ze) {
  if (net->ne == 0) return;
  // each edge scalar is computed by the rank owning esrc (zero elsewhere; summed by the all-reduce)
  CUDA_CHECK(cudaMemsetAsync(d_ze, 0, (size_t)net->ne * 2 * sizeof(double), net->ctx->stream));
  std::vector<EdgePair> ep(net->ne);
  for (int e = 0; e < net->ne; ++e) {
    if (itn_is_local(net, net->esrc[e])) ep[e] = {net->M[2 * e].p, net->M[2 * e + 1].p, (int)net->M[2 * e].n};
    else ep[e] = {nullptr, nullptr, 0};
  }
  DevBuf b(net->ctx, ep.size() * sizeof(EdgePair));
  const EdgePair* d = itn_upload(net->ctx, ep, b);
  if (net->cplx) k_edge_scalar<true><<<net->ne, 32, 0, net->ctx->stream>>>(d, d_ze);
  else k_edge_scalar<false><<<net->ne, 32, 0, net->ctx->stream>>>(d, d_ze);
  ITN_LAUNCH_CHECK(net->ctx);
}

static void region_scalars_host(itn_net* net, std::vector<std::complex<double>>& zv, std::vector<std::complex<double>>& ze) {
  require_all_set(net);
  DevBuf dz(net->ctx, (size_t)(net->nv + net->ne) * 2 * sizeof(double));
  vertex_scalars_dev(net, dz.as<double>());
  edge_scalars_dev(net, dz.as<double>() + 2 * (size_t)net->nv);
  itn_dist_allreduce_sum(net->ctx, dz.as<double>(), (net->nv + net->ne) * 2);
  std::vector<double> h((size_t)(net->nv + net->ne) * 2);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), dz.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, net->ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(net->ctx->stream));
  zv.resize(net->nv);
  ze.resize(net->ne);
  for (int v = 0; v < net->nv; ++v) zv[v] = {h[2 * v], h[2 * v + 1]};
  for (int e = 0; e < net->ne; ++e) ze[e] = {h[2 * (net->nv + e)], h[2 * (net->nv + e) + 1]};
}

extern "C" int itn_region_scalars(itn_net* net, void* z_v, void* z_e) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  set_device(net->ctx);
  itn_flush_pending(net);
  std::vector<std::complex<double>> zv, ze;
  region_scalars_host(net, zv, ze);
  if (net->cplx) {
    if (z_v) memcpy(z_v, zv.data(), zv.size() * sizeof(std::complex<double>));
    if (z_e) memcpy(z_e, ze.data(), ze.size() * sizeof(std::complex<double>));
  } else {
    if (z_v) for (int v = 0; v < net->nv; ++v) ((double*)z_v)[v] = zv[v].real();
    if (z_e) for (int e = 0; e < net->ne; ++e) ((double*)z_e)[e] = ze[e].real();
  }
  API_END
}

extern "C" int itn_logscalar(itn_net* net, double out[2]) {
  API_BEGIN
  ITN_REQUIRE(net && out, ITN_EINVAL, "NULL argument");
  set_device(net->ctx);
  itn_flush_pending(net);
  std::vector<std::complex<double>> zv, ze;
  region_scalars_host(net, zv, ze);
  // logscalar (abstractbeliefpropagationcache.jl:397-408): the O(nv + ne) log-sum over the device-computed scalars
  for (auto& z : ze)
    if (z == std::complex<double>(0.0, 0.0)) {
      out[0] = -INFINITY;
      out[1] = 0.0;
      return ITN_OK;
    }
  std::complex<double> acc(0.0, 0.0);
  for (auto& z : zv) acc += std::log(z);
  for (auto& z : ze) acc -= std::log(z);
  out[0] = acc.real();
  out[1] = acc.imag();
  API_END
}

static int rescale_impl(itn_net* net, const int32_t* verts, int nverts);

extern "C" int itn_rescale(itn_net* net) { return rescale_impl(net, nullptr, -1); }

extern "C" int itn_rescale_verts(itn_net* net, const int32_t* verts, int n) {
  if (n < 0 || (n > 0 && !verts)) {
    itn_set_error("itn_rescale_verts: verts must hold n >= 0 vertex ids");
    return ITN_EINVAL;
  }
  return rescale_impl(net, verts, n);
}

// nverts < 0: every vertex; otherwise only the listed vertices (ket and bra of each) are rescaled
static int rescale_impl(itn_net* net, const int32_t* verts, int nverts) {
  API_BEGIN
  ITN_REQUIRE(net, ITN_EINVAL, "net is NULL");
  std::vector<char> sel;
  if (nverts >= 0) {
    sel.assign(net->nv, 0);
    for (int i = 0; i < nverts; ++i) {
      ITN_REQUIRE(verts[i] >= 0 && verts[i] < net->nv, ITN_EINVAL, "vertex out of range");
      sel[verts[i]] = 1;
    }
  }
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "not defined for a bilinear form network (a bra layer is set): only BP updates, region scalars and logscalar are");
  set_device(net->ctx);
  itn_flush_pending(net);
  itn_ctx* ctx = net->ctx;
  require_all_set(net);
  {
    // every rank rescales the message pairs it stores (both owners of a cut edge do the same arithmetic)
    std::vector<EdgePair> ep;
    for (int e = 0; e < net->ne; ++e)
      if (msg_stored(net, 2 * e)) ep.push_back({net->M[2 * e].p, net->M[2 * e + 1].p, (int)net->M[2 * e].n});
    if (!ep.empty()) {
      DevBuf b(ctx, ep.size() * sizeof(EdgePair));
      const EdgePair* d = itn_upload(ctx, ep, b);
      if (net->cplx) k_rescale_msgs<true><<<(unsigned)ep.size(), 32, 0, ctx->stream>>>(d);
      else k_rescale_msgs<false><<<(unsigned)ep.size(), 32, 0, ctx->stream>>>(d);
      ITN_LAUNCH_CHECK(ctx);
    }
  }
  DevBuf dz(ctx, (size_t)net->nv * 2 * sizeof(double));
  vertex_scalars_dev(net, dz.as<double>());
  std::vector<ScaleJob> sj;
  long long maxn = 0;
  for (int v = 0; v < net->nv; ++v) {
    if (!itn_is_local(net, v) || (!sel.empty() && !sel[v])) continue;
    sj.push_back({net->T[v].p, net->T[v].n * net->planes(), dz.as<double>() + 2 * v, 1.0});
    maxn = std::max<long long>(maxn, sj.back().n);
  }
  if (!sj.empty()) {
    DevBuf sb(ctx, sj.size() * sizeof(ScaleJob));
    const ScaleJob* dj = itn_upload(ctx, sj, sb);
    unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn + 1023) / 1024, 32));
    k_scale<<<dim3((unsigned)sj.size(), gy), 256, 0, ctx->stream>>>(dj);
    ITN_LAUNCH_CHECK(ctx);
  }
  for (int v = 0; v < net->nv; ++v)
    if (itn_is_local(net, v) && (sel.empty() || sel[v])) net->touch(v);
  net->topo_version++;
  API_END
}

// ------------------------------------------------------------------------------------------------
// observables
// ------------------------------------------------------------------------------------------------
static void upload_planar(itn_net* net, const void* host, long long n_each, int count, double* dev) {
  // host: `count` interleaved arrays of n_each elements -> device planar per array
  const int P = net->planes();
  std::vector<double> tmp((size_t)n_each * P * count);
  const double* h = (const double*)host;
  for (int c = 0; c < count; ++c)
    for (long long i = 0; i < n_each; ++i) {
      if (net->cplx) {
        tmp[(size_t)c * n_each * 2 + i] = h[((size_t)c * n_each + i) * 2];
        tmp[(size_t)c * n_each * 2 + n_each + i] = h[((size_t)c * n_each + i) * 2 + 1];
      } else {
        tmp[(size_t)c * n_each + i] = h[(size_t)c * n_each + i];
      }
    }
  CUDA_CHECK(cudaMemcpyAsync(dev, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, net->ctx->stream));
}

extern "C" int itn_expect1(itn_net* net, const int32_t* verts, int n, const void* ops, void* out) {
  API_BEGIN
  ITN_REQUIRE(net && verts && ops && out && n >= 0, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "not defined for a bilinear form network (a bra layer is set): only BP updates, region scalars and logscalar are");
  if (n == 0) return ITN_OK;
  set_device(net->ctx);
  itn_flush_pending(net);
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  // all ops must share the layout d_v x d_v; they are packed back to back with their own d
  size_t op_elems = 0, rho_elems = 0;
  for (int i = 0; i < n; ++i) {
    ITN_REQUIRE(verts[i] >= 0 && verts[i] < net->nv, ITN_EINVAL, "vertex out of range");
    int d = net->sdim[verts[i]];
    op_elems += (size_t)d * d;
    rho_elems += (size_t)d * d;
  }
  DevBuf dops(ctx, op_elems * P * sizeof(double)), drho(ctx, rho_elems * P * sizeof(double));
  DevBuf dout(ctx, (size_t)n * 2 * sizeof(double));
  std::vector<JobSpec> specs;
  std::vector<ExpectJob> ej(n);
  size_t off = 0, hoff = 0;
  for (int i = 0; i < n; ++i) {
    int v = verts[i], d = net->sdim[v];
    double* rho = drho.as<double>() + off * P;
    double* op = dops.as<double>() + off * P;
    upload_planar(net, (const char*)ops + hoff * P * sizeof(double), (long long)d * d, 1, op);
    if (itn_is_local(net, v)) {
      for (int e : net->inc[v]) ITN_REQUIRE(net->M[net->msg_into(v, e)].p, ITN_EINVAL, "an incoming message is not set");
      specs.push_back({v, 1u, rho});
      ej[i] = {rho, op, d};
    } else {
      ej[i] = {nullptr, op, d};  // computed by the owning rank, summed in by the all-reduce below
    }
    off += (size_t)d * d;
    hoff += (size_t)d * d;
  }
  itn_run_vertex_jobs(net, specs);
  DevBuf jb(ctx, ej.size() * sizeof(ExpectJob));
  const ExpectJob* dj = itn_upload(ctx, ej, jb);
  if (net->cplx) k_expect_finalize<true><<<(n + 127) / 128, 128, 0, ctx->stream>>>(dj, n, dout.as<double>());
  else k_expect_finalize<false><<<(n + 127) / 128, 128, 0, ctx->stream>>>(dj, n, dout.as<double>());
  ITN_LAUNCH_CHECK(ctx);
  itn_dist_allreduce_sum(ctx, dout.as<double>(), n * 2);
  std::vector<double> h((size_t)n * 2);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), dout.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (net->cplx) memcpy(out, h.data(), h.size() * sizeof(double));
  else for (int i = 0; i < n; ++i) ((double*)out)[i] = h[2 * i];
  API_END
}

extern "C" int itn_rdm2(itn_net* net, const int32_t* eids, int n, void* out) {
  API_BEGIN
  ITN_REQUIRE(net && eids && out && n >= 0, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "not defined for a bilinear form network (a bra layer is set): only BP updates, region scalars and logscalar are");
  if (n == 0) return ITN_OK;
  set_device(net->ctx);
  itn_flush_pending(net);
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  const bool multi = ctx->nranks > 1;
  // Partitioned network: every rank is called with the same list.  The bond environment of a site is computed where the
  // site lives; the rank that owns esrc combines the two environments of an edge (for an edge that crosses a cut the
  // other rank ships its (d chi)^2 environment, as itn_apply2 does), and one all-reduce hands every rank every matrix.
  int maxD2 = 0;
  std::map<int, std::pair<size_t, size_t>> seg;  // peer -> (doubles to send, doubles to receive)
  for (int i = 0; i < n; ++i) {
    int e = eids[i];
    ITN_REQUIRE(e >= 0 && e < net->ne, ITN_EINVAL, "edge id out of range");
    int u = net->esrc[e], v = net->edst[e];
    int D = net->sdim[u] * net->sdim[v];
    ITN_REQUIRE(D * D <= 256, ITN_EUNSUPPORTED, "rdm2 supports d_u*d_v <= 16");
    maxD2 = std::max(maxD2, D * D);
    const bool lu = itn_is_local(net, u), lv = itn_is_local(net, v);
    const size_t nv2 = (size_t)P * net->sdim[v] * net->edim[e] * net->sdim[v] * net->edim[e];
    if (lu && !lv) seg[net->owner[v]].second += nv2;
    if (!lu && lv) seg[net->owner[u]].first += nv2;
  }
  size_t tot_s = 0, tot_r = 0;
  std::map<int, std::pair<size_t, size_t>> segoff, cur;
  for (auto& kv : seg) {
    segoff[kv.first] = {tot_s, tot_r};
    cur[kv.first] = {0, 0};
    tot_s += kv.second.first;
    tot_r += kv.second.second;
  }
  DevBuf sbuf(ctx, tot_s * sizeof(double)), rbuf(ctx, tot_r * sizeof(double));
  size_t env_elems = 0;
  for (int i = 0; i < n; ++i) {
    int e = eids[i], u = net->esrc[e], v = net->edst[e];
    long long nu = (long long)net->sdim[u] * net->edim[e], nv_ = (long long)net->sdim[v] * net->edim[e];
    if (itn_is_local(net, u)) env_elems += (size_t)(nu * nu);
    if (itn_is_local(net, u) && itn_is_local(net, v)) env_elems += (size_t)(nv_ * nv_);
  }
  DevBuf denv(ctx, std::max<size_t>(env_elems, 1) * P * sizeof(double));
  DevBuf dout(ctx, (size_t)n * maxD2 * 2 * sizeof(double));
  CUDA_CHECK(cudaMemsetAsync(dout.p, 0, (size_t)n * maxD2 * 2 * sizeof(double), ctx->stream));
  std::vector<JobSpec> specs;
  std::vector<Rdm2Job> rj;
  std::vector<int> rj_gate;
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    int e = eids[i], u = net->esrc[e], v = net->edst[e];
    const bool lu = itn_is_local(net, u), lv = itn_is_local(net, v);
    if (!lu && !lv) continue;
    long long nu = (long long)net->sdim[u] * net->edim[e], nv_ = (long long)net->sdim[v] * net->edim[e];
    for (int w : {u, v})
      if (itn_is_local(net, w))
        for (int f : net->inc[w])
          if (f != e) ITN_REQUIRE(net->M[net->msg_into(w, f)].p, ITN_EINVAL, "an incoming message is not set");
    double *eu = nullptr, *ev = nullptr;
    if (lu) {
      eu = denv.as<double>() + off * P;
      off += (size_t)(nu * nu);
      specs.push_back({u, 1u | (1u << (net->slot(u, e) + 1)), eu});
      if (lv) {
        ev = denv.as<double>() + off * P;
        off += (size_t)(nv_ * nv_);
        specs.push_back({v, 1u | (1u << (net->slot(v, e) + 1)), ev});
      } else {  // arrives from the rank that owns v
        const int peer = net->owner[v];
        ev = rbuf.as<double>() + segoff[peer].second + cur[peer].second;
        cur[peer].second += (size_t)P * nv_ * nv_;
      }
      rj.push_back({eu, ev, net->sdim[u], net->sdim[v], net->edim[e]});
      rj_gate.push_back(i);
    } else {  // guest: the environment of v goes to the owner of u
      const int peer = net->owner[u];
      ev = sbuf.as<double>() + segoff[peer].first + cur[peer].first;
      cur[peer].first += (size_t)P * nv_ * nv_;
      specs.push_back({v, 1u | (1u << (net->slot(v, e) + 1)), ev});
    }
  }
  itn_run_vertex_jobs(net, specs);
  if (multi) {
    std::vector<P2PSeg> xs;
    for (auto& kv : seg)
      xs.push_back({kv.first, sbuf.as<double>() + segoff[kv.first].first, kv.second.first,
                    rbuf.as<double>() + segoff[kv.first].second, kv.second.second});
    itn_dist_p2p(ctx, xs);
  }
  if (!rj.empty()) {
    // the finalize kernel writes row blockIdx.x of its output: give every owned gate its own row of dout
    DevBuf jb(ctx, rj.size() * sizeof(Rdm2Job)), tmp(ctx, rj.size() * (size_t)maxD2 * 2 * sizeof(double));
    const Rdm2Job* dj = itn_upload(ctx, rj, jb);
    if (net->cplx) k_rdm2_finalize<true><<<(unsigned)rj.size(), 256, 0, ctx->stream>>>(dj, tmp.as<double>(), maxD2 * 2);
    else k_rdm2_finalize<false><<<(unsigned)rj.size(), 256, 0, ctx->stream>>>(dj, tmp.as<double>(), maxD2 * 2);
    ITN_LAUNCH_CHECK(ctx);
    for (size_t k = 0; k < rj.size(); ++k)
      CUDA_CHECK(cudaMemcpyAsync(dout.as<double>() + (size_t)rj_gate[k] * maxD2 * 2, tmp.as<double>() + k * (size_t)maxD2 * 2,
                                 (size_t)maxD2 * 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  itn_dist_allreduce_sum(ctx, dout.as<double>(), n * maxD2 * 2);
  std::vector<double> h((size_t)n * maxD2 * 2);
  CUDA_CHECK(cudaMemcpyAsync(h.data(), dout.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  double* o = (double*)out;
  size_t oo = 0;
  for (int i = 0; i < n; ++i) {
    int e = eids[i];
    int D = net->sdim[net->esrc[e]] * net->sdim[net->edst[e]];
    for (int t = 0; t < D * D; ++t) {
      if (net->cplx) {
        o[2 * (oo + t)] = h[(size_t)i * maxD2 * 2 + 2 * t];
        o[2 * (oo + t) + 1] = h[(size_t)i * maxD2 * 2 + 2 * t + 1];
      } else {
        o[oo + t] = h[(size_t)i * maxD2 * 2 + 2 * t];
      }
    }
    oo += (size_t)D * D;
  }
  API_END
}

// ------------------------------------------------------------------------------------------------
// one-site gates
// ------------------------------------------------------------------------------------------------
extern "C" int itn_apply1(itn_net* net, const int32_t* verts, int n, const void* gates, int normalize) {
  API_BEGIN
  ITN_REQUIRE(net && verts && gates && n >= 0, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "not defined for a bilinear form network (a bra layer is set): only BP updates, region scalars and logscalar are");
  if (n == 0) return ITN_OK;
  set_device(net->ctx);
  itn_flush_pending(net);
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  size_t g_elems = 0;
  std::vector<char> seen(net->nv, 0);
  for (int i = 0; i < n; ++i) {
    int v = verts[i];
    ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "Gate being applied does not share indices with tensor network.");
    ITN_REQUIRE(!seen[v], ITN_EINVAL, "a batch of one-site gates must act on distinct vertices");
    seen[v] = 1;
    // multi-GPU: every rank is called with the same batch; vertices stored by another rank are skipped (their gate
    // still consumes its place in `gates`), the contract of itn_net_set_tensors / itn_expect1 / itn_apply2
    if (itn_is_local(net, v)) ITN_REQUIRE(net->T[v].p, ITN_EINVAL, "site tensor is not set");
    g_elems += (size_t)net->sdim[v] * net->sdim[v];
  }
  DevBuf dg(ctx, std::max<size_t>(g_elems, 1) * P * sizeof(double));
  std::vector<Apply1Job> jobs;
  std::vector<NormJob> nj;
  std::vector<int> done;
  size_t off = 0;
  long long maxn = 0;
  for (int i = 0; i < n; ++i) {
    int v = verts[i], d = net->sdim[v];
    ITN_REQUIRE(d <= 8, ITN_EUNSUPPORTED, "one-site gates support site dimensions up to 8");
    if (itn_is_local(net, v)) {
      double* g = dg.as<double>() + off * P;
      upload_planar(net, (const char*)gates + off * P * sizeof(double), (long long)d * d, 1, g);
      jobs.push_back({net->T[v].p, net->T[v].n, g, d, normalize});
      nj.push_back({net->T[v].p, net->T[v].n * P});
      maxn = std::max<long long>(maxn, net->T[v].n);
      done.push_back(v);
    }
    off += (size_t)d * d;
  }
  if (jobs.empty()) return ITN_OK;
  n = (int)jobs.size();
  DevBuf jb(ctx, jobs.size() * sizeof(Apply1Job));
  const Apply1Job* dj = itn_upload(ctx, jobs, jb);
  unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn / 2 + 255) / 256, 64));
  if (net->cplx) k_apply1<true><<<dim3(n, gy), 256, 0, ctx->stream>>>(dj, nullptr);
  else k_apply1<false><<<dim3(n, gy), 256, 0, ctx->stream>>>(dj, nullptr);
  ITN_LAUNCH_CHECK(ctx);
  if (normalize) {
    DevBuf nb(ctx, nj.size() * sizeof(NormJob));
    const NormJob* dn = itn_upload(ctx, nj, nb);
    k_normalize<<<n, 256, 0, ctx->stream>>>(dn);
    ITN_LAUNCH_CHECK(ctx);
  }
  for (int v : done) net->touch(v);
  net->topo_version++;
  API_END
}

// ------------------------------------------------------------------------------------------------
// Pairwise contraction of two host tensors on the device: the `contract` of the tensors inside one partition
// (NDTensors contract called from src/caches/abstractbeliefpropagationcache.jl:232-233 on a multi-site partition).
// The host mirror uses it to merge the site tensors of a partition into one super-site tensor (generalised
// partitions); this is set-up work on small tensors, not a tuned kernel: one thread per output
// element, a serial loop over the contracted multi-index.
// ------------------------------------------------------------------------------------------------
namespace {
struct DotSpec {
  int nfa, nfb, nc;
  int fa_dim[16], fb_dim[16], c_dim[16];
  long long fa_str[16], fb_str[16], ca_str[16], cb_str[16];
  long long nout, ncon;
};

template <bool C>
__global__ void __launch_bounds__(256) k_tensordot(const double* __restrict__ a, const double* __restrict__ b,
                                                   double* __restrict__ out, DotSpec S) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < S.nout; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx, oa = 0, ob = 0;
    for (int q = 0; q < S.nfa; ++q) {
      oa += (r % S.fa_dim[q]) * S.fa_str[q];
      r /= S.fa_dim[q];
    }
    for (int q = 0; q < S.nfb; ++q) {
      ob += (r % S.fb_dim[q]) * S.fb_str[q];
      r /= S.fb_dim[q];
    }
    double accr = 0.0, acci = 0.0;
    for (long long c = 0; c < S.ncon; ++c) {
      long long rc = c, pa = oa, pb = ob;
      for (int q = 0; q < S.nc; ++q) {
        const long long k = rc % S.c_dim[q];
        rc /= S.c_dim[q];
        pa += k * S.ca_str[q];
        pb += k * S.cb_str[q];
      }
      if (C) {
        const double xr = a[2 * pa], xi = a[2 * pa + 1], yr = b[2 * pb], yi = b[2 * pb + 1];
        accr += xr * yr - xi * yi;
        acci += xr * yi + xi * yr;
      } else {
        accr += a[pa] * b[pb];
      }
    }
    if (C) {
      out[2 * idx] = accr;
      out[2 * idx + 1] = acci;
    } else {
      out[idx] = accr;
    }
  }
}
}  // namespace

extern "C" int itn_tensordot(itn_ctx* ctx, int dtype, const void* a_host, int nda, const int32_t* dims_a, const void* b_host,
                             int ndb, const int32_t* dims_b, int npairs, const int32_t* axes_a, const int32_t* axes_b,
                             void* out_host) {
  API_BEGIN
  ITN_REQUIRE(ctx && a_host && b_host && out_host && dims_a && dims_b, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(dtype == ITN_F64 || dtype == ITN_C128, ITN_EUNSUPPORTED, "dtype must be 0 (Float64) or 1 (ComplexF64)");
  ITN_REQUIRE(nda >= 0 && ndb >= 0 && nda <= 16 && ndb <= 16 && npairs >= 0 && npairs <= nda && npairs <= ndb, ITN_EINVAL,
              "bad tensor ranks");
  ITN_REQUIRE(npairs == 0 || (axes_a && axes_b), ITN_EINVAL, "NULL argument");
  set_device(ctx);
  DotSpec S;
  memset(&S, 0, sizeof(S));
  std::vector<long long> sa(nda), sb(ndb);
  long long na = 1, nb = 1;
  for (int i = 0; i < nda; ++i) {
    ITN_REQUIRE(dims_a[i] >= 1, ITN_ESHAPE, "extents must be positive");
    sa[i] = na;
    na *= dims_a[i];
  }
  for (int i = 0; i < ndb; ++i) {
    ITN_REQUIRE(dims_b[i] >= 1, ITN_ESHAPE, "extents must be positive");
    sb[i] = nb;
    nb *= dims_b[i];
  }
  std::vector<char> ca(nda, 0), cb(ndb, 0);
  S.ncon = 1;
  for (int q = 0; q < npairs; ++q) {
    const int x = axes_a[q], y = axes_b[q];
    ITN_REQUIRE(x >= 0 && x < nda && y >= 0 && y < ndb && !ca[x] && !cb[y], ITN_EINVAL, "bad contraction axes");
    ITN_REQUIRE(dims_a[x] == dims_b[y], ITN_ESHAPE, "contracted extents differ");
    ca[x] = cb[y] = 1;
    S.c_dim[q] = dims_a[x];
    S.ca_str[q] = sa[x];
    S.cb_str[q] = sb[y];
    S.ncon *= dims_a[x];
  }
  S.nc = npairs;
  S.nout = 1;
  for (int i = 0; i < nda; ++i)
    if (!ca[i]) {
      S.fa_dim[S.nfa] = dims_a[i];
      S.fa_str[S.nfa++] = sa[i];
      S.nout *= dims_a[i];
    }
  for (int i = 0; i < ndb; ++i)
    if (!cb[i]) {
      S.fb_dim[S.nfb] = dims_b[i];
      S.fb_str[S.nfb++] = sb[i];
      S.nout *= dims_b[i];
    }
  const size_t item = (dtype == ITN_C128 ? 2 : 1) * sizeof(double);
  DevBuf da(ctx, (size_t)na * item), db(ctx, (size_t)nb * item), dout(ctx, (size_t)S.nout * item);
  CUDA_CHECK(cudaMemcpyAsync(da.p, a_host, (size_t)na * item, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(db.p, b_host, (size_t)nb * item, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((S.nout + 255) / 256, 148 * 8));
  if (dtype == ITN_C128) k_tensordot<true><<<grid, 256, 0, ctx->stream>>>(da.as<double>(), db.as<double>(), dout.as<double>(), S);
  else k_tensordot<false><<<grid, 256, 0, ctx->stream>>>(da.as<double>(), db.as<double>(), dout.as<double>(), S);
  ITN_LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(out_host, dout.p, (size_t)S.nout * item, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}
