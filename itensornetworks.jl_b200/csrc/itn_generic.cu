// Generic (any degree, any per-bond extent, real or complex) vertex-contraction kernels.
//
// A job contracts one site tensor with its conjugate and the incoming messages on every closed
// bond, leaving a set of modes open:
//   out[o, o'] = sum_x B[x, o] * conj(A'[x, o'])        (x = closed multi-index, o = open multi-index)
// with A' = A permuted to [closed..., open...] and B = A' with every closed bond's incoming message
// absorbed (one mode product per bond).  This restates
//   updated_message(::Algorithm"contract")  src/caches/abstractbeliefpropagationcache.jl:225-239  (open = {k})
//   region_scalar(bpc, vertex)              src/caches/beliefpropagationcache.jl:107-113          (open = {})
//   expect numerator/denominator            src/expect.jl:5-19                                    (open = {site})
//   two-site RDM environment                test/test_belief_propagation.jl:64-91                 (open = {site, k})
// with the closed-form pairwise order "absorb z-1 messages into the ket, close with the bra"
// instead of a per-call optimaltree search (abstract...cache.jl:232).
//
// Storage is planar (re plane, im plane).  These kernels are the shape-agnostic path and the on-GPU
// second opinion for the DMMA fast path (itn_fast.cu); FP64 FMA pipe, not tensor cores.
#include <algorithm>
#include <climits>
#include <cstring>
#include <functional>

#include "itn_internal.h"

namespace {

constexpr int kThreads = 256;

template <bool C>
__global__ void __launch_bounds__(kThreads) k_permute(const VJob* __restrict__ jobs, int bra) {
  const VJob& J = jobs[blockIdx.x];
  if (J.identity_perm) return;
  if (bra && !J.b) return;
  const long long n = J.n;
  const double* __restrict__ a = bra ? J.b : J.a;
  double* __restrict__ ap = bra ? J.bp : J.ap;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.y * blockDim.x) {
    long long r = i, off = 0;
#pragma unroll 1
    for (int m = 0; m < J.nm; ++m) {
      int d = J.dims[m];
      long long q = r / d;
      off += (r - q * d) * J.pstride[m];
      r = q;
    }
    ap[off] = a[i];
    if (C) ap[n + off] = a[n + i];
  }
}

// out[l, b, r] = sum_a in[l, a, r] * m(a, b);   m(a,b) = M[a + K*b]  (or M[b + N*a] if trans), optionally conj
template <bool C>
__global__ void __launch_bounds__(kThreads) k_modeprod(const VJob* __restrict__ jobs, int step, int smem_elems) {
  extern __shared__ double sm[];
  const VJob& J = jobs[blockIdx.x];
  if (step >= J.nsteps) return;
  const ModeStep S = J.steps[step];
  const double* __restrict__ src = (step == 0) ? J.ap : J.w[(step - 1) & 1];
  double* __restrict__ dst = J.w[step & 1];
  const long long L = S.L, R = S.R;
  const int K = S.K, N = S.N;
  const long long n_in = L * K * R, n_out = L * N * R;
  const int kn = K * N;
  const bool use_sm = kn <= smem_elems;
  // stage m as mm[a + K*b] (already transposed / conjugated)
  if (use_sm) {
    for (int i = threadIdx.x; i < kn; i += blockDim.x) {
      int a = i % K, b = i / K;
      int srcidx = S.trans ? (b + N * a) : i;
      sm[i] = S.m[srcidx];
      if (C) sm[kn + i] = S.conj ? -S.m[S.mplane + srcidx] : S.m[S.mplane + srcidx];
    }
    __syncthreads();
  }
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < n_out;
       idx += (long long)gridDim.y * blockDim.x) {
    long long l = idx % L;
    long long t = idx / L;
    int b = (int)(t % N);
    long long r = t / N;
    const double* __restrict__ s = src + l + L * (long long)K * r;
    double accr = 0.0, acci = 0.0;
    if (use_sm) {
      const double* mr = sm + (long long)K * b;
      const double* mi = sm + kn + (long long)K * b;
      for (int a = 0; a < K; ++a) {
        double xr = s[L * a];
        if (C) {
          double xi = s[n_in + L * a];
          accr += xr * mr[a] - xi * mi[a];
          acci += xr * mi[a] + xi * mr[a];
        } else {
          accr += xr * mr[a];
        }
      }
    } else {
      for (int a = 0; a < K; ++a) {
        int mi_ = S.trans ? (b + N * a) : (a + K * b);
        double mr = S.m[mi_];
        double xr = s[L * a];
        if (C) {
          double mi = S.conj ? -S.m[S.mplane + mi_] : S.m[S.mplane + mi_];
          double xi = s[n_in + L * a];
          accr += xr * mr - xi * mi;
          acci += xr * mi + xi * mr;
        } else {
          accr += xr * mr;
        }
      }
    }
    dst[idx] = accr;
    if (C) dst[n_out + idx] = acci;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[o + No*o'] = sum_x B[x + X*o] * conj(A'[x + X*o'])
template <bool C>
__global__ void __launch_bounds__(kThreads) k_gram(const VJob* __restrict__ jobs) {
  __shared__ double part[8 * 64 * 2];  // S <= 8 splits, No*No <= 64 when S > 1
  const VJob& J = jobs[blockIdx.x];
  const long long X = J.X;
  const int No = J.No;
  const long long n = J.n;
  const double* __restrict__ B = (J.nsteps == 0) ? J.ap : J.w[(J.nsteps - 1) & 1];
  const double* __restrict__ A = J.b ? J.bp : J.ap;  // bilinear form: close with the bra layer
  double* __restrict__ out = J.out;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t1 = (No + 3) >> 2, T = t1 * t1;
  const int S = (T >= nwarps) ? 1 : (nwarps / T);
  const int n2 = No * No;
  for (int item = warp; item < T * S; item += nwarps) {
    const int tile = item / S, split = item % S;
    const int o0 = (tile % t1) * 4, p0 = (tile / t1) * 4;
    const long long xlo = X * split / S, xhi = X * (split + 1) / S;
    double ar[4][4], ai[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.0;
    for (long long x = xlo + lane; x < xhi; x += 32) {
      double br[4], bi[4], cr[4], ci[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bool ok = (o0 + i) < No;
        br[i] = ok ? B[x + X * (o0 + i)] : 0.0;
        bi[i] = (C && ok) ? B[n + x + X * (o0 + i)] : 0.0;
        bool ok2 = (p0 + i) < No;
        cr[i] = ok2 ? A[x + X * (p0 + i)] : 0.0;
        ci[i] = (C && ok2) ? A[n + x + X * (p0 + i)] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // b * conj(c) = (br*cr + bi*ci) + i (bi*cr - br*ci)
          ar[i][j] += br[i] * cr[j];
          if (C) {
            ar[i][j] += bi[i] * ci[j];
            ai[i][j] += bi[i] * cr[j] - br[i] * ci[j];
          }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double vr = warp_sum(ar[i][j]);
        double vi = C ? warp_sum(ai[i][j]) : 0.0;
        if (lane == 0 && (o0 + i) < No && (p0 + j) < No) {
          int oi = (o0 + i) + No * (p0 + j);
          if (S == 1) {
            out[oi] = vr;
            if (C) out[n2 + oi] = vi;
          } else {
            part[(split * 64 + oi) * 2] = vr;
            part[(split * 64 + oi) * 2 + 1] = vi;
          }
        }
      }
  }
  if (S > 1) {
    __syncthreads();
    for (int oi = threadIdx.x; oi < n2; oi += blockDim.x) {
      double vr = 0.0, vi = 0.0;
      for (int s = 0; s < S; ++s) {
        vr += part[(s * 64 + oi) * 2];
        vi += part[(s * 64 + oi) * 2 + 1];
      }
      out[oi] = vr;
      if (C) out[n2 + oi] = vi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// DMMA versions of the two kernels above (FP64 mma.sync m16n8k8, ComplexF64 as four real products on the planes).
// Fragments are loaded straight from global memory: every element of the big operand is used by exactly one warp,
// so there is nothing to share through shared memory except the small message matrix.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// out[l, b, r] = sum_a in[l, a, r] * M[a + K b].  Rows of the product are the flattened (l, r) pairs, 16 per warp tile;
// the message matrix sits in shared memory, zero padded to multiples of 8 (row stride K8 + 4: conflict-free fragments).
template <bool C>
__global__ void __launch_bounds__(kThreads) k_modeprod_mma(const VJob* __restrict__ jobs, int step) {
  extern __shared__ double sm[];
  const VJob& J = jobs[blockIdx.x];
  if (step >= J.nsteps) return;
  const ModeStep S = J.steps[step];
  const double* __restrict__ src = (step == 0) ? J.ap : J.w[(step - 1) & 1];
  double* __restrict__ dst = J.w[step & 1];
  const long long L = S.L, R = S.R;
  const int K = S.K, N = S.N;
  const int K8 = (K + 7) & ~7, N8 = (N + 7) & ~7, ld = K8 + 4;
  const long long n_in = L * K * R, n_out = L * N * R;
  double* Mr = sm;
  double* Mi = sm + (size_t)ld * N8;
  for (int i = threadIdx.x; i < ld * N8; i += blockDim.x) {
    const int k = i % ld, b = i / ld;
    const bool ok = k < K && b < N;
    Mr[i] = ok ? S.m[k + K * b] : 0.0;
    if (C) Mi[i] = ok ? S.m[S.mplane + k + K * b] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long rows = L * R;
  const long long ntiles = (rows + 15) >> 4;
  for (long long tile = (long long)blockIdx.y * (kThreads / 32) + warp; tile < ntiles;
       tile += (long long)gridDim.y * (kThreads / 32)) {
    long long ib[2], ob[2];
    bool ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long rho = tile * 16 + g + 8 * h;
      ok[h] = rho < rows;
      const long long l = ok[h] ? rho % L : 0, r = ok[h] ? rho / L : 0;
      ib[h] = l + L * K * r;
      ob[h] = l + L * N * r;
    }
    for (int n0 = 0; n0 < N8; n0 += 16) {
      const int nbs = (N8 - n0) >= 16 ? 2 : 1;
      double cr[2][4], ci[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int v = 0; v < 4; ++v) cr[nb][v] = ci[nb][v] = 0.0;
      for (int k0 = 0; k0 < K8; k0 += 8) {
        double ar[4], ai[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int h = v & 1, k = k0 + t + 4 * (v >> 1);
          const bool lo = ok[h] && k < K;
          const long long a = ib[h] + L * k;
          ar[v] = lo ? src[a] : 0.0;
          ai[v] = (C && lo) ? src[n_in + a] : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          if (nb >= nbs) break;
          double br[2], bi[2], nbi[2];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int o = (k0 + t + 4 * v) + ld * (n0 + 8 * nb + g);
            br[v] = Mr[o];
            bi[v] = C ? Mi[o] : 0.0;
            nbi[v] = -bi[v];
          }
          dmma_16x8x8(cr[nb], ar, br);
          if (C) {
            dmma_16x8x8(ci[nb], ar, bi);
            dmma_16x8x8(cr[nb], ai, nbi);
            dmma_16x8x8(ci[nb], ai, br);
          }
        }
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        if (nb >= nbs) break;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int h = v >> 1, b = n0 + 8 * nb + 2 * t + (v & 1);
          if (ok[h] && b < N) {
            const long long o = ob[h] + L * b;
            dst[o] = cr[nb][v];
            if (C) dst[n_out + o] = ci[nb][v];
          }
        }
      }
    }
  }
}

// out[o + No*o'] = sum_x B[x + X*o] * conj(A'[x + X*o']).  Work items are (16-row tile, pair of 8-column blocks, split of x);
// a warp owns one item at a time, partial sums of the splits meet in shared memory in a fixed order.
template <bool C>
__global__ void __launch_bounds__(kThreads) k_gram_mma(const VJob* __restrict__ jobs) {
  extern __shared__ double sm[];  // [S][2][No16 * No16]
  const VJob& J = jobs[blockIdx.x];
  const long long X = J.X;
  const int No = J.No;
  const long long n = J.n;
  const double* __restrict__ B = (J.nsteps == 0) ? J.ap : J.w[(J.nsteps - 1) & 1];
  const double* __restrict__ A = J.b ? J.bp : J.ap;  // bilinear form: close with the bra layer
  double* __restrict__ out = J.out;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int No16 = (No + 15) & ~15;
  const int mt = No16 >> 4, np = No16 >> 4;  // 16-row tiles, pairs of 8-column blocks
  const int T = mt * np;
  const int S = (T >= nwarps) ? 1 : (nwarps / T);
  const int n2 = No * No, p2 = No16 * No16;
  for (int item = warp; item < T * S; item += nwarps) {
    const int tile = item / S, split = item % S;
    const int o0 = (tile % mt) * 16, p0 = (tile / mt) * 16;
    const long long xlo = (X * split / S) & ~7ll, xhi = (split == S - 1) ? X : ((X * (split + 1) / S) & ~7ll);
    double cr[2][4], ci[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int v = 0; v < 4; ++v) cr[nb][v] = ci[nb][v] = 0.0;
    for (long long x0 = xlo; x0 < xhi; x0 += 8) {
      double ar[4], ai[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int o = o0 + g + 8 * (v & 1);
        const long long x = x0 + t + 4 * (v >> 1);
        const bool ok = o < No && x < xhi;
        ar[v] = ok ? B[x + X * o] : 0.0;
        ai[v] = (C && ok) ? B[n + x + X * o] : 0.0;
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        double br[2], bi[2], nbi[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int p = p0 + 8 * nb + g;
          const long long x = x0 + t + 4 * v;
          const bool ok = p < No && x < xhi;
          br[v] = ok ? A[x + X * p] : 0.0;
          bi[v] = (C && ok) ? A[n + x + X * p] : 0.0;
          nbi[v] = -bi[v];
        }
        // b conj(a): re += br_B ar_A + bi_B ai_A ; im += bi_B ar_A - br_B ai_A   (ar/ai = B fragment, br/bi = A' fragment)
        dmma_16x8x8(cr[nb], ar, br);
        if (C) {
          dmma_16x8x8(cr[nb], ai, bi);
          dmma_16x8x8(ci[nb], ai, br);
          dmma_16x8x8(ci[nb], ar, nbi);
        }
      }
    }
    double* pr = sm + (size_t)split * 2 * p2;
    double* pi = pr + p2;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int o = o0 + g + 8 * (v >> 1), p = p0 + 8 * nb + 2 * t + (v & 1);
        pr[o + No16 * p] = cr[nb][v];
        if (C) pi[o + No16 * p] = ci[nb][v];
      }
  }
  __syncthreads();
  for (int oi = threadIdx.x; oi < n2; oi += blockDim.x) {
    const int o = oi % No, p = oi / No;
    double vr = 0.0, vi = 0.0;
    for (int s2 = 0; s2 < S; ++s2) {
      vr += sm[(size_t)s2 * 2 * p2 + o + No16 * p];
      if (C) vi += sm[(size_t)s2 * 2 * p2 + p2 + o + No16 * p];
    }
    out[oi] = vr;
    if (C) out[n2 + oi] = vi;
  }
}

// ------------------------------------------------------------------------------------------------
// Vertex-level synchronous sweep on the canonical layout (no permutation): all z outgoing messages of a vertex with the
// partial absorptions shared between them (divide and conquer over the bond set: z = 3 -> 5 mode products instead of 6,
// z = 4 -> 8 instead of 12, z = 6 -> 16 instead of 30), every step on DMMA.
//   mode product   dst[l, b, r] = sum_a src[l, a, r] M[a + K b]                      (MpOp)
//   close          out[o + chi o'] = sum_{l, r} B[l, o, r] conj(A[l, o', r])          (GrOp, any position of the open bond)
// ------------------------------------------------------------------------------------------------
struct MpOp {
  const double* src;
  double* dst;
  const double* m;  // planar K x K
  long long L, R;
  int K;
};
struct GrOp {
  const double* B;
  const double* A;
  double* out;  // planar chi x chi
  long long L, R;
  int chi;
  double* part;  // gridDim.y > 1: gridDim.y partial results (planar chi x chi each), summed in order by k_gr_reduce
};

template <bool C>
__global__ void __launch_bounds__(kThreads) k_mp_mma(const MpOp* __restrict__ ops) {
  extern __shared__ double sm[];
  const MpOp S = ops[blockIdx.x];
  const double* __restrict__ src = S.src;
  double* __restrict__ dst = S.dst;
  const long long L = S.L, R = S.R;
  const int K = S.K;
  const int K8 = (K + 7) & ~7, ld = K8 + 4;
  const long long nel = L * K * R;
  double* Mr = sm;
  double* Mi = sm + (size_t)ld * K8;
  for (int i = threadIdx.x; i < ld * K8; i += blockDim.x) {
    const int k = i % ld, b = i / ld;
    const bool ok = k < K && b < K;
    Mr[i] = ok ? S.m[k + K * b] : 0.0;
    if (C) Mi[i] = ok ? S.m[(long long)K * K + k + K * b] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long rows = L * R;
  const long long ntiles = (rows + 15) >> 4;
  for (long long tile = (long long)blockIdx.y * (kThreads / 32) + warp; tile < ntiles;
       tile += (long long)gridDim.y * (kThreads / 32)) {
    long long ib[2];
    bool ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long rho = tile * 16 + g + 8 * h;
      ok[h] = rho < rows;
      const long long l = ok[h] ? rho % L : 0, r = ok[h] ? rho / L : 0;
      ib[h] = l + L * K * r;
    }
    for (int n0 = 0; n0 < K8; n0 += 16) {
      const int nbs = (K8 - n0) >= 16 ? 2 : 1;
      double cr[2][4], ci[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int v = 0; v < 4; ++v) cr[nb][v] = ci[nb][v] = 0.0;
      for (int k0 = 0; k0 < K8; k0 += 8) {
        double ar[4], ai[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int h = v & 1, k = k0 + t + 4 * (v >> 1);
          const bool lo = ok[h] && k < K;
          const long long a = ib[h] + L * k;
          ar[v] = lo ? src[a] : 0.0;
          ai[v] = (C && lo) ? src[nel + a] : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          if (nb >= nbs) break;
          double br[2], bi[2], nbi[2];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int o = (k0 + t + 4 * v) + ld * (n0 + 8 * nb + g);
            br[v] = Mr[o];
            bi[v] = C ? Mi[o] : 0.0;
            nbi[v] = -bi[v];
          }
          dmma_16x8x8(cr[nb], ar, br);
          if (C) {
            dmma_16x8x8(ci[nb], ar, bi);
            dmma_16x8x8(cr[nb], ai, nbi);
            dmma_16x8x8(ci[nb], ai, br);
          }
        }
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        if (nb >= nbs) break;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int h = v >> 1, b = n0 + 8 * nb + 2 * t + (v & 1);
          if (ok[h] && b < K) {
            const long long o = ib[h] + L * b;
            dst[o] = cr[nb][v];
            if (C) dst[nel + o] = ci[nb][v];
          }
        }
      }
    }
  }
}

template <bool C>
__global__ void __launch_bounds__(kThreads) k_gr_mma(const GrOp* __restrict__ ops) {
  extern __shared__ double sm[];  // [S][2][No16 * No16]
  const GrOp J = ops[blockIdx.x];
  const long long L = J.L, X = J.L * J.R;
  const int No = J.chi;
  const long long n = X * No, LK = L * No;
  const double* __restrict__ B = J.B;
  const double* __restrict__ A = J.A;
  double* __restrict__ out = J.out;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int No16 = (No + 15) & ~15;
  const int mt = No16 >> 4;
  const int T = mt * mt;
  const int S = (T >= nwarps) ? 1 : (nwarps / T);
  const int n2 = No * No, p2 = No16 * No16;
  // small graphs: the closed index range is also split across gridDim.y CTAs (partials, fixed-order reduction)
  const int Y = gridDim.y, yb = blockIdx.y;
  const long long Xlo = (X * yb / Y) & ~7ll, Xhi = (yb == Y - 1) ? X : ((X * (yb + 1) / Y) & ~7ll);
  const long long Xc = Xhi - Xlo;
  if (Y > 1) out = J.part + (size_t)yb * (C ? 2 : 1) * n2;
  for (int item = warp; item < T * S; item += nwarps) {
    const int tile = item / S, split = item % S;
    const int o0 = (tile % mt) * 16, p0 = (tile / mt) * 16;
    const long long xlo = Xlo + ((Xc * split / S) & ~7ll), xhi = (split == S - 1) ? Xhi : (Xlo + ((Xc * (split + 1) / S) & ~7ll));
    double cr[2][4], ci[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int v = 0; v < 4; ++v) cr[nb][v] = ci[nb][v] = 0.0;
    for (long long x0 = xlo; x0 < xhi; x0 += 8) {
      // element (x, o) lives at (x mod L) + L chi (x div L) + L o
      long long xb[2];
      bool xok[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const long long x = x0 + t + 4 * v;
        xok[v] = x < xhi;
        const long long r = x / L;
        xb[v] = (x - r * L) + LK * r;
      }
      double ar[4], ai[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int o = o0 + g + 8 * (v & 1);
        const bool ok = o < No && xok[v >> 1];
        const long long a = xb[v >> 1] + L * o;
        ar[v] = ok ? B[a] : 0.0;
        ai[v] = (C && ok) ? B[n + a] : 0.0;
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        double br[2], bi[2], nbi[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int p = p0 + 8 * nb + g;
          const bool ok = p < No && xok[v];
          const long long a = xb[v] + L * p;
          br[v] = ok ? A[a] : 0.0;
          bi[v] = (C && ok) ? A[n + a] : 0.0;
          nbi[v] = -bi[v];
        }
        dmma_16x8x8(cr[nb], ar, br);
        if (C) {
          dmma_16x8x8(cr[nb], ai, bi);
          dmma_16x8x8(ci[nb], ai, br);
          dmma_16x8x8(ci[nb], ar, nbi);
        }
      }
    }
    double* pr = sm + (size_t)split * 2 * p2;
    double* pi = pr + p2;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int o = o0 + g + 8 * (v >> 1), p = p0 + 8 * nb + 2 * t + (v & 1);
        pr[o + No16 * p] = cr[nb][v];
        if (C) pi[o + No16 * p] = ci[nb][v];
      }
  }
  __syncthreads();
  for (int oi = threadIdx.x; oi < n2; oi += blockDim.x) {
    const int o = oi % No, p = oi / No;
    double vr = 0.0, vi = 0.0;
    for (int s2 = 0; s2 < S; ++s2) {
      vr += sm[(size_t)s2 * 2 * p2 + o + No16 * p];
      if (C) vi += sm[(size_t)s2 * 2 * p2 + p2 + o + No16 * p];
    }
    out[oi] = vr;
    if (C) out[n2 + oi] = vi;
  }
}

// out = sum over the Y partials of an operation, in index order
template <bool C>
__global__ void __launch_bounds__(256) k_gr_reduce(const GrOp* __restrict__ ops, int Y) {
  const GrOp J = ops[blockIdx.x];
  const int tot = (C ? 2 : 1) * J.chi * J.chi;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    double acc = 0.0;
    for (int y = 0; y < Y; ++y) acc += J.part[(size_t)y * tot + i];
    J.out[i] = acc;
  }
}

// Block-wide deterministic sum of up to 3 values.
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* sh) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    sh[warp * 3] = a;
    sh[warp * 3 + 1] = b;
    sh[warp * 3 + 2] = c;
  }
  __syncthreads();
  double ta = 0, tb = 0, tc = 0;
  for (int w = 0; w < nw; ++w) {
    ta += sh[w * 3];
    tb += sh[w * 3 + 1];
    tc += sh[w * 3 + 2];
  }
  a = ta;
  b = tb;
  c = tc;
  __syncthreads();
}

// Frobenius-normalise the staged message, message_diff against the old one, write to dest.
// message_diff (abstractbeliefpropagationcache.jl:32-36): 1 - |<a^, b^>|^2.
//
// herm (norm networks <psi|psi>, bra = ket): the message is replaced by its Hermitian part (m + m^H) / 2 first.  In exact
// arithmetic it IS Hermitian (every update maps Hermitian messages to a Hermitian one); in floating point the update is
// multilinear in the z - 1 incoming messages, so a complex phase e^{i phi_e} on message e propagates as
// phi_out = sum of phi_in: the phases obey the non-backtracking matrix of the graph and grow by a factor ~ (z - 1) per
// sweep (measured: 1e-16 -> O(1) in ~35 sweeps on the square lattice, faster at z = 6), which the Frobenius
// normalisation of the reference (:234-237) does not remove.  The Hermitian part differs from the reference's message by
// rounding (1e-16) for as long as the reference's own phases are still small, and stays put afterwards.
template <bool C>
__global__ void __launch_bounds__(128) k_commit(const CommitJob* __restrict__ jobs, int normalize, int herm,
                                                double* __restrict__ diffs) {
  __shared__ double sh[16];
  const CommitJob J = jobs[blockIdx.x];
  const int n2 = J.n2;
  int chi = 0;
  if (herm) {
    chi = (int)(sqrt((double)n2) + 0.5);
    if (chi * chi != n2) herm = 0;  // not a square matrix (never the case for a message)
  }
  auto load = [&](int i, double& nr, double& ni) {
    nr = J.staged[i];
    ni = C ? J.staged[n2 + i] : 0.0;
    if (herm) {
      const int c = i / chi, r = i - c * chi, t = c + chi * r;
      nr = 0.5 * (nr + J.staged[t]);
      ni = C ? 0.5 * (ni - J.staged[n2 + t]) : 0.0;
    }
  };
  double ss = 0.0, so = 0.0, dr = 0.0;
  double di = 0.0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    double nr, ni;
    load(i, nr, ni);
    ss += nr * nr + ni * ni;
    if (J.old) {
      double orr = J.old[i], oi = C ? J.old[n2 + i] : 0.0;
      so += orr * orr + oi * oi;
      dr += nr * orr + ni * oi;   // conj(new) * old
      di += nr * oi - ni * orr;
    }
  }
  block_sum3(ss, so, dr, sh);
  double dummy1 = 0, dummy2 = 0;
  block_sum3(di, dummy1, dummy2, sh);
  const double nrm = sqrt(ss);
  const double scale = (normalize && nrm != 0.0) ? 1.0 / nrm : 1.0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    double nr, ni;
    load(i, nr, ni);
    J.dest[i] = nr * scale;
    if (C) J.dest[n2 + i] = ni * scale;
  }
  if (diffs && threadIdx.x == 0) {
    double f = (dr * dr + di * di) / (ss * so);
    diffs[blockIdx.x] = J.old ? 1.0 - f : 0.0;
  }
}

}  // namespace

int itn_open_extent(const itn_net* net, int v, uint32_t open_mask) {
  long long no = 1;
  if (open_mask & 1u) no *= net->sdim[v];
  for (size_t k = 0; k < net->inc[v].size(); ++k)
    if (open_mask & (1u << (k + 1))) no *= net->edim[net->inc[v][k]];
  return (int)no;
}

void itn_run_commit(itn_net* net, const std::vector<CommitJob>& jobs, int normalize, double* d_diffs) {
  if (jobs.empty()) return;
  itn_ctx* ctx = net->ctx;
  const int herm = net->has_bra() ? 0 : 1;  // bilinear forms <phi|psi> have no Hermitian messages
  DevBuf buf(ctx, jobs.size() * sizeof(CommitJob));
  const CommitJob* d = itn_upload(ctx, jobs, buf);
  if (net->cplx)
    k_commit<true><<<(unsigned)jobs.size(), 128, 0, ctx->stream>>>(d, normalize, herm, d_diffs);
  else
    k_commit<false><<<(unsigned)jobs.size(), 128, 0, ctx->stream>>>(d, normalize, herm, d_diffs);
  ITN_LAUNCH_CHECK(ctx);
}

// the same with descriptors that already live on the device (uploaded once per itn_bp_update call: a pageable upload
// inside the sweep loop makes the host wait for the stream and the GPU idle while the next sweep is being enqueued)
void itn_run_commit_dev(itn_net* net, const CommitJob* d_jobs, size_t n, int normalize, double* d_diffs) {
  if (n == 0) return;
  itn_ctx* ctx = net->ctx;
  const int herm = net->has_bra() ? 0 : 1;
  if (net->cplx)
    k_commit<true><<<(unsigned)n, 128, 0, ctx->stream>>>(d_jobs, normalize, herm, d_diffs);
  else
    k_commit<false><<<(unsigned)n, 128, 0, ctx->stream>>>(d_jobs, normalize, herm, d_diffs);
  ITN_LAUNCH_CHECK(ctx);
}

void itn_run_vertex_jobs(itn_net* net, const std::vector<JobSpec>& specs) {
  if (specs.empty()) return;
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  // Build descriptors (workspace pointers filled per batch).
  std::vector<VJob> jobs(specs.size());
  for (size_t j = 0; j < specs.size(); ++j) {
    const JobSpec& sp = specs[j];
    const int v = sp.v;
    ITN_REQUIRE(v >= 0 && v < net->nv, ITN_EINVAL, "vertex out of range");
    ITN_REQUIRE(net->T[v].p != nullptr, ITN_EINVAL, "site tensor of vertex " + std::to_string(v) + " is not set");
    const int z = (int)net->inc[v].size();
    ITN_REQUIRE(z + 1 <= ITN_MAX_MODES, ITN_EUNSUPPORTED, "vertex degree too large");
    VJob J;
    memset(&J, 0, sizeof(J));
    J.a = sp.tensor ? sp.tensor : net->T[v].p;
    J.b = (net->has_bra() && net->Tb[v].p) ? net->Tb[v].p : nullptr;
    J.n = net->T[v].n;
    J.nm = z + 1;
    J.dims[0] = net->sdim[v];
    for (int k = 0; k < z; ++k) J.dims[k + 1] = net->edim[net->inc[v][k]];
    // order: closed modes first (ascending), then open modes (ascending)
    int order[ITN_MAX_MODES], pos = 0;
    for (int m = 0; m < J.nm; ++m)
      if (!(sp.open_mask & (1u << m))) order[pos++] = m;
    const int nclosed = pos;
    for (int m = 0; m < J.nm; ++m)
      if (sp.open_mask & (1u << m)) order[pos++] = m;
    long long stride = 1;
    bool ident = true;
    long long X = 1, No = 1;
    int pdims[ITN_MAX_MODES];
    for (int q = 0; q < J.nm; ++q) {
      int m = order[q];
      if (m != q) ident = false;
      J.pstride[m] = stride;
      pdims[q] = J.dims[m];
      stride *= J.dims[m];
      if (q < nclosed) X *= J.dims[m]; else No *= J.dims[m];
    }
    J.identity_perm = ident ? 1 : 0;
    J.X = X;
    J.No = (int)No;
    // mode-product chain over closed bond modes
    long long Lacc = 1;
    for (int q = 0; q < nclosed; ++q) {
      int m = order[q];
      if (m >= 1 && !sp.no_messages) {
        int e = net->inc[v][m - 1];
        const DevTensor& msg = net->M[net->msg_into(v, e)];
        const double* ovr = sp.mats ? sp.mats[m - 1] : nullptr;
        ITN_REQUIRE(ovr != nullptr || msg.p != nullptr, ITN_EINVAL,
                    "message into vertex " + std::to_string(v) + " on edge " + std::to_string(e) +
                        " is not set (on trees use the forest-cover sequence)");
        ModeStep& S = J.steps[J.nsteps++];
        S.L = Lacc;
        S.K = S.N = pdims[q];
        S.R = J.n / (Lacc * pdims[q]);
        S.m = ovr ? ovr : msg.p;
        S.mplane = (long long)pdims[q] * pdims[q];
        S.trans = 0;
        S.conj = 0;
      }
      Lacc *= pdims[q];
    }
    J.out = sp.out;
    jobs[j] = J;
  }
  // Batches bounded by the scratch budget: each job needs up to 3 planar copies.
  size_t lo = 0;
  const size_t N = jobs.size();
  while (lo < N) {
    size_t hi = lo;
    size_t bytes = 0;
    long long maxn = 0;
    int maxsteps = 0, maxkn = 0;
    while (hi < N) {
      const VJob& J = jobs[hi];
      size_t need = (size_t)J.n * P * sizeof(double) * ((J.identity_perm ? 0 : (J.b ? 2 : 1)) + (J.nsteps >= 2 ? 2 : J.nsteps));
      if (hi > lo && bytes + need > ctx->ws_budget) break;
      bytes += need;
      ++hi;
    }
    DevBuf ws(ctx, std::max<size_t>(bytes, 16));
    char* base = ws.as<char>();
    size_t off = 0;
    std::vector<VJob> batch(jobs.begin() + lo, jobs.begin() + hi);
    for (VJob& J : batch) {
      size_t tb = (size_t)J.n * P * sizeof(double);
      if (J.identity_perm) {
        J.ap = const_cast<double*>(J.a);
        J.bp = const_cast<double*>(J.b);
      } else {
        J.ap = (double*)(base + off);
        off += tb;
        if (J.b) {
          J.bp = (double*)(base + off);
          off += tb;
        }
      }
      for (int i = 0; i < std::min(J.nsteps, 2); ++i) {
        J.w[i] = (double*)(base + off);
        off += tb;
      }
      maxn = std::max<long long>(maxn, J.n);
      maxsteps = std::max(maxsteps, J.nsteps);
      for (int s = 0; s < J.nsteps; ++s) maxkn = std::max(maxkn, J.steps[s].K * J.steps[s].N);
    }
    DevBuf jb(ctx, batch.size() * sizeof(VJob));
    const VJob* dj = itn_upload(ctx, batch, jb);
    const unsigned nb = (unsigned)batch.size();
    unsigned gy = (unsigned)std::min<long long>((maxn + kThreads * 4 - 1) / (kThreads * 4), 64);
    if (gy < 1) gy = 1;
    // keep the total grid reasonable for huge batches
    while (gy > 1 && (unsigned long long)gy * nb > 148ull * 64ull) gy = (gy + 1) / 2;
    dim3 grid(nb, gy);
    bool any_perm = false, any_bra_perm = false;
    for (const VJob& J : batch) {
      any_perm |= !J.identity_perm;
      any_bra_perm |= !J.identity_perm && J.b;
    }
    if (any_perm) {
      if (net->cplx) k_permute<true><<<grid, kThreads, 0, ctx->stream>>>(dj, 0);
      else k_permute<false><<<grid, kThreads, 0, ctx->stream>>>(dj, 0);
      ITN_LAUNCH_CHECK(ctx);
    }
    if (any_bra_perm) {
      if (net->cplx) k_permute<true><<<grid, kThreads, 0, ctx->stream>>>(dj, 1);
      else k_permute<false><<<grid, kThreads, 0, ctx->stream>>>(dj, 1);
      ITN_LAUNCH_CHECK(ctx);
    }
    int maxK = 0, maxNo = 0;
    bool plain = true;
    for (const VJob& J : batch) {
      maxNo = std::max(maxNo, J.No);
      for (int s = 0; s < J.nsteps; ++s) {
        maxK = std::max(maxK, std::max(J.steps[s].K, J.steps[s].N));
        plain = plain && !J.steps[s].trans && !J.steps[s].conj;
      }
    }
    const bool use_mma = ctx->path_mode != 2 && plain && maxK <= 64 && maxNo <= 64;
    if (use_mma) {
      // DMMA kernels: the message matrix padded to multiples of 8 in shared memory; 16-row tiles of flattened (l, r)
      const int K8 = (maxK + 7) & ~7;
      const size_t msm = (size_t)(K8 + 4) * K8 * 2 * sizeof(double);
      if (msm > 48 * 1024) {
        if (net->cplx) CUDA_CHECK(cudaFuncSetAttribute(k_modeprod_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
        else CUDA_CHECK(cudaFuncSetAttribute(k_modeprod_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
      }
      unsigned gym = (unsigned)std::max<long long>(1, std::min<long long>(maxn / std::max(maxK, 1) / 128 + 1, 148));
      while (gym > 1 && (unsigned long long)gym * nb > 148ull * 32ull) gym = (gym + 1) / 2;
      dim3 gridm(nb, gym);
      for (int s = 0; s < maxsteps; ++s) {
        if (net->cplx) k_modeprod_mma<true><<<gridm, kThreads, msm, ctx->stream>>>(dj, s);
        else k_modeprod_mma<false><<<gridm, kThreads, msm, ctx->stream>>>(dj, s);
        ITN_LAUNCH_CHECK(ctx);
      }
      size_t gsm = 0;
      for (const VJob& J : batch) {
        const int No16 = (J.No + 15) & ~15;
        const int T = (No16 / 16) * (No16 / 16), S = T >= kThreads / 32 ? 1 : (kThreads / 32) / T;
        gsm = std::max(gsm, (size_t)S * 2 * No16 * No16 * sizeof(double));
      }
      if (gsm > 48 * 1024) {
        if (net->cplx) CUDA_CHECK(cudaFuncSetAttribute(k_gram_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
        else CUDA_CHECK(cudaFuncSetAttribute(k_gram_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
      }
      if (net->cplx) k_gram_mma<true><<<nb, kThreads, gsm, ctx->stream>>>(dj);
      else k_gram_mma<false><<<nb, kThreads, gsm, ctx->stream>>>(dj);
      ITN_LAUNCH_CHECK(ctx);
      lo = hi;
      continue;
    }
    int smem_elems = maxkn;
    size_t smem_bytes = (size_t)smem_elems * P * sizeof(double);
    if (smem_bytes > 96 * 1024) {  // too large to stage: kernels read the matrix from global memory
      smem_elems = 0;
      smem_bytes = 0;
    }
    if (smem_bytes > 48 * 1024) {
      if (net->cplx) CUDA_CHECK(cudaFuncSetAttribute(k_modeprod<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      else CUDA_CHECK(cudaFuncSetAttribute(k_modeprod<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    }
    for (int s = 0; s < maxsteps; ++s) {
      if (net->cplx) k_modeprod<true><<<grid, kThreads, smem_bytes, ctx->stream>>>(dj, s, smem_elems);
      else k_modeprod<false><<<grid, kThreads, smem_bytes, ctx->stream>>>(dj, s, smem_elems);
      ITN_LAUNCH_CHECK(ctx);
    }
    if (net->cplx) k_gram<true><<<nb, kThreads, 0, ctx->stream>>>(dj);
    else k_gram<false><<<nb, kThreads, 0, ctx->stream>>>(dj);
    ITN_LAUNCH_CHECK(ctx);
    lo = hi;
  }
}

// Mode products on tensors in canonical order (no permutation, no close): dst = src x_{mode_1} M_1 x_{mode_2} M_2 ...
// Used for the environment-support projectors of the simple update (itn_linalg.cu).
void itn_run_modeprods(itn_ctx* ctx, bool cplx, const std::vector<ModeProdSpec>& specs, std::vector<const double*>& result) {
  result.assign(specs.size(), nullptr);
  if (specs.empty()) return;
  const int P = cplx ? 2 : 1;
  std::vector<VJob> jobs(specs.size());
  long long maxn = 0;
  int maxsteps = 0, maxkn = 0;
  for (size_t j = 0; j < specs.size(); ++j) {
    const ModeProdSpec& sp = specs[j];
    VJob J;
    memset(&J, 0, sizeof(J));
    J.a = sp.src;
    J.ap = const_cast<double*>(sp.src);
    J.identity_perm = 1;
    J.n = sp.n;
    J.nm = sp.nm;
    J.w[0] = sp.w0;
    J.w[1] = sp.w1;
    for (int m = 0; m < sp.nm; ++m) J.dims[m] = sp.dims[m];
    for (int t = 0; t < sp.nsteps; ++t) {
      const int m = sp.mode[t];
      long long L = 1;
      for (int q = 0; q < m; ++q) L *= sp.dims[q];
      ModeStep& S = J.steps[J.nsteps++];
      S.L = L;
      S.K = S.N = sp.dims[m];
      S.R = sp.n / (L * sp.dims[m]);
      S.m = sp.mat[t];
      S.mplane = (long long)sp.dims[m] * sp.dims[m];
      S.trans = sp.trans[t] ? 1 : 0;
      S.conj = 0;
      maxkn = std::max(maxkn, S.K * S.N);
    }
    maxn = std::max(maxn, J.n);
    maxsteps = std::max(maxsteps, J.nsteps);
    result[j] = J.nsteps == 0 ? sp.src : J.w[(J.nsteps - 1) & 1];
    jobs[j] = J;
  }
  DevBuf jb(ctx, jobs.size() * sizeof(VJob));
  const VJob* dj = itn_upload(ctx, jobs, jb);
  unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn + kThreads * 4 - 1) / (kThreads * 4), 64));
  while (gy > 1 && (unsigned long long)gy * jobs.size() > 148ull * 64ull) gy = (gy + 1) / 2;
  dim3 grid((unsigned)jobs.size(), gy);
  int smem_elems = maxkn;
  size_t smem_bytes = (size_t)smem_elems * P * sizeof(double);
  if (smem_bytes > 96 * 1024) {
    smem_elems = 0;
    smem_bytes = 0;
  }
  if (smem_bytes > 48 * 1024) {
    if (cplx) CUDA_CHECK(cudaFuncSetAttribute(k_modeprod<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    else CUDA_CHECK(cudaFuncSetAttribute(k_modeprod<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  }
  for (int t = 0; t < maxsteps; ++t) {
    if (cplx) k_modeprod<true><<<grid, kThreads, smem_bytes, ctx->stream>>>(dj, t, smem_elems);
    else k_modeprod<false><<<grid, kThreads, smem_bytes, ctx->stream>>>(dj, t, smem_elems);
    ITN_LAUNCH_CHECK(ctx);
  }
}

// ------------------------------------------------------------------------------------------------
// vertex-level synchronous sweeps (k_mp_mma / k_gr_mma)
// ------------------------------------------------------------------------------------------------
bool itn_vertex_sweep_ok(const itn_net* net, int v) {
  if (net->ctx->path_mode == 2) return false;
  const int z = (int)net->inc[v].size();
  if (z < 2) return false;
  for (int e : net->inc[v])
    if (net->edim[e] > 64) return false;
  return true;
}

void itn_run_vertex_sweeps(itn_net* net, const std::vector<SweepSpec>& specs) {
  if (specs.empty()) return;
  itn_ctx* ctx = net->ctx;
  const int P = net->planes();
  size_t lo = 0;
  while (lo < specs.size()) {
    // batch bounded by the scratch budget: a vertex of degree z needs at most z temporaries
    size_t hi = lo, bytes = 0;
    while (hi < specs.size()) {
      const int v = specs[hi].v;
      const size_t need = (size_t)net->T[v].n * P * sizeof(double) * net->inc[v].size();
      if (hi > lo && bytes + need > ctx->ws_budget) break;
      bytes += need;
      ++hi;
    }
    DevBuf ws(ctx, std::max<size_t>(bytes, 16));
    // rounds[i]: the i-th operation of every vertex of the batch (operations of one vertex run in order)
    std::vector<std::vector<MpOp>> mp_rounds;
    std::vector<std::vector<GrOp>> gr_rounds;
    int maxK = 1;
    long long maxn = 0;
    size_t woff = 0;
    for (size_t si = lo; si < hi; ++si) {
      const SweepSpec& sp = specs[si];
      const int v = sp.v;
      const int z = (int)net->inc[v].size();
      ITN_REQUIRE(net->T[v].p != nullptr, ITN_EINVAL, "site tensor of vertex " + std::to_string(v) + " is not set");
      const long long n = net->T[v].n;
      maxn = std::max(maxn, n);
      std::vector<long long> Ls(z), Rs(z);
      std::vector<int> chis(z);
      std::vector<const double*> msg(z);
      long long acc = net->sdim[v];
      for (int k = 0; k < z; ++k) {
        const int e = net->inc[v][k];
        chis[k] = net->edim[e];
        Ls[k] = acc;
        acc *= chis[k];
        maxK = std::max(maxK, chis[k]);
        const DevTensor& m = net->M[net->msg_into(v, e)];
        ITN_REQUIRE(m.p != nullptr, ITN_EINVAL,
                    "message into vertex " + std::to_string(v) + " on edge " + std::to_string(e) + " is not set");
        msg[k] = m.p;
      }
      for (int k = 0; k < z; ++k) Rs[k] = n / (Ls[k] * chis[k]);
      std::vector<double*> slots(z);
      for (int k = 0; k < z; ++k) {
        slots[k] = (double*)(ws.as<char>() + woff);
        woff += (size_t)n * P * sizeof(double);
      }
      int round = 0, top = 0;
      auto put_mp = [&](const MpOp& op) {
        if ((int)mp_rounds.size() <= round) mp_rounds.resize(round + 1), gr_rounds.resize(round + 1);
        mp_rounds[round++].push_back(op);
      };
      auto put_gr = [&](const GrOp& op) {
        if ((int)gr_rounds.size() <= round) mp_rounds.resize(round + 1), gr_rounds.resize(round + 1);
        gr_rounds[round++].push_back(op);
      };
      // solve(T, S): T has every bond outside S absorbed; emits the outgoing message of every bond in S
      std::function<void(const double*, int, int)> solve = [&](const double* Tn, int s_lo, int s_hi) {
        if (s_hi - s_lo == 1) {
          const int k = s_lo;
          put_gr({Tn, net->bra(v), sp.out[k], Ls[k], Rs[k], chis[k]});
          return;
        }
        const int mid = (s_lo + s_hi) / 2;
        const int saved = top;
        auto absorb = [&](int a_lo, int a_hi) {
          const double* cur = Tn;
          for (int k = a_lo; k < a_hi; ++k) {
            double* dst = slots[top++];
            put_mp({cur, dst, msg[k], Ls[k], Rs[k], chis[k]});
            cur = dst;
          }
          return cur;
        };
        const double* T1 = absorb(mid, s_hi);  // absorb the upper half, emit the lower half
        solve(T1, s_lo, mid);
        top = saved;
        const double* T2 = absorb(s_lo, mid);
        solve(T2, mid, s_hi);
        top = saved;
      };
      solve(net->T[v].p, 0, z);
    }
    const int K8 = (maxK + 7) & ~7;
    const size_t msm = (size_t)(K8 + 4) * K8 * 2 * sizeof(double);
    const int No16 = (maxK + 15) & ~15;
    const int Tt = (No16 / 16) * (No16 / 16), Ss = Tt >= kThreads / 32 ? 1 : (kThreads / 32) / Tt;
    size_t gsm = (size_t)Ss * 2 * No16 * No16 * sizeof(double);
    gsm = std::max(gsm, (size_t)(kThreads / 32) * 2 * 256 * sizeof(double));
    if (net->cplx) {
      if (msm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_mp_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
      if (gsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_gr_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    } else {
      if (msm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_mp_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
      if (gsm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_gr_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    }
    // rounds with few closes (small graphs: heavy-hex has 127 vertices) split the closed index range of every close
    // over Y CTAs so that the launch covers the SMs; partial results are summed in a fixed order
    std::vector<int> ysplit(gr_rounds.size(), 1);
    size_t part_doubles = 0;
    for (size_t r = 0; r < gr_rounds.size(); ++r) {
      const size_t ng = gr_rounds[r].size();
      if (ng == 0 || ng >= 148) continue;
      long long xmin = LLONG_MAX;
      for (const GrOp& op : gr_rounds[r]) xmin = std::min(xmin, op.L * op.R);
      const int y = (int)std::min<long long>(std::min<long long>(8, 296 / (long long)ng), xmin / 512);
      if (y < 2) continue;
      ysplit[r] = y;
      for (const GrOp& op : gr_rounds[r]) part_doubles += (size_t)y * P * op.chi * op.chi;
    }
    DevBuf partb(ctx, std::max<size_t>(part_doubles, 1) * sizeof(double));
    {
      double* pp = partb.as<double>();
      for (size_t r = 0; r < gr_rounds.size(); ++r)
        for (GrOp& op : gr_rounds[r]) {
          op.part = nullptr;
          if (ysplit[r] > 1) {
            op.part = pp;
            pp += (size_t)ysplit[r] * P * op.chi * op.chi;
          }
        }
    }
    // one upload for all operation tables
    size_t nmp = 0, ngr = 0;
    for (auto& r : mp_rounds) nmp += r.size();
    for (auto& r : gr_rounds) ngr += r.size();
    std::vector<MpOp> all_mp;
    std::vector<GrOp> all_gr;
    all_mp.reserve(nmp);
    all_gr.reserve(ngr);
    for (auto& r : mp_rounds) all_mp.insert(all_mp.end(), r.begin(), r.end());
    for (auto& r : gr_rounds) all_gr.insert(all_gr.end(), r.begin(), r.end());
    DevBuf mb(ctx, std::max<size_t>(nmp, 1) * sizeof(MpOp)), gb(ctx, std::max<size_t>(ngr, 1) * sizeof(GrOp));
    const MpOp* dmp = itn_upload(ctx, all_mp, mb);
    const GrOp* dgr = itn_upload(ctx, all_gr, gb);
    size_t mo = 0, go = 0;
    for (size_t r = 0; r < mp_rounds.size(); ++r) {
      const unsigned nm = (unsigned)mp_rounds[r].size(), ng = (unsigned)gr_rounds[r].size();
      if (nm) {
        unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>(maxn / std::max(maxK, 1) / 128 + 1, 148));
        while (gy > 1 && (unsigned long long)gy * nm > 148ull * 32ull) gy = (gy + 1) / 2;
        if (net->cplx) k_mp_mma<true><<<dim3(nm, gy), kThreads, msm, ctx->stream>>>(dmp + mo);
        else k_mp_mma<false><<<dim3(nm, gy), kThreads, msm, ctx->stream>>>(dmp + mo);
        ITN_LAUNCH_CHECK(ctx);
        mo += nm;
      }
      if (ng) {
        const unsigned Y = (unsigned)ysplit[r];
        if (net->cplx) k_gr_mma<true><<<dim3(ng, Y), kThreads, gsm, ctx->stream>>>(dgr + go);
        else k_gr_mma<false><<<dim3(ng, Y), kThreads, gsm, ctx->stream>>>(dgr + go);
        ITN_LAUNCH_CHECK(ctx);
        if (Y > 1) {
          if (net->cplx) k_gr_reduce<true><<<ng, 256, 0, ctx->stream>>>(dgr + go, (int)Y);
          else k_gr_reduce<false><<<ng, 256, 0, ctx->stream>>>(dgr + go, (int)Y);
          ITN_LAUNCH_CHECK(ctx);
        }
        go += ng;
      }
    }
    lo = hi;
  }
}
