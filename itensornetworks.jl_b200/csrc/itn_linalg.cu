// Batched small dense linear algebra and the simple-update gate (apply.jl:33-95) on the device.
//
// simple_update_bp, restated so that the big site tensors are touched only twice per side
// (one environment contraction, one rebuild) instead of absorb / QR / un-absorb:
//
//   reference (apply.jl:40-93)                         this engine
//   --------------------------------------------       ---------------------------------------------
//   S_j = sqrt(E_j), W_j = E_j^-1/2 (eigen)             E_j <- (E_j + E_j^H)/2          (k_hermitize)
//   A~ = A x_j S_j                                     C = A~^T conj(A~) = bond environment of (s, l) with the
//   Q, R = qr(A~)  (rows: outer bonds, cols: (s,l))        messages E_j absorbed            (generic kernels)
//                                                      C = V L V^H (Jacobi); R = L^1/2 V^T, R^+ = conj(V) L^-1/2
//   theta = R1 R2 ; theta' = gate theta                 same, on the r x (s,l) factors     (k_su_theta)
//   U S V = svd(theta'), truncate (maxdim, cutoff)      one-sided Jacobi SVD + ITensors truncation rule
//   R1' = U sqrt(S), R2' = sqrt(S) V                    same
//   A' = (Q x_j conj(W_j)) R' = A x_j (S_j W_j^H) R^+ R'   A' = A . (R^+ R')   on the fused (s, l) index
//
// S_j W_j^H = P_j is the projector onto the support of E_j (eigenvalues below the reference's eigen cutoff,
// 10 eps relative, are dropped): the identity for full-rank messages; otherwise A is replaced by A x_j P_j
// before the rebuild (k_eig_fn with f = 1 builds P_j, itn_run_modeprods applies it).  Q is never formed: with R R^+ = 1 the product Q R' equals A~ R^+ R', and the
// sqrt / inverse-sqrt gauges cancel.  Singular values, truncation error, kept dimension and the
// contracted pair A1'.A2' are identical to the reference's in exact arithmetic (they do not depend on
// the orthonormal basis chosen for the r index).
//
// map_eigvals (apply.jl:21-25) runs on the same Jacobi kernel.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

#include "itn_internal.h"

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// one-sided (Hestenes) Jacobi SVD, one CTA per matrix, planar storage, column-major
//   in : A (m x n)                     out: us = A V = U diag(sigma) (columns not normalised), V (n x n, optional),
//                                           sigma[n] sorted descending, perm[n] (sorted rank -> column)
// A group of LP lanes owns one column pair of a round (32 / LP pairs per warp), so that a 64 x 64 matrix keeps
// 16 warps busy and three matrices share an SM.  Column norms are recomputed exactly at the start of every sweep
// and at the end (singular values); inside a sweep they follow the rotation identities a' = a - t|g|, b' = b + t|g|.
// ------------------------------------------------------------------------------------------------
struct SvdJob {
  double* a;       // planar m x n input (im plane at +m*n); overwritten with U*Sigma unless `us` is set
  double* v;       // planar n x n (im plane at +n*n), or null: V is not accumulated
  double* sigma;   // n
  int* perm;       // n
  int m, n;
  double* us;      // optional separate output for U*Sigma (a stays intact)
  const int* skip; // optional device flag: non-zero -> this matrix needs no decomposition
};

constexpr int kSvdMaxThreads = 1024;
constexpr size_t kGlueSmemMax = 200 * 1024;  // dynamic shared memory the staged glue kernels may ask for
constexpr int kSvdMaxSweeps = 30;

template <int LP>
__device__ __forceinline__ double gsum(double v, unsigned mask) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

template <bool C, int LP, bool CACHE>
__global__ void __launch_bounds__(CACHE ? 512 : kSvdMaxThreads) k_jacobi_svd(const SvdJob* __restrict__ jobs, int smem_doubles) {
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ double s_norm[256];  // squared column norms
  __shared__ double s_red[33];
  const SvdJob J = jobs[blockIdx.x];
  if (J.skip && *J.skip) return;
  const int m = J.m, n = J.n;
  if (n == 0 || m == 0) return;
  const long long mn = (long long)m * n, nn = (long long)n * n;
  const int P = C ? 2 : 1;
  const bool has_v = J.v != nullptr;
  double* dst = J.us ? J.us : J.a;
  const bool in_smem = (mn + (has_v ? nn : 0)) * P <= smem_doubles;
  double* Ar = in_smem ? sm : dst;
  double* Ai = C ? Ar + mn : nullptr;
  double* Vr = has_v ? (in_smem ? sm + P * mn : J.v) : nullptr;
  double* Vi = (C && has_v) ? Vr + nn : nullptr;
  const int nthreads = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  constexpr int PW = 32 / LP;
  const int sub = lane / LP, sl = lane % LP;
  const unsigned gmask = (LP == 32) ? 0xffffffffu : (((1u << LP) - 1u) << (sub * LP));
  if (in_smem || dst != J.a)
    for (long long i = tid; i < mn * P; i += nthreads) Ar[i] = J.a[i];
  if (has_v)
    for (long long i = tid; i < nn; i += nthreads) {
      Vr[i] = (i % n == i / n) ? 1.0 : 0.0;
      if (C) Vi[i] = 0.0;
    }
  __syncthreads();
  auto column_norms = [&]() {
    for (int j = warp * PW + sub; j < n; j += nwarps * PW) {
      double a2 = 0.0;
      for (int i = sl; i < m; i += LP) {
        const double r = Ar[(long long)j * m + i], im = C ? Ai[(long long)j * m + i] : 0.0;
        a2 += r * r + im * im;
      }
      a2 = gsum<LP>(a2, gmask);
      if (sl == 0) s_norm[j] = a2;
    }
  };
  column_norms();
  __syncthreads();
  // columns whose norm is below ~1e-19 ||A||_F are numerically null: rotating them only amplifies underflow
  // noise (they appear whenever rank(A) < n, e.g. wide matrices and rank-deficient Gram matrices)
  if (tid == 0) {
    double t = 0.0;
    for (int j = 0; j < n; ++j) t += s_norm[j];
    s_red[32] = t * (2.220446049250313e-19 * 2.220446049250313e-19);
  }
  __syncthreads();
  const double tiny = s_red[32];
  const int ne = n + (n & 1);  // even number of players; index n (if any) is a bye
  // rotate while |<a_p, a_q>| > tol |a_p| |a_q|; sqrt(m) eps is the rounding level of the inner product (as in LAPACK xGESVJ)
  const double tol = sqrt((double)m) * 2.220446049250313e-16;
  for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    if (sweep > 0) column_norms();
    __syncthreads();
    for (int round = 0; round < ne - 1; ++round) {
      for (int k = warp * PW + sub; k < ne / 2; k += nwarps * PW) {
        int p, q;
        if (k == 0) {
          p = ne - 1;
          q = round;
        } else {
          p = (round + k) % (ne - 1);
          q = (round - k + (ne - 1)) % (ne - 1);
        }
        if (p >= n || q >= n) continue;
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        const double alpha = s_norm[p], beta = s_norm[q];
        if (alpha <= tiny || beta <= tiny) continue;
        double* apr = Ar + (long long)p * m;
        double* aqr = Ar + (long long)q * m;
        double* api = C ? Ai + (long long)p * m : nullptr;
        double* aqi = C ? Ai + (long long)q * m : nullptr;
        double gr = 0, gi = 0;
        double rpr[4], rpi[4], rqr[4], rqi[4];
        if (CACHE) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = sl + LP * j;
            const bool ok = i < m;
            rpr[j] = ok ? apr[i] : 0.0;
            rqr[j] = ok ? aqr[i] : 0.0;
            rpi[j] = (C && ok) ? api[i] : 0.0;
            rqi[j] = (C && ok) ? aqi[i] : 0.0;
            gr += rpr[j] * rqr[j] + rpi[j] * rqi[j];  // conj(a_p) . a_q
            gi += rpr[j] * rqi[j] - rpi[j] * rqr[j];
          }
        } else {
          for (int i = sl; i < m; i += LP) {
            const double pr = apr[i], qr = aqr[i];
            const double pi = C ? api[i] : 0.0, qi = C ? aqi[i] : 0.0;
            gr += pr * qr + pi * qi;  // conj(a_p) . a_q
            gi += pr * qi - pi * qr;
          }
        }
        gr = gsum<LP>(gr, gmask);
        gi = C ? gsum<LP>(gi, gmask) : 0.0;
        const double g2 = gr * gr + gi * gi;
        if (g2 == 0.0 || !(g2 > tol * tol * alpha * beta)) continue;
        const double ginv = rsqrt(g2);
        const double gabs = g2 * ginv;
        // phase e^{-i phi} applied to column q so that the inner product becomes real positive
        const double er = gr * ginv, ei = -gi * ginv;
        const double zeta = (beta - alpha) * 0.5 * ginv;
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = rsqrt(1.0 + t * t), s = c * t;
        if (CACHE) {
          // m <= 4 LP: the rows of this lane group are still in registers from the inner product
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = sl + LP * j;
            if (i < m) {
              const double pr = rpr[j], qr0 = rqr[j];
              const double pi = C ? rpi[j] : 0.0, qi0 = C ? rqi[j] : 0.0;
              const double qr = qr0 * er - qi0 * ei, qi = qr0 * ei + qi0 * er;
              apr[i] = c * pr - s * qr;
              aqr[i] = s * pr + c * qr;
              if (C) {
                api[i] = c * pi - s * qi;
                aqi[i] = s * pi + c * qi;
              }
            }
          }
        } else {
          for (int i = sl; i < m; i += LP) {
            const double pr = apr[i], qr0 = aqr[i];
            const double pi = C ? api[i] : 0.0, qi0 = C ? aqi[i] : 0.0;
            const double qr = qr0 * er - qi0 * ei, qi = qr0 * ei + qi0 * er;
            apr[i] = c * pr - s * qr;
            aqr[i] = s * pr + c * qr;
            if (C) {
              api[i] = c * pi - s * qi;
              aqi[i] = s * pi + c * qi;
            }
          }
        }
        if (has_v) {
          double* vpr = Vr + (long long)p * n;
          double* vqr = Vr + (long long)q * n;
          double* vpi = C ? Vi + (long long)p * n : nullptr;
          double* vqi = C ? Vi + (long long)q * n : nullptr;
          for (int i = sl; i < n; i += LP) {
            const double pr = vpr[i], qr0 = vqr[i];
            const double pi = C ? vpi[i] : 0.0, qi0 = C ? vqi[i] : 0.0;
            const double qr = qr0 * er - qi0 * ei, qi = qr0 * ei + qi0 * er;
            vpr[i] = c * pr - s * qr;
            vqr[i] = s * pr + c * qr;
            if (C) {
              vpi[i] = c * pi - s * qi;
              vqi[i] = s * pi + c * qi;
            }
          }
        }
        if (sl == 0) {
          s_norm[p] = fmax(alpha - t * gabs, 0.0);
          s_norm[q] = beta + t * gabs;
          // rotations at the rounding level of the inner product are applied but do not keep the iteration alive
          if (g2 > 64.0 * tol * tol * alpha * beta) s_rot = 1;
        }
      }
      __syncthreads();
    }
    const int any = s_rot;
    __syncthreads();
    if (!any) break;
  }
  // singular values = exact column norms; stable descending order
  column_norms();
  __syncthreads();
  for (int j = tid; j < n; j += nthreads) {
    const double sj = s_norm[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const double sk = s_norm[k];
      rank += (sk > sj) || (sk == sj && k < j);
    }
    J.sigma[rank] = sqrt(sj);
    J.perm[rank] = j;
  }
  if (in_smem) {
    for (long long i = tid; i < mn * P; i += nthreads) dst[i] = sm[i];
    if (has_v)
      for (long long i = tid; i < nn * P; i += nthreads) J.v[i] = sm[P * mn + i];
  }
}

template <bool C, int LP, bool CACHE>
void launch_jacobi(itn_ctx* ctx, const SvdJob* dj, unsigned njobs, int pairs, size_t smem) {
  constexpr int PW = 32 / LP;
  const int warps = std::min(CACHE ? 16 : 32, std::max(1, (pairs + PW - 1) / PW));
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_svd<C, LP, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_svd<C, LP, CACHE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  k_jacobi_svd<C, LP, CACHE><<<njobs, warps * 32, smem, ctx->stream>>>(dj, (int)(smem / sizeof(double)));
  ITN_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// k_jacobi64: the same one-sided Jacobi iteration (same pair ordering, thresholds and norm updates as k_jacobi_svd)
// specialised for the bond matrix of a gate: m, n <= 64, no V.  The matrix sits in shared memory as 64 x n with the rows
// zero-padded to 64 (a zero row changes no inner product).  Eight lanes own one column pair of a round (four pairs per
// warp, eight warps for 32 pairs); lane sl holds rows {2 sl, 2 sl + 1} + 16 j, j = 0..3, of both columns in registers
// from the inner product to the rotation, so a round moves every column through shared memory exactly once in each
// direction with 16-byte accesses (each 8-lane group touches 128 contiguous bytes: conflict free).  Control flow is
// uniform inside a warp up to the shuffles (full-mask butterflies over offsets 4, 2, 1 stay inside a group); pair
// indices need no division; shared memory is addressed directly (no generic loads).
// ------------------------------------------------------------------------------------------------
template <bool C>
__global__ void __launch_bounds__(256, 2) k_jacobi64(const SvdJob* __restrict__ jobs) {
  extern __shared__ double sm[];  // [planes][n cols][64 rows]
  __shared__ double s_norm[64];
  __shared__ double s_tiny;
  __shared__ int s_rot;
  const SvdJob J = jobs[blockIdx.x];
  if (J.skip && *J.skip) return;
  const int m = J.m, n = J.n;
  if (n == 0 || m == 0) return;
  double* dst = J.us ? J.us : J.a;
  const int nthreads = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const int sub = lane >> 3, sl = lane & 7;
  double* Ar = sm;
  double* Ai = sm + 64 * 64;
  {
    const double* src_r = J.a;
    const double* src_i = J.a + (size_t)m * n;
    for (int idx = tid; idx < n * 64; idx += nthreads) {
      const int j = idx >> 6, i = idx & 63;
      const bool ok = i < m;
      Ar[idx] = ok ? src_r[(size_t)j * m + i] : 0.0;
      if (C) Ai[idx] = ok ? src_i[(size_t)j * m + i] : 0.0;
    }
  }
  __syncthreads();
  auto column_norms = [&]() {
    for (int jb = warp * 4; jb < n; jb += nwarps * 4) {  // uniform trip count inside a warp
      const int j = jb + sub;
      double a2 = 0.0;
      if (j < n) {
        const double2* cr = reinterpret_cast<const double2*>(Ar + j * 64) + sl;
        const double2* ci = reinterpret_cast<const double2*>(Ai + j * 64) + sl;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const double2 r = cr[8 * t];
          a2 += r.x * r.x + r.y * r.y;
          if (C) {
            const double2 im = ci[8 * t];
            a2 += im.x * im.x + im.y * im.y;
          }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      if (j < n && sl == 0) s_norm[j] = a2;
    }
  };
  column_norms();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int j = 0; j < n; ++j) t += s_norm[j];
    s_tiny = t * (2.220446049250313e-19 * 2.220446049250313e-19);
  }
  __syncthreads();
  const double tiny = s_tiny;
  const int ne = n + (n & 1), npairs = ne >> 1;
  const int k = warp * 4 + sub;  // pair slot of this lane group (the launch provides >= npairs groups)
  const double tol = sqrt((double)m) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    if (sweep > 0) column_norms();
    __syncthreads();
    for (int round = 0; round < ne - 1; ++round) {
      int p, q;
      if (k == 0) {
        p = ne - 1;
        q = round;
      } else {
        p = round + k;
        if (p >= ne - 1) p -= ne - 1;
        q = round - k;
        if (q < 0) q += ne - 1;
      }
      const bool valid = k < npairs && p < n && q < n;
      if (!valid) p = q = 0;  // addresses stay in range; nothing is loaded or stored for a bye
      if (p > q) {
        const int t = p;
        p = q;
        q = t;
      }
      const double alpha = s_norm[p], beta = s_norm[q];
      double2* cpr = reinterpret_cast<double2*>(Ar + p * 64) + sl;
      double2* cqr = reinterpret_cast<double2*>(Ar + q * 64) + sl;
      double2* cpi = reinterpret_cast<double2*>(Ai + p * 64) + sl;
      double2* cqi = reinterpret_cast<double2*>(Ai + q * 64) + sl;
      double2 pr[4], qr[4], pi[4], qi[4];
      double gr = 0.0, gi = 0.0;
      if (valid) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          pr[t] = cpr[8 * t];
          qr[t] = cqr[8 * t];
          if (C) {
            pi[t] = cpi[8 * t];
            qi[t] = cqi[8 * t];
          }
        }
        double gr1 = 0.0, gi1 = 0.0;  // two accumulators per part: shorter dependent FMA chains
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // conj(a_p) . a_q
          gr = fma(pr[t].x, qr[t].x, gr);
          gr1 = fma(pr[t].y, qr[t].y, gr1);
          if (C) {
            gr = fma(pi[t].x, qi[t].x, gr);
            gr1 = fma(pi[t].y, qi[t].y, gr1);
            gi = fma(pr[t].x, qi[t].x, gi);
            gi1 = fma(pr[t].y, qi[t].y, gi1);
            gi = fma(-pi[t].x, qr[t].x, gi);
            gi1 = fma(-pi[t].y, qr[t].y, gi1);
          }
        }
        gr += gr1;
        gi += gi1;
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        gr += __shfl_xor_sync(0xffffffffu, gr, o);
        if (C) gi += __shfl_xor_sync(0xffffffffu, gi, o);
      }
      const double g2 = gr * gr + gi * gi;
      const double thr = tol2 * alpha * beta;
      if (valid && alpha > tiny && beta > tiny && g2 > thr && g2 > 0.0) {
        const double ginv = rsqrt(g2);
        const double gabs = g2 * ginv;
        // phase e^{-i phi} applied to column q so that the inner product becomes real positive
        const double er = gr * ginv, ei = -gi * ginv;
        const double zeta = (beta - alpha) * 0.5 * ginv;
        const double w = fma(zeta, zeta, 1.0);
        const double den = fabs(zeta) + w * rsqrt(w);  // >= 1
        const double rden = rsqrt(den);
        const double t = copysign(rden * rden, zeta);  // sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), sign(0) = +
        const double c = rsqrt(fma(t, t, 1.0)), s = c * t;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double2 npr, nqr, npi, nqi;
          {
            const double a = pr[u].x, ai = C ? pi[u].x : 0.0, b0 = qr[u].x, b0i = C ? qi[u].x : 0.0;
            const double b = C ? (b0 * er - b0i * ei) : b0 * er, bi = C ? (b0 * ei + b0i * er) : 0.0;
            npr.x = c * a - s * b;
            nqr.x = s * a + c * b;
            npi.x = c * ai - s * bi;
            nqi.x = s * ai + c * bi;
          }
          {
            const double a = pr[u].y, ai = C ? pi[u].y : 0.0, b0 = qr[u].y, b0i = C ? qi[u].y : 0.0;
            const double b = C ? (b0 * er - b0i * ei) : b0 * er, bi = C ? (b0 * ei + b0i * er) : 0.0;
            npr.y = c * a - s * b;
            nqr.y = s * a + c * b;
            npi.y = c * ai - s * bi;
            nqi.y = s * ai + c * bi;
          }
          cpr[8 * u] = npr;
          cqr[8 * u] = nqr;
          if (C) {
            cpi[8 * u] = npi;
            cqi[8 * u] = nqi;
          }
        }
        if (sl == 0) {
          s_norm[p] = fmax(alpha - t * gabs, 0.0);
          s_norm[q] = beta + t * gabs;
          // rotations at the rounding level of the inner product are applied but do not keep the iteration alive
          if (g2 > 64.0 * thr) s_rot = 1;
        }
      }
      __syncthreads();
    }
    const int any = s_rot;
    __syncthreads();
    if (!any) break;
  }
  // singular values = exact column norms; stable descending order
  column_norms();
  __syncthreads();
  for (int j = tid; j < n; j += nthreads) {
    const double sj = s_norm[j];
    int rank = 0;
    for (int kk = 0; kk < n; ++kk) {
      const double sk = s_norm[kk];
      rank += (sk > sj) || (sk == sj && kk < j);
    }
    J.sigma[rank] = sqrt(sj);
    J.perm[rank] = j;
  }
  {
    double* dst_r = dst;
    double* dst_i = dst + (size_t)m * n;
    for (int idx = tid; idx < n * 64; idx += nthreads) {
      const int j = idx >> 6, i = idx & 63;
      if (i < m) {
        dst_r[(size_t)j * m + i] = Ar[idx];
        if (C) dst_i[(size_t)j * m + i] = Ai[idx];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_jacobi64oe: k_jacobi64 with the odd-even transposition ordering (Brent-Luk): the columns sit on a line of n
// positions, steps alternate between the pairs (2k, 2k+1) and (2k+1, 2k+2), and the two columns of a pair exchange
// positions after their rotation, so that n consecutive steps bring every pair of columns together exactly once.
// Lane group k keeps the column at position 2k+1 in REGISTERS for the whole decomposition: in every step it loads only
// its partner (position 2k or 2k+2) from shared memory, and after the rotation it keeps the rotated partner (which now
// sits at position 2k+1) and stores its own rotated column into the partner's slot.  Shared memory therefore holds the
// even positions only (32 KB for 64 ComplexF64 columns) and a step moves ONE column per pair in each direction, half
// the traffic of k_jacobi64, whose ncu profile is shared-memory bound (profiles/r1r_k_jacobi64_ncu.txt).
// Thresholds, rotation formulas and norm updates are those of k_jacobi_svd / k_jacobi64.
// ------------------------------------------------------------------------------------------------
// OCC = resident CTAs per SM asked for: 2 keeps the partner column in registers between the inner product and the
// rotation (128 registers); 3 reads it again from shared memory (<= 85 registers, one more matrix in flight per SM).
// M = padded row count (64 or 128), LP = M / 8 lanes per column (each lane holds 8 rows), NMAX = maximal column count
// (64 or 128): the gate's bond matrix is 64 x 64 on the chi = 16 square lattice and 128 x 64 / 64 x 128 on heavy-hex chi = 32.
template <bool C, int OCC, int M, int LP, int NMAX>
__global__ void __launch_bounds__(NMAX / 2 * LP, OCC) k_jacobi64oe(const SvdJob* __restrict__ jobs) {
  static_assert(M == 8 * LP && (LP == 8 || LP == 16), "each lane holds four double2 per plane");
  constexpr int PW = 32 / LP;     // lane groups per warp
  constexpr int NS = NMAX / 2;    // shared-memory slots
  extern __shared__ double sm[];  // [planes][NS slots][M rows]: the columns at the even positions
  __shared__ double s_norm[NMAX]; // squared column norms by position
  __shared__ double s_tiny;
  __shared__ int s_rot;
  const SvdJob J = jobs[blockIdx.x];
  if (J.skip && *J.skip) return;
  const int m = J.m, n = J.n;
  if (n == 0 || m == 0) return;
  double* dst = J.us ? J.us : J.a;
  const int nthreads = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LP, sl = lane % LP;
  const int k = warp * PW + sub;        // lane group: owns position 2k+1 (registers) and loads / stores slots k, k+1
  const bool has_reg = 2 * k + 1 < n;  // this group holds a column
  double* Ar = sm;
  double* Ai = sm + NS * M;
  const size_t mn = (size_t)m * n;
  for (int idx = tid; idx < ((n + 1) >> 1) * M; idx += nthreads) {  // even positions -> shared memory
    const int slot = idx / M, i = idx % M;
    const bool ok = i < m;
    Ar[idx] = ok ? J.a[(size_t)(2 * slot) * m + i] : 0.0;
    if (C) Ai[idx] = ok ? J.a[mn + (size_t)(2 * slot) * m + i] : 0.0;
  }
  double2 pr[4], pi[4];  // the column at position 2k+1: rows {2 sl, 2 sl + 1} + 2 LP t
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i0 = 2 * sl + 2 * LP * t;
    const size_t base = (size_t)(2 * k + 1) * m;
    pr[t].x = (has_reg && i0 < m) ? J.a[base + i0] : 0.0;
    pr[t].y = (has_reg && i0 + 1 < m) ? J.a[base + i0 + 1] : 0.0;
    pi[t].x = (C && has_reg && i0 < m) ? J.a[mn + base + i0] : 0.0;
    pi[t].y = (C && has_reg && i0 + 1 < m) ? J.a[mn + base + i0 + 1] : 0.0;
  }
  __syncthreads();
  auto column_norms = [&]() {  // uniform control flow: every lane takes part in both reductions
    double a2 = 0.0, b2 = 0.0;
    if (2 * k < n) {
      const double2* cr = reinterpret_cast<const double2*>(Ar + k * M) + sl;
      const double2* ci = reinterpret_cast<const double2*>(Ai + k * M) + sl;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2 r = cr[LP * t];
        a2 += r.x * r.x + r.y * r.y;
        if (C) {
          const double2 im = ci[LP * t];
          a2 += im.x * im.x + im.y * im.y;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      b2 += pr[t].x * pr[t].x + pr[t].y * pr[t].y;
      if (C) b2 += pi[t].x * pi[t].x + pi[t].y * pi[t].y;
    }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) {
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if (sl == 0) {
      if (2 * k < n) s_norm[2 * k] = a2;
      if (has_reg) s_norm[2 * k + 1] = b2;
    }
  };
  column_norms();
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int j = 0; j < n; ++j) t += s_norm[j];
    s_tiny = t * (2.220446049250313e-19 * 2.220446049250313e-19);
  }
  __syncthreads();
  const double tiny = s_tiny;
  const double tol = sqrt((double)m) * 2.220446049250313e-16;
  const double tol2 = tol * tol;
  int step = 0;  // parity continues across sweeps: any n consecutive steps visit every pair once
  for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    if (sweep > 0) column_norms();
    __syncthreads();
    for (int it = 0; it < n; ++it, ++step) {
      const int slot = k + (step & 1);  // even step: pair (2k, 2k+1); odd step: pair (2k+1, 2k+2)
      const int ppos = 2 * slot;
      const bool valid = has_reg && ppos < n;
      double2* cqr = reinterpret_cast<double2*>(Ar + (valid ? slot : 0) * M) + sl;
      double2* cqi = reinterpret_cast<double2*>(Ai + (valid ? slot : 0) * M) + sl;
      double2 qr[4], qi[4];
      double gr = 0.0, gi = 0.0;
      if (valid) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          qr[t] = cqr[LP * t];
          if (C) qi[t] = cqi[LP * t];
        }
        double gr1 = 0.0, gi1 = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {  // conj(a_p) . a_q
          gr = fma(pr[t].x, qr[t].x, gr);
          gr1 = fma(pr[t].y, qr[t].y, gr1);
          if (C) {
            gr = fma(pi[t].x, qi[t].x, gr);
            gr1 = fma(pi[t].y, qi[t].y, gr1);
            gi = fma(pr[t].x, qi[t].x, gi);
            gi1 = fma(pr[t].y, qi[t].y, gi1);
            gi = fma(-pi[t].x, qr[t].x, gi);
            gi1 = fma(-pi[t].y, qr[t].y, gi1);
          }
        }
        gr += gr1;
        gi += gi1;
      }
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) {
        gr += __shfl_xor_sync(0xffffffffu, gr, o);
        if (C) gi += __shfl_xor_sync(0xffffffffu, gi, o);
      }
      if (valid) {
        const double alpha = s_norm[2 * k + 1], beta = s_norm[ppos];  // p = the register column, q = the partner
        const double g2 = gr * gr + gi * gi;
        const double thr = tol2 * alpha * beta;
        double na = alpha, nb = beta;
        if (alpha > tiny && beta > tiny && g2 > thr && g2 > 0.0) {
          // Rotation angle without chained reciprocal square roots: with delta = beta - alpha and
          // h = (4 |g|^2 + delta^2)^-1/2:  cos(2 theta) = |delta| h,  sin(2 theta) = 2 |g| h, hence
          // c^2 = (1 + |delta| h) / 2,  s = sign(delta) |g| h / c,  t |g| = sign(delta) |g|^2 h / c^2
          // (the same smaller-angle rotation as t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = delta / (2 |g|));
          // rsqrt(g2) and h are independent, only rsqrt(c^2) waits for h.
          const double delta = beta - alpha;
          const double ginv = rsqrt(g2);
          const double h = rsqrt(fma(delta, delta, 4.0 * g2));
          const double er = gr * ginv, ei = -gi * ginv;
          const double cc = fma(0.5 * fabs(delta), h, 0.5);
          const double rc = rsqrt(cc);
          const double c = cc * rc;
          const double s = copysign(g2 * ginv * h * rc, delta);
          const double tg = copysign(g2 * h * rc * rc, delta);  // t |g|
          na = fmax(alpha - tg, 0.0);
          nb = beta + tg;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            double2 npr, nqr, npi, nqi;
            if (OCC > 2) {  // the partner was not kept in registers
              qr[u] = cqr[LP * u];
              if (C) qi[u] = cqi[LP * u];
            }
            {
              const double a = pr[u].x, ai = C ? pi[u].x : 0.0, b0 = qr[u].x, b0i = C ? qi[u].x : 0.0;
              const double b = C ? (b0 * er - b0i * ei) : b0 * er, bi = C ? (b0 * ei + b0i * er) : 0.0;
              npr.x = c * a - s * b;
              nqr.x = s * a + c * b;
              npi.x = c * ai - s * bi;
              nqi.x = s * ai + c * bi;
            }
            {
              const double a = pr[u].y, ai = C ? pi[u].y : 0.0, b0 = qr[u].y, b0i = C ? qi[u].y : 0.0;
              const double b = C ? (b0 * er - b0i * ei) : b0 * er, bi = C ? (b0 * ei + b0i * er) : 0.0;
              npr.y = c * a - s * b;
              nqr.y = s * a + c * b;
              npi.y = c * ai - s * bi;
              nqi.y = s * ai + c * bi;
            }
            // position exchange: the rotated register column goes to the partner's slot, the rotated partner stays
            cqr[LP * u] = npr;
            pr[u] = nqr;
            if (C) {
              cqi[LP * u] = npi;
              pi[u] = nqi;
            }
          }
          if (sl == 0 && g2 > 64.0 * thr) s_rot = 1;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // no rotation: the two columns still exchange positions
            if (OCC > 2) {
              qr[u] = cqr[LP * u];
              if (C) qi[u] = cqi[LP * u];
            }
            cqr[LP * u] = pr[u];
            pr[u] = qr[u];
            if (C) {
              cqi[LP * u] = pi[u];
              pi[u] = qi[u];
            }
          }
        }
        if (sl == 0) {
          s_norm[ppos] = na;       // the former register column now sits in the slot
          s_norm[2 * k + 1] = nb;  // the former partner is the register column
        }
      }
      __syncthreads();
    }
    const int any = s_rot;
    __syncthreads();
    if (!any) break;
  }
  column_norms();
  __syncthreads();
  for (int j = tid; j < n; j += nthreads) {
    const double sj = s_norm[j];
    int rank = 0;
    for (int kk = 0; kk < n; ++kk) {
      const double sk = s_norm[kk];
      rank += (sk > sj) || (sk == sj && kk < j);
    }
    J.sigma[rank] = sqrt(sj);
    J.perm[rank] = j;
  }
  for (int idx = tid; idx < ((n + 1) >> 1) * M; idx += nthreads) {
    const int slot = idx / M, i = idx % M;
    if (i < m) {
      dst[(size_t)(2 * slot) * m + i] = Ar[idx];
      if (C) dst[mn + (size_t)(2 * slot) * m + i] = Ai[idx];
    }
  }
  if (has_reg) {
    const size_t base = (size_t)(2 * k + 1) * m;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i0 = 2 * sl + 2 * LP * t;
      if (i0 < m) {
        dst[base + i0] = pr[t].x;
        if (C) dst[mn + base + i0] = pi[t].x;
      }
      if (i0 + 1 < m) {
        dst[base + i0 + 1] = pr[t].y;
        if (C) dst[mn + base + i0 + 1] = pi[t].y;
      }
    }
  }
}

template <bool C, int OCC, int M, int LP, int NMAX>
void launch_jacobi64oe(itn_ctx* ctx, const SvdJob* dj, unsigned njobs, int maxn) {
  constexpr int PW = 32 / LP;
  const int groups = (maxn + 1) / 2;
  const int warps = std::max(1, (groups + PW - 1) / PW);
  const size_t smem = (size_t)(C ? 2 : 1) * (NMAX / 2) * M * sizeof(double);
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi64oe<C, OCC, M, LP, NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi64oe<C, OCC, M, LP, NMAX>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  k_jacobi64oe<C, OCC, M, LP, NMAX><<<njobs, warps * 32, smem, ctx->stream>>>(dj);
  ITN_LAUNCH_CHECK(ctx);
}

template <bool C>
void launch_jacobi64(itn_ctx* ctx, const SvdJob* dj, unsigned njobs, int maxn) {
  const int npairs = (maxn + 1) / 2;
  const int warps = std::max(1, (npairs + 3) / 4);
  const size_t smem = (size_t)(C ? 2 : 1) * 64 * 64 * sizeof(double);
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi64<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaFuncSetAttribute(k_jacobi64<C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  k_jacobi64<C><<<njobs, warps * 32, smem, ctx->stream>>>(dj);
  ITN_LAUNCH_CHECK(ctx);
}

// variant: 0 = auto, 1 = generic kernel only, 2 = round-robin k_jacobi64, 3 = odd-even kernel at 3 CTAs per SM
// (a per-call argument: itn_svd_batch's second opinions; every other caller uses the default)
void run_jacobi(itn_ctx* ctx, bool cplx, const std::vector<SvdJob>& jobs, int g_jacobi_variant = 0) {
  if (jobs.empty()) return;
  int maxn = 0, maxm = 0;
  size_t need = 0;
  bool any_v = false;
  for (auto& j : jobs) {
    ITN_REQUIRE(j.n <= 256, ITN_EUNSUPPORTED, "Jacobi SVD supports at most 256 columns");
    maxn = std::max(maxn, j.n);
    maxm = std::max(maxm, j.m);
    any_v = any_v || j.v != nullptr;
    need = std::max(need, ((size_t)j.m * j.n + (j.v ? (size_t)j.n * j.n : 0)) * (cplx ? 2 : 1));
  }
  if (!any_v && maxm <= 64 && maxn <= 64 && maxn >= 2 && g_jacobi_variant != 1) {
    DevBuf jb(ctx, jobs.size() * sizeof(SvdJob));
    const SvdJob* dj = itn_upload(ctx, jobs, jb);
    if (g_jacobi_variant == 2) {
      if (cplx) launch_jacobi64<true>(ctx, dj, (unsigned)jobs.size(), maxn);
      else launch_jacobi64<false>(ctx, dj, (unsigned)jobs.size(), maxn);
    } else if (g_jacobi_variant == 3) {
      if (cplx) launch_jacobi64oe<true, 3, 64, 8, 64>(ctx, dj, (unsigned)jobs.size(), maxn);
      else launch_jacobi64oe<false, 3, 64, 8, 64>(ctx, dj, (unsigned)jobs.size(), maxn);
    } else {
      if (cplx) launch_jacobi64oe<true, 2, 64, 8, 64>(ctx, dj, (unsigned)jobs.size(), maxn);
      else launch_jacobi64oe<false, 2, 64, 8, 64>(ctx, dj, (unsigned)jobs.size(), maxn);
    }
    return;
  }
  if (!any_v && maxm <= 128 && maxn <= 128 && maxn >= 2 && g_jacobi_variant == 0) {
    // heavy-hex chi = 32 layers mix 128 x 64, 64 x 128 and smaller bond matrices: one launch per shape class, whatever
    // exceeds both instances (128 x 128) stays on the shape-generic kernel
    std::vector<SvdJob> small, tall, wide, rest;
    for (auto& j : jobs) {
      if (j.m <= 64 && j.n <= 64) small.push_back(j);
      else if (j.n <= 64) tall.push_back(j);
      else if (j.m <= 64) wide.push_back(j);
      else rest.push_back(j);
    }
    if (rest.size() != jobs.size()) {
      auto nmax = [](const std::vector<SvdJob>& v) {
        int r = 0;
        for (auto& j : v) r = std::max(r, j.n);
        return r;
      };
      if (!small.empty()) run_jacobi(ctx, cplx, small);
      if (!tall.empty()) {
        DevBuf jb(ctx, tall.size() * sizeof(SvdJob));
        const SvdJob* dj = itn_upload(ctx, tall, jb);
        if (cplx) launch_jacobi64oe<true, 1, 128, 16, 64>(ctx, dj, (unsigned)tall.size(), nmax(tall));
        else launch_jacobi64oe<false, 1, 128, 16, 64>(ctx, dj, (unsigned)tall.size(), nmax(tall));
      }
      if (!wide.empty()) {
        DevBuf jb(ctx, wide.size() * sizeof(SvdJob));
        const SvdJob* dj = itn_upload(ctx, wide, jb);
        if (cplx) launch_jacobi64oe<true, 1, 64, 8, 128>(ctx, dj, (unsigned)wide.size(), nmax(wide));
        else launch_jacobi64oe<false, 1, 64, 8, 128>(ctx, dj, (unsigned)wide.size(), nmax(wide));
      }
      if (!rest.empty()) run_jacobi(ctx, cplx, rest);
      return;
    }
  }
  size_t smem = std::min<size_t>(need * sizeof(double), 200 * 1024);
  DevBuf jb(ctx, jobs.size() * sizeof(SvdJob));
  const SvdJob* dj = itn_upload(ctx, jobs, jb);
  const int pairs = (maxn + 1) / 2;
  const unsigned nj = (unsigned)jobs.size();
  if (maxm <= 64) {
    // (CACHE = true keeps the rows of a column pair in registers between the inner product and the rotation; at 126
    //  registers per thread it halves the occupancy and is slower than re-reading shared memory, so it is not selected)
    if (cplx) launch_jacobi<true, 16, false>(ctx, dj, nj, pairs, smem);
    else launch_jacobi<false, 16, false>(ctx, dj, nj, pairs, smem);
  } else {
    if (cplx) launch_jacobi<true, 32, false>(ctx, dj, nj, pairs, smem);
    else launch_jacobi<false, 32, false>(ctx, dj, nj, pairs, smem);
  }
}

// ------------------------------------------------------------------------------------------------
// batched Cholesky of small Hermitian matrices, one CTA (n <= 64 threads, thread i = row i) per matrix:
//   K = conj(H) - shift_rel * trace(H) * 1,  H = (C + C^H) / 2;   K = L L^H
// With R = L^H this is conj(C) = R^H R, i.e. C = R^T conj(R): the same R factor (up to the unitary freedom of its row
// index) as the eigen route R = Lambda^1/2 V^T of k_su_build_R, at a fraction of the cost; R^+ = R^-1 = L^-H.
// ok = 0 when a pivot falls below n eps max_i K_ii (numerically rank deficient): the caller's eigen route takes over.
// With R == null the kernel is a definiteness test only (support test of the BP environments).
// ------------------------------------------------------------------------------------------------
struct CholJob {
  const double* c;  // planar n x n
  double* R;        // planar n x n, or null
  double* Rp;       // planar n x n
  int* ok;
  int n;
  double shift_rel;
  int* cond = nullptr;  // optional: set to 1 when the factorisation succeeded but min pivot / max pivot < kCholCondRatio
};
// Pivot ratio of the Gram matrix below which the R factor is refined by a second pass (CholeskyQR2, itn_apply2): the
// ratio bounds kappa(C) = kappa(A~)^2 from below, and the updated pair of a gate carries 30 kappa^2 eps without it.
constexpr double kCholCondRatio = 1e-4;

template <bool C>
__global__ void __launch_bounds__(64) k_chol(const CholJob* __restrict__ jobs) {
  extern __shared__ double sm[];
  __shared__ double s_piv, s_max, s_tr, s_pmin, s_pmax;
  __shared__ int s_fail;
  const CholJob J = jobs[blockIdx.x];
  const int n = J.n, n2 = n * n;
  const int i = threadIdx.x;
  double* Kr = sm;
  double* Ki = sm + n2;
  // K = conj((C + C^H) / 2), lower triangle
  if (i < n)
    for (int j = 0; j <= i; ++j) {
      Kr[i + n * j] = 0.5 * (J.c[i + n * j] + J.c[j + n * i]);
      if (C) Ki[i + n * j] = -0.5 * (J.c[n2 + i + n * j] - J.c[n2 + j + n * i]);
    }
  __syncthreads();
  if (i == 0) {
    double mx = 0.0, tr = 0.0;
    for (int k = 0; k < n; ++k) {
      mx = fmax(mx, Kr[k + n * k]);
      tr += Kr[k + n * k];
    }
    s_max = mx;
    s_tr = tr;
    s_fail = 0;
    s_pmin = 1e300;
    s_pmax = 0.0;
  }
  __syncthreads();
  const double shift = J.shift_rel * s_tr;
  const double thr = s_max * n * 2.220446049250313e-16;
  if (i < n && shift != 0.0) Kr[i + n * i] -= shift;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (i == k) {
      const double d = Kr[k + n * k];
      if (!(d > thr)) s_fail = 1;
      s_piv = d;
      s_pmin = fmin(s_pmin, d);
      s_pmax = fmax(s_pmax, d);
    }
    __syncthreads();
    if (s_fail) break;
    const double rs = rsqrt(s_piv);
    if (i == k) {
      Kr[k + n * k] = s_piv * rs;
      if (C) Ki[k + n * k] = 0.0;
    } else if (i > k && i < n) {
      Kr[i + n * k] *= rs;
      if (C) Ki[i + n * k] *= rs;
    }
    __syncthreads();
    if (i > k && i < n) {
      const double lr = Kr[i + n * k], li = C ? Ki[i + n * k] : 0.0;
      for (int j = k + 1; j <= i; ++j) {
        const double mr = Kr[j + n * k], mi = C ? Ki[j + n * k] : 0.0;
        // K[i,j] -= L[i,k] conj(L[j,k])
        Kr[i + n * j] -= lr * mr + li * mi;
        if (C) Ki[i + n * j] -= li * mr - lr * mi;
      }
    }
    __syncthreads();
  }
  if (s_fail) {
    if (i == 0) *J.ok = 0;
    return;
  }
  if (i == 0) {
    *J.ok = 1;
    if (J.cond) *J.cond = (s_pmin < kCholCondRatio * s_pmax) ? 1 : 0;
  }
  if (!J.R || i >= n) return;
  // R[a, o] = conj(L[o, a]) (upper triangular)
  for (int a = 0; a < n; ++a) {
    const bool up = i >= a;  // thread i = column o
    J.R[a + n * i] = up ? Kr[i + n * a] : 0.0;
    if (C) J.R[n2 + a + n * i] = up ? -Ki[i + n * a] : 0.0;
  }
  // X = L^-1, column i computed by thread i;  R^+[o, a] = conj(X[a, o]): thread o = i writes its own column of X
  // into row o of R^+ (coalesced across threads) and reads back only what it wrote itself
  double* Pr = J.Rp;
  double* Pi = J.Rp + n2;
  for (int a = 0; a < n; ++a) {
    double xr = 0.0, xi = 0.0;
    if (a == i) {
      xr = 1.0 / Kr[a + n * a];
    } else if (a > i) {
      double sr = 0.0, si = 0.0;
      for (int k = i; k < a; ++k) {
        const double lr = Kr[a + n * k], li = C ? Ki[a + n * k] : 0.0;
        const double yr = Pr[i + n * k], yi = C ? -Pi[i + n * k] : 0.0;  // X[k, i] = conj(R^+[i, k])
        sr += lr * yr - li * yi;
        si += lr * yi + li * yr;
      }
      const double inv = 1.0 / Kr[a + n * a];
      xr = -sr * inv;
      xi = -si * inv;
    }
    Pr[i + n * a] = xr;
    if (C) Pi[i + n * a] = -xi;
  }
}

void run_chol(itn_ctx* ctx, bool cplx, const std::vector<CholJob>& jobs) {
  if (jobs.empty()) return;
  int maxn = 0;
  for (auto& j : jobs) maxn = std::max(maxn, j.n);
  ITN_REQUIRE(maxn <= 64, ITN_EINVAL, "batched Cholesky supports n <= 64");
  const size_t smem = (size_t)maxn * maxn * (cplx ? 2 : 1) * sizeof(double);
  DevBuf jb(ctx, jobs.size() * sizeof(CholJob));
  const CholJob* dj = itn_upload(ctx, jobs, jb);
  const int threads = maxn <= 32 ? 32 : 64;
  if (cplx) {
    CUDA_CHECK(cudaFuncSetAttribute(k_chol<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chol<true><<<(unsigned)jobs.size(), threads, smem, ctx->stream>>>(dj);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(k_chol<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chol<false><<<(unsigned)jobs.size(), threads, smem, ctx->stream>>>(dj);
  }
  ITN_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------
// map_eigvals
// ------------------------------------------------------------------------------------------------
struct HermJob {
  const double* src;  // planar chi x chi
  double* dst;
  int chi;
};
// dst = (src + src^H) / 2
template <bool C>
__global__ void k_hermitize(const HermJob* __restrict__ jobs) {
  const HermJob J = jobs[blockIdx.x];
  const int n = J.chi, n2 = n * n;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int r = i % n, c = i / n, t = c + n * r;
    J.dst[i] = 0.5 * (J.src[i] + J.src[t]);
    if (C) J.dst[n2 + i] = 0.5 * (J.src[n2 + i] - J.src[n2 + t]);
  }
}

// ITensors / NDTensors truncate! rule on weights p[0..n) sorted by decreasing magnitude.
__device__ int truncate_spectrum(double* p, int n, int maxdim, double cutoff, double* truncerr_out) {
  for (int i = n - 1; i >= 0; --i) {
    if (p[i] >= 0.0) break;
    p[i] = 0.0;
  }
  if (n == 1) {
    *truncerr_out = 0.0;
    return 1;
  }
  int md = (maxdim > 0 && maxdim < n) ? maxdim : n;
  int k = n;
  double terr = 0.0;
  while (k > md) {
    terr += p[k - 1];
    --k;
  }
  double scale = 0.0;
  for (int i = 0; i < n; ++i) scale += p[i];
  if (scale == 0.0) scale = 1.0;
  if (cutoff >= 0.0) {
    while (k > 1 && terr + p[k - 1] <= cutoff * scale) {
      terr += p[k - 1];
      --k;
    }
  }
  *truncerr_out = terr / scale;
  return k < 1 ? 1 : k;
}

struct EigFnJob {
  const double* h;    // hermitised input (planar)
  const double* us;   // Jacobi output U*Sigma (planar)
  const double* v;    // Jacobi output V
  const double* sigma;
  const int* perm;
  double* out;        // planar chi x chi
  int chi;
  int* deficient;     // optional: set to 1 when eigenvalues were dropped by the cutoff
  const int* skip;    // optional device flag: non-zero -> nothing to do (the Cholesky test proved full support)
  double shift;       // > 0: the decomposed matrix was H + shift * 1 (positive definite), lambda_j = sigma_j - shift
  int* mixed;         // optional (shift == 0): set to 1 when a right singular vector is not an eigenvector (a +-lambda
                      // pair of an indefinite H mixes in the SVD); the caller then repeats the job with a shift
};
// out = V_kept f(lambda) V_kept^H with lambda_j = sigma_j * sign(Re <v_j, (U Sigma)_j>); diagonal inputs
// short-circuit to f(diag) (map_diag, apply.jl:22).  fn: 0 sqrt, 1 inv sqrt, 2 inv, 3 one (support projector).
template <bool C>
__global__ void __launch_bounds__(256) k_eig_fn(const EigFnJob* __restrict__ jobs, int fn, double cutoff) {
  __shared__ double lam[256];
  __shared__ double fr[256], fi[256];
  __shared__ int s_keep, s_diag;
  const EigFnJob J = jobs[blockIdx.x];
  if (J.skip && *J.skip) return;
  const int n = J.chi, n2 = n * n;
  const int tid = threadIdx.x;
  if (tid == 0) s_diag = 1;
  __syncthreads();
  for (int i = tid; i < n2; i += blockDim.x) {
    if (i % n != i / n && (J.h[i] != 0.0 || (C && J.h[n2 + i] != 0.0))) s_diag = 0;
  }
  __syncthreads();
  auto apply_f = [&](double l, double& re, double& im) {
    // f on a real eigenvalue; complex dtype follows the principal branch for negative arguments
    double sr, si;
    if (fn == 3) {
      re = 1.0;
      im = 0.0;
      return;
    }
    if (l >= 0.0) {
      sr = sqrt(l);
      si = 0.0;
    } else if (C) {
      sr = 0.0;
      si = sqrt(-l);
    } else {
      sr = nan("");
      si = 0.0;
    }
    if (fn == 0) {
      re = sr;
      im = si;
    } else if (fn == 1) {
      if (si == 0.0) {
        re = 1.0 / sr;
        im = 0.0;
      } else {
        const double den = sr * sr + si * si;
        re = sr / den;
        im = -si / den;
      }
    } else {
      re = 1.0 / l;
      im = 0.0;
    }
  };
  if (s_diag) {
    for (int i = tid; i < n2; i += blockDim.x) {
      double re = 0.0, im = 0.0;
      if (i % n == i / n) apply_f(J.h[i], re, im);
      J.out[i] = re;
      if (C) J.out[n2 + i] = im;
    }
    return;
  }
  __shared__ int s_perm[256];
  for (int j = tid; j < n; j += blockDim.x) {
    const int col = J.perm[j];
    s_perm[j] = col;
    if (J.shift > 0.0) {
      lam[j] = J.sigma[j] - J.shift;
      continue;
    }
    double d = 0.0;
    for (int i = 0; i < n; ++i) {
      d += J.v[col * n + i] * J.us[col * n + i];
      if (C) d += J.v[n2 + col * n + i] * J.us[n2 + col * n + i];
    }
    lam[j] = d < 0.0 ? -J.sigma[j] : J.sigma[j];
    // <v_j, H v_j> = +sigma_j for every significant j when H is positive semi-definite.  A negative eigenvalue may pair up
    // with a positive one of (nearly) the same magnitude, and the right singular vectors of such a pair are an arbitrary
    // rotation of the two eigenvectors (d anywhere in [-sigma, sigma]): any significant d below sigma / 2 sends the
    // matrix to the shifted route, i.e. every indefinite input takes it
    if (J.mixed && J.sigma[j] > 1e-8 * J.sigma[0] && d < 0.5 * J.sigma[j]) *J.mixed = 1;
  }
  __syncthreads();
  if (tid == 0 && J.shift > 0.0) {
    // eigenvalues of the shifted decomposition come sorted by value: restore the order by decreasing magnitude
    for (int i = 1; i < n; ++i) {
      const double l = lam[i];
      const int c = s_perm[i];
      int k = i - 1;
      while (k >= 0 && fabs(lam[k]) < fabs(l)) {
        lam[k + 1] = lam[k];
        s_perm[k + 1] = s_perm[k];
        --k;
      }
      lam[k + 1] = l;
      s_perm[k + 1] = c;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int keep = n;
    if (cutoff >= 0.0) {
      double terr;
      double p[256];
      for (int i = 0; i < n; ++i) p[i] = lam[i];
      keep = truncate_spectrum(p, n, 0, cutoff, &terr);
    }
    s_keep = keep;
    if (J.deficient && keep < n) *J.deficient = 1;
  }
  __syncthreads();
  const int keep = s_keep;
  for (int j = tid; j < keep; j += blockDim.x) apply_f(lam[j], fr[j], fi[j]);
  __syncthreads();
  for (int i = tid; i < n2; i += blockDim.x) {
    const int r = i % n, c = i / n;
    double ar = 0.0, ai = 0.0;
    for (int j = 0; j < keep; ++j) {
      const int col = s_perm[j];
      const double vr = J.v[col * n + r], vi = C ? J.v[n2 + col * n + r] : 0.0;
      const double wr = J.v[col * n + c], wi = C ? -J.v[n2 + col * n + c] : 0.0;  // conj(V[c, j])
      const double pr = vr * wr - vi * wi, pi = vr * wi + vi * wr;
      ar += pr * fr[j] - pi * fi[j];
      ai += pr * fi[j] + pi * fr[j];
    }
    J.out[i] = ar;
    if (C) J.out[n2 + i] = ai;
  }
}

// host interleaved <-> device planar for batches of equally sized matrices
template <bool C>
__global__ void k_split(const double* __restrict__ in, double* __restrict__ out, int each, int count) {
  const long long tot = (long long)each * count;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / each, r = i % each;
    if (C) {
      out[b * 2 * each + r] = in[2 * i];
      out[b * 2 * each + each + r] = in[2 * i + 1];
    } else {
      out[i] = in[i];
    }
  }
}
template <bool C>
__global__ void k_merge(const double* __restrict__ in, double* __restrict__ out, int each, int count) {
  const long long tot = (long long)each * count;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / each, r = i % each;
    if (C) {
      out[2 * i] = in[b * 2 * each + r];
      out[2 * i + 1] = in[b * 2 * each + each + r];
    } else {
      out[i] = in[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// simple update glue kernels (one CTA per edge)
// ------------------------------------------------------------------------------------------------
struct SuEdge {
  // per side (0: esrc, 1: edst)
  const double* gus[2];   // Jacobi output of the bond environment C (n x n): U*Sigma (unused), kept for symmetry
  const double* gv[2];    // eigenvectors V (n x n)
  const double* gsig[2];  // eigenvalues, sorted descending
  const int* gperm[2];
  double* R[2];           // r x n
  double* Rp[2];          // n x r
  double* T[2];           // n x (d * chi_new), column-major: T[(s + d*l) + n*(s' + d*l')]
  int d[2], n[2], r[2];
  int chi;                // current bond dimension
  const double* gate;     // planar d1 d2 d1 d2: g[s1' + d1*(s2' + d2*(s1 + d1*s2))]
  const int* rok[2];      // device flag: R / R^+ of this side already came from the Cholesky route
  double* theta0;         // planar (r1 d1) x (r2 d2): theta' as built (input of the SVD, kept for V)
  double* theta;          // planar (r1 d1) x (r2 d2): U Sigma after the SVD
  double* tv;             // V of theta (kept columns only, rebuilt from theta0 and U Sigma)
  double* tsig;
  int* tperm;
  int newdim;             // filled on the host after truncation
  // wide theta' (m < n columns): the Jacobi kernel decomposes theta'^H (n x m, no null columns); tt0 = theta'^H,
  // tt = theta'^H U = V Sigma (columns in the kernel's order); k_su_urec turns them into theta (U Sigma) and tv (V)
  double* tt0;
  double* tt;
  int transposed;
};

// R[i, o] = sqrt(lambda_i) V[o, i];  R^+[o, i] = conj(V[o, i]) / sqrt(lambda_i)   (C = V L V^H = conj(A~^H A~))
template <bool C>
__global__ void __launch_bounds__(256) k_su_build_R(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x >> 1];
  const int side = blockIdx.x & 1;
  if (E.rok[side] && *E.rok[side]) return;
  const int n = E.n[side], r = E.r[side];
  const long long n2 = (long long)n * n, rn = (long long)r * n;
  const double lmax = E.gsig[side][0];
  const double thr = lmax * n * 2.220446049250313e-16;
  for (long long idx = threadIdx.x; idx < rn; idx += blockDim.x) {
    const int i = (int)(idx % r), o = (int)(idx / r);
    const int col = E.gperm[side][i];
    const double lam = E.gsig[side][i];
    const double sq = sqrt(lam);
    const double inv = (lam > thr && lam > 0.0) ? 1.0 / sq : 0.0;
    const double vr = E.gv[side][(long long)col * n + o];
    const double vi = C ? E.gv[side][n2 + (long long)col * n + o] : 0.0;
    E.R[side][i + (long long)r * o] = sq * vr;
    E.Rp[side][o + (long long)n * i] = inv * vr;
    if (C) {
      E.R[side][rn + i + (long long)r * o] = sq * vi;
      E.Rp[side][rn + o + (long long)n * i] = -inv * vi;
    }
  }
}

// theta'[(r1, s1'), (r2, s2')] = sum gate[s1', s2', s1, s2] sum_l R1[r1, (s1, l)] R2[r2, (s2, l)]
template <bool C>
__global__ void __launch_bounds__(256) k_su_theta(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x];
  const int d1 = E.d[0], d2 = E.d[1], r1 = E.r[0], r2 = E.r[1], chi = E.chi;
  const int m = r1 * d1, nc = r2 * d2;
  const long long mn = (long long)m * nc;
  const long long p1 = (long long)r1 * E.n[0], p2 = (long long)r2 * E.n[1];
  const int gsz = d1 * d2 * d1 * d2;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < mn; idx += (long long)gridDim.y * blockDim.x) {
    const int row = (int)(idx % m), colx = (int)(idx / m);
    const int a = row % r1, s1p = row / r1, b = colx % r2, s2p = colx / r2;
    double accr = 0.0, acci = 0.0;
    for (int s1 = 0; s1 < d1; ++s1)
      for (int s2 = 0; s2 < d2; ++s2) {
        double tr = 0.0, ti = 0.0;
        for (int l = 0; l < chi; ++l) {
          const long long i1 = a + (long long)r1 * (s1 + d1 * l), i2 = b + (long long)r2 * (s2 + d2 * l);
          const double xr = E.R[0][i1], yr = E.R[1][i2];
          if (C) {
            const double xi = E.R[0][p1 + i1], yi = E.R[1][p2 + i2];
            tr += xr * yr - xi * yi;
            ti += xr * yi + xi * yr;
          } else {
            tr += xr * yr;
          }
        }
        const int gi = s1p + d1 * (s2p + d2 * (s1 + d1 * s2));
        const double gr = E.gate[gi], gim = C ? E.gate[gsz + gi] : 0.0;
        accr += gr * tr - gim * ti;
        acci += gr * ti + gim * tr;
      }
    E.theta0[idx] = accr;
    if (C) E.theta0[mn + idx] = acci;
  }
}

// V of the kept singular triplets, rebuilt from theta' = U S V^H:  V[:, col] = theta'^H (U S)[:, col] / sigma^2.
// (The Jacobi SVD of theta' does not accumulate V; for the kept columns sigma / sigma_max >= sqrt(cutoff), and
// the product R1' R2' = U_k U_k^H theta' does not depend on the accuracy of the small-sigma rows at all.)
template <bool C>
__global__ void __launch_bounds__(256) k_su_vrec(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x];
  if (E.transposed) return;  // k_su_urec
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1], nd = E.newdim;
  const long long mn = (long long)m * nc, nn = (long long)nc * nc;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < nc * nd; idx += gridDim.y * blockDim.x) {
    const int j = idx % nc, lp = idx / nc;
    const int col = E.tperm[lp];
    const double sig = E.tsig[lp];
    const double f = sig > 0.0 ? 1.0 / (sig * sig) : 0.0;
    double accr = 0.0, acci = 0.0;
    for (int i = 0; i < m; ++i) {
      const double ar = E.theta0[i + (long long)m * j], ur = E.theta[i + (long long)m * col];
      if (C) {
        const double ai = E.theta0[mn + i + (long long)m * j], ui = E.theta[mn + i + (long long)m * col];
        accr += ar * ur + ai * ui;  // conj(a) * u
        acci += ar * ui - ai * ur;
      } else {
        accr += ar * ur;
      }
    }
    E.tv[j + (long long)nc * col] = f * accr;
    if (C) E.tv[nn + j + (long long)nc * col] = f * acci;
  }
}

// wide theta': tt0 = theta'^H (conjugate transpose, n x m)
template <bool C>
__global__ void __launch_bounds__(256) k_su_transpose(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x];
  if (!E.transposed) return;
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1];
  const long long mn = (long long)m * nc;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < mn; idx += (long long)gridDim.y * blockDim.x) {
    const int i = (int)(idx % m), j = (int)(idx / m);
    E.tt0[j + (long long)nc * i] = E.theta0[idx];
    if (C) E.tt0[mn + j + (long long)nc * i] = -E.theta0[mn + idx];
  }
}

// wide theta' after the decomposition of theta'^H:  V[:, col] = tt[:, col] / sigma,  (U Sigma)[:, col] = theta' V[:, col]
template <bool C>
__global__ void __launch_bounds__(256) k_su_urec(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x];
  if (!E.transposed) return;
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1], nd = E.newdim;
  const long long mn = (long long)m * nc, nn = (long long)nc * nc;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < (m + nc) * nd; idx += gridDim.y * blockDim.x) {
    const int lp = idx / (m + nc), q = idx % (m + nc);
    const int col = E.tperm[lp];
    const double sig = E.tsig[lp];
    const double f = sig > 0.0 ? 1.0 / sig : 0.0;
    if (q < nc) {  // V
      const int j = q;
      E.tv[j + (long long)nc * col] = f * E.tt[j + (long long)nc * col];
      if (C) E.tv[nn + j + (long long)nc * col] = f * E.tt[mn + j + (long long)nc * col];
    } else {       // U Sigma
      const int i = q - nc;
      double accr = 0.0, acci = 0.0;
      for (int j = 0; j < nc; ++j) {
        const double ar = E.theta0[i + (long long)m * j], wr = E.tt[j + (long long)nc * col];
        if (C) {
          const double ai = E.theta0[mn + i + (long long)m * j], wi = E.tt[mn + j + (long long)nc * col];
          accr += ar * wr - ai * wi;
          acci += ar * wi + ai * wr;
        } else {
          accr += ar * wr;
        }
      }
      E.theta[i + (long long)m * col] = f * accr;
      if (C) E.theta[mn + i + (long long)m * col] = f * acci;
    }
  }
}

struct SuTrunc {
  const double* tsig;
  int ncand;    // min(m, n)
  double* row;  // result row of the gate: [new bond dimension, truncation error, kept singular values (stride entries)]
};
__global__ void k_su_truncate(const SuTrunc* __restrict__ jobs, int n, int maxdim, double cutoff, int stride) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const SuTrunc T = jobs[j];
  double p[512];
  const int nc = T.ncand;
  for (int i = 0; i < nc; ++i) p[i] = T.tsig[i] * T.tsig[i];
  double terr = 0.0;
  const int keep = truncate_spectrum(p, nc, maxdim, cutoff, &terr);
  T.row[0] = (double)keep;
  T.row[1] = terr;
  for (int i = 0; i < stride; ++i) T.row[2 + i] = i < keep ? T.tsig[i] : 0.0;
}

// T_side[(s, l), (s', l')] = sum_r R^+[(s, l), r] R'[r, s', l'],
//   R1'[r, s', l'] = (U S)[(r, s'), l'] / sqrt(sigma_l'),  R2'[r, s', l'] = sqrt(sigma_l') conj(V[(r, s'), l'])
template <bool C>
__global__ void __launch_bounds__(256) k_su_T(const SuEdge* __restrict__ edges) {
  const SuEdge E = edges[blockIdx.x >> 1];
  const int side = blockIdx.x & 1;
  const int n = E.n[side], r = E.r[side], d = E.d[side], nd = E.newdim;
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1];
  const long long mn = (long long)m * nc, nn = (long long)nc * nc, rn = (long long)r * n;
  const long long tot = (long long)n * d * nd;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.y * blockDim.x) {
    const int o = (int)(idx % n);
    const int q = (int)(idx / n);
    const int sp = q % d, lp = q / d;
    const int col = E.tperm[lp];
    const double sig = E.tsig[lp];
    const double f = side == 0 ? (sig > 0.0 ? 1.0 / sqrt(sig) : 0.0) : sqrt(sig);
    double accr = 0.0, acci = 0.0;
    for (int a = 0; a < r; ++a) {
      const double pr = E.Rp[side][o + (long long)n * a];
      const double pi = C ? E.Rp[side][rn + o + (long long)n * a] : 0.0;
      double xr, xi;
      if (side == 0) {
        const long long i = (a + (long long)r * sp) + (long long)m * col;
        xr = E.theta[i];
        xi = C ? E.theta[mn + i] : 0.0;
      } else {
        const long long i = (a + (long long)r * sp) + (long long)nc * col;
        xr = E.tv[i];
        xi = C ? -E.tv[nn + i] : 0.0;
      }
      accr += pr * xr - pi * xi;
      acci += pr * xi + pi * xr;
    }
    E.T[side][idx] = f * accr;
    if (C) E.T[side][tot + idx] = f * acci;
  }
}

// ---- thin sides: fewer outer-bond states than (site, gate bond) states, X < n (degree <= 2 sites at d chi > chi^(z-1)) ----
// The bond environment C = A~^T conj(A~) (n x n) then has rank X, and the reference's QR of the X x n matrix A~ returns
// an X x n R factor (apply.jl:74-75 with rows = outer bonds).  Any square-root factor of the outer messages serves as
// well as the Hermitian one: with M_j = S_j S_j^H (Cholesky, S_j = R_j^T of k_chol) the tensor A x_j S_j, read as an
// X x n matrix, IS such an R (R^T conj(R) = C), and R^+ = R^H (R R^H)^-1 with one X x X Cholesky.  No eigen-
// decomposition of the rank-deficient C is needed; if a Cholesky meets a non-positive pivot the flag stays 0 and the
// eigen route (k_jacobi_svd + k_su_build_R) takes over as before.
struct ThinJob {
  const double* at;  // A x_j S_j over the outer bonds, canonical planar [s, bonds...]
  long long n_t;     // its element count (offset of the imaginary plane)
  double* R;         // planar X x n:  R[i + X o], i = outer multi-index, o = s + d l
  double* Rp;        // planar n x X:  R^+[o + n i]
  double* G;         // planar X x X:  R R^H
  const double* Gp;  // planar X x X:  k_chol's R^+ output for G  (conj(G)^-1 = Gp Gp^H)
  const int* okG;    // k_chol flag of G
  const int* ok_env; // k_chol flags of the outer messages (n_env of them, contiguous)
  int n_env;
  int* ok;           // out: 1 when R / R^+ are valid
  int d, chi, X, n;
  long long lo;      // canonical stride of the gate bond (product of the extents below it, site included)
};

template <bool C>
__global__ void __launch_bounds__(256) k_thin_R(const ThinJob* __restrict__ jobs) {
  const ThinJob J = jobs[blockIdx.x];
  const long long lo_rest = J.lo / J.d, rn = (long long)J.X * J.n;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < J.n_t; idx += (long long)gridDim.y * blockDim.x) {
    const int s = (int)(idx % J.d);
    long long r = idx / J.d;
    const long long below = r % lo_rest;
    r /= lo_rest;
    const int l = (int)(r % J.chi);
    const long long above = r / J.chi;
    const long long i = below + lo_rest * above;
    const int o = s + J.d * l;
    J.R[i + (long long)J.X * o] = J.at[idx];
    if (C) J.R[rn + i + (long long)J.X * o] = J.at[J.n_t + idx];
  }
}

template <bool C>
__global__ void __launch_bounds__(256) k_thin_gram(const ThinJob* __restrict__ jobs) {
  const ThinJob J = jobs[blockIdx.x];
  const int X = J.X, n = J.n;
  const long long rn = (long long)X * n;
  for (int idx = threadIdx.x; idx < X * X; idx += blockDim.x) {
    const int i = idx % X, ip = idx / X;
    double ar = 0.0, ai = 0.0;
    for (int o = 0; o < n; ++o) {
      const double xr = J.R[i + (long long)X * o], yr = J.R[ip + (long long)X * o];
      if (C) {
        const double xi = J.R[rn + i + (long long)X * o], yi = J.R[rn + ip + (long long)X * o];
        ar += xr * yr + xi * yi;  // x conj(y)
        ai += xi * yr - xr * yi;
      } else {
        ar += xr * yr;
      }
    }
    J.G[idx] = ar;
    if (C) J.G[(long long)X * X + idx] = ai;
  }
}

// R^+ = R^H G^-1,  G^-1 = conj(Gp Gp^H)
template <bool C>
__global__ void __launch_bounds__(256) k_thin_pinv(const ThinJob* __restrict__ jobs) {
  extern __shared__ double sm[];  // [planes][X x X]: G^-1
  __shared__ int s_ok;
  const ThinJob J = jobs[blockIdx.x];
  const int X = J.X, n = J.n, x2 = X * X;
  if (threadIdx.x == 0) {
    int ok = *J.okG;
    for (int q = 0; q < J.n_env; ++q) ok = ok && J.ok_env[q];
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) return;  // *J.ok stays 0: eigen route
  double* Wr = sm;
  double* Wi = sm + x2;
  for (int idx = threadIdx.x; idx < x2; idx += blockDim.x) {
    const int a = idx % X, b = idx / X;  // W[a, b] = sum_k Gp[a, k] conj(Gp[b, k]);  G^-1 = conj(W)
    double wr = 0.0, wi = 0.0;
    for (int k = 0; k < X; ++k) {
      const double pr = J.Gp[a + X * k], qr = J.Gp[b + X * k];
      if (C) {
        const double pi = J.Gp[x2 + a + X * k], qi = J.Gp[x2 + b + X * k];
        wr += pr * qr + pi * qi;
        wi += pi * qr - pr * qi;
      } else {
        wr += pr * qr;
      }
    }
    Wr[idx] = wr;
    if (C) Wi[idx] = -wi;
  }
  __syncthreads();
  const long long rn = (long long)X * n;
  for (int idx = threadIdx.x; idx < n * X; idx += blockDim.x) {
    const int o = idx % n, i = idx / n;
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < X; ++k) {  // sum_k conj(R[k, o]) Ginv[k, i]
      const double xr = J.R[k + (long long)X * o], gr = Wr[k + X * i];
      if (C) {
        const double xi = -J.R[rn + k + (long long)X * o], gi = Wi[k + X * i];
        ar += xr * gr - xi * gi;
        ai += xr * gi + xi * gr;
      } else {
        ar += xr * gr;
      }
    }
    J.Rp[o + (long long)n * i] = ar;
    if (C) J.Rp[rn + o + (long long)n * i] = ai;
  }
  if (threadIdx.x == 0) *J.ok = 1;
}

// R^+ of a thin side, refined: the explicit G^-1 = conj(Gp Gp^H) of k_thin_pinv carries kappa(R)^2 eps.  With
//   Y = R^H conj(Gp)            (n x X, orthonormal columns up to kappa^2 eps: Y^H Y = Gp^T G conj(Gp) ~ 1)
//   R^+ = Y (Y^H Y)^-1 Gp^T     (R R^+ = 1 holds for ANY invertible Gp; the inverted matrix is ~ 1, so the result
//                                carries kappa eps from forming Y and nothing else)
// the pseudo-inverse is what CholeskyQR2 gives for R^H.  (Y^H Y)^-1 by Gauss-Jordan without pivoting (the matrix is
// Hermitian positive definite and within kappa^2 eps of the identity).  One CTA per job, everything in shared memory.
template <bool C>
__global__ void __launch_bounds__(256) k_thin_pinv2(const ThinJob* __restrict__ jobs) {
  extern __shared__ double sm[];
  __shared__ int s_ok;
  __shared__ double s_pr, s_pi;
  const ThinJob J = jobs[blockIdx.x];
  const int X = J.X, n = J.n, x2 = X * X, nx = n * X;
  if (threadIdx.x == 0) {
    int ok = *J.okG;
    for (int q = 0; q < J.n_env; ++q) ok = ok && J.ok_env[q];
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) return;  // *J.ok stays 0: eigen route
  constexpr int P = C ? 2 : 1;
  double* Yr = sm;                 // [P][n x X]
  double* Yi = sm + nx;
  double* Ar = sm + P * nx;        // [P][X x 2X]: augmented (Y^H Y | 1) -> (1 | W); row a, column b at a + X b
  double* Ai = Ar + 2 * x2;
  double* Mr = Ar + P * 2 * x2;    // [P][X x X]: M = W Gp^T
  double* Mi = Mr + x2;
  const long long rn = (long long)X * n;
  // Y[o, i] = sum_k conj(R[k, o]) conj(Gp[k, i])
  for (int idx = threadIdx.x; idx < nx; idx += blockDim.x) {
    const int o = idx % n, i = idx / n;
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < X; ++k) {
      const double xr = J.R[k + (long long)X * o], gr = J.Gp[k + X * i];
      if (C) {
        const double xi = J.R[rn + k + (long long)X * o], gi = J.Gp[x2 + k + X * i];
        ar += xr * gr - xi * gi;     // conj(x) conj(g) = conj(x g)
        ai -= xr * gi + xi * gr;
      } else {
        ar += xr * gr;
      }
    }
    Yr[idx] = ar;
    if (C) Yi[idx] = ai;
  }
  __syncthreads();
  // A = (Y^H Y | 1)
  for (int idx = threadIdx.x; idx < 2 * x2; idx += blockDim.x) {
    const int a = idx % X, b = idx / X;
    double ar = 0.0, ai = 0.0;
    if (b < X) {
      for (int o = 0; o < n; ++o) {
        const double pr = Yr[o + n * a], qr = Yr[o + n * b];
        if (C) {
          const double pi = Yi[o + n * a], qi = Yi[o + n * b];
          ar += pr * qr + pi * qi;   // conj(p) q
          ai += pr * qi - pi * qr;
        } else {
          ar += pr * qr;
        }
      }
    } else {
      ar = (b - X == a) ? 1.0 : 0.0;
    }
    Ar[idx] = ar;
    if (C) Ai[idx] = ai;
  }
  __syncthreads();
  // Gauss-Jordan: after step k column k of the left half is e_k
  for (int k = 0; k < X; ++k) {
    if (threadIdx.x == 0) {
      const double pr = Ar[k + X * k], pi = C ? Ai[k + X * k] : 0.0;
      const double den = pr * pr + pi * pi;
      s_pr = pr / den;               // 1 / pivot
      s_pi = -pi / den;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * X; b += blockDim.x) {  // scale row k
      const double vr = Ar[k + X * b], vi = C ? Ai[k + X * b] : 0.0;
      Ar[k + X * b] = vr * s_pr - vi * s_pi;
      if (C) Ai[k + X * b] = vr * s_pi + vi * s_pr;
    }
    __syncthreads();
    // rows a != k: row_a -= A[a, k] row_k, columns b > k only (the others are already final or unused)
    for (int idx = threadIdx.x; idx < X * (2 * X - k - 1); idx += blockDim.x) {
      const int a = idx % X, b = k + 1 + idx / X;
      if (a == k) continue;
      const double fr = Ar[a + X * k], fi = C ? Ai[a + X * k] : 0.0;
      const double vr = Ar[k + X * b], vi = C ? Ai[k + X * b] : 0.0;
      Ar[a + X * b] -= fr * vr - fi * vi;
      if (C) Ai[a + X * b] -= fr * vi + fi * vr;
    }
    __syncthreads();
  }
  // M = W Gp^T:  M[a, i] = sum_b W[a, b] Gp[i, b]
  for (int idx = threadIdx.x; idx < x2; idx += blockDim.x) {
    const int a = idx % X, i = idx / X;
    double ar = 0.0, ai = 0.0;
    for (int b = 0; b < X; ++b) {
      const double wr = Ar[a + X * (X + b)], gr = J.Gp[i + X * b];
      if (C) {
        const double wi = Ai[a + X * (X + b)], gi = J.Gp[x2 + i + X * b];
        ar += wr * gr - wi * gi;
        ai += wr * gi + wi * gr;
      } else {
        ar += wr * gr;
      }
    }
    Mr[idx] = ar;
    if (C) Mi[idx] = ai;
  }
  __syncthreads();
  // R^+ = Y M
  for (int idx = threadIdx.x; idx < nx; idx += blockDim.x) {
    const int o = idx % n, i = idx / n;
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < X; ++k) {
      const double yr = Yr[o + n * k], mr = Mr[k + X * i];
      if (C) {
        const double yi = Yi[o + n * k], mi = Mi[k + X * i];
        ar += yr * mr - yi * mi;
        ai += yr * mi + yi * mr;
      } else {
        ar += yr * mr;
      }
    }
    J.Rp[o + (long long)n * i] = ar;
    if (C) J.Rp[rn + o + (long long)n * i] = ai;
  }
  if (threadIdx.x == 0) *J.ok = 1;
}

// ---- shared-memory versions of the three glue kernels (one CTA per gate / gate side; operands staged once) ----------
// The plain kernels above read every operand element from global memory once per output element; at 2048 gates per
// layer they cost ~1 ms each although they move < 0.3 GB.  Selected when the operands of every gate of the batch fit.

// theta': thread per (a, b) computes t[s1][s2] = sum_l R1[a,(s1,l)] R2[b,(s2,l)] once, then all d1 d2 gate outputs.
template <bool C>
__global__ void __launch_bounds__(256) k_su_theta_s(const SuEdge* __restrict__ edges) {
  extern __shared__ double sm[];
  const SuEdge E = edges[blockIdx.x];
  const int d1 = E.d[0], d2 = E.d[1], r1 = E.r[0], r2 = E.r[1], chi = E.chi;
  const int m = r1 * d1, nc = r2 * d2;
  const long long mn = (long long)m * nc;
  const int p1 = r1 * E.n[0], p2 = r2 * E.n[1];
  const int gsz = d1 * d2 * d1 * d2;
  constexpr int P = C ? 2 : 1;
  double* R1 = sm;              // [P][p1]
  double* R2 = R1 + P * p1;     // [P][p2]
  double* G = R2 + P * p2;      // [P][gsz]
  for (int i = threadIdx.x; i < P * p1; i += blockDim.x) R1[i] = E.R[0][i];
  for (int i = threadIdx.x; i < P * p2; i += blockDim.x) R2[i] = E.R[1][i];
  for (int i = threadIdx.x; i < P * gsz; i += blockDim.x) G[i] = E.gate[i];
  __syncthreads();
  for (int ab = threadIdx.x; ab < r1 * r2; ab += blockDim.x) {
    const int a = ab % r1, b = ab / r1;
    double tr[16], ti[16];
    for (int s2 = 0; s2 < d2; ++s2)
      for (int s1 = 0; s1 < d1; ++s1) {
        double xr_ = 0.0, xi_ = 0.0;
        for (int l = 0; l < chi; ++l) {
          const int i1 = a + r1 * (s1 + d1 * l), i2 = b + r2 * (s2 + d2 * l);
          const double xr = R1[i1], yr = R2[i2];
          if (C) {
            const double xi = R1[p1 + i1], yi = R2[p2 + i2];
            xr_ += xr * yr - xi * yi;
            xi_ += xr * yi + xi * yr;
          } else {
            xr_ += xr * yr;
          }
        }
        tr[s1 + d1 * s2] = xr_;
        ti[s1 + d1 * s2] = xi_;
      }
    for (int s2p = 0; s2p < d2; ++s2p)
      for (int s1p = 0; s1p < d1; ++s1p) {
        double accr = 0.0, acci = 0.0;
        for (int s2 = 0; s2 < d2; ++s2)
          for (int s1 = 0; s1 < d1; ++s1) {
            const int gi = s1p + d1 * (s2p + d2 * (s1 + d1 * s2));
            const double gr = G[gi], gim = C ? G[gsz + gi] : 0.0;
            const double xr = tr[s1 + d1 * s2], xi = ti[s1 + d1 * s2];
            accr += gr * xr - gim * xi;
            acci += gr * xi + gim * xr;
          }
        const long long idx = (a + (long long)r1 * s1p) + (long long)m * (b + (long long)r2 * s2p);
        E.theta0[idx] = accr;
        if (C) E.theta0[mn + idx] = acci;
      }
  }
}

// V[:, col] = theta'^H (U S)[:, col] / sigma^2 for the kept columns; theta' staged with a padded column stride
template <bool C>
__global__ void __launch_bounds__(256) k_su_vrec_s(const SuEdge* __restrict__ edges) {
  extern __shared__ double sm[];
  const SuEdge E = edges[blockIdx.x];
  if (E.transposed) return;  // k_su_urec
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1], nd = E.newdim;
  const long long mn = (long long)m * nc, nn = (long long)nc * nc;
  constexpr int P = C ? 2 : 1;
  const int ms = m + 1;            // padded stride: lanes walk j, rows of one column stay contiguous
  double* Th = sm;                 // [P][nc][ms]
  double* Us = Th + P * nc * ms;   // [P][nd][m]   kept columns of U S, in sorted order
  for (int idx = threadIdx.x; idx < m * nc; idx += blockDim.x) {
    const int i = idx % m, j = idx / m;
    Th[j * ms + i] = E.theta0[idx];
    if (C) Th[nc * ms + j * ms + i] = E.theta0[mn + idx];
  }
  for (int idx = threadIdx.x; idx < m * nd; idx += blockDim.x) {
    const int i = idx % m, lp = idx / m;
    const long long src = i + (long long)m * E.tperm[lp];
    Us[idx] = E.theta[src];
    if (C) Us[nd * m + idx] = E.theta[mn + src];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nc * nd; idx += blockDim.x) {
    const int j = idx % nc, lp = idx / nc;
    const int col = E.tperm[lp];
    const double sig = E.tsig[lp];
    const double f = sig > 0.0 ? 1.0 / (sig * sig) : 0.0;
    const double* tr = Th + j * ms;
    const double* ti = Th + nc * ms + j * ms;
    const double* ur = Us + lp * m;
    const double* ui = Us + nd * m + lp * m;
    double accr = 0.0, acci = 0.0;
    for (int i = 0; i < m; ++i) {
      if (C) {
        accr += tr[i] * ur[i] + ti[i] * ui[i];  // conj(a) * u
        acci += tr[i] * ui[i] - ti[i] * ur[i];
      } else {
        accr += tr[i] * ur[i];
      }
    }
    E.tv[j + (long long)nc * col] = f * accr;
    if (C) E.tv[nn + j + (long long)nc * col] = f * acci;
  }
}

// T_side = R^+ R' with R^+ and the kept, scaled columns of R' staged
template <bool C>
__global__ void __launch_bounds__(256) k_su_T_s(const SuEdge* __restrict__ edges) {
  extern __shared__ double sm[];
  const SuEdge E = edges[blockIdx.x >> 1];
  const int side = blockIdx.x & 1;
  const int n = E.n[side], r = E.r[side], d = E.d[side], nd = E.newdim;
  const int m = E.r[0] * E.d[0], nc = E.r[1] * E.d[1];
  const long long mn = (long long)m * nc, nn = (long long)nc * nc;
  const int rn = r * n, q = d * nd, rq = r * q;
  constexpr int P = C ? 2 : 1;
  double* Rp = sm;            // [P][n x r]
  double* Rn = Rp + P * rn;   // [P][r x (d nd)]:  R'[a, sp, lp]
  for (int i = threadIdx.x; i < P * rn; i += blockDim.x) Rp[i] = E.Rp[side][i];
  for (int idx = threadIdx.x; idx < rq; idx += blockDim.x) {
    const int a = idx % r, sq = idx / r, sp = sq % d, lp = sq / d;
    const int col = E.tperm[lp];
    const double sig = E.tsig[lp];
    const double f = side == 0 ? (sig > 0.0 ? 1.0 / sqrt(sig) : 0.0) : sqrt(sig);
    double xr, xi;
    if (side == 0) {
      const long long i = (a + (long long)r * sp) + (long long)m * col;
      xr = E.theta[i];
      xi = C ? E.theta[mn + i] : 0.0;
    } else {
      const long long i = (a + (long long)r * sp) + (long long)nc * col;
      xr = E.tv[i];
      xi = C ? -E.tv[nn + i] : 0.0;
    }
    Rn[idx] = f * xr;
    if (C) Rn[rq + idx] = f * xi;
  }
  __syncthreads();
  const long long tot = (long long)n * q;
  for (int idx = threadIdx.x; idx < n * q; idx += blockDim.x) {
    const int o = idx % n, sq = idx / n;
    double accr = 0.0, acci = 0.0;
    for (int a = 0; a < r; ++a) {
      const double pr = Rp[o + n * a], xr = Rn[a + r * sq];
      if (C) {
        const double pi = Rp[rn + o + n * a], xi = Rn[rq + a + r * sq];
        accr += pr * xr - pi * xi;
        acci += pr * xi + pi * xr;
      } else {
        accr += pr * xr;
      }
    }
    E.T[side][idx] = accr;
    if (C) E.T[side][tot + idx] = acci;
  }
}

// CholeskyQR2: R <- R2 R1 (n x n, column-major R[i + n o]) and R^+ <- R1^+ R2^+ when the second factorisation succeeded
struct ComposeJob {
  double* R;         // in: R1, out: R2 R1
  double* Rp;        // in: R1^+, out: R1^+ R2^+
  const double* R2;
  const double* Rp2;
  const int* ok2;
  int n;
};
template <bool C>
__global__ void __launch_bounds__(256) k_su_compose(const ComposeJob* __restrict__ jobs) {
  extern __shared__ double sm[];
  const ComposeJob J = jobs[blockIdx.x];
  if (!*J.ok2) return;
  const int n = J.n, n2 = n * n;
  constexpr int PLN = C ? 2 : 1;
  double* a = sm;              // R1, then R1^+
  double* b = sm + PLN * n2;   // R2, then R2^+
  for (int which = 0; which < 2; ++which) {
    double* dst = which ? J.Rp : J.R;
    const double* lhs = which ? a : b;  // R = R2 R1; R^+ = R1^+ R2^+
    const double* rhs = which ? b : a;
    __syncthreads();
    for (int i = threadIdx.x; i < PLN * n2; i += blockDim.x) {
      a[i] = dst[i];
      b[i] = (which ? J.Rp2 : J.R2)[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < n2; idx += blockDim.x) {
      const int r = idx % n, c = idx / n;
      double xr = 0.0, xi = 0.0;
      for (int k = 0; k < n; ++k) {
        const double lr = lhs[r + n * k], rr = rhs[k + n * c];
        if (C) {
          const double li = lhs[n2 + r + n * k], ri = rhs[n2 + k + n * c];
          xr += lr * rr - li * ri;
          xi += lr * ri + li * rr;
        } else {
          xr += lr * rr;
        }
      }
      dst[idx] = xr;
      if (C) dst[n2 + idx] = xi;
    }
  }
}

struct SuSite {
  const double* a;   // old tensor, canonical planar [s, bonds...]
  double* out;       // new tensor
  const double* T;   // n x (d * chi_new)
  long long n_old, n_new;
  long long lo;      // product of extents below the shared bond (incl. site): stride of l
  long long hi;      // product of extents above the shared bond
  int d, chi, chi_new;
};
// A'[s', lo.., l', hi..] = sum_{s, l} A[s, lo.., l, hi..] T[(s, l), (s', l')]
template <bool C>
__global__ void __launch_bounds__(256) k_su_rebuild(const SuSite* __restrict__ sites) {
  const SuSite S = sites[blockIdx.x];
  const int d = S.d, chi = S.chi, cn = S.chi_new, n = d * chi;
  const long long lo_rest = S.lo / d;  // extents between site and the shared bond
  const long long tsz = (long long)n * d * cn;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < S.n_new; idx += (long long)gridDim.y * blockDim.x) {
    const int sp = (int)(idx % d);
    long long r = idx / d;
    const long long mid = r % lo_rest;
    r /= lo_rest;
    const int lp = (int)(r % cn);
    const long long top = r / cn;
    const long long base = d * mid + S.lo * (long long)chi * top;
    const long long tcol = (long long)n * (sp + d * lp);
    double accr = 0.0, acci = 0.0;
    for (int l = 0; l < chi; ++l)
      for (int s = 0; s < d; ++s) {
        const long long ia = base + s + S.lo * l;
        const double ar = S.a[ia], tr = S.T[tcol + s + d * l];
        if (C) {
          const double ai = S.a[S.n_old + ia], ti = S.T[tsz + tcol + s + d * l];
          accr += ar * tr - ai * ti;
          acci += ar * ti + ai * tr;
        } else {
          accr += ar * tr;
        }
      }
    S.out[idx] = accr;
    if (C) S.out[S.n_new + idx] = acci;
  }
}

struct DiagMsgJob {
  double* m;            // planar chi x chi
  const double* svals;  // null: identity
  int chi;
};
template <bool C>
__global__ void k_diag_msg(const DiagMsgJob* __restrict__ jobs) {
  const DiagMsgJob J = jobs[blockIdx.x];
  const int n2 = J.chi * J.chi;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int r = i % J.chi, c = i / J.chi;
    J.m[i] = (r == c) ? (J.svals ? J.svals[r] : 1.0) : 0.0;
    if (C) J.m[n2 + i] = 0.0;
  }
}

struct NormJob2 {
  double* p;
  long long n;
};
__global__ void __launch_bounds__(256) k_normalize2(const NormJob2* __restrict__ jobs) {
  __shared__ double sh[8];
  __shared__ double tot;
  const NormJob2 J = jobs[blockIdx.x];
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
  const double f = 1.0 / sqrt(tot);
  for (long long i = threadIdx.x; i < J.n; i += blockDim.x) J.p[i] *= f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// map_eigvals entry points
// ------------------------------------------------------------------------------------------------
void itn_dev_map_eigvals(itn_ctx* ctx, bool cplx, int fn, int chi, int n, const double* const* in_ptrs,
                         double* const* out_ptrs, double cutoff) {
  // in_ptrs / out_ptrs are HOST arrays of device pointers (planar chi x chi matrices)
  if (n == 0) return;
  ITN_REQUIRE(chi >= 1 && chi <= 256, ITN_EUNSUPPORTED, "map_eigvals supports 1 <= chi <= 256");
  const int P = cplx ? 2 : 1;
  const size_t n2 = (size_t)chi * chi;
  DevBuf h(ctx, n * n2 * P * sizeof(double)), us(ctx, n * n2 * P * sizeof(double)), v(ctx, n * n2 * P * sizeof(double));
  DevBuf sig(ctx, (size_t)n * chi * sizeof(double)), perm(ctx, (size_t)n * chi * sizeof(int));
  std::vector<HermJob> hj(n);
  std::vector<SvdJob> sj(n);
  std::vector<EigFnJob> ej(n);
  for (int i = 0; i < n; ++i) {
    double* hi = h.as<double>() + i * n2 * P;
    double* ui = us.as<double>() + i * n2 * P;
    double* vi = v.as<double>() + i * n2 * P;
    hj[i] = {in_ptrs[i], hi, chi};
    sj[i] = {ui, vi, sig.as<double>() + (size_t)i * chi, perm.as<int>() + (size_t)i * chi, chi, chi};
    ej[i] = {hi, ui, vi, sig.as<double>() + (size_t)i * chi, perm.as<int>() + (size_t)i * chi, out_ptrs[i], chi, nullptr};
  }
  DevBuf hb(ctx, hj.size() * sizeof(HermJob)), eb(ctx, ej.size() * sizeof(EigFnJob));
  const HermJob* dh = itn_upload(ctx, hj, hb);
  if (cplx) k_hermitize<true><<<n, 128, 0, ctx->stream>>>(dh);
  else k_hermitize<false><<<n, 128, 0, ctx->stream>>>(dh);
  ITN_LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(us.p, h.p, n * n2 * P * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  run_jacobi(ctx, cplx, sj);
  DevBuf mixed(ctx, (size_t)n * sizeof(int));
  CUDA_CHECK(cudaMemsetAsync(mixed.p, 0, (size_t)n * sizeof(int), ctx->stream));
  for (int i = 0; i < n; ++i) ej[i].mixed = mixed.as<int>() + i;
  const EigFnJob* de = itn_upload(ctx, ej, eb);
  if (cplx) k_eig_fn<true><<<n, 256, 0, ctx->stream>>>(de, fn, cutoff);
  else k_eig_fn<false><<<n, 256, 0, ctx->stream>>>(de, fn, cutoff);
  ITN_LAUNCH_CHECK(ctx);
  // Indefinite Hermitian input: the right singular vectors of a +-lambda pair are not eigenvectors (H = [[0,1],[1,0]]
  // has V = 1), which k_eig_fn reports per matrix.  Those matrices are decomposed again as H + shift * 1 with
  // shift = 2 ||H||_F > spectral radius: positive definite, so its SVD IS its eigen-decomposition, lambda = sigma - shift.
  // (The unshifted route stays the default: one-sided Jacobi keeps the small eigenvalues of the PSD BP messages to
  // high relative accuracy, which the inverse square roots of the simple update need.)
  std::vector<int> hm(n);
  CUDA_CHECK(cudaMemcpyAsync(hm.data(), mixed.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<int> redo;
  for (int i = 0; i < n; ++i)
    if (hm[i]) redo.push_back(i);
  if (redo.empty()) return;
  std::vector<double> hh(n2 * P);
  std::vector<SvdJob> sj2;
  std::vector<EigFnJob> ej2;
  for (int i : redo) {
    double* hi = h.as<double>() + i * n2 * P;
    double* ui = us.as<double>() + i * n2 * P;
    CUDA_CHECK(cudaMemcpyAsync(hh.data(), hi, n2 * P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    double fro = 0.0;
    for (double x : hh) fro += x * x;
    const double shift = 2.0 * std::sqrt(fro);
    for (int k = 0; k < chi; ++k) hh[(size_t)k * chi + k] += shift;
    CUDA_CHECK(cudaMemcpyAsync(ui, hh.data(), n2 * P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sj2.push_back(sj[i]);
    EigFnJob e = ej[i];
    e.shift = shift;
    e.mixed = nullptr;
    ej2.push_back(e);
  }
  run_jacobi(ctx, cplx, sj2);
  DevBuf eb2(ctx, ej2.size() * sizeof(EigFnJob));
  const EigFnJob* de2 = itn_upload(ctx, ej2, eb2);
  if (cplx) k_eig_fn<true><<<(unsigned)ej2.size(), 256, 0, ctx->stream>>>(de2, fn, cutoff);
  else k_eig_fn<false><<<(unsigned)ej2.size(), 256, 0, ctx->stream>>>(de2, fn, cutoff);
  ITN_LAUNCH_CHECK(ctx);
}

#define API_BEGIN try {
#define API_END                              \
  }                                          \
  catch (const ItnError& e) {                \
    itn_set_error(e.what());                 \
    return e.code;                           \
  }                                          \
  catch (const std::bad_alloc&) {            \
    itn_set_error("host allocation failed"); \
    return ITN_ENOMEM;                       \
  }                                          \
  catch (const std::exception& e) {          \
    itn_set_error(e.what());                 \
    return ITN_EINVAL;                       \
  }                                          \
  return ITN_OK;

extern "C" int itn_map_eigvals(itn_ctx* ctx, int dtype, int fn, int chi, int n, const void* host_in, void* host_out,
                               double cutoff) {
  API_BEGIN
  ITN_REQUIRE(ctx && host_in && host_out, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(dtype == ITN_F64 || dtype == ITN_C128, ITN_EUNSUPPORTED, "dtype must be 0 (Float64) or 1 (ComplexF64)");
  ITN_REQUIRE(fn >= 0 && fn <= 2, ITN_EINVAL, "fn must be 0 (sqrt), 1 (inv sqrt) or 2 (inv)");
  ITN_REQUIRE(n >= 0 && chi >= 1, ITN_EINVAL, "bad batch shape");
  if (n == 0) return ITN_OK;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const bool cplx = dtype == ITN_C128;
  const int P = cplx ? 2 : 1;
  const size_t n2 = (size_t)chi * chi, bytes = (size_t)n * n2 * P * sizeof(double);
  DevBuf raw(ctx, bytes), in(ctx, bytes), out(ctx, bytes);
  CUDA_CHECK(cudaMemcpyAsync(raw.p, host_in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned g = (unsigned)std::min<size_t>((n * n2 + 255) / 256, 2048);
  if (cplx) k_split<true><<<g, 256, 0, ctx->stream>>>(raw.as<double>(), in.as<double>(), (int)n2, n);
  else k_split<false><<<g, 256, 0, ctx->stream>>>(raw.as<double>(), in.as<double>(), (int)n2, n);
  ITN_LAUNCH_CHECK(ctx);
  std::vector<const double*> ip(n);
  std::vector<double*> op(n);
  for (int i = 0; i < n; ++i) {
    ip[i] = in.as<double>() + i * n2 * P;
    op[i] = out.as<double>() + i * n2 * P;
  }
  itn_dev_map_eigvals(ctx, cplx, fn, chi, n, ip.data(), op.data(), cutoff);
  if (cplx) k_merge<true><<<g, 256, 0, ctx->stream>>>(out.as<double>(), raw.as<double>(), (int)n2, n);
  else k_merge<false><<<g, 256, 0, ctx->stream>>>(out.as<double>(), raw.as<double>(), (int)n2, n);
  ITN_LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(host_out, raw.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  API_END
}

// Batched singular values (and U*Sigma) of host matrices through the same Jacobi kernels the gate path uses:
// instrumentation for tests/test_gpu_linalg.py and tools/svd_bench.py.
extern "C" int itn_svd_batch(itn_ctx* ctx, int dtype, int m, int n, int batch, const void* host_in, double* host_sigma,
                             void* host_us, int variant, double* device_ms) {
  API_BEGIN
  ITN_REQUIRE(ctx && host_in && host_sigma, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(dtype == ITN_F64 || dtype == ITN_C128, ITN_EUNSUPPORTED, "dtype must be 0 (Float64) or 1 (ComplexF64)");
  ITN_REQUIRE(m >= 1 && n >= 1 && n <= 256 && batch >= 0, ITN_EINVAL, "bad batch shape");
  ITN_REQUIRE(variant >= 0 && variant <= 3, ITN_EINVAL, "variant must be 0 (auto), 1 (generic kernel), 2 (round-robin m, n <= 64 kernel) or 3 (odd-even kernel, 3 CTAs per SM)");
  if (batch == 0) return ITN_OK;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const bool cplx = dtype == ITN_C128;
  const int P = cplx ? 2 : 1;
  const size_t mn = (size_t)m * n, bytes = (size_t)batch * mn * P * sizeof(double);
  DevBuf raw(ctx, bytes), in(ctx, bytes), us(ctx, bytes);
  DevBuf sig(ctx, (size_t)batch * n * sizeof(double)), perm(ctx, (size_t)batch * n * sizeof(int));
  CUDA_CHECK(cudaMemcpyAsync(raw.p, host_in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned g = (unsigned)std::min<size_t>((batch * mn + 255) / 256, 2048);
  if (cplx) k_split<true><<<g, 256, 0, ctx->stream>>>(raw.as<double>(), in.as<double>(), (int)mn, batch);
  else k_split<false><<<g, 256, 0, ctx->stream>>>(raw.as<double>(), in.as<double>(), (int)mn, batch);
  ITN_LAUNCH_CHECK(ctx);
  std::vector<SvdJob> jobs(batch);
  for (int i = 0; i < batch; ++i)
    jobs[i] = {in.as<double>() + i * mn * P, nullptr, sig.as<double>() + (size_t)i * n, perm.as<int>() + (size_t)i * n, m, n,
               us.as<double>() + i * mn * P, nullptr};
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  try {
    CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    run_jacobi(ctx, cplx, jobs, variant);
    CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
  } catch (...) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    throw;
  }
  CUDA_CHECK(cudaMemcpyAsync(host_sigma, sig.p, (size_t)batch * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (host_us) {
    if (cplx) k_merge<true><<<g, 256, 0, ctx->stream>>>(us.as<double>(), raw.as<double>(), (int)mn, batch);
    else k_merge<false><<<g, 256, 0, ctx->stream>>>(us.as<double>(), raw.as<double>(), (int)mn, batch);
    ITN_LAUNCH_CHECK(ctx);
    CUDA_CHECK(cudaMemcpyAsync(host_us, raw.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  if (se == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CUDA_CHECK(se);
  if (device_ms) *device_ms = ms;
  API_END
}

// ------------------------------------------------------------------------------------------------
// apply2: simple update on a vertex-disjoint batch of edges
//
// Multi-GPU (graph-partitioned network, every rank calls with the same list): site-level work - hermitised
// environments, the bond environment C of a side, the support projectors and the rebuild A' = A . T - runs on the rank
// that owns the site; edge-level work - R factors of both sides, theta', its SVD, truncation and the T factors - runs
// on the rank that owns esrc ("owner").  For an edge that crosses a cut the other rank ("guest") ships its C
// (n x n, 16 KiB at chi = 16) to the owner and receives its T (n x d chi') back: two grouped ncclSend/ncclRecv
// exchanges per layer plus one all-reduce that gives every rank the new bond dimensions, truncation errors and
// singular values of all gates.
// ------------------------------------------------------------------------------------------------
extern "C" int itn_apply2(itn_net* net, const int32_t* eids, int n, const void* gates, int maxdim, double cutoff,
                          int normalize, int msg_mode, int32_t* newdim_out, double* truncerr_out, double* svals_out,
                          int svals_stride) {
  API_BEGIN
  ITN_REQUIRE(net && eids && gates && n >= 0, ITN_EINVAL, "NULL argument");
  ITN_REQUIRE(msg_mode == 0 || msg_mode == 1, ITN_EINVAL, "msg_mode must be 0 (identity) or 1 (singular values)");
  ITN_REQUIRE(!net->has_bra(), ITN_EUNSUPPORTED, "gates are not defined for a bilinear form network (a bra layer is set)");
  if (n == 0) return ITN_OK;
  CUDA_CHECK(cudaSetDevice(net->ctx->device));
  itn_flush_uploads(net);
  itn_ctx* ctx = net->ctx;
  // sites on the tile path read and write the tile-major layouts only: their canonical copies may stay unwritten
  // (itn_net::canon_stale); every other reader below calls itn_canon_ensure first
  if (net->n_canon_stale && ctx->path_mode != 0) itn_canon_ensure_all(net);
  const bool cplx = net->cplx;
  const int P = net->planes();
  const bool multi = ctx->nranks > 1;
  // ---- validation (errors mirror src/apply.jl:118-129,140-144) ----
  std::vector<char> used(net->nv, 0);
  for (int i = 0; i < n; ++i) {
    const int e = eids[i];
    ITN_REQUIRE(e >= 0 && e < net->ne, ITN_EINVAL, "Vertices where the gates are being applied must be neighbors for now.");
    for (int v : {net->esrc[e], net->edst[e]}) {
      ITN_REQUIRE(!used[v], ITN_EINVAL, "a batch of two-site gates must be vertex-disjoint");
      used[v] = 1;
      if (!itn_is_local(net, v)) continue;
      ITN_REQUIRE(net->T[v].p, ITN_EINVAL, "site tensor is not set");
      for (int f : net->inc[v])
        if (f != e)
          ITN_REQUIRE(net->M[net->msg_into(v, f)].p, ITN_EINVAL,
                      "environment message into vertex " + std::to_string(v) + " is not set (update the BP cache first)");
    }
  }
  HostTrace trace("apply2");
  trace.mark("validate");
  // ---- geometry of every gate (global metadata), role of this rank ----
  enum { NONE = 0, OWNER = 1, GUEST = 2 };
  struct Geo {
    int e, v[2], k[2], d[2], nn[2], r[2], chi, m, nc, cand;
    long long X[2];
    bool loc[2];
    int role, peer;  // peer: the other rank of a cut edge, -1 otherwise
    int oi;          // index among the gates this rank owns
  };
  std::vector<Geo> geo(n);
  int max_cand = 1, n_own = 0;
  for (int i = 0; i < n; ++i) {
    Geo& g = geo[i];
    g.e = eids[i];
    g.v[0] = net->esrc[g.e];
    g.v[1] = net->edst[g.e];
    g.chi = net->edim[g.e];
    for (int s = 0; s < 2; ++s) {
      const int v = g.v[s];
      g.k[s] = net->slot(v, g.e);
      g.d[s] = net->sdim[v];
      g.nn[s] = g.d[s] * g.chi;
      g.X[s] = net->tensor_elems(v) / g.nn[s];
      g.r[s] = (int)std::min<long long>(g.X[s], g.nn[s]);
      g.loc[s] = itn_is_local(net, v);
      ITN_REQUIRE(g.nn[s] <= 256 && g.d[s] <= 8, ITN_EUNSUPPORTED, "simple update supports d <= 8 and d*chi <= 256");
    }
    g.m = g.r[0] * g.d[0];
    g.nc = g.r[1] * g.d[1];
    ITN_REQUIRE(g.nc <= 256 && g.m <= 512, ITN_EUNSUPPORTED, "simple update supports bond matrices up to 512 x 256");
    g.cand = std::min(g.m, g.nc);
    max_cand = std::max(max_cand, g.cand);
    g.role = g.loc[0] ? OWNER : (g.loc[1] ? GUEST : NONE);
    g.peer = -1;
    if (g.role == OWNER && !g.loc[1]) g.peer = net->owner[g.v[1]];
    if (g.role == GUEST) g.peer = net->owner[g.v[0]];
    g.oi = g.role == OWNER ? n_own++ : -1;
  }
  ITN_REQUIRE(!svals_out || svals_stride >= 1, ITN_EINVAL, "svals_stride must be positive");
  // result row of a gate: kept dimension, truncation error, kept singular values (never more than maxdim of them)
  const int stride = std::max(maxdim > 0 ? std::min(max_cand, maxdim) : max_cand, 1);
  const int RS = 2 + stride;  // doubles per result row
  // ---- communication segments of the cut edges (same gate order on both ranks) ----
  struct Seg {
    size_t sendC = 0, recvC = 0, sendT = 0, recvT = 0;        // doubles
    size_t sendC_off = 0, recvC_off = 0, sendT_off = 0, recvT_off = 0;
  };
  std::map<int, Seg> segs;
  auto csize = [&](const Geo& g) { return (size_t)P * g.nn[1] * g.nn[1]; };
  auto tsize = [&](const Geo& g) { return (size_t)P * g.nn[1] * g.d[1] * g.cand; };
  for (const Geo& g : geo) {
    if (g.peer < 0) continue;
    Seg& sg = segs[g.peer];
    if (g.role == GUEST) {
      sg.sendC += csize(g);
      sg.recvT += tsize(g);
    } else {
      sg.recvC += csize(g);
      sg.sendT += tsize(g);
    }
  }
  size_t tot_sendC = 0, tot_recvC = 0, tot_sendT = 0, tot_recvT = 0;
  for (auto& kv : segs) {
    Seg& sg = kv.second;
    sg.sendC_off = tot_sendC; tot_sendC += sg.sendC;
    sg.recvC_off = tot_recvC; tot_recvC += sg.recvC;
    sg.sendT_off = tot_sendT; tot_sendT += sg.sendT;
    sg.recvT_off = tot_recvT; tot_recvT += sg.recvT;
  }
  DevBuf sendC(ctx, tot_sendC * sizeof(double)), recvC(ctx, tot_recvC * sizeof(double));
  DevBuf sendT(ctx, tot_sendT * sizeof(double)), recvT(ctx, tot_recvT * sizeof(double));
  std::map<int, Seg> cur = segs;  // running offsets while carving
  for (auto& kv : cur) kv.second.sendC = kv.second.recvC = kv.second.sendT = kv.second.recvT = 0;
  // ---- sizes and workspace ----
  size_t ws_doubles = 0, env_doubles = 0, sig_total = 0, env_count = 0, env_sig = 0;
  for (const Geo& g : geo) {
    for (int s = 0; s < 2; ++s) {
      if (g.role == OWNER) {
        // C, V, R, R^+
        ws_doubles += (size_t)P * (2 * (size_t)g.nn[s] * g.nn[s] + 2 * (size_t)g.r[s] * g.nn[s]);
        sig_total += g.nn[s];
        ws_doubles += (size_t)P * g.nn[s] * g.d[s] * g.cand;  // T
      }
      if (g.loc[s])
        for (int f : net->inc[g.v[s]])
          if (f != g.e) {
            env_doubles += (size_t)P * net->edim[f] * net->edim[f];
            ++env_count;
            env_sig += net->edim[f];
          }
    }
    if (g.role == OWNER) {
      ws_doubles += (size_t)P * (4 * (size_t)g.m * g.nc + (size_t)g.nc * g.nc);  // theta0, theta, (tt0, tt), tv
      sig_total += g.nc;
    }
  }
  DevBuf ws(ctx, std::max<size_t>(ws_doubles, 1) * sizeof(double)), envs(ctx, std::max<size_t>(env_doubles, 1) * sizeof(double));
  DevBuf env_us(ctx, std::max<size_t>(env_doubles, 1) * sizeof(double)), env_v(ctx, std::max<size_t>(env_doubles, 1) * sizeof(double));
  DevBuf env_pi(ctx, std::max<size_t>(env_doubles, 1) * sizeof(double));
  DevBuf env_s(ctx, std::max<size_t>(env_sig, 1) * sizeof(double)), env_p(ctx, std::max<size_t>(env_sig, 1) * sizeof(int));
  DevBuf env_flag(ctx, std::max<size_t>(env_count, 1) * sizeof(int));
  CUDA_CHECK(cudaMemsetAsync(env_flag.p, 0, std::max<size_t>(env_count, 1) * sizeof(int), ctx->stream));
  // Cholesky fast routes (run_chol): rok[2 oi + s] = 1 when R / R^+ of that side came from the Cholesky factor of its
  // bond environment, env_ok[k] = 1 when environment k is safely positive definite (nothing for the eigen cutoff to drop)
  DevBuf rok(ctx, std::max<size_t>(2 * (size_t)n_own, 1) * sizeof(int)), env_ok(ctx, std::max<size_t>(env_count, 1) * sizeof(int));
  CUDA_CHECK(cudaMemsetAsync(rok.p, 0, std::max<size_t>(2 * (size_t)n_own, 1) * sizeof(int), ctx->stream));
  CUDA_CHECK(cudaMemsetAsync(env_ok.p, 0, std::max<size_t>(env_count, 1) * sizeof(int), ctx->stream));
  // cond[2 oi + s] = 1: the Cholesky pivots of that bond environment span more than 1 / kCholCondRatio (CholeskyQR2 below)
  DevBuf cond(ctx, std::max<size_t>(2 * (size_t)n_own, 1) * sizeof(int));
  CUDA_CHECK(cudaMemsetAsync(cond.p, 0, std::max<size_t>(2 * (size_t)n_own, 1) * sizeof(int), ctx->stream));
  struct Q2Cand {
    int gate, side, ovr;  // ovr: index into `overrides`
  };
  std::vector<Q2Cand> q2cand;
  static const bool no_q2 = getenv("ITN_NO_CHOLQR2") != nullptr;
  DevBuf sigs(ctx, std::max<size_t>(sig_total, 1) * sizeof(double)), perms(ctx, std::max<size_t>(sig_total, 1) * sizeof(int));
  // result rows of all gates: filled by the owner of each gate, summed over ranks
  DevBuf res(ctx, (size_t)n * RS * sizeof(double));
  CUDA_CHECK(cudaMemsetAsync(res.p, 0, (size_t)n * RS * sizeof(double), ctx->stream));
  // gates of the edges this rank owns: host interleaved -> planar
  size_t gate_elems = 0;
  std::vector<size_t> goff(n), gsrc(n);
  {
    size_t src_off = 0;
    for (int i = 0; i < n; ++i) {
      const size_t ge = (size_t)geo[i].d[0] * geo[i].d[1] * geo[i].d[0] * geo[i].d[1];
      gsrc[i] = src_off;
      src_off += ge;
      goff[i] = gate_elems;
      if (geo[i].role == OWNER) gate_elems += ge;
    }
  }
  DevBuf d_gates(ctx, std::max<size_t>(gate_elems, 1) * P * sizeof(double));
  std::vector<double> tmp;  // lives until the results have been read back (the stream is synchronised there)
  if (gate_elems) {
    tmp.resize(gate_elems * P);
    const double* h = (const double*)gates;
    for (int i = 0; i < n; ++i) {
      if (geo[i].role != OWNER) continue;
      const size_t ge = (size_t)geo[i].d[0] * geo[i].d[1] * geo[i].d[0] * geo[i].d[1];
      for (size_t t = 0; t < ge; ++t) {
        if (cplx) {
          tmp[goff[i] * 2 + t] = h[(gsrc[i] + t) * 2];
          tmp[goff[i] * 2 + ge + t] = h[(gsrc[i] + t) * 2 + 1];
        } else {
          tmp[goff[i] + t] = h[gsrc[i] + t];
        }
      }
    }
    CUDA_CHECK(cudaMemcpyAsync(d_gates.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  // ---- carve workspace, build jobs ----
  struct EnvRef {
    int gate, side, slot;
    const double* pi;
  };
  std::vector<EnvRef> env_refs;
  std::vector<SvdJob> ej_svd, gj, tj;
  std::vector<EigFnJob> ej_fn;
  std::vector<CholJob> chol_r, chol_env;
  // thin sides (X < n): R = A x_j S_j from the Cholesky factors of the outer messages (see k_thin_R)
  std::vector<ThinJob> thin;
  std::vector<CholJob> thin_chol_env, thin_chol_G;
  std::vector<ModeProdSpec> thin_specs;
  std::vector<double*> thin_scratch;
  size_t thin_flags_n = 0;
  for (const Geo& g : geo)
    if (g.role == OWNER)
      for (int s = 0; s < 2; ++s)
        if (g.loc[s] && g.r[s] < g.nn[s]) thin_flags_n += net->inc[g.v[s]].size() + 1;
  DevBuf thin_flags(ctx, std::max<size_t>(thin_flags_n, 1) * sizeof(int));
  CUDA_CHECK(cudaMemsetAsync(thin_flags.p, 0, std::max<size_t>(thin_flags_n, 1) * sizeof(int), ctx->stream));
  size_t thin_flag_off = 0;
  int thin_maxX = 1;
  std::vector<SuEdge> se(n_own);
  std::vector<HermJob> hj;
  std::vector<JobSpec> specs;
  std::vector<std::vector<const double*>> overrides;  // per local side: matrices per bond slot
  overrides.reserve(2 * (size_t)n);
  std::vector<FastBenvJob> fast_env;
  std::vector<char> fast_site(2 * (size_t)n, 0);
  std::vector<SuTrunc> tr(n_own);
  std::vector<double*> guestT(n, nullptr);  // guest gates: where the owner's T factor arrives
  bool any_transposed = false;
  double* w = ws.as<double>();
  double* envp = envs.as<double>();
  size_t env_sig_off = 0;
  double* sp = sigs.as<double>();
  int* pp = perms.as<int>();
  for (int i = 0; i < n; ++i) {
    const Geo& g = geo[i];
    if (g.role == NONE) continue;
    SuEdge* E = g.role == OWNER ? &se[g.oi] : nullptr;
    if (E) {
      memset(E, 0, sizeof(*E));
      E->chi = g.chi;
    }
    for (int s = 0; s < 2; ++s) {
      const int v = g.v[s];
      double* Cm = nullptr;
      if (E) {
        const size_t n2 = (size_t)g.nn[s] * g.nn[s], rn = (size_t)g.r[s] * g.nn[s];
        if (g.loc[s]) {
          Cm = w; w += P * n2;
        } else {  // the guest's bond environment arrives here
          Seg& sg = segs[g.peer];
          Cm = recvC.as<double>() + sg.recvC_off + cur[g.peer].recvC;
          cur[g.peer].recvC += csize(g);
        }
        double* Vm = w; w += P * n2;
        E->R[s] = w; w += P * rn;
        E->Rp[s] = w; w += P * rn;
        E->gus[s] = Cm;
        E->gv[s] = Vm;
        E->gsig[s] = sp;
        E->gperm[s] = pp;
        E->d[s] = g.d[s];
        E->n[s] = g.nn[s];
        E->r[s] = g.r[s];
        int* okp = rok.as<int>() + 2 * (size_t)g.oi + s;
        SvdJob gjob = {Cm, Vm, sp, pp, g.nn[s], g.nn[s], nullptr, okp};
        gj.push_back(gjob);
        E->rok[s] = okp;
        if (g.r[s] == g.nn[s] && g.nn[s] <= 64) {
          chol_r.push_back({Cm, E->R[s], E->Rp[s], okp, g.nn[s], 0.0});
          if (g.loc[s] && !no_q2) {
            chol_r.back().cond = cond.as<int>() + 2 * (size_t)g.oi + s;
            q2cand.push_back({i, s, (int)overrides.size()});  // the overrides of this side are appended just below
          }
        }
        sp += g.nn[s];
        pp += g.nn[s];
      } else if (s == 1) {  // guest: C of the local side goes to the owner
        Seg& sg = segs[g.peer];
        Cm = sendC.as<double>() + sg.sendC_off + cur[g.peer].sendC;
        cur[g.peer].sendC += csize(g);
        guestT[i] = recvT.as<double>() + sg.recvT_off + cur[g.peer].recvT;
        cur[g.peer].recvT += tsize(g);
      }
      if (!g.loc[s]) continue;
      const bool tile_site = itn_fast_gate_site_ok(net, v);
      if (!tile_site) itn_canon_ensure(net, v);
      // hermitised environments (map_eigvals symmetrises its argument, apply.jl:9-15 with ishermitian = true)
      overrides.emplace_back(net->inc[v].size(), nullptr);
      for (size_t j = 0; j < net->inc[v].size(); ++j) {
        const int f = net->inc[v][j];
        if (f == g.e) continue;
        const int c = net->edim[f];
        hj.push_back({net->M[net->msg_into(v, f)].p, envp, c});
        overrides.back()[j] = envp;
        {
          const size_t off = envp - envs.as<double>();
          double* us = env_us.as<double>() + off;
          double* vv = env_v.as<double>() + off;
          double* pi = env_pi.as<double>() + off;
          const size_t k = env_refs.size();
          SvdJob ejob = {us, vv, env_s.as<double>() + env_sig_off, env_p.as<int>() + env_sig_off, c, c, nullptr,
                         env_ok.as<int>() + k};
          ej_svd.push_back(ejob);
          ej_fn.push_back({envp, us, vv, env_s.as<double>() + env_sig_off, env_p.as<int>() + env_sig_off, pi, c,
                           env_flag.as<int>() + k, env_ok.as<int>() + k});
          // E - 100 eps tr(E) positive definite => every eigenvalue exceeds the reference's relative cutoff 10 eps
          if (c <= 64) chol_env.push_back({envp, nullptr, nullptr, env_ok.as<int>() + k, c, 100.0 * 2.220446049250313e-16});
          env_refs.push_back({i, s, (int)j, pi});
          env_sig_off += c;
        }
        envp += (size_t)P * c * c;
      }
      if (E && g.r[s] < g.nn[s] && g.X[s] <= 64) {
        bool ok = true;
        for (int f : net->inc[v])
          if (f != g.e) ok = ok && net->edim[f] <= 64;
        if (ok) {
          const int X = (int)g.X[s], nn = g.nn[s];
          const size_t x2 = (size_t)X * X;
          ModeProdSpec sp;
          memset(&sp, 0, sizeof(sp));
          sp.src = net->T[v].p;
          sp.n = net->T[v].n;
          sp.nm = (int)net->inc[v].size() + 1;
          sp.dims[0] = net->sdim[v];
          for (size_t j = 0; j < net->inc[v].size(); ++j) sp.dims[j + 1] = net->edim[net->inc[v][j]];
          int* flags = thin_flags.as<int>() + thin_flag_off;
          int n_env = 0;
          for (size_t j = 0; j < net->inc[v].size(); ++j) {
            const int f = net->inc[v][j];
            if (f == g.e) continue;
            const int c = net->edim[f];
            double* fac = (double*)itn_dev_alloc(ctx, (size_t)2 * c * c * P * sizeof(double));  // R_j and its inverse
            thin_scratch.push_back(fac);
            thin_chol_env.push_back({overrides.back()[j], fac, fac + (size_t)c * c * P, flags + n_env, c, 0.0});
            sp.mode[sp.nsteps] = (int)j + 1;
            sp.mat[sp.nsteps] = fac;
            sp.trans[sp.nsteps] = 1;  // S_j = R_j^T
            sp.nsteps++;
            ++n_env;
          }
          if (sp.nsteps > 0) {
            sp.w0 = (double*)itn_dev_alloc(ctx, (size_t)sp.n * P * sizeof(double));
            thin_scratch.push_back(sp.w0);
          }
          if (sp.nsteps > 1) {
            sp.w1 = (double*)itn_dev_alloc(ctx, (size_t)sp.n * P * sizeof(double));
            thin_scratch.push_back(sp.w1);
          }
          thin_specs.push_back(sp);
          double* gb = (double*)itn_dev_alloc(ctx, 3 * x2 * P * sizeof(double));  // G, its Cholesky factor, the inverse factor
          thin_scratch.push_back(gb);
          int* okG = flags + n_env;
          thin_chol_G.push_back({gb, gb + x2 * P, gb + 2 * x2 * P, okG, X, 0.0});
          long long lo = g.d[s];
          for (int j = 0; j < g.k[s]; ++j) lo *= net->edim[net->inc[v][j]];
          ThinJob tj = {nullptr, net->T[v].n, E->R[s], E->Rp[s], gb, gb + 2 * x2 * P, okG, flags, n_env,
                        rok.as<int>() + 2 * (size_t)g.oi + s, g.d[s], g.chi, X, nn, lo};
          thin.push_back(tj);
          thin_flag_off += (size_t)n_env + 1;
          thin_maxX = std::max(thin_maxX, X);
        }
      }
      if (tile_site) {
        // degree 4, all bonds 16, d = 2: bond environment on the DMMA tile path
        fast_env.push_back({v, g.k[s], overrides.back().data(), Cm});
        fast_site[2 * (size_t)i + s] = 1;
      } else {
        JobSpec spx;
        spx.v = v;
        spx.open_mask = 1u | (1u << (g.k[s] + 1));
        spx.out = Cm;
        spx.mats = overrides.back().data();
        specs.push_back(spx);
      }
    }
    if (!E) continue;
    E->gate = d_gates.as<double>() + goff[i] * P;
    E->theta0 = w; w += (size_t)P * g.m * g.nc;
    E->theta = w; w += (size_t)P * g.m * g.nc;
    E->tv = w; w += (size_t)P * g.nc * g.nc;
    E->tsig = sp;
    E->tperm = pp;
    E->tt0 = w; w += (size_t)P * g.m * g.nc;
    E->tt = w; w += (size_t)P * g.m * g.nc;
    // wide theta' (every heavy-hex edge whose esrc is the degree-2 site): decompose theta'^H, which has no null columns
    E->transposed = (g.m < g.nc && g.nc <= 128 && g.m <= 64) ? 1 : 0;
    any_transposed = any_transposed || E->transposed;
    {
      SvdJob tjob = {E->theta0, nullptr, sp, pp, g.m, g.nc, E->theta, nullptr};
      if (E->transposed) tjob = {E->tt0, nullptr, sp, pp, g.nc, g.m, E->tt, nullptr};
      tj.push_back(tjob);
    }
    sp += g.nc;
    pp += g.nc;
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !g.loc[1]) {  // the guest's T factor is built in the send buffer
        Seg& sg = segs[g.peer];
        E->T[s] = sendT.as<double>() + sg.sendT_off + cur[g.peer].sendT;
        cur[g.peer].sendT += tsize(g);
      } else {
        E->T[s] = w;
        w += (size_t)P * g.nn[s] * g.d[s] * g.cand;
      }
    }
    tr[g.oi] = {E->tsig, g.cand, res.as<double>() + (size_t)i * RS};
  }
  trace.mark("plan");
  // ---- 1. hermitise environments, bond environments C_side of the local sides ----
  if (!hj.empty()) {
    DevBuf hb(ctx, hj.size() * sizeof(HermJob));
    const HermJob* dh = itn_upload(ctx, hj, hb);
    if (cplx) k_hermitize<true><<<(unsigned)hj.size(), 128, 0, ctx->stream>>>(dh);
    else k_hermitize<false><<<(unsigned)hj.size(), 128, 0, ctx->stream>>>(dh);
    ITN_LAUNCH_CHECK(ctx);
  }
  itn_run_vertex_jobs(net, specs);
  itn_fast_bond_envs(net, fast_env);
  if (multi) {  // guests ship their bond environments to the owners
    std::vector<P2PSeg> xs;
    for (auto& kv : segs)
      xs.push_back({kv.first, sendC.as<double>() + kv.second.sendC_off, kv.second.sendC,
                    recvC.as<double>() + kv.second.recvC_off, kv.second.recvC});
    itn_dist_p2p(ctx, xs);
  }
  trace.mark("launch_env");
  // ---- 2. R factors (Cholesky; eigen route where r < n or C is rank deficient), environment support ----
  {
    // the three independent Cholesky batches (outer messages of thin sides, full-rank bond environments, support tests)
    // in one launch
    std::vector<CholJob> all(thin_chol_env);
    all.insert(all.end(), chol_r.begin(), chol_r.end());
    all.insert(all.end(), chol_env.begin(), chol_env.end());
    run_chol(ctx, cplx, all);
  }
  // pivot-ratio flags of the bond environments travel to the host while the rest of step 2 is enqueued (CholeskyQR2 below)
  // (page-locked destination: a copy to pageable memory would block the host right here)
  const int* hcond = nullptr;
  struct EventGuard {  // destroyed on every exit, also when a later launch throws
    cudaEvent_t ev = nullptr;
    ~EventGuard() {
      if (ev) cudaEventDestroy(ev);
    }
  } cond_guard;
  cudaEvent_t& cond_ev = cond_guard.ev;
  if (!q2cand.empty()) {
    if (ctx->pinned_flags_n < 2 * (size_t)n_own) {
      CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // an earlier copy may still target the old block
      if (ctx->pinned_flags) cudaFreeHost(ctx->pinned_flags);
      ctx->pinned_flags = nullptr;
      ctx->pinned_flags_n = 0;
      CUDA_CHECK(cudaMallocHost((void**)&ctx->pinned_flags, 2 * (size_t)n_own * sizeof(int)));
      ctx->pinned_flags_n = 2 * (size_t)n_own;
    }
    hcond = ctx->pinned_flags;
    CUDA_CHECK(cudaEventCreateWithFlags(&cond_ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaMemcpyAsync(ctx->pinned_flags, cond.p, 2 * (size_t)n_own * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaEventRecord(cond_ev, ctx->stream));
  }
  if (!thin.empty()) {
    std::vector<const double*> tres;
    itn_run_modeprods(ctx, cplx, thin_specs, tres);
    long long tmax = 0;
    for (size_t q = 0; q < thin.size(); ++q) {
      thin[q].at = tres[q];
      tmax = std::max(tmax, thin[q].n_t);
    }
    DevBuf tb(ctx, thin.size() * sizeof(ThinJob));
    const ThinJob* dt = itn_upload(ctx, thin, tb);
    const unsigned nt = (unsigned)thin.size();
    const unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((tmax + 1023) / 1024, 32));
    if (cplx) k_thin_R<true><<<dim3(nt, gy), 256, 0, ctx->stream>>>(dt);
    else k_thin_R<false><<<dim3(nt, gy), 256, 0, ctx->stream>>>(dt);
    ITN_LAUNCH_CHECK(ctx);
    if (cplx) k_thin_gram<true><<<nt, 256, 0, ctx->stream>>>(dt);
    else k_thin_gram<false><<<nt, 256, 0, ctx->stream>>>(dt);
    ITN_LAUNCH_CHECK(ctx);
    run_chol(ctx, cplx, thin_chol_G);
    // refined pseudo-inverse (k_thin_pinv2) when its operands fit shared memory, the plain one otherwise
    int thin_maxn = 1;
    for (const ThinJob& t : thin) thin_maxn = std::max(thin_maxn, t.n);
    const size_t psm2 = (size_t)P * ((size_t)thin_maxn * thin_maxX + 3 * (size_t)thin_maxX * thin_maxX) * sizeof(double);
    static const bool no_pinv2 = getenv("ITN_NO_CHOLQR2") != nullptr;
    if (psm2 <= kGlueSmemMax && !no_pinv2) {
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_thin_pinv2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm2));
        k_thin_pinv2<true><<<nt, 256, psm2, ctx->stream>>>(dt);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_thin_pinv2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm2));
        k_thin_pinv2<false><<<nt, 256, psm2, ctx->stream>>>(dt);
      }
    } else {
      const size_t psm = (size_t)P * thin_maxX * thin_maxX * sizeof(double);
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_thin_pinv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        k_thin_pinv<true><<<nt, 256, psm, ctx->stream>>>(dt);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_thin_pinv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        k_thin_pinv<false><<<nt, 256, psm, ctx->stream>>>(dt);
      }
    }
    ITN_LAUNCH_CHECK(ctx);
    for (double* p : thin_scratch) itn_dev_free(ctx, p);  // stream ordered
    thin_scratch.clear();
  }
  run_jacobi(ctx, cplx, gj);
  if (!ej_svd.empty()) {
    // projector onto the support of every environment: eigenvalues below 10 eps (relative) are dropped, as in
    // map_eigvals(sqrt / inv o sqrt, env; cutoff = 10 eps) (apply.jl:36,40-69)
    CUDA_CHECK(cudaMemcpyAsync(env_us.p, envs.p, env_doubles * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    run_jacobi(ctx, cplx, ej_svd);
    DevBuf fb(ctx, ej_fn.size() * sizeof(EigFnJob));
    const EigFnJob* de = itn_upload(ctx, ej_fn, fb);
    const double eig_cutoff = 10.0 * 2.220446049250313e-16;
    if (cplx) k_eig_fn<true><<<(unsigned)ej_fn.size(), 256, 0, ctx->stream>>>(de, 3, eig_cutoff);
    else k_eig_fn<false><<<(unsigned)ej_fn.size(), 256, 0, ctx->stream>>>(de, 3, eig_cutoff);
    ITN_LAUNCH_CHECK(ctx);
  }
  // CholeskyQR2 for ill-conditioned sides.  R from chol(A~^H A~) knows the small directions of A~ to kappa^2 eps only.
  // Where the pivots say so, the factorisation is repeated on A1 = A R1^+ (a near-isometry on the (s, l) index, formed
  // with the rebuild kernel): C2 = bond environment of A1 with the same messages, R2 = chol(C2), R = R2 R1 and
  // R^+ = R1^+ R2^+.  One small read-back decides (no device idle time: see above); well-conditioned layers pay nothing else.
  if (!q2cand.empty()) {
    // the host waits for the flags only; the device keeps working on the Jacobi / support jobs enqueued above
    CUDA_CHECK(cudaEventSynchronize(cond_ev));
    std::vector<SuSite> s2;
    std::vector<JobSpec> spec2;
    std::vector<CholJob> chol2;
    std::vector<ComposeJob> comp;
    std::vector<double*> scratch2;
    long long maxn2 = 0;
    int maxnn = 1;
    size_t nflag = 0;
    for (const Q2Cand& q : q2cand) nflag += hcond[2 * (size_t)geo[q.gate].oi + q.side] ? 1 : 0;
    DevBuf ok2(ctx, std::max<size_t>(nflag, 1) * sizeof(int));
    CUDA_CHECK(cudaMemsetAsync(ok2.p, 0, std::max<size_t>(nflag, 1) * sizeof(int), ctx->stream));
    size_t fi = 0;
    for (const Q2Cand& q : q2cand) {
      const Geo& g = geo[q.gate];
      const int sd = q.side;
      if (!hcond[2 * (size_t)g.oi + sd]) continue;
      const int v = g.v[sd], nn = g.nn[sd];
      SuEdge& E = se[g.oi];
      itn_canon_ensure(net, v);
      double* a1 = (double*)itn_dev_alloc(ctx, (size_t)net->T[v].n * P * sizeof(double));
      double* mats = (double*)itn_dev_alloc(ctx, (size_t)3 * nn * nn * P * sizeof(double));  // C2, R2, R2^+
      scratch2.push_back(a1);
      scratch2.push_back(mats);
      SuSite S;
      S.a = net->T[v].p;
      S.out = a1;
      S.T = E.Rp[sd];
      S.n_old = S.n_new = net->T[v].n;
      long long lo = g.d[sd];
      for (int j = 0; j < g.k[sd]; ++j) lo *= net->edim[net->inc[v][j]];
      S.lo = lo;
      S.hi = net->T[v].n / (lo * g.chi);
      S.d = g.d[sd];
      S.chi = S.chi_new = g.chi;
      s2.push_back(S);
      maxn2 = std::max(maxn2, S.n_new);
      JobSpec spx;
      spx.v = v;
      spx.open_mask = 1u | (1u << (g.k[sd] + 1));
      spx.out = mats;
      spx.mats = overrides[q.ovr].data();
      spx.tensor = a1;
      spec2.push_back(spx);
      const size_t n2 = (size_t)nn * nn * P;
      chol2.push_back({mats, mats + n2, mats + 2 * n2, ok2.as<int>() + fi, nn, 0.0});
      comp.push_back({E.R[sd], E.Rp[sd], mats + n2, mats + 2 * n2, ok2.as<int>() + fi, nn});
      maxnn = std::max(maxnn, nn);
      ++fi;
    }
    if (!s2.empty()) {
      ctx->cholqr2_sides += (int64_t)s2.size();
      DevBuf sb(ctx, s2.size() * sizeof(SuSite));
      const SuSite* ds = itn_upload(ctx, s2, sb);
      unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn2 + 255) / 256, 128));
      while (gy > 1 && (unsigned long long)gy * s2.size() > 148ull * 32ull) gy = (gy + 1) / 2;
      if (cplx) k_su_rebuild<true><<<dim3((unsigned)s2.size(), gy), 256, 0, ctx->stream>>>(ds);
      else k_su_rebuild<false><<<dim3((unsigned)s2.size(), gy), 256, 0, ctx->stream>>>(ds);
      ITN_LAUNCH_CHECK(ctx);
      itn_run_vertex_jobs(net, spec2);
      run_chol(ctx, cplx, chol2);
      DevBuf cb(ctx, comp.size() * sizeof(ComposeJob));
      const ComposeJob* dc = itn_upload(ctx, comp, cb);
      const size_t csm = (size_t)2 * P * maxnn * maxnn * sizeof(double);
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_compose<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        k_su_compose<true><<<(unsigned)comp.size(), 256, csm, ctx->stream>>>(dc);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_compose<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        k_su_compose<false><<<(unsigned)comp.size(), 256, csm, ctx->stream>>>(dc);
      }
      ITN_LAUNCH_CHECK(ctx);
      for (double* p : scratch2) itn_dev_free(ctx, p);  // stream ordered
    }
  }
  trace.mark("launch_R");
  // ---- 3. theta', SVD, truncation (owned gates) ----
  // staged-operand version of k_su_theta when R1, R2 and the gate of every owned edge fit (0: plain kernel)
  size_t theta_smem = 0;
  for (const Geo& g : geo) {
    if (g.role != OWNER) continue;
    const size_t need = (size_t)P * ((size_t)g.r[0] * g.nn[0] + (size_t)g.r[1] * g.nn[1] + (size_t)g.d[0] * g.d[1] * g.d[0] * g.d[1]) * sizeof(double);
    if (g.d[0] * g.d[1] > 16 || need > kGlueSmemMax) {
      theta_smem = 0;
      break;
    }
    theta_smem = std::max(theta_smem, need);
  }
  DevBuf seb(ctx, std::max<size_t>(se.size(), 1) * sizeof(SuEdge));
  const SuEdge* dse = n_own ? itn_upload(ctx, se, seb) : nullptr;
  if (n_own) {
    if (cplx) k_su_build_R<true><<<2 * n_own, 256, 0, ctx->stream>>>(dse);
    else k_su_build_R<false><<<2 * n_own, 256, 0, ctx->stream>>>(dse);
    ITN_LAUNCH_CHECK(ctx);
    if (theta_smem) {
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_theta_s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)theta_smem));
        k_su_theta_s<true><<<n_own, 256, theta_smem, ctx->stream>>>(dse);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_theta_s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)theta_smem));
        k_su_theta_s<false><<<n_own, 256, theta_smem, ctx->stream>>>(dse);
      }
    } else {
      if (cplx) k_su_theta<true><<<dim3(n_own, 8), 256, 0, ctx->stream>>>(dse);
      else k_su_theta<false><<<dim3(n_own, 8), 256, 0, ctx->stream>>>(dse);
    }
    ITN_LAUNCH_CHECK(ctx);
    if (any_transposed) {
      if (cplx) k_su_transpose<true><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
      else k_su_transpose<false><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
      ITN_LAUNCH_CHECK(ctx);
    }
    run_jacobi(ctx, cplx, tj);
    DevBuf trb(ctx, tr.size() * sizeof(SuTrunc));
    const SuTrunc* dtr = itn_upload(ctx, tr, trb);
    k_su_truncate<<<(n_own + 31) / 32, 32, 0, ctx->stream>>>(dtr, n_own, maxdim, cutoff, stride);
    ITN_LAUNCH_CHECK(ctx);
  }
  trace.mark("launch_svd");
  itn_dist_allreduce_sum(ctx, res.as<double>(), n * RS);  // every rank learns the outcome of every gate
  std::vector<double> hres((size_t)n * RS);
  CUDA_CHECK(cudaMemcpyAsync(hres.data(), res.p, hres.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int> eflag(std::max<size_t>(env_count, 1), 0);
  CUDA_CHECK(cudaMemcpyAsync(eflag.data(), env_flag.p, eflag.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  trace.mark("sync_wait");
  std::vector<int> newdim(n);
  for (int i = 0; i < n; ++i) newdim[i] = (int)(hres[(size_t)i * RS] + 0.5);
  // sites with a rank-deficient environment: A <- A x_j P_j before the rebuild
  std::vector<ModeProdSpec> pspecs;
  std::vector<std::pair<int, int>> pspec_site;  // (gate, side)
  std::vector<double*> pscratch;
  std::vector<const double*> presult;
  for (size_t k = 0; k < env_refs.size(); ++k) {
    if (!eflag[k]) continue;
    const EnvRef& r = env_refs[k];
    const int v = geo[r.gate].v[r.side];
    int idx = -1;
    for (size_t q = 0; q < pspec_site.size(); ++q)
      if (pspec_site[q] == std::make_pair(r.gate, r.side)) idx = (int)q;
    if (idx < 0) {
      itn_canon_ensure(net, v);
      ModeProdSpec sp;
      memset(&sp, 0, sizeof(sp));
      sp.src = net->T[v].p;
      sp.n = net->T[v].n;
      sp.nm = (int)net->inc[v].size() + 1;
      sp.dims[0] = net->sdim[v];
      for (size_t j = 0; j < net->inc[v].size(); ++j) sp.dims[j + 1] = net->edim[net->inc[v][j]];
      sp.w0 = (double*)itn_dev_alloc(ctx, (size_t)sp.n * P * sizeof(double));
      pscratch.push_back(sp.w0);
      sp.w1 = (double*)itn_dev_alloc(ctx, (size_t)sp.n * P * sizeof(double));
      pscratch.push_back(sp.w1);
      pspecs.push_back(sp);
      pspec_site.push_back({r.gate, r.side});
      idx = (int)pspecs.size() - 1;
    }
    ModeProdSpec& sp = pspecs[idx];
    sp.mode[sp.nsteps] = r.slot + 1;
    sp.mat[sp.nsteps] = r.pi;
    sp.nsteps++;
  }
  itn_run_modeprods(ctx, cplx, pspecs, presult);
  trace.mark("projectors");
  // ---- 4. T factors (owner), shipped to the guests; new site tensors of the local sides ----
  if (n_own) {
    for (int i = 0; i < n; ++i)
      if (geo[i].role == OWNER) se[geo[i].oi].newdim = newdim[i];
    CUDA_CHECK(cudaMemcpyAsync((void*)dse, se.data(), se.size() * sizeof(SuEdge), cudaMemcpyHostToDevice, ctx->stream));
    // staged operand sizes with the bond dimensions kept by this layer
    size_t vrec_smem = 0, T_smem = 0;
    for (int i = 0; i < n; ++i) {
      const Geo& g = geo[i];
      if (g.role != OWNER) continue;
      vrec_smem = std::max(vrec_smem, (size_t)P * ((size_t)(g.m + 1) * g.nc + (size_t)g.m * newdim[i]) * sizeof(double));
      for (int s = 0; s < 2; ++s)
        T_smem = std::max(T_smem, (size_t)P * ((size_t)g.r[s] * g.nn[s] + (size_t)g.r[s] * g.d[s] * newdim[i]) * sizeof(double));
    }
    if (vrec_smem <= kGlueSmemMax) {
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_vrec_s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vrec_smem));
        k_su_vrec_s<true><<<n_own, 256, vrec_smem, ctx->stream>>>(dse);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_vrec_s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vrec_smem));
        k_su_vrec_s<false><<<n_own, 256, vrec_smem, ctx->stream>>>(dse);
      }
    } else {
      if (cplx) k_su_vrec<true><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
      else k_su_vrec<false><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
    }
    ITN_LAUNCH_CHECK(ctx);
    if (any_transposed) {
      if (cplx) k_su_urec<true><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
      else k_su_urec<false><<<dim3(n_own, 4), 256, 0, ctx->stream>>>(dse);
      ITN_LAUNCH_CHECK(ctx);
    }
    if (T_smem <= kGlueSmemMax) {
      if (cplx) {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_T_s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_smem));
        k_su_T_s<true><<<2 * n_own, 256, T_smem, ctx->stream>>>(dse);
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_su_T_s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_smem));
        k_su_T_s<false><<<2 * n_own, 256, T_smem, ctx->stream>>>(dse);
      }
    } else {
      if (cplx) k_su_T<true><<<dim3(2 * n_own, 4), 256, 0, ctx->stream>>>(dse);
      else k_su_T<false><<<dim3(2 * n_own, 4), 256, 0, ctx->stream>>>(dse);
    }
    ITN_LAUNCH_CHECK(ctx);
  }
  if (multi) {
    std::vector<P2PSeg> xs;
    for (auto& kv : segs)
      xs.push_back({kv.first, sendT.as<double>() + kv.second.sendT_off, kv.second.sendT,
                    recvT.as<double>() + kv.second.recvT_off, kv.second.recvT});
    itn_dist_p2p(ctx, xs);
  }
  std::vector<SuSite> sites, slow_sites;
  std::vector<int> site_v;
  std::vector<FastRebuildJob> fast_reb;
  std::vector<NormJob2> nj;
  long long maxn = 0;
  // one allocation for all new site tensors of the layer, one for the new messages (shared, reference counted)
  auto align32 = [](size_t x) { return (x + 31) & ~(size_t)31; };
  size_t slab_doubles = 0, mslab_doubles = 0;
  for (int i = 0; i < n; ++i) {
    const Geo& g = geo[i];
    for (int s = 0; s < 2; ++s)
      if (g.loc[s]) slab_doubles += align32((size_t)(net->T[g.v[s]].n / g.chi * newdim[i]) * P);
    if (g.role != NONE) mslab_doubles += 2 * align32((size_t)newdim[i] * newdim[i] * P);
  }
  DevSlab* tslab = slab_doubles ? new DevSlab() : nullptr;
  DevSlab* mslab = mslab_doubles ? new DevSlab() : nullptr;
  try {
    if (tslab) tslab->base = itn_dev_alloc(ctx, slab_doubles * sizeof(double));
    if (mslab) mslab->base = itn_dev_alloc(ctx, mslab_doubles * sizeof(double));
  } catch (...) {
    if (tslab && tslab->base) itn_dev_free(ctx, tslab->base);
    delete tslab;
    delete mslab;
    for (double* p : pscratch) itn_dev_free(ctx, p);
    throw;
  }
  {
    size_t toff = 0;
    for (int i = 0; i < n; ++i) {
      const Geo& g = geo[i];
      for (int s = 0; s < 2; ++s) {
        if (!g.loc[s]) continue;
        const int v = g.v[s];
        SuSite S;
        S.a = net->T[v].p;
        for (size_t q = 0; q < pspec_site.size(); ++q)
          if (pspec_site[q] == std::make_pair(i, s)) S.a = presult[q];
        S.T = g.role == OWNER ? se[g.oi].T[s] : guestT[i];
        S.d = g.d[s];
        S.chi = g.chi;
        S.chi_new = newdim[i];
        S.n_old = net->T[v].n;
        S.n_new = net->T[v].n / g.chi * newdim[i];
        long long lo = g.d[s];
        for (int j = 0; j < g.k[s]; ++j) lo *= net->edim[net->inc[v][j]];
        S.lo = lo;
        S.hi = net->T[v].n / (lo * g.chi);
        S.out = (double*)tslab->base + toff;
        toff += align32((size_t)S.n_new * P);
        sites.push_back(S);
        site_v.push_back(v);
        if (fast_site[2 * (size_t)i + s] && S.a == net->T[v].p && newdim[i] <= 16) {
          fast_reb.push_back({v, g.k[s], newdim[i], S.T, S.out, /*lazy=*/!normalize});
        } else {
          itn_canon_ensure(net, v);
          slow_sites.push_back(S);
        }
        nj.push_back({S.out, S.n_new * P});
        maxn = std::max(maxn, S.n_new);
      }
    }
  }
  trace.mark("alloc_new");
  {
    itn_fast_rebuild(net, fast_reb);
    trace.mark("tile_rebuild");
    if (!slow_sites.empty()) {
      DevBuf sb(ctx, slow_sites.size() * sizeof(SuSite));
      const SuSite* ds = itn_upload(ctx, slow_sites, sb);
      unsigned gy = (unsigned)std::max<long long>(1, std::min<long long>((maxn + 255) / 256, 128));
      while (gy > 1 && (unsigned long long)gy * slow_sites.size() > 148ull * 32ull) gy = (gy + 1) / 2;
      if (cplx) k_su_rebuild<true><<<dim3((unsigned)slow_sites.size(), gy), 256, 0, ctx->stream>>>(ds);
      else k_su_rebuild<false><<<dim3((unsigned)slow_sites.size(), gy), 256, 0, ctx->stream>>>(ds);
      ITN_LAUNCH_CHECK(ctx);
    }
    if (normalize && !nj.empty()) {
      DevBuf nb(ctx, nj.size() * sizeof(NormJob2));
      const NormJob2* dn = itn_upload(ctx, nj, nb);
      k_normalize2<<<(unsigned)nj.size(), 256, 0, ctx->stream>>>(dn);
      ITN_LAUNCH_CHECK(ctx);
    }
  }
  for (double* p : pscratch) itn_dev_free(ctx, p);  // stream ordered: freed after the rebuild kernel
  trace.mark("launch_rebuild");
  // ---- 5. commit: swap tensors, new bond dimensions, reset the messages on the gated edges ----
  std::vector<DiagMsgJob> dj;
  size_t moff = 0;
  for (size_t si = 0; si < sites.size(); ++si) {
    const int v = site_v[si];
    itn_tensor_free(ctx, net->T[v]);
    net->T[v].p = sites[si].out;
    net->T[v].n = sites[si].n_new;
    net->T[v].slab = tslab;
    tslab->refs++;
    net->touch(v);
  }
  for (int i = 0; i < n; ++i) {
    const Geo& g = geo[i];
    net->edim[g.e] = newdim[i];  // every rank keeps the bond dimensions of the whole graph
    if (g.role == NONE) continue;
    for (int dd = 0; dd < 2; ++dd) {
      DevTensor& m = net->M[2 * g.e + dd];
      const long long n2 = (long long)newdim[i] * newdim[i];
      if (m.p) itn_tensor_free(ctx, m);
      m.p = (double*)mslab->base + moff;
      moff += align32((size_t)n2 * P);
      m.n = n2;
      m.slab = mslab;
      mslab->refs++;
      dj.push_back({m.p, msg_mode == 1 ? res.as<double>() + (size_t)i * RS + 2 : nullptr, newdim[i]});
    }
  }
  if (!dj.empty()) {
    DevBuf db(ctx, dj.size() * sizeof(DiagMsgJob));
    const DiagMsgJob* dd = itn_upload(ctx, dj, db);
    if (cplx) k_diag_msg<true><<<(unsigned)dj.size(), 128, 0, ctx->stream>>>(dd);
    else k_diag_msg<false><<<(unsigned)dj.size(), 128, 0, ctx->stream>>>(dd);
    ITN_LAUNCH_CHECK(ctx);
  }
  net->topo_version++;
  if (!normalize) itn_fast_commit_direct(net);  // k_normalize2 rescaled the canonical tensors only
  for (int i = 0; i < n; ++i) {
    if (newdim_out) newdim_out[i] = newdim[i];
    if (truncerr_out) truncerr_out[i] = hres[(size_t)i * RS + 1];
    if (svals_out)
      for (int t = 0; t < svals_stride; ++t) svals_out[(size_t)i * svals_stride + t] = t < stride ? hres[(size_t)i * RS + 2 + t] : 0.0;
  }
  trace.mark("commit");
  // no synchronisation here: everything destined for host pointers was complete at the read-back above; the rebuild of
  // the site tensors finishes in stream order while the host prepares the next call
  API_END
}

extern "C" int itn_apply_layers(itn_net* net, int nlayers, const int32_t* layer_ptr, const int32_t* eids, const void* gates,
                                int maxdim, double cutoff, int normalize, int msg_mode, const int32_t* seq_src,
                                const int32_t* seq_dst, int nseq, const int32_t* group_ptr, int ngroups, int bp_maxiter,
                                double bp_tol, int bp_normalize, int32_t* newdim_n, double* truncerr_n, double* svals,
                                int svals_stride, int32_t* bp_iters_total) {
  if (!net || nlayers < 0 || (nlayers > 0 && (!layer_ptr || !eids || !gates))) {
    itn_set_error("NULL argument");
    return ITN_EINVAL;
  }
  if (bp_iters_total) *bp_iters_total = 0;
  const size_t item = (net->cplx ? 2 : 1) * sizeof(double);
  size_t goff = 0;  // bytes into `gates`
  for (int l = 0; l < nlayers; ++l) {
    const int lo = layer_ptr[l], hi = layer_ptr[l + 1];
    if (hi < lo) {
      itn_set_error("layer_ptr must be non-decreasing");
      return ITN_EINVAL;
    }
    int st = itn_apply2(net, eids + lo, hi - lo, (const char*)gates + goff, maxdim, cutoff, normalize, msg_mode,
                        newdim_n ? newdim_n + lo : nullptr, truncerr_n ? truncerr_n + lo : nullptr,
                        svals ? svals + (size_t)lo * svals_stride : nullptr, svals_stride);
    if (st != ITN_OK) return st;
    for (int i = lo; i < hi; ++i) {
      const int e = eids[i];
      const size_t d1 = net->sdim[net->esrc[e]], d2 = net->sdim[net->edst[e]];
      goff += d1 * d2 * d1 * d2 * item;
    }
    if (bp_maxiter > 0 && nseq > 0) {
      int32_t it = 0;
      st = itn_bp_update(net, seq_src, seq_dst, nseq, group_ptr, ngroups, bp_maxiter, bp_tol, bp_normalize, &it, nullptr);
      if (st != ITN_OK) return st;
      if (bp_iters_total) *bp_iters_total += it;
    }
  }
  return ITN_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU plumbing lives in itn_dist.cu
// ------------------------------------------------------------------------------------------------
