// Block path for synchronous BP sweeps: any degree 2..8, any per-bond extent <= 32, real or complex, on the canonical
// tensor layout (no tile-major copies).
//
// Restates updated_message(::Algorithm"contract") (src/caches/abstractbeliefpropagationcache.jl:225-239) for ALL z
// outgoing messages of a vertex, with the bond set split into a fast group G1 (the first ceil(z/2) bonds, small strides)
// and a slow group G2 (the rest), and three passes over the tensor:
//
//   pass 1  P = A x_{G1} M                                              (blocks: all of G1, a chunk of G2's flat index)
//   pass 2  out_k = < P x_{G2 \ k} M | A >  for k in G2,   S = A x_{G2} M    (blocks: a chunk of (site, G1), all of G2)
//   pass 3  out_k = < S x_{G1 \ k} M | A >  for k in G1                   (blocks as in pass 1)
//
// "x_j M" is a mode product with the message arriving on bond j, "< B | A >" closes every index except bond k against the
// conjugated site tensor.  Inside a pass a CTA stages one block of the tensor(s) in shared memory ONCE (cp.async.bulk, one
// bulk copy per contiguous row, completion on an mbarrier) and runs the whole operation list of the pass on it --
// several mode products and closes per byte moved, where the shape-generic kernels of itn_generic.cu make one pass over
// HBM per mode product.  The "all but one" products inside a group share their partial products by divide and conquer
// (z = 6: 22 units of d chi^7 multiply-adds instead of the 36 of six independent updates; z = 4: 12 of 16; z = 3: 8 of 9).
//
// Arithmetic: FP64 tensor-core DMMA in the m8n8k4 shape (mma.sync.aligned.m8n8k4.f64: measured at the full 37 TFLOP/s
// issue rate on B200, tools/fp64_peak.cu).  The k = 4 shape is what makes chi = 6 affordable: a complex mode product is
// the REAL product [X_re | X_im] (fibres x 2 chi) times [[M_re, M_im], [-M_im, M_re]] (2 chi x 2 chi), so the reduction
// length is 2 chi = 12 = three k4 steps with no padding (m16n8k8 pads 12 -> 16 and 6 -> 8: 1.78x the flops; here only the
// output rows 6 -> 8 pad: 1.33x).  The message fragments sit in REGISTERS for the whole operation (the message is the A
// operand, 8 output rows x 4 reduction entries per step); the tensor is the B operand, 8 fibres per tile, read from and
// written back to shared memory.  Closes put the bond index on both M and N and reduce over fibres (k = 4 fibres per step).
//
// Shared-memory addressing is table driven: for every mode of the pass the host lists the base address of every fibre of
// the block, ordered so that the four fibres a half-warp touches together with the four reduction entries of a step fall
// into 16 different banks whenever the block geometry allows it (plan_fibres below; padding of the row and plane strides
// is part of the search).
#include <algorithm>
#include <array>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>

#include "itn_internal.h"

namespace {

constexpr int kBT = 256;       // threads per CTA (8 warps)
constexpr int kNW = kBT / 32;
constexpr int kMaxGM = 4;      // bonds per group
constexpr int kMaxOps = 24;      // z = 8: 17 operations in pass 2 (12 for the closes of four bonds, 4 + 1 for S)
constexpr int kRedDoubles = 1024;  // close: partial tiles of the fibre splits before the fixed-order sum (chi <= 16)
constexpr unsigned kNoFibre = 0xFFFFu;

enum { OP_MP = 0, OP_CLOSE = 1, OP_STORE = 2 };

struct BlkOp {
  unsigned char type, src, dst, mode;  // MP: dst = src x_mode M; CLOSE: out_mode = <src | buffer 0>; STORE: src -> Wout
};
struct BlkMode {
  int chi;    // extent of the bond
  int S;      // shared-memory stride of the bond index inside the block
  int ntile;  // fibre tiles (8 fibres each; the table is padded with kNoFibre)
  int tab;    // offset of the fibre table of this mode (entries)
  int slot;   // bond slot at the vertex (message / output index)
  int exact;  // no padding anywhere: the fibre count is a multiple of 8 and the stacked reduction length a multiple of 4
};
struct BlkPass {
  int nblk;           // blocks per vertex
  int nrows, rowlen;  // rows per plane and doubles per row moved between HBM and shared memory (a row is contiguous in both)
  long long grow;     // HBM row stride (doubles)
  long long gblk;     // HBM offset between consecutive blocks (doubles)
  long long gplane;   // HBM plane stride = elements of the tensor
  int nlev;           // shared-memory position of row r: sum over levels of digit_l(r) * lev_s[l], r in mixed radix lev_n
  int lev_n[4];       //   (padded strides keep every bond stride away from multiples of 8 doubles = the bank period / 2)
  int lev_s[4];
  int PL;             // shared-memory plane stride (doubles)
  int bufsz;          // doubles per buffer (planes * PL)
  int nbuf;
  int bulk;           // 1: cp.async.bulk per row, 0: per-thread copies (rows that are not multiples of 16 bytes)
  int load_p;         // the pass also loads the partially absorbed tensor (buffer 1)
  int nmodes;
  BlkMode modes[kMaxGM];
  int nops;
  BlkOp ops[kMaxOps];
  int tab_len;
  int rowtab;              // offset (entries) of the row-position table inside the pass table (nrows entries)
  int msg_smem;            // 1: the messages of the pass are staged in shared memory (msg_off, doubles), 0: read from L2
  int msg_off[kMaxGM];
  int msg_len;             // doubles reserved for the staged messages
  int wl;                  // 1: warp-local operation lists (every warp owns whole batch slices of the block, see k_block)
  int nclose;              // closes of the pass
  int buf_total;           // doubles reserved for the tensor buffers (warp-local: also the scratch of the final cross-warp sum)
  int red_len;             // doubles of the cross-warp scratch behind the buffers (0 in warp-local passes)
};
struct BlkVertex {
  const double* X;     // site tensor (canonical planar)
  const double* Pin;   // partially absorbed tensor read by the pass (pass 2: P, pass 3: S)
  double* Wout;        // partially absorbed tensor written by the pass (pass 1: P, pass 2: S)
  const double* msg[8];
  double* part[8];     // per bond slot: nblk partial messages (planar chi x chi each), summed by k_block_reduce
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
// m16n8k4: rows 0-7 of A (a0) feed c[0], c[1], rows 8-15 (a1) feed c[2], c[3]; columns 2t, 2t + 1 as in m8n8k4.  Complex
// products put the real-part rows in the lower half and the imaginary-part rows in the upper half: one instruction does
// the work of two m8n8k4 (all FP64 MMA shapes run at the same 37 TFLOP/s on B200, tools/fp64_peak.cu).
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(b));
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared-memory accesses by 32-bit byte address: the tile loops keep one base per reduction step / output row and add the
// fibre's byte offset, one integer add per access (generic pointers cost the compiler two to three, and it rebuilds the
// bases inside the loop when registers are short).  volatile keeps the loads of a tile ahead of its in-place stores.
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }

// Per-operation constants, computed once per CTA (one thread per operation) instead of once per warp and operation.
struct alignas(16) OpConst {  // three 16-byte loads per warp and operation
  int type, src, dst;  // buffer offsets in doubles
  int frag;            // offset of the fragment-ordered message copy in shared memory, -1: gather from global memory
  int tab, ntile, S, chi, exact, slot, mode, pad_;
};

// ---- dst = src x_mode M  (message = A operand in registers, tensor = B operand, 8 fibres per tile) ----------------------
// Complex products run as real ones on the stacked reduction index.  Two stackings:
//   packed  (KS <= 4, chi <= 8): [re 0..chi-1 | im 0..chi-1] back to back, ceil(2 chi / 4) steps: chi = 6 needs 3 steps, not 4.
//           Fragments A_re = [M_re ; -M_im] (rows of the real part of the output) and A_im = [M_im ; M_re]: one m16n8k4 per
//           step with A = (A_re, A_im).
//   aligned (KS >= 8): each plane padded to KS / 2 steps of its own.  Only M_re and M_im are kept (chi = 32: 64 registers):
//           c1 += (M_re, M_im) x X_re and c2 += (M_im, M_re) x X_im are two independent chains, and
//           out_re = c1.lo - c2.lo, out_im = c1.hi + c2.hi.
// One message-fragment value: what lane (g, t) of the warp owning row tile mt holds for reduction step ks
// (which = 0: first array, 1: second array).
template <bool C, int KS>
__device__ __forceinline__ double mp_fragment(const double* __restrict__ msg, int chi, int mt, int ks, int which, int lane) {
  constexpr bool AL = C && KS >= 8;
  const int g = lane >> 2, t = lane & 3;
  const int b = mt * 8 + g, kk = 4 * ks + t;
  const int K2 = (C && !AL) ? 2 * chi : chi;
  if (b >= chi || kk >= K2) return 0.0;
  const int chi2 = chi * chi;
  if (!C) return which ? 0.0 : msg[kk + chi * b];
  if (AL) return which ? msg[chi2 + kk + chi * b] : msg[kk + chi * b];
  if (kk < chi) return which ? msg[chi2 + kk + chi * b] : msg[kk + chi * b];
  return which ? msg[(kk - chi) + chi * b] : -msg[chi2 + (kk - chi) + chi * b];
}
// shared-memory offset of the reduction entry lane t reads in step ks, -1: beyond the end
template <bool C, int KS>
__device__ __forceinline__ int mp_koff(int chi, int S, int PL, int ks, int t) {
  constexpr bool AL = C && KS >= 8;
  const int K2 = (C && !AL) ? 2 * chi : chi;
  const int kk = 4 * ks + t;
  if (kk >= K2) return -1;
  return (!C || AL || kk < chi) ? kk * S : PL + (kk - chi) * S;
}

template <bool C, int KS, bool EXACT>
__device__ __forceinline__ void mp_tiles(const OpConst& O, const int PL, const double (&A0)[(C && KS >= 8) ? KS / 2 : KS],
                                         const double (&A1)[C ? ((KS >= 8) ? KS / 2 : KS) : 1], const int (&koff)[(C && KS >= 8) ? KS / 2 : KS],
                                         const double* src, double* dst, const unsigned short* __restrict__ tb, const int fsub,
                                         const int fstep, const int b, const int lane) {
  constexpr bool AL = C && KS >= 8;
  constexpr int KH = AL ? KS / 2 : KS;
  const int chi = O.chi, S = O.S;
  const int g = lane >> 2, t = lane & 3;
  const int K2 = (C && !AL) ? 2 * chi : chi;
  const int ksj = EXACT ? KH : ((K2 + 3) >> 2);  // exact modes fill every step of the kernel instance
  const bool bok = b < chi;
  const unsigned PL8 = 8u * (unsigned)PL;
  unsigned ak[KH];  // byte address of reduction entry (ks, lane & 3) of fibre 0
#pragma unroll
  for (int ks = 0; ks < KH; ++ks) ak[ks] = smem_u32(src) + 8u * (unsigned)(koff[ks] < 0 ? 0 : koff[ks]);
  unsigned d0 = smem_u32(dst) + 8u * (unsigned)(b * S), d1 = d0 + PL8;
  // opaque to the optimiser: otherwise it re-associates (f + b S) * 8 + base inside the loop, two instructions per store
  asm volatile("" : "+r"(d0), "+r"(d1));
  // fb: the table entry (fibre base in doubles, kNoFibre for padding)
  auto load1 = [&](unsigned fb, int ks, unsigned plane_off8) -> double {
    if (EXACT) return lds64(ak[ks] + (fb << 3) + plane_off8);
    return (koff[ks] >= 0 && fb != kNoFibre) ? lds64(ak[ks] + (fb << 3) + plane_off8) : 0.0;
  };
  // results of one tile: lane holds (row b; fibres 2t, 2t + 1)
  auto store_tile = [&](unsigned fb, double r0, double r1, double i0, double i1) {
    const unsigned f0 = __shfl_sync(0xffffffffu, fb, 8 * t), f1 = __shfl_sync(0xffffffffu, fb, 8 * t + 4);
    if (src == dst) __syncwarp();  // in place (one row tile per fibre): every lane has read its fibres
    if (bok) {
      if (EXACT || f0 != kNoFibre) {
        sts64(d0 + (f0 << 3), r0);
        if (C) sts64(d1 + (f0 << 3), i0);
      }
      if (EXACT || f1 != kNoFibre) {
        sts64(d0 + (f1 << 3), r1);
        if (C) sts64(d1 + (f1 << 3), i1);
      }
    }
  };
  if constexpr (AL) {
    for (int ft = fsub; ft < O.ntile; ft += fstep) {
      const unsigned fb = tb[ft * 8 + g];
      double xr[KH], xi[KH];
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          xr[ks] = load1(fb, ks, 0);
          xi[ks] = load1(fb, ks, PL8);
        }
      double c1[4] = {0.0, 0.0, 0.0, 0.0}, c2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          dmma1684(c1, A0[ks], A1[ks], xr[ks]);
          dmma1684(c2, A1[ks], A0[ks], xi[ks]);
        }
      store_tile(fb, c1[0] - c2[0], c1[1] - c2[1], c1[2] + c2[2], c1[3] + c2[3]);
    }
  } else if constexpr (C) {
    int ft = fsub;
    // two independent tiles per iteration: two accumulator chains in flight behind the fixed DMMA latency
    for (; ft + fstep < O.ntile; ft += 2 * fstep) {
      const unsigned fbA = tb[ft * 8 + g], fbB = tb[(ft + fstep) * 8 + g];
      double xa[KH], xb[KH];
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          xa[ks] = load1(fbA, ks, 0);
          xb[ks] = load1(fbB, ks, 0);
        }
      double ca[4] = {0.0, 0.0, 0.0, 0.0}, cb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          dmma1684(ca, A0[ks], A1[ks], xa[ks]);
          dmma1684(cb, A0[ks], A1[ks], xb[ks]);
        }
      store_tile(fbA, ca[0], ca[1], ca[2], ca[3]);
      store_tile(fbB, cb[0], cb[1], cb[2], cb[3]);
    }
    for (; ft < O.ntile; ft += fstep) {
      const unsigned fbA = tb[ft * 8 + g];
      double xa[KH];
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) xa[ks] = load1(fbA, ks, 0);
      double ca[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) dmma1684(ca, A0[ks], A1[ks], xa[ks]);
      store_tile(fbA, ca[0], ca[1], ca[2], ca[3]);
    }
  } else {
    int ft = fsub;
    for (; ft + fstep < O.ntile; ft += 2 * fstep) {
      const unsigned fbA = tb[ft * 8 + g], fbB = tb[(ft + fstep) * 8 + g];
      double xa[KH], xb[KH];
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          xa[ks] = load1(fbA, ks, 0);
          xb[ks] = load1(fbB, ks, 0);
        }
      double ca[2] = {0.0, 0.0}, cb[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) {
          dmma884(ca, A0[ks], xa[ks]);
          dmma884(cb, A0[ks], xb[ks]);
        }
      store_tile(fbA, ca[0], ca[1], 0.0, 0.0);
      store_tile(fbB, cb[0], cb[1], 0.0, 0.0);
    }
    for (; ft < O.ntile; ft += fstep) {
      const unsigned fbA = tb[ft * 8 + g];
      double xa[KH];
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) xa[ks] = load1(fbA, ks, 0);
      double ca[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KH; ++ks)
        if (ks < ksj) dmma884(ca, A0[ks], xa[ks]);
      store_tile(fbA, ca[0], ca[1], 0.0, 0.0);
    }
  }
}

// msm: the fragment-ordered message copies staged by the prologue ([row tile][step][lane][array]); skoff: the reduction
// offsets per (mode, step, lane & 3).  Without a staged copy the fragments are gathered from the message in global memory.
template <bool C, int KS>
__device__ __forceinline__ void op_mp(const OpConst& O, const int PL, const double* __restrict__ msg,
                                      const double* __restrict__ msm, const int* __restrict__ skoff, const double* bufs_c,
                                      double* bufs, const unsigned short* __restrict__ tab, const int warp, const int lane) {
  constexpr bool AL = C && KS >= 8;
  constexpr int KH = AL ? KS / 2 : KS;  // fragments per array
  const int chi = O.chi;
  const int mtj = (chi + 7) >> 3;
  const int mtl = mtj > 2 ? 2 : mtj - 1;  // warps are dealt out over 1, 2 or 4 row tiles (log2)
  const int mt = warp & ((1 << mtl) - 1), fsub = warp >> mtl, fstep = kNW >> mtl;
  if (mt >= mtj) return;
  double A0[KH], A1[C ? KH : 1];
  int koff[KH];
  if (O.frag >= 0) {
    // [row tile][step][lane][array]: both arrays of a lane in one 16-byte load
    const double* fr = msm + O.frag + (size_t)mt * KH * 64 + 2 * lane;
#pragma unroll
    for (int ks = 0; ks < KH; ++ks) {
      if constexpr (C) {
        const double2 v = *reinterpret_cast<const double2*>(fr + ks * 64);
        A0[ks] = v.x;
        A1[ks] = v.y;
      } else {
        A0[ks] = fr[ks * 64];
      }
    }
  } else {
#pragma unroll
    for (int ks = 0; ks < KH; ++ks) {
      A0[ks] = mp_fragment<C, KS>(msg, chi, mt, ks, 0, lane);
      if (C) A1[ks] = mp_fragment<C, KS>(msg, chi, mt, ks, 1, lane);
    }
  }
  const int* ko = skoff + O.mode * 64 + (lane & 3);
#pragma unroll
  for (int ks = 0; ks < KH; ++ks) koff[ks] = ko[ks * 4];
  const int b = mt * 8 + (lane >> 2);
  if (O.exact) mp_tiles<C, KS, true>(O, PL, A0, A1, koff, bufs_c + O.src, bufs + O.dst, tab + O.tab, fsub, fstep, b, lane);
  else mp_tiles<C, KS, false>(O, PL, A0, A1, koff, bufs_c + O.src, bufs + O.dst, tab + O.tab, fsub, fstep, b, lane);
}

// ---- out[b + chi b'] = sum over the fibres of the block  W[f, b] conj(X[f, b'])  (4 fibres per k step) -------------------
// Complex: ca += (W_re, W_im) x X_re, cb += (W_im, W_re) x X_im (two independent m16n8k4 chains per tile pair);
// out_re = ca.lo + cb.lo, out_im = ca.hi - cb.hi.
template <bool C, int MT>
__device__ __forceinline__ void op_close(const OpConst& O, const int PL, const double* W, const double* X, double* red,
                                         double* __restrict__ out, const unsigned short* __restrict__ tab, const int warp,
                                         const int lane, const int tid, const bool half_red) {
  constexpr int NPW = MT == 4 ? 2 : 1;
  const int chi = O.chi, S = O.S;
  const int mtj = (chi + 7) >> 3;
  const int mtd = mtj == 3 ? 4 : mtj;
  const int g = lane >> 2, t = lane & 3;
  int mt, nt0, fs, FS;
  if (mtd == 1) {
    mt = 0; nt0 = 0; fs = warp; FS = kNW;
  } else if (mtd == 2) {
    mt = (warp & 3) >> 1; nt0 = warp & 1; fs = warp >> 2; FS = 2;
  } else {
    mt = warp >> 1; nt0 = (warp & 1) * 2; fs = 0; FS = 1;
  }
  const int npw = mtd == 4 ? 2 : 1;
  const int CH = mtd * 8;
  double acc[NPW][2][2];  // [pair][re / im][column 2t, 2t + 1]
  const unsigned short* tb = tab + O.tab;
  const int nks = O.ntile * 2;
  const int bm = mt * 8 + g;
  const int bmS = bm * S;
  const bool mok = bm < chi;
  if constexpr (C) {
    double ca[NPW][4], cb[NPW][4];
#pragma unroll
    for (int q = 0; q < NPW; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) ca[q][j] = cb[q][j] = 0.0;
    const unsigned PL8 = 8u * (unsigned)PL;
    const unsigned wb = smem_u32(W) + 8u * (unsigned)bmS;
    unsigned xb[NPW];
#pragma unroll
    for (int q = 0; q < NPW; ++q) xb[q] = smem_u32(X) + 8u * (unsigned)(((nt0 + q) * 8 + g) * S);
#pragma unroll 2
    for (int ks = fs; ks < nks; ks += FS) {
      const unsigned fb = tb[ks * 4 + t];
      const unsigned f8 = fb << 3;
      const bool ok = O.exact || fb != kNoFibre;
      const bool okm = ok && mok;
      const double wre = okm ? lds64(wb + f8) : 0.0;
      const double wim = okm ? lds64(wb + PL8 + f8) : 0.0;
#pragma unroll
      for (int q = 0; q < NPW; ++q) {
        if (q < npw) {
          const int bn = (nt0 + q) * 8 + g;
          const bool okn = ok && bn < chi;
          const double xre = okn ? lds64(xb[q] + f8) : 0.0;
          const double xim = okn ? lds64(xb[q] + PL8 + f8) : 0.0;
          dmma1684(ca[q], wre, wim, xre);
          dmma1684(cb[q], wim, wre, xim);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
      acc[q][0][0] = ca[q][0] + cb[q][0];
      acc[q][0][1] = ca[q][1] + cb[q][1];
      acc[q][1][0] = ca[q][2] - cb[q][2];
      acc[q][1][1] = ca[q][3] - cb[q][3];
    }
  } else {
    double acc2[NPW][2];
#pragma unroll
    for (int q = 0; q < NPW; ++q) acc[q][0][0] = acc[q][0][1] = acc[q][1][0] = acc[q][1][1] = acc2[q][0] = acc2[q][1] = 0.0;
    auto step = [&](int ks, double (&a)[2], int q) {
      const unsigned fb = tb[ks * 4 + t];
      const bool ok = O.exact || fb != kNoFibre;
      const double w = (ok && mok) ? W[fb + bmS] : 0.0;
      const int bn = (nt0 + q) * 8 + g;
      const double x = (ok && bn < chi) ? X[fb + bn * S] : 0.0;
      dmma884(a, w, x);
    };
    int ks = fs;
    for (; ks + FS < nks; ks += 2 * FS) {
#pragma unroll
      for (int q = 0; q < NPW; ++q)
        if (q < npw) {
          step(ks, acc[q][0], q);
          step(ks + FS, acc2[q], q);
        }
    }
    if (ks < nks) {
#pragma unroll
      for (int q = 0; q < NPW; ++q)
        if (q < npw) step(ks, acc[q][0], q);
    }
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
      acc[q][0][0] += acc2[q][0];
      acc[q][0][1] += acc2[q][1];
    }
  }
  if (mtd == 4) {
    // one warp owns a pair of 8 x 8 tiles over ALL fibres of the block: no cross-warp sum, straight to the partial buffer
    const int n2 = chi * chi;
#pragma unroll
    for (int q = 0; q < NPW; ++q) {
      const int row = mt * 8 + g;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = (nt0 + q) * 8 + 2 * t + j;
        if (row < chi && col < chi) {
          out[row + chi * col] = acc[q][0][j];
          if (C) out[n2 + row + chi * col] = acc[q][1][j];
        }
      }
    }
    return;
  }
  // partial tiles of the fibre splits meet in shared memory and are summed in split order
  const int ch2 = CH * CH;
  if (half_red && mtd == 1) {
    // half-size scratch (a plan that fits three CTAs per SM with it): warps 4..7 deposit their tiles, warps 0..3 add
    // them to their own (each lane updates the elements it owns: no barrier between its read and its write), then the
    // four sums are added in warp order -- ((w0 + w4) + (w1 + w5)) + ..., fixed, hence deterministic
    double* r = red + (size_t)(warp & 3) * (C ? 2 : 1) * ch2;
    const int row = g, col = 2 * t;
    if (warp >= 4) {
      r[row + CH * col] = acc[0][0][0];
      r[row + CH * (col + 1)] = acc[0][0][1];
      if (C) {
        r[ch2 + row + CH * col] = acc[0][1][0];
        r[ch2 + row + CH * (col + 1)] = acc[0][1][1];
      }
    }
    __syncthreads();
    if (warp < 4) {
      r[row + CH * col] += acc[0][0][0];
      r[row + CH * (col + 1)] += acc[0][0][1];
      if (C) {
        r[ch2 + row + CH * col] += acc[0][1][0];
        r[ch2 + row + CH * (col + 1)] += acc[0][1][1];
      }
    }
    __syncthreads();
    const int n2 = chi * chi;
    for (int i = tid; i < (C ? 2 : 1) * n2; i += kBT) {
      const int p = i / n2, o = i - p * n2;
      const int rw = o % chi, cl = o / chi;
      double a = 0.0;
      for (int s = 0; s < 4; ++s) a += red[(size_t)(s * (C ? 2 : 1) + p) * ch2 + rw + CH * cl];
      out[i] = a;
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < NPW; ++q) {
    if (q < npw) {
      double* r = red + (size_t)fs * (C ? 2 : 1) * ch2;
      const int row = mt * 8 + g, col = (nt0 + q) * 8 + 2 * t;
      r[row + CH * col] = acc[q][0][0];
      r[row + CH * (col + 1)] = acc[q][0][1];
      if (C) {
        r[ch2 + row + CH * col] = acc[q][1][0];
        r[ch2 + row + CH * (col + 1)] = acc[q][1][1];
      }
    }
  }
  __syncthreads();
  const int n2 = chi * chi;
  for (int i = tid; i < (C ? 2 : 1) * n2; i += kBT) {
    const int p = i / n2, o = i - p * n2;
    const int row = o % chi, col = o / chi;
    double a = 0.0;
    for (int s = 0; s < FS; ++s) a += red[(size_t)(s * (C ? 2 : 1) + p) * ch2 + row + CH * col];
    out[i] = a;
  }
}

template <bool C>
__device__ __forceinline__ void op_close_wl(const OpConst& O, const int PL, const double* W, const double* X,
                                            const unsigned short* __restrict__ tab, const int warp, const int lane,
                                            double (&r)[C ? 4 : 2]) {
  const int chi = O.chi;
  const int g = lane >> 2, t = lane & 3;
  const unsigned short* tb = tab + O.tab;
  const bool mok = g < chi;
  const int gS = g * O.S;
  if constexpr (C) {
    const unsigned PL8 = 8u * (unsigned)PL;
    const unsigned wb = smem_u32(W) + 8u * (unsigned)gS, xb = smem_u32(X) + 8u * (unsigned)gS;
    double ca[4] = {0.0, 0.0, 0.0, 0.0}, cb[4] = {0.0, 0.0, 0.0, 0.0};
    for (int ft = warp; ft < O.ntile; ft += kNW) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned fb = tb[ft * 8 + h * 4 + t];
        const unsigned f8 = fb << 3;
        const bool ok = (O.exact || fb != kNoFibre) && mok;
        const double wre = ok ? lds64(wb + f8) : 0.0;
        const double wim = ok ? lds64(wb + PL8 + f8) : 0.0;
        const double xre = ok ? lds64(xb + f8) : 0.0;
        const double xim = ok ? lds64(xb + PL8 + f8) : 0.0;
        dmma1684(ca, wre, wim, xre);
        dmma1684(cb, wim, wre, xim);
      }
    }
    r[0] = ca[0] + cb[0];
    r[1] = ca[1] + cb[1];
    r[2] = ca[2] - cb[2];
    r[3] = ca[3] - cb[3];
  } else {
    double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
    for (int ft = warp; ft < O.ntile; ft += kNW) {
      const unsigned f0 = tb[ft * 8 + t], f1 = tb[ft * 8 + 4 + t];
      const bool ok0 = (O.exact || f0 != kNoFibre) && mok, ok1 = (O.exact || f1 != kNoFibre) && mok;
      const double w0 = ok0 ? W[f0 + gS] : 0.0, x0 = ok0 ? X[f0 + gS] : 0.0;
      const double w1 = ok1 ? W[f1 + gS] : 0.0, x1 = ok1 ? X[f1 + gS] : 0.0;
      dmma884(a0, w0, x0);
      dmma884(a1, w1, x1);
    }
    r[0] = a0[0] + a1[0];
    r[1] = a0[1] + a1[1];
  }
}

constexpr int kMaxClose = 4;  // closes per pass (z = 8: four bonds per group)

// WL (warp-local, every extent of the vertex <= 8): the host deals the batch slices of the block -- the values of the
// indices no operation of the pass contracts (site index and chunk index in passes 1 / 3, the chunk of the flat
// (site, G1) index in pass 2) -- out to the 8 warps and orders the fibre tables so that tile ft belongs to warp ft % 8 and
// touches only that warp's slices.  Every mode product then reads what the same warp wrote: the operation list runs with
// __syncwarp between operations instead of a CTA barrier after each, products run in place (fewer buffers, larger blocks),
// closes stay in registers, and ONE cross-warp sum per pass (scratch aliased on the dead tensor buffers) ends the CTA.
template <bool C, int KS, int MT, int NB, bool WL>
__global__ void __launch_bounds__(kBT, NB) k_block(const __grid_constant__ BlkPass P, const BlkVertex* __restrict__ gv,
                                                  const unsigned short* __restrict__ gtab) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ OpConst sOp[kMaxOps];
  __shared__ int sKoff[kMaxGM * 64];  // [mode][step (16)][lane & 3]
  __shared__ int sCloseOp[kMaxClose];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vi = blockIdx.x / P.nblk, blk = blockIdx.x - vi * P.nblk;
  const BlkVertex* __restrict__ V = gv + vi;
  double* bufs = sm;
  double* red = bufs + P.buf_total;
  double* msm = red + P.red_len;                                      // staged messages (P.msg_smem)
  unsigned short* tab = (unsigned short*)(msm + P.msg_len);
  constexpr int PLN = C ? 2 : 1;
  const unsigned mb = smem_u32(&mbar);
  if (P.bulk == 1 && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // per-operation constants and per-mode reduction offsets
  if (tid < P.nops) {
    const BlkOp op = P.ops[tid];
    const BlkMode M = P.modes[op.mode];
    OpConst o;
    o.type = op.type;
    o.src = op.src * P.bufsz;
    o.dst = op.dst * P.bufsz;
    o.frag = P.msg_smem ? P.msg_off[op.mode] : -1;
    o.tab = M.tab;
    o.ntile = M.ntile;
    o.S = M.S;
    o.chi = M.chi;
    o.exact = M.exact;
    o.slot = M.slot;
    o.mode = op.mode;
    // A close ends with all its reads of the tensor buffers behind a barrier of its own; its trailing sum reads the
    // cross-warp scratch only, and the barrier that ends the following mode product separates it from the next close
    // (extents above 16 close without a cross-warp sum and without a barrier of their own: they keep the trailing one)
    o.pad_ = (op.type == OP_CLOSE && M.chi <= 16 && (tid + 1 == P.nops || P.ops[tid + 1].type == OP_MP)) ? 0 : 1;  // barrier after the op
    sOp[tid] = o;
  }
  if (WL && tid == 0) {
    int c = 0;
    for (int oi = 0; oi < P.nops; ++oi)
      if (P.ops[oi].type == OP_CLOSE && c < kMaxClose) sCloseOp[c++] = oi;
  }
  if (tid < kMaxGM * 64) {
    const int k = tid >> 6, ks = (tid >> 2) & 15, t = tid & 3;
    sKoff[tid] = (k < P.nmodes && ks < KS) ? mp_koff<C, KS>(P.modes[k].chi, P.modes[k].S, P.PL, ks, t) : -1;
  }
  __syncthreads();
  // ---- stage the block(s): buffer 0 = site tensor, buffer 1 = partially absorbed tensor.  Long rows travel as bulk
  // copies (cp.async.bulk, completion on the mbarrier; the copy engine takes one row per instruction), short rows as
  // 16-byte cp.async of all threads in parallel.  Fibre tables and messages follow while the copies are in flight. -----------
  const long long goff = (long long)blk * P.gblk;
  const int nin = P.load_p ? 2 : 1;
  const double* gX = V->X;
  const double* gP = P.load_p ? V->Pin : nullptr;
  if (P.bulk == 1) {
    const unsigned rowbytes = (unsigned)P.rowlen * 8u;
    if (tid == 0) {
      const unsigned total = rowbytes * (unsigned)(nin * PLN * P.nrows);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(total) : "memory");
    }
    for (int row = tid; row < P.nrows; row += kBT) {
      const int pos = gtab[P.rowtab + row];
      for (int which = 0; which < nin; ++which)
        for (int pl = 0; pl < PLN; ++pl) {
          const double* g = (which ? gP : gX) + (long long)pl * P.gplane + goff + (long long)row * P.grow;
          const unsigned dsts = smem_u32(bufs + (size_t)which * P.bufsz + (size_t)pl * P.PL + pos);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dsts),
                       "l"(g), "r"(rowbytes), "r"(mb)
                       : "memory");
        }
    }
  } else if (P.bulk == 2) {
    // 16-byte chunks: chunk c of row r, rows of `half` chunks; lanes walk the chunks of consecutive rows
    const int half = P.rowlen >> 1;
    const int tot = P.nrows * half;
    // (row, c) of chunk i advance by (kBT / half, kBT % half) per iteration: one division per thread, not per chunk
    const int drow = kBT / half, dc = kBT - drow * half;
    int row = tid / half, c = tid - row * half;
    const unsigned sb = smem_u32(bufs);
    for (int i = tid; i < tot; i += kBT) {
      const int pos = gtab[P.rowtab + row] + 2 * c;
      const long long go = goff + (long long)row * P.grow + 2 * c;
      for (int which = 0; which < nin; ++which)
        for (int pl = 0; pl < PLN; ++pl) {
          const double* g = (which ? gP : gX) + (long long)pl * P.gplane + go;
          const unsigned dsts = sb + 8u * (unsigned)(which * P.bufsz + pl * P.PL + pos);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dsts), "l"(g) : "memory");
        }
      row += drow;
      c += dc;
      if (c >= half) {
        c -= half;
        ++row;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  } else {
    const long long tot = (long long)nin * PLN * P.nrows * P.rowlen;
    for (long long i = tid; i < tot; i += kBT) {
      const int r = (int)(i / P.rowlen), c = (int)(i - (long long)r * P.rowlen);
      const int which = r / (PLN * P.nrows), rr = r - which * (PLN * P.nrows);
      const int pl = rr / P.nrows, row = rr - pl * P.nrows;
      bufs[(size_t)which * P.bufsz + (size_t)pl * P.PL + gtab[P.rowtab + row] + c] =
          (which ? gP : gX)[(long long)pl * P.gplane + goff + (long long)row * P.grow + c];
    }
  }
  for (int i = tid; i < P.tab_len; i += kBT) tab[i] = gtab[i];
  if (P.msg_smem) {
    // fragment-ordered copies of the messages: [row tile][step][lane][array], what op_mp keeps in registers per operation
    constexpr int KHs = (C && KS >= 8) ? KS / 2 : KS;
#pragma unroll
    for (int k = 0; k < kMaxGM; ++k)
      if (k < P.nmodes) {
        const double* __restrict__ gm = V->msg[P.modes[k].slot];
        const int chi = P.modes[k].chi;
        const int n = ((chi + 7) >> 3) * KHs * 64;
        double* d = msm + P.msg_off[k];
        for (int i = tid; i < n; i += kBT) {
          const int which = i & 1, ln = (i >> 1) & 31, r = i >> 6;
          const int ks = r % KHs, mt = r / KHs;
          d[i] = mp_fragment<C, KS>(gm, chi, mt, ks, which, ln);
        }
      }
  }
  if (P.bulk == 1) {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(mb), "r"(0u)
          : "memory");
    }
  } else if (P.bulk == 2) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  }
  __syncthreads();
  // ---- the operation list of the pass --------------------------------------------------------------------------------
  constexpr int NR = C ? 4 : 2;
  double hacc[WL ? kMaxClose : 1][NR];
#pragma unroll
  for (int c = 0; c < (WL ? kMaxClose : 1); ++c)
#pragma unroll
    for (int j = 0; j < NR; ++j) hacc[c][j] = 0.0;
  int ci = 0;
  for (int oi = 0; oi < P.nops; ++oi) {
    const OpConst O = sOp[oi];
    if (O.type == OP_MP) {
      op_mp<C, KS>(O, P.PL, V->msg[O.slot], msm, sKoff, bufs, bufs, tab, warp, lane);
      if (WL) __syncwarp();
    } else if (O.type == OP_CLOSE) {
      if constexpr (WL) {
        double r[NR];
        op_close_wl<C>(O, P.PL, bufs + O.src, bufs, tab, warp, lane, r);
#pragma unroll
        for (int c = 0; c < kMaxClose; ++c)
          if (c == ci) {
#pragma unroll
            for (int j = 0; j < NR; ++j) hacc[c][j] = r[j];
          }
        ++ci;
      } else {
        double* out = V->part[O.slot] + (size_t)blk * PLN * O.chi * O.chi;
        op_close<C, MT>(O, P.PL, bufs + O.src, bufs, red, out, tab, warp, lane, tid, P.red_len < kRedDoubles);
      }
    } else {
      if (WL) __syncthreads();  // rows of the store cross the warps' slices
      const double* b = bufs + O.src;
      double* gW = V->Wout;
      const unsigned short* rt = tab + P.rowtab;
      if (P.bulk) {
        // 16-byte moves, row positions from the table: a warp per row (long rows) or several rows per warp (short rows)
        const int half = P.rowlen >> 1;
        if (half >= 32) {
          for (int pl = 0; pl < PLN; ++pl)
            for (int row = warp; row < P.nrows; row += kNW) {
              const double* sp = b + (size_t)pl * P.PL + rt[row];
              double* gp = gW + (long long)pl * P.gplane + goff + (long long)row * P.grow;
              for (int c = lane; c < half; c += 32)
                *reinterpret_cast<double2*>(gp + 2 * c) = *reinterpret_cast<const double2*>(sp + 2 * c);
            }
        } else {
          const int rpw = 32 / half, lr = lane / half, lc = lane - lr * half;
          for (int pl = 0; pl < PLN; ++pl)
            for (int row0 = warp * rpw; row0 < P.nrows; row0 += kNW * rpw) {
              const int row = row0 + lr;
              if (lr < rpw && row < P.nrows) {
                const double* sp = b + (size_t)pl * P.PL + rt[row];
                double* gp = gW + (long long)pl * P.gplane + goff + (long long)row * P.grow;
                *reinterpret_cast<double2*>(gp + 2 * lc) = *reinterpret_cast<const double2*>(sp + 2 * lc);
              }
            }
        }
      } else {
        const long long tot = (long long)PLN * P.nrows * P.rowlen;
        for (long long i = tid; i < tot; i += kBT) {
          const int r = (int)(i / P.rowlen), c = (int)(i - (long long)r * P.rowlen);
          const int pl = r / P.nrows, row = r - pl * P.nrows;
          gW[(long long)pl * P.gplane + goff + (long long)row * P.grow + c] = b[(size_t)pl * P.PL + rt[row] + c];
        }
      }
    }
    if (!WL && O.pad_) __syncthreads();
  }
  if constexpr (WL) {
    // one cross-warp sum for all closes of the pass: [close][warp][plane][8 x 8], on the (dead) tensor buffers
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int c = 0; c < kMaxClose; ++c)
      if (c < P.nclose) {
        double* q = bufs + (size_t)(c * kNW + warp) * (PLN * 64) + g + 16 * t;
        q[0] = hacc[c][0];
        q[8] = hacc[c][1];
        if (C) {
          q[64] = hacc[c][2];
          q[72] = hacc[c][3];
        }
      }
    __syncthreads();
    for (int c = 0; c < P.nclose; ++c) {
      const OpConst O = sOp[sCloseOp[c]];
      const int chi = O.chi, n2 = chi * chi;
      double* out = V->part[O.slot] + (size_t)blk * PLN * n2;
      for (int i = tid; i < PLN * n2; i += kBT) {
        const int p = i / n2, o = i - p * n2;
        const int col = o / chi, row = o - col * chi;
        const double* q = bufs + (size_t)c * kNW * (PLN * 64) + p * 64 + row + 8 * col;
        double a = 0.0;
#pragma unroll
        for (int s = 0; s < kNW; ++s) a += q[s * (PLN * 64)];
        out[i] = a;
      }
    }
  }
}


// staged message = sum of the per-block partials, in block order (deterministic)
struct BlkRedJob {
  const double* part;
  double* out;
  int nblk, n;  // n = planes * chi^2
};
__global__ void __launch_bounds__(128) k_block_reduce(const BlkRedJob* __restrict__ jobs) {
  const BlkRedJob J = jobs[blockIdx.x];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < J.n; i += gridDim.y * blockDim.x) {
    double a = 0.0;
    for (int b = 0; b < J.nblk; ++b) a += J.part[(size_t)b * J.n + i];
    J.out[i] = a;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// host side: buckets, block geometry, fibre tables, operation lists
// ------------------------------------------------------------------------------------------------------------------------
struct Signature {
  int d, z;
  std::array<int, 8> chi;
  bool operator<(const Signature& o) const { return std::tie(d, z, chi) < std::tie(o.d, o.z, o.chi); }
};

struct ModeGeom {
  int chi, S, slot;
  std::vector<int> bases;  // base address of every fibre of the block
};

inline int mod16(int x) { return ((x % 16) + 16) % 16; }

// Orders the fibres of one mode into tiles of 8 so that the shared-memory accesses of op_mp / op_close are conflict free
// when the geometry allows it.  Access patterns, per half-warp (16 lanes of 8 bytes = one wavefront when all banks differ):
//   loads   fibres n = 0..3 (and 4..7) of a tile  x  the four reduction entries of a k step  (op_mp B operand; op_close
//           reads the same quads with the roles of fibre and entry exchanged)
//   stores  fibres n = 0, 2, 4, 6 (and 1, 3, 5, 7) of a tile  x  four consecutive output rows
// Returns the table (padded with kNoFibre) and the number of excess wavefronts it could not avoid.
int plan_fibres(const ModeGeom& mg, bool cplx, bool aligned, int PL, std::vector<unsigned short>& table) {
  const int chi = mg.chi, S = mg.S;
  const int K2 = (cplx && !aligned) ? 2 * chi : chi;
  const int ksj = (K2 + 3) / 4;
  std::vector<std::array<int, 4>> pats;  // bank offsets of the four entries of every k step (-1: not loaded)
  for (int ks = 0; ks < ksj; ++ks) {
    std::array<int, 4> p;
    for (int t = 0; t < 4; ++t) {
      const int kk = 4 * ks + t;
      p[t] = kk >= K2 ? -1 : mod16((!cplx || aligned || kk < chi) ? kk * S : PL + (kk - chi) * S);
    }
    pats.push_back(p);
  }
  std::array<int, 4> spat;  // stores and close loads: four consecutive bond indices
  for (int b = 0; b < 4; ++b) spat[b] = b < chi ? mod16(b * S) : -1;
  auto cost_pat = [&](const int* r4, const std::array<int, 4>& p) {
    int cnt[16] = {0}, mx = 0;
    for (int i = 0; i < 4; ++i)
      if (r4[i] >= 0)
        for (int t = 0; t < 4; ++t)
          if (p[t] >= 0) mx = std::max(mx, ++cnt[(r4[i] + p[t]) & 15]);
    return std::max(0, mx - 1);
  };
  auto load_cost = [&](const int* r4) {
    int c = cost_pat(r4, spat);
    for (const auto& p : pats) c += cost_pat(r4, p);
    return c;
  };
  auto store_cost = [&](const int* r4) { return cost_pat(r4, spat); };
  const int nf = (int)mg.bases.size();
  std::vector<std::vector<int>> cls(16);
  for (int f = nf - 1; f >= 0; --f) cls[mod16(mg.bases[f])].push_back(f);
  // quad types: 4 residues (first = 0) with conflict-free loads, realised greedily at every shift
  std::vector<std::array<int, 4>> quads;  // fibre ids
  for (int b = 0; b < 16; ++b)
    for (int c = b; c < 16; ++c)
      for (int d = c; d < 16; ++d) {
        int r4[4] = {0, b, c, d};
        if (load_cost(r4) != 0) continue;
        for (int sft = 0; sft < 16; ++sft) {
          int need[16] = {0};
          for (int i = 0; i < 4; ++i) need[(r4[i] + sft) & 15]++;
          while (true) {
            bool ok = true;
            for (int r = 0; r < 16; ++r) ok = ok && (int)cls[r].size() >= need[r];
            if (!ok) break;
            std::array<int, 4> q;
            for (int i = 0; i < 4; ++i) {
              auto& cl = cls[(r4[i] + sft) & 15];
              q[i] = cl.back();
              cl.pop_back();
            }
            quads.push_back(q);
          }
        }
      }
  // leftovers: arbitrary quads (conflicts counted below)
  std::vector<int> left;
  for (auto& c : cls) left.insert(left.end(), c.begin(), c.end());
  std::sort(left.begin(), left.end(), std::greater<int>());
  while (!left.empty()) {
    std::array<int, 4> q = {-1, -1, -1, -1};
    for (int i = 0; i < 4 && !left.empty(); ++i) {
      q[i] = left.back();
      left.pop_back();
    }
    quads.push_back(q);
  }
  auto res = [&](int f) { return f < 0 ? -1 : mod16(mg.bases[f]); };
  // pair quads into tiles [A0 A1 A2 A3 | B0 B1 B2 B3]; the order inside A and B is chosen so that the store quads
  // {A0, A2, B0, B2} and {A1, A3, B1, B3} are conflict free too when possible.  Costs are memoised per residue pattern.
  static const int perms[24][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 1, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {0, 3, 2, 1},
                                   {1, 0, 2, 3}, {1, 0, 3, 2}, {1, 2, 0, 3}, {1, 2, 3, 0}, {1, 3, 0, 2}, {1, 3, 2, 0},
                                   {2, 0, 1, 3}, {2, 0, 3, 1}, {2, 1, 0, 3}, {2, 1, 3, 0}, {2, 3, 0, 1}, {2, 3, 1, 0},
                                   {3, 0, 1, 2}, {3, 0, 2, 1}, {3, 1, 0, 2}, {3, 1, 2, 0}, {3, 2, 0, 1}, {3, 2, 1, 0}};
  auto key_of = [&](const std::array<int, 4>& q) {
    int k = 0;
    for (int i = 0; i < 4; ++i) k = k * 17 + (res(q[i]) + 1);
    return k;
  };
  std::map<int, std::vector<int>> by_key;  // residue pattern -> quads not paired yet
  for (int i = (int)quads.size() - 1; i >= 0; --i) by_key[key_of(quads[i])].push_back(i);
  struct Best {
    int cost, pa, pb;
  };
  std::map<std::pair<int, int>, Best> memo;
  auto pair_cost = [&](int qa, int qb) {
    const std::pair<int, int> k(key_of(quads[qa]), key_of(quads[qb]));
    auto it = memo.find(k);
    if (it != memo.end()) return it->second;
    Best best{1 << 30, 0, 0};
    for (int pa = 0; pa < 24 && best.cost > 0; ++pa)
      for (int pb = 0; pb < 24; ++pb) {
        int ra[4] = {res(quads[qa][perms[pa][0]]), res(quads[qa][perms[pa][2]]), res(quads[qb][perms[pb][0]]), res(quads[qb][perms[pb][2]])};
        int rb[4] = {res(quads[qa][perms[pa][1]]), res(quads[qa][perms[pa][3]]), res(quads[qb][perms[pb][1]]), res(quads[qb][perms[pb][3]])};
        const int c = store_cost(ra) + store_cost(rb);
        if (c < best.cost) {
          best = {c, pa, pb};
          if (c == 0) break;
        }
      }
    memo[k] = best;
    return best;
  };
  int excess = 0;
  table.clear();
  std::vector<char> used(quads.size(), 0);
  for (size_t a = 0; a < quads.size(); ++a) {
    if (used[a]) continue;
    used[a] = 1;
    auto& own = by_key[key_of(quads[a])];
    own.erase(std::find(own.begin(), own.end(), (int)a));
    int partner = -1;
    Best pb{1 << 30, 0, 0};
    for (auto& kv : by_key) {
      if (kv.second.empty()) continue;
      const Best c = pair_cost((int)a, kv.second.back());
      if (c.cost < pb.cost) {
        pb = c;
        partner = kv.second.back();
        if (c.cost == 0) break;
      }
    }
    int f8[8];
    for (int i = 0; i < 4; ++i) {
      f8[i] = quads[a][partner >= 0 ? perms[pb.pa][i] : i];
      f8[4 + i] = partner >= 0 ? quads[partner][perms[pb.pb][i]] : -1;
    }
    if (partner >= 0) {
      used[partner] = 1;
      auto& pv = by_key[key_of(quads[partner])];
      pv.erase(std::find(pv.begin(), pv.end(), partner));
    }
    for (int i = 0; i < 8; ++i) table.push_back(f8[i] < 0 ? (unsigned short)kNoFibre : (unsigned short)mg.bases[f8[i]]);
    int la[4] = {res(f8[0]), res(f8[1]), res(f8[2]), res(f8[3])}, lb[4] = {res(f8[4]), res(f8[5]), res(f8[6]), res(f8[7])};
    int sa[4] = {res(f8[0]), res(f8[2]), res(f8[4]), res(f8[6])}, sb[4] = {res(f8[1]), res(f8[3]), res(f8[5]), res(f8[7])};
    excess += load_cost(la) + load_cost(lb) + store_cost(sa) + store_cost(sb);
  }
  return excess;
}

struct PassPlan {
  BlkPass desc;
  std::vector<unsigned short> table;
  size_t smem = 0;
  int excess = 0;
  double eff = 1.0;  // warp-local plans: real fibres / fibres of the padded per-warp tiles (averaged over the modes)
  int chunk = 0, split = 0, pad_chunk = 0, pad_plane = 0;  // the PassChoice this plan was made with
};

// Emits "all but one" closes for the bond set [lo, hi) of the pass (local mode indices) from buffer T by divide and conquer.
struct Emitter {
  BlkPass* P;
  std::vector<char> busy;  // buffers in use
  bool ok = true;
  int maxbuf = 0;
  int alloc(int not_this = -1) {
    for (int i = 1; i < (int)busy.size(); ++i)
      if (!busy[i] && i != not_this) {
        busy[i] = 1;
        maxbuf = std::max(maxbuf, i + 1);
        return i;
      }
    ok = false;
    return 1;
  }
  void put(unsigned char type, int src, int dst, int mode) {
    if (P->nops >= kMaxOps) {
      ok = false;
      return;
    }
    P->ops[P->nops++] = {type, (unsigned char)src, (unsigned char)dst, (unsigned char)mode};
  }
  // absorbs modes [a_lo, a_hi) into T; returns the buffer holding the result (T itself when the range is empty)
  int absorb(int T, int a_lo, int a_hi) {
    int cur = T;
    for (int k = a_lo; k < a_hi; ++k) {
      const int dst = alloc();
      put(OP_MP, cur, dst, k);
      if (cur != T) busy[cur] = 0;
      cur = dst;
    }
    return cur;
  }
  void solve(int T, int lo, int hi) {
    if (hi - lo == 1) {
      put(OP_CLOSE, T, 0, lo);
      return;
    }
    const int mid = (lo + hi) / 2;
    const int t1 = absorb(T, mid, hi);
    solve(t1, lo, mid);
    if (t1 != T) busy[t1] = 0;
    const int t2 = absorb(T, lo, mid);
    solve(t2, mid, hi);
    if (t2 != T) busy[t2] = 0;
  }
};

// Warp-local passes: one row tile per fibre (every extent <= 8), so a mode product may run in place.  A buffer the
// caller does not need any more ("owned") is overwritten; the first product on a buffer that must survive goes to a new one.
struct EmitterWL {
  BlkPass* P;
  std::vector<char> busy;
  bool ok = true;
  int maxbuf = 0;
  int alloc() {
    for (int i = 1; i < (int)busy.size(); ++i)
      if (!busy[i]) {
        busy[i] = 1;
        maxbuf = std::max(maxbuf, i + 1);
        return i;
      }
    ok = false;
    return 1;
  }
  void put(unsigned char type, int src, int dst, int mode) {
    if (P->nops >= kMaxOps) {
      ok = false;
      return;
    }
    P->ops[P->nops++] = {type, (unsigned char)src, (unsigned char)dst, (unsigned char)mode};
  }
  int absorb(int T, int a_lo, int a_hi, bool owned) {
    int cur = T;
    bool mine = owned;
    for (int k = a_lo; k < a_hi; ++k) {
      if (mine) {
        put(OP_MP, cur, cur, k);
      } else {
        const int dst = alloc();
        put(OP_MP, cur, dst, k);
        cur = dst;
        mine = true;
      }
    }
    return cur;
  }
  void solve(int T, int lo, int hi, bool owned) {
    if (hi - lo == 1) {
      put(OP_CLOSE, T, 0, lo);
      return;
    }
    const int mid = (lo + hi) / 2;
    const int t1 = absorb(T, mid, hi, false);
    solve(t1, lo, mid, true);
    busy[t1] = 0;
    const int t2 = absorb(T, lo, mid, owned);
    solve(t2, mid, hi, true);
    if (t2 != T) busy[t2] = 0;
  }
};

// dynamic shared memory that still lets two CTAs share an SM: 228 KB per SM, 1 KB reserved per CTA, and the kernel's
// 2176 bytes of static shared memory (sOp, sKoff, mbar) -- 113 KB, the round-2 value, silently dropped the chi = 8 pass 2
// and the chi = 16 pass 3 (114.6 / 113.3 KB) to one CTA per SM
bool block_debug() {
  static const bool on = getenv("ITN_BLOCK_DEBUG") != nullptr;
  return on;
}

constexpr size_t kStaticSmem = 2304;  // ptxas: sOp, sKoff, sCloseOp, mbar
constexpr size_t kSmemTwoCtasMax = 228 * 1024 / 2 - 1024 - kStaticSmem;
constexpr size_t kSmemThreeCtasMax = 228 * 1024 / 3 - 1024 - kStaticSmem;
constexpr size_t kSmemOneCta = 227 * 1024 - kStaticSmem;
// Resident CTAs per SM the planner sizes the blocks for.  The instances with at most four k4 steps (chi <= 8 complex,
// chi <= 16 real) also exist in a <= 80 register build (__launch_bounds__(256, 3)) that lets three CTAs share an SM when
// the block is sized for a third of the shared memory; ITN_BLOCK_CTAS = 3 selects that sizing (experiments).
int target_ctas(int ks_inst) {
  static const int env = getenv("ITN_BLOCK_CTAS") ? atoi(getenv("ITN_BLOCK_CTAS")) : 0;
  if (ks_inst > 4) return 2;
  // measured on B200 with each variant in its own instance (NB = 2: 100 / 113 registers, NB = 3: 79 / 80): 16^3 chi = 6
  // sweep 28.5 ms with two CTAs of large blocks, 30.1 ms with three CTAs of small ones; 32 x 32 chi = 8: 0.578 / 0.581 ms
  return env == 3 ? 3 : 2;
}
thread_local size_t kSmemTwoCtas = kSmemTwoCtasMax;  // budget of the plan being made (set by choose_pass)

size_t pass_smem(const BlkPass& P, size_t tab_len) {
  return ((size_t)P.buf_total + P.red_len + P.msg_len) * sizeof(double) + ((tab_len * 2 + 15) & ~(size_t)15) + 64;
}

// k * S mod 16 distinct for the (up to) four consecutive bond indices a half-warp touches together
bool good_stride(long long S, int chi) {
  const int n = std::min(chi, 4);
  int seen = 0;
  for (int k = 0; k < n; ++k) {
    const int r = mod16((int)((k * S) % 16));
    if (seen & (1 << r)) return false;
    seen |= 1 << r;
  }
  return true;
}
// smallest stride >= base (same parity) that is good for a bond of extent chi
long long pad_to_good(long long base, int chi) {
  for (int p = 0; p < 16; p += 2)
    if (good_stride(base + p, chi)) return base + p;
  return base;
}

struct PassChoice {
  int chunk = 1;
  int split = 0;      // fast passes: bonds below `split` stay inside the contiguous row
  int pad_chunk = 0;  // extra doubles between consecutive chunk slices (fast) -- tried by full evaluation
  int pad_plane = 0;
  bool wl = false;    // warp-local operation lists (every extent of the vertex <= 8)
  bool half_red = false;  // barrier plans with one 8 x 8 tile per close: half-size cross-warp scratch (one more barrier per close)
};

// Builds the plan of one pass.  fast (passes 1 and 3): the block holds every index of (site, G1) and `chunk` values of G2's
// flat index; slow (pass 2): `chunk` values of the flat (site, G1) index and every index of G2.  Rows (contiguous in HBM
// and in shared memory) are placed in shared memory with padded strides so that no bond stride is a multiple of 8 doubles.
bool plan_pass(const Signature& sg, int h, int which, bool cplx, int ks_inst, const PassChoice& ch, bool with_tables,
               PassPlan& out) {
  const bool aligned = cplx && ks_inst >= 8;
  const int z = sg.z, d = sg.d;
  long long L = d, XR = 1;
  for (int k = 0; k < h; ++k) L *= sg.chi[k];
  for (int k = h; k < z; ++k) XR *= sg.chi[k];
  const long long n = L * XR;
  const bool fast = which != 1;
  const int chunk = ch.chunk;
  BlkPass& P = out.desc;
  memset(&P, 0, sizeof(P));
  P.gplane = n;
  struct Axis {
    int n;
    long long s;
    int mode;  // local mode index or -1
  };
  std::vector<Axis> axes;
  std::vector<int> mode_slot;
  long long top;  // doubles per plane
  if (fast) {
    if (XR % chunk) return false;
    const int j = ch.split;
    long long rowlen = d;
    axes.push_back({d, 1, -1});
    long long S = d;
    for (int k = 0; k < j; ++k) {
      axes.push_back({sg.chi[k], S, k});
      S *= sg.chi[k];
      rowlen = S;
    }
    P.nlev = 0;
    long long nrows = 1;
    for (int k = j; k < h; ++k) {
      S = pad_to_good(S, sg.chi[k]);
      axes.push_back({sg.chi[k], S, k});
      if (P.nlev >= 3) return false;
      P.lev_n[P.nlev] = sg.chi[k];
      P.lev_s[P.nlev] = (int)S;
      P.nlev++;
      nrows *= sg.chi[k];
      S *= sg.chi[k];
    }
    S += ch.pad_chunk;
    axes.push_back({chunk, S, -1});
    P.lev_n[P.nlev] = chunk;
    P.lev_s[P.nlev] = (int)S;
    P.nlev++;
    nrows *= chunk;
    top = S * chunk;
    P.nblk = (int)(XR / chunk);
    P.nrows = (int)nrows;
    P.rowlen = (int)rowlen;
    P.grow = rowlen;
    P.gblk = L * chunk;
    // no padding anywhere: the whole block is one contiguous run
    bool dense = true;
    {
      long long expect = rowlen;
      for (int l = 0; l < P.nlev; ++l) {
        dense = dense && P.lev_s[l] == expect;
        expect *= P.lev_n[l];
      }
    }
    if (dense) {
      P.nrows = 1;
      P.rowlen = (int)(L * chunk);
      P.grow = 0;
      P.nlev = 0;
    }
    for (int k = 0; k < h; ++k) mode_slot.push_back(k);
  } else {
    if (L % chunk) return false;
    axes.push_back({chunk, 1, -1});
    long long S = chunk;
    P.nlev = 0;
    for (int k = h; k < z; ++k) {
      S = pad_to_good(S, sg.chi[k]);
      axes.push_back({sg.chi[k], S, k - h});
      if (P.nlev >= 4) return false;
      P.lev_n[P.nlev] = sg.chi[k];
      P.lev_s[P.nlev] = (int)S;
      P.nlev++;
      S *= sg.chi[k];
    }
    top = S;
    P.nblk = (int)(L / chunk);
    P.nrows = (int)XR;
    P.rowlen = chunk;
    P.grow = L;
    P.gblk = chunk;
    for (int k = h; k < z; ++k) mode_slot.push_back(k);
  }
  P.PL = (int)top + ch.pad_plane;
  P.bufsz = (cplx ? 2 : 1) * P.PL;
  if ((long long)P.bufsz > 60000 || top > 30000) return false;  // 16-bit fibre tables
  bool even = P.rowlen % 2 == 0 && P.grow % 2 == 0 && P.gblk % 2 == 0 && n % 2 == 0 && P.PL % 2 == 0;
  for (int l = 0; l < P.nlev; ++l) even = even && P.lev_s[l] % 2 == 0;
  // 1: bulk copies (rows of at least 512 bytes: one instruction of the copy engine per row); 2: 16-byte cp.async of all
  // threads (short rows: a warp issues 32 bulk copies one after the other, 32 cp.async at once); 0: 8-byte loads
  P.bulk = even ? (P.rowlen * 8 >= 512 ? 1 : 2) : 0;
  P.load_p = which != 0;
  const int nm = (int)mode_slot.size();
  P.nmodes = nm;
  // operation list.  In-place products need one row tile per fibre (every extent of the vertex <= 8): then a tile's fibres
  // are read and written by the same warp also when the tiles of an operation are dealt out across the warps, so the
  // barrier plans use the in-place operation lists too (one buffer less: larger blocks, fewer fixed costs per flop)
  static const bool no_inplace = getenv("ITN_BLOCK_NO_INPLACE") != nullptr;
  const bool inplace = ch.wl || (!no_inplace && (cplx ? ks_inst <= 4 : ks_inst <= 2));
  if (inplace) {
    EmitterWL em;
    em.P = &P;
    em.busy.assign(8, 0);
    em.busy[0] = 1;
    em.maxbuf = 1;
    if (which == 0) {
      for (int k = 0; k < nm; ++k) em.put(OP_MP, 0, 0, k);  // P = X x_{G1} M in place
      em.put(OP_STORE, 0, 0, 0);
    } else {
      em.busy[1] = 1;
      em.maxbuf = 2;
      em.solve(1, 0, nm, true);  // the partially absorbed tensor is dead after its closes
      if (which == 1) {
        for (int k = 0; k < nm; ++k) em.put(OP_MP, 0, 0, k);  // S = X x_{G2} M in place: every close has read X
        em.put(OP_STORE, 0, 0, 0);
      }
    }
    if (!em.ok) return false;
    P.nbuf = em.maxbuf;
  } else {
    Emitter em;
    em.P = &P;
    em.busy.assign(8, 0);
    em.busy[0] = 1;
    if (which == 0) {
      // P = X x_{G1} M: ping-pong between buffers 1 and 0 (X is dead after its first product)
      int cur = 0;
      for (int k = 0; k < nm; ++k) {
        const int dst = cur == 0 ? 1 : 0;
        em.put(OP_MP, cur, dst, k);
        cur = dst;
      }
      em.put(OP_STORE, cur, 0, 0);
      em.maxbuf = 2;
    } else {
      em.busy[1] = 1;
      em.maxbuf = 2;
      em.solve(1, 0, nm);
      if (which == 1) {
        // S = X x_{G2} M; the buffer of P is free now, X (buffer 0) is only read
        em.busy[1] = 0;
        int cur = 0;
        for (int k = 0; k < nm; ++k) {
          const int dst = em.alloc(cur);
          em.put(OP_MP, cur, dst, k);
          if (cur != 0) em.busy[cur] = 0;
          cur = dst;
        }
        em.put(OP_STORE, cur, 0, 0);
      }
    }
    if (!em.ok) return false;
    P.nbuf = em.maxbuf;
  }
  P.wl = ch.wl ? 1 : 0;
  P.nclose = 0;
  for (int o = 0; o < P.nops; ++o) P.nclose += P.ops[o].type == OP_CLOSE;
  if (ch.wl && P.nclose > kMaxClose) return false;
  {
    const long long scratch = ch.wl ? (long long)P.nclose * kNW * (cplx ? 2 : 1) * 64 : 0;
    P.buf_total = (int)std::max<long long>((long long)P.nbuf * P.bufsz, scratch);
    P.red_len = ch.wl ? 0 : (ch.half_red ? kRedDoubles / 2 : kRedDoubles);
  }
  // warp-local: the batch axes (no mode of the pass) are dealt out to the warps in contiguous, balanced runs
  std::vector<int> owner;  // batch value -> warp
  long long nbatch = 1;
  if (ch.wl) {
    for (const Axis& a : axes)
      if (a.mode < 0) nbatch *= a.n;
    owner.resize(nbatch);
    for (int w = 0; w < kNW; ++w)
      for (long long v = w * nbatch / kNW; v < (w + 1) * nbatch / kNW; ++v) owner[v] = w;
  }
  double eff_sum = 0.0;
  // fibre tables
  out.table.clear();
  out.excess = 0;
  size_t tab_len = 0;
  long long nreal = 1;
  for (const Axis& a : axes) nreal *= a.n;
  for (int k = 0; k < nm; ++k) {
    ModeGeom m;
    m.slot = mode_slot[k];
    m.chi = 0;
    for (const Axis& a : axes)
      if (a.mode == k) {
        m.chi = a.n;
        m.S = (int)a.s;
      }
    P.modes[k].chi = m.chi;
    P.modes[k].S = m.S;
    P.modes[k].slot = m.slot;
    {
      // exact: no padding fibres, and the (stacked) reduction length fills every k4 step of the kernel instance, so the
      // inner loops of the kernel have compile-time trip counts and no predicates
      const int K2 = (cplx && !aligned) ? 2 * m.chi : m.chi;
      const int KHs = aligned ? ks_inst / 2 : ks_inst;
      P.modes[k].exact = ((nreal / m.chi) % 8 == 0 && K2 == 4 * KHs) ? 1 : 0;
    }
    const long long nf = nreal / m.chi;
    if (ch.wl) {
      // fibres of the largest per-warp share -> tiles per warp; the table interleaves the warps' tiles (tile ft: warp ft % 8)
      long long maxb = 0;
      for (int w = 0; w < kNW; ++w) maxb = std::max(maxb, (w + 1) * nbatch / kNW - w * nbatch / kNW);
      const long long per_slice = nf / nbatch;
      const int ntw_est = (int)((maxb * per_slice + 7) / 8);
      eff_sum += (double)nf / (64.0 * ntw_est);
      bool uniform = true;
      for (int w = 0; w < kNW; ++w) uniform = uniform && ((w + 1) * nbatch / kNW - w * nbatch / kNW) == maxb;
      P.modes[k].exact = (P.modes[k].exact && uniform && (maxb * per_slice) % 8 == 0) ? 1 : 0;
      if (with_tables) {
        std::vector<const Axis*> others;
        for (const Axis& a : axes)
          if (a.mode != k) others.push_back(&a);
        std::vector<std::vector<int>> wb(kNW);
        for (long long f = 0; f < nf; ++f) {
          long long q = f, off = 0, bv = 0, bmul = 1;
          for (const Axis* a : others) {
            const long long dg = q % a->n;
            off += dg * a->s;
            q /= a->n;
            if (a->mode < 0) {
              bv += dg * bmul;
              bmul *= a->n;
            }
          }
          wb[owner[bv]].push_back((int)off);
        }
        std::vector<std::vector<unsigned short>> wt(kNW);
        size_t ntw = 0;
        for (int w = 0; w < kNW; ++w) {
          if (wb[w].empty()) continue;
          ModeGeom mw = m;
          mw.bases = wb[w];
          out.excess += plan_fibres(mw, cplx, aligned, P.PL, wt[w]);
          ntw = std::max(ntw, wt[w].size() / 8);
        }
        P.modes[k].tab = (int)out.table.size();
        P.modes[k].ntile = (int)(ntw * kNW);
        for (size_t j = 0; j < ntw; ++j)
          for (int w = 0; w < kNW; ++w)
            for (int i = 0; i < 8; ++i)
              out.table.push_back(j * 8 + i < wt[w].size() ? wt[w][j * 8 + i] : (unsigned short)kNoFibre);
        tab_len = out.table.size();
      } else {
        tab_len += (size_t)ntw_est * kNW * 8;
      }
    } else if (with_tables) {
      // every combination of the other axes, first axis fastest
      std::vector<const Axis*> others;
      for (const Axis& a : axes)
        if (a.mode != k) others.push_back(&a);
      m.bases.resize(nf);
      for (long long f = 0; f < nf; ++f) {
        long long q = f, off = 0;
        for (const Axis* a : others) {
          off += (q % a->n) * a->s;
          q /= a->n;
        }
        m.bases[f] = (int)off;
      }
      std::vector<unsigned short> t;
      out.excess += plan_fibres(m, cplx, aligned, P.PL, t);
      P.modes[k].tab = (int)out.table.size();
      P.modes[k].ntile = (int)t.size() / 8;
      out.table.insert(out.table.end(), t.begin(), t.end());
      tab_len = out.table.size();
    } else {
      tab_len += (size_t)((nreal / m.chi + 15) & ~7ll);
    }
  }
  out.eff = ch.wl ? eff_sum / std::max(nm, 1) : 1.0;
  // shared-memory position of every row (the kernel's stores read it instead of dividing)
  P.rowtab = (int)out.table.size();
  if (with_tables) {
    for (int row = 0; row < P.nrows; ++row) {
      int off = 0, q = row;
      for (int l = 0; l < P.nlev; ++l) {
        off += (q % P.lev_n[l]) * P.lev_s[l];
        q /= P.lev_n[l];
      }
      out.table.push_back((unsigned short)off);
    }
    while (out.table.size() % 8) out.table.push_back(0);
    tab_len = out.table.size();
  } else {
    tab_len += (size_t)((P.nrows + 15) & ~7);
  }
  P.tab_len = (int)out.table.size();
  // messages in shared memory (the fragments of every mode product are re-read from them at the start of the operation)
  // unless that costs the second CTA per SM
  {
    int len = 0;
    // fragment-ordered copies, sized by the kernel instance of the vertex (its largest extent, `ks_inst`)
    const int KHs = (cplx && ks_inst >= 8) ? ks_inst / 2 : ks_inst;
    for (int k = 0; k < nm; ++k) {
      P.msg_off[k] = len;
      len += ((P.modes[k].chi + 7) / 8) * KHs * 64;
    }
    const size_t without = pass_smem(P, tab_len);
    const size_t with = without + (size_t)len * sizeof(double);
    const bool fits = with <= kSmemTwoCtas || (without > kSmemTwoCtas && with <= kSmemOneCta);
    P.msg_smem = fits ? 1 : 0;
    P.msg_len = fits ? len : 0;
  }
  out.smem = pass_smem(P, tab_len);
  out.chunk = ch.chunk;
  out.split = ch.split;
  out.pad_chunk = ch.pad_chunk;
  out.pad_plane = ch.pad_plane;
  return true;
}

// Chooses the block of a pass: the largest chunk that leaves two CTAs per SM (one if it must), shrunk while the launch
// would not cover the SMs; then the paddings with the fewest excess wavefronts.
bool choose_pass(const Signature& sg, int h, int which, bool cplx, int ks_inst, size_t nverts, PassPlan& best, bool wl) {
  const bool aligned = cplx && ks_inst >= 8;
  kSmemTwoCtas = target_ctas(ks_inst) == 3 ? kSmemThreeCtasMax : kSmemTwoCtasMax;
  const int z = sg.z, d = sg.d;
  long long L = d, XR = 1;
  for (int k = 0; k < h; ++k) L *= sg.chi[k];
  for (int k = h; k < z; ++k) XR *= sg.chi[k];
  const bool fast = which != 1;
  const long long range = fast ? XR : L;
  PassChoice ch;
  ch.wl = wl;
  // blocks sized for three CTAs per SM always take the half-size cross-warp scratch
  ch.half_red = !wl && target_ctas(ks_inst) == 3 && (cplx ? ks_inst <= 4 : ks_inst <= 2);
  if (fast) {
    // bonds whose natural stride is fine stay inside the contiguous row
    long long S = d;
    ch.split = h;
    for (int k = 0; k < h; ++k) {
      if (k >= 1 && !good_stride(S, sg.chi[k])) {
        ch.split = k;
        break;
      }
      S *= sg.chi[k];
    }
  }
  int best_chunk = -1;
  bool best_two = false;
  double best_score = -1;
  for (long long c = 1; c <= range; ++c) {
    if (range % c) continue;
    if (!fast && (c % 2) && (L % 2 == 0) && c != range) continue;  // 16-byte rows
    PassPlan pp;
    PassChoice t = ch;
    t.chunk = (int)c;
    t.pad_chunk = fast ? 6 : 0;  // room for the paddings tried below
    t.pad_plane = 6;
    if (!plan_pass(sg, h, which, cplx, ks_inst, t, false, pp)) continue;
    if (pp.smem > kSmemOneCta) continue;
    if (pp.smem > kSmemTwoCtas && pp.smem < kSmemTwoCtas + kSmemTwoCtas / 8) {
      // the estimate (room for every padding, table lengths rounded up) is within reach of two CTAs per SM: plan exactly
      PassPlan ex;
      PassChoice t0 = ch;
      t0.chunk = (int)c;
      const bool okx = plan_pass(sg, h, which, cplx, ks_inst, t0, true, ex);
      if (block_debug()) fprintf(stderr, "      exact plan of chunk %lld: ok %d smem %zu msgs %d\n", c, (int)okx, ex.smem, ex.desc.msg_smem);
      if (okx && ex.smem <= kSmemTwoCtas) pp = std::move(ex);
    }
    const long long nb = (fast ? L : XR) * c;
    const long long ctas = (long long)pp.desc.nblk * (long long)nverts;
    const bool two = pp.smem <= kSmemTwoCtas;
    double score = std::min<double>((double)nb, 4096.0) * pp.eff;
    if (!two) score *= 0.2;  // one CTA of 8 warps per SM hides neither the staging nor the DMMA latency
    if (!fast && c * 8 < 64) score *= 0.5 + c / 16.0;  // short HBM rows waste sectors and bulk-copy issue slots
    if (ctas < 296) score *= ((double)ctas / 296.0) * 0.9 + 0.1;
    if (block_debug())
      fprintf(stderr, "      pass %d candidate chunk %lld: smem %zu (%s, msgs %s) score %.0f\n", which + 1, c, pp.smem,
              two ? "2 CTAs/SM" : "1 CTA/SM", pp.desc.msg_smem ? "staged" : "global", score);
    if (score > best_score) {
      best_score = score;
      best_chunk = (int)c;
      best_two = two;
    }
  }
  if (best_chunk < 0) return false;
  ch.chunk = best_chunk;
  if (wl && !(getenv("ITN_BLOCK_WL") && atoi(getenv("ITN_BLOCK_WL")) == 2)) {
    // a warp-local plan is only taken when every mode stays exact (geometry_for): known before any table is built,
    // so the padding search below (a full table per candidate and warp) is skipped for the signatures that fail
    PassPlan pp;
    if (!plan_pass(sg, h, which, cplx, ks_inst, ch, false, pp)) return false;
    for (int k = 0; k < pp.desc.nmodes; ++k)
      if (!pp.desc.modes[k].exact) return false;
  }
  const bool want_two = best_two;
  bool mixed_steps = false;  // packed complex stacking with a k step that straddles the planes
  if (cplx && !aligned)
    for (int k = fast ? 0 : h; k < (fast ? h : z); ++k) mixed_steps = mixed_steps || (sg.chi[k] % 4) != 0;
  bool have = false;
  double best_ex = 1e300;
  for (int pc : {0, 2, 4, 6}) {
    if (!fast && pc) break;
    for (int pl : {0, 2, 4, 6, 8, 10, 12, 14}) {
      if (!mixed_steps && pl) break;
      PassPlan pp;
      PassChoice t = ch;
      t.pad_chunk = pc;
      t.pad_plane = pl;
      if (!plan_pass(sg, h, which, cplx, ks_inst, t, true, pp)) continue;
      if (pp.smem > (want_two ? kSmemTwoCtas : kSmemOneCta)) continue;
      const double ex = (double)pp.excess + 1e-3 * (pc + pl);
      if (ex < best_ex) {
        best_ex = ex;
        best = std::move(pp);
        have = true;
      }
      if (have && best.excess == 0) break;
    }
    if (have && best.excess == 0) break;
  }
  // a barrier plan of a small-extent vertex that misses three CTAs per SM by less than half the cross-warp scratch
  // takes the half-size scratch (z = 6, chi = 6 pass 2: 77.4 -> 73.3 KB; 24 instead of 16 resident warps)
  static const bool no_half = getenv("ITN_BLOCK_NO_HALF_RED") != nullptr;
  const bool small = cplx ? ks_inst <= 4 : ks_inst <= 2;
  if (have && !wl && small && !no_half && best.smem > kSmemThreeCtasMax &&
      best.smem <= kSmemThreeCtasMax + (kRedDoubles / 2) * sizeof(double)) {
    PassChoice t = ch;
    t.chunk = best.chunk;
    t.split = best.split;
    t.pad_chunk = best.pad_chunk;
    t.pad_plane = best.pad_plane;
    t.half_red = true;
    PassPlan hp;
    if (plan_pass(sg, h, which, cplx, ks_inst, t, true, hp) && hp.smem <= kSmemThreeCtasMax && hp.excess <= best.excess)
      best = std::move(hp);
  }
  return have;
}

void kernel_instance(bool cplx, int chimax, int& KS, int& MT) {
  if (cplx) {
    KS = chimax <= 2 ? 1 : chimax <= 4 ? 2 : chimax <= 6 ? 3 : chimax <= 8 ? 4 : chimax <= 16 ? 8 : 16;
    MT = chimax <= 8 ? 1 : chimax <= 16 ? 2 : 4;
  } else {
    KS = chimax <= 4 ? 1 : chimax <= 8 ? 2 : chimax <= 16 ? 4 : 8;
    MT = chimax <= 8 ? 1 : chimax <= 16 ? 2 : 4;
  }
}

// Geometry of a signature: independent of the network, shared by every network of the process.
struct Geometry {
  Signature sig;
  int h = 0;
  int KS = 0, MT = 0;  // kernel instance; KS = 0: the signature has no plan (falls back to the shape-generic kernels)
  PassPlan pass[3];
  long long nelem = 0;
};

struct GeomKey {
  Signature sig;
  bool cplx;
  int fill;  // 0: few vertices (the block size is shrunk so that the launch still covers the SMs), 1: plenty
  bool operator<(const GeomKey& o) const { return std::tie(sig, cplx, fill) < std::tie(o.sig, o.cplx, o.fill); }
};

const Geometry* geometry_for(const Signature& s, bool cplx, size_t nverts) {
  static std::map<GeomKey, std::unique_ptr<Geometry>> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  // the vertex count matters only while the launch does not fill the machine
  const int fill = nverts >= 2048 ? 1 : 0;
  GeomKey key{s, cplx, fill ? 1 << 30 : (int)nverts};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second.get();
  std::unique_ptr<Geometry> g(new Geometry());
  g->sig = s;
  g->h = (s.z + 1) / 2;
  int chimax = 0;
  g->nelem = s.d;
  for (int k = 0; k < s.z; ++k) {
    chimax = std::max(chimax, s.chi[k]);
    g->nelem *= s.chi[k];
  }
  kernel_instance(cplx, chimax, g->KS, g->MT);
  // warp-local passes (k_block<..., WL = true>): every extent <= 8 and a block whose batch slices fill the warps' tiles
  static const bool wl_off = getenv("ITN_BLOCK_WL") && atoi(getenv("ITN_BLOCK_WL")) == 0;
  const bool wl_ok = !wl_off && chimax <= 8;
  bool ok = true;
  // (measured, 8^3 chi = 6: warp-local plans whose per-warp tables need padding tiles or lose the conflict-free fibre
  // order -- a slice holds one parity of the site index -- execute 20 % more instructions and 5x the bank conflicts,
  // which costs more than the barriers they save; chi = 8 and 4 plan exactly and gain 5 %)
  for (int w = 0; w < 3 && ok; ++w) {
    ok = choose_pass(s, g->h, w, cplx, g->KS, nverts, g->pass[w], false);
    if (ok && wl_ok) {
      PassPlan wp;
      bool exact = true;
      if (choose_pass(s, g->h, w, cplx, g->KS, nverts, wp, true)) {
        for (int k = 0; k < wp.desc.nmodes; ++k) exact = exact && wp.desc.modes[k].exact;
        const bool force = getenv("ITN_BLOCK_WL") && atoi(getenv("ITN_BLOCK_WL")) == 2;  // experiments: always
        if (force || (exact && wp.excess <= g->pass[w].excess + (int)(wp.table.size() / 80))) g->pass[w] = std::move(wp);
      }
    }
  }
  if (!ok) g->KS = 0;
  if (ok && block_debug()) {
    fprintf(stderr, "[itn block] d=%d z=%d chi=", s.d, s.z);
    for (int k = 0; k < s.z; ++k) fprintf(stderr, "%d%s", s.chi[k], k + 1 < s.z ? "," : "");
    fprintf(stderr, " h=%d KS=%d MT=%d verts=%zu\n", g->h, g->KS, g->MT, nverts);
    for (int w = 0; w < 3; ++w) {
      const BlkPass& P = g->pass[w].desc;
      fprintf(stderr, "   pass %d: nblk=%d rows=%d x %d levels=", w + 1, P.nblk, P.nrows, P.rowlen);
      for (int l = 0; l < P.nlev; ++l) fprintf(stderr, "%dx%d ", P.lev_n[l], P.lev_s[l]);
      fprintf(stderr, "strides=");
      for (int k = 0; k < P.nmodes; ++k) fprintf(stderr, "%d ", P.modes[k].S);
      fprintf(stderr, "PL=%d nbuf=%d bulk=%d ops=%d smem=%zu excess wavefronts=%d of %zu quads wl=%d eff=%.2f exact=", P.PL, P.nbuf,
              P.bulk, P.nops, g->pass[w].smem, g->pass[w].excess, g->pass[w].table.size() / 4, P.wl, g->pass[w].eff);
      for (int k = 0; k < P.nmodes; ++k) fprintf(stderr, "%d", P.modes[k].exact);
      fprintf(stderr, "\n");
    }
  }
  const Geometry* r = g.get();
  cache[key] = std::move(g);
  return r;
}

struct Bucket {
  const Geometry* geo = nullptr;
  std::vector<int> verts;
  // device copies of the geometry (per network: they live on its device)
  unsigned short* d_tab[3] = {nullptr, nullptr, nullptr};
  // per update() call
  DevBuf* scratch = nullptr;  // P and S of every vertex
  DevBuf* part = nullptr;
  DevBuf* dverts = nullptr;
  DevBuf* dred = nullptr;
  size_t nred = 0;
};

struct BlockCache {
  std::map<const Geometry*, Bucket> buckets;
  std::vector<Bucket*> active;
  std::vector<int> vbucket;
  // the buckets of a sweep are independent (they read the pre-sweep messages and write their own staged outputs): all
  // but the largest run on side streams, forked from and joined to the context stream with events, so that the few CTAs
  // of the corner / rim / degree-2 buckets fill the SMs the main bucket leaves idle instead of running one after another
  std::vector<cudaStream_t> side;
  cudaEvent_t fork = nullptr;
  std::vector<cudaEvent_t> join;
  size_t pending_joins = 0;
};

bool signature_of(const itn_net* net, int v, Signature& s) {
  const int z = (int)net->inc[v].size();
  if (z < 2 || z > 8 || !net->T[v].p) return false;
  s.d = net->sdim[v];
  s.z = z;
  s.chi.fill(0);
  for (int k = 0; k < z; ++k) {
    s.chi[k] = net->edim[net->inc[v][k]];
    if (s.chi[k] > 32 || s.chi[k] < 1) return false;
  }
  return s.d >= 1 && s.d <= 8;
}

template <bool C, int KS, int MT, int NB, bool WL = false>
void launch_inst(itn_ctx* ctx, cudaStream_t st, unsigned grid, size_t smem, const BlkPass* dp, const BlkVertex* dv,
                 const unsigned short* dt) {
  CUDA_CHECK(cudaFuncSetAttribute(k_block<C, KS, MT, NB, WL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_CHECK(cudaFuncSetAttribute(k_block<C, KS, MT, NB, WL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  k_block<C, KS, MT, NB, WL><<<grid, kBT, smem, st>>>(*dp, dv, dt);  // the descriptor travels as a kernel parameter
  ITN_LAUNCH_CHECK(ctx);
}

// NB = resident CTAs per SM the instance is compiled for: the small instances exist for 2 (all the registers they want)
// and for 3 (at most 80 registers); a pass whose block leaves room for three CTAs in shared memory takes the latter
void launch_pass(itn_ctx* ctx, cudaStream_t st, bool cplx, int KS, unsigned grid, size_t smem, const BlkPass* dp,
                 const BlkVertex* dv, const unsigned short* dt) {
  const bool three = KS <= 4 && smem <= kSmemThreeCtasMax;
#define ITN_BLK(Cx, K, M)                                                                              \
  if (dp->wl)                                                                                          \
    return three ? launch_inst<Cx, K, M, 3, true>(ctx, st, grid, smem, dp, dv, dt)                      \
                 : launch_inst<Cx, K, M, 2, true>(ctx, st, grid, smem, dp, dv, dt);                     \
  return three ? launch_inst<Cx, K, M, 3>(ctx, st, grid, smem, dp, dv, dt)                              \
               : launch_inst<Cx, K, M, 2>(ctx, st, grid, smem, dp, dv, dt)
  if (cplx) {
    switch (KS) {
      case 1: ITN_BLK(true, 1, 1);
      case 2: ITN_BLK(true, 2, 1);
      case 3: ITN_BLK(true, 3, 1);
      case 4: ITN_BLK(true, 4, 1);
      case 8: return launch_inst<true, 8, 2, 2>(ctx, st, grid, smem, dp, dv, dt);
      default: return launch_inst<true, 16, 4, 2>(ctx, st, grid, smem, dp, dv, dt);
    }
  } else {
    switch (KS) {
      case 1: ITN_BLK(false, 1, 1);
      case 2: ITN_BLK(false, 2, 1);
      case 4:
        return three ? launch_inst<false, 4, 2, 3>(ctx, st, grid, smem, dp, dv, dt)
                     : launch_inst<false, 4, 2, 2>(ctx, st, grid, smem, dp, dv, dt);
      default: return launch_inst<false, 8, 4, 2>(ctx, st, grid, smem, dp, dv, dt);
    }
  }
#undef ITN_BLK
}

void end_call(Bucket& b) {
  delete b.scratch;
  delete b.part;
  delete b.dverts;
  delete b.dred;
  b.scratch = b.part = b.dverts = b.dred = nullptr;
}

}  // namespace

void itn_block_release(itn_net* net) {
  if (!net->block) return;
  BlockCache* bc = (BlockCache*)net->block;
  for (auto& kv : bc->buckets) {
    Bucket& b = kv.second;
    end_call(b);
    for (auto& t : b.d_tab) itn_dev_free(net->ctx, t);
  }
  for (cudaStream_t st : bc->side) cudaStreamDestroy(st);
  for (cudaEvent_t ev : bc->join) cudaEventDestroy(ev);
  if (bc->fork) cudaEventDestroy(bc->fork);
  delete bc;
  net->block = nullptr;
}

// Plans a synchronous sweep on the block path: handled[i] = 2 for the message jobs it takes (vertices all of whose
// outgoing messages are listed exactly once, none of them taken by the tile path, and whose signature has a plan).
int itn_block_bp_plan(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv, std::vector<char>& handled) {
  if (net->ctx->path_mode != 0 || net->has_bra()) return 0;
  static const bool off = getenv("ITN_NO_BLOCK") != nullptr;
  if (off) return 0;
  BlockCache* bc = (BlockCache*)net->block;
  if (!bc) net->block = bc = new BlockCache();
  if (handled.size() != dids.size()) handled.assign(dids.size(), 0);
  std::vector<int> cnt(net->nv, 0);
  {
    std::vector<char> listed(net->M.size(), 0);
    for (size_t i = 0; i < dids.size(); ++i) {
      cnt[srcv[i]] += (listed[dids[i]] || handled[i]) ? 100 : 1;
      listed[dids[i]] = 1;
    }
  }
  for (Bucket* b : bc->active) b->verts.clear();
  bc->active.clear();
  bc->vbucket.assign(net->nv, -1);
  // count the vertices of every signature first: the geometry depends on how many blocks the launch will have
  std::map<Signature, std::vector<int>> by_sig;
  for (int v = 0; v < net->nv; ++v) {
    if (cnt[v] != (int)net->inc[v].size()) continue;
    Signature s;
    if (signature_of(net, v, s)) by_sig[s].push_back(v);
  }
  int n = 0;
  for (auto& kv : by_sig) {
    const Geometry* geo = geometry_for(kv.first, net->cplx, kv.second.size());
    if (geo->KS == 0) continue;
    Bucket& b = bc->buckets[geo];
    b.geo = geo;
    b.verts = kv.second;
    for (int v : b.verts) bc->vbucket[v] = (int)bc->active.size();
    bc->active.push_back(&b);
  }
  for (size_t i = 0; i < dids.size(); ++i)
    if (!handled[i] && bc->vbucket[srcv[i]] >= 0) {
      handled[i] = 2;
      ++n;
    }
  return n;
}

// Prepares the planned sweep for one itn_bp_update call: device tables, scratch tensors, per-vertex descriptors.  The
// un-normalised new message of job i (handled[i] == 2) will be written to staged[i] by every itn_block_bp_run.
void itn_block_bp_begin(itn_net* net, const std::vector<int>& dids, const std::vector<int>& srcv,
                        const std::vector<char>& handled, double* const* staged) {
  BlockCache* bc = (BlockCache*)net->block;
  if (!bc || bc->active.empty()) return;
  itn_ctx* ctx = net->ctx;
  const int PLN = net->planes();
  std::vector<std::array<double*, 8>> st(net->nv);
  for (auto& a : st) a.fill(nullptr);
  for (size_t i = 0; i < dids.size(); ++i)
    if (handled[i] == 2) st[srcv[i]][net->slot(srcv[i], dids[i] / 2)] = staged[i];
  for (Bucket* bp : bc->active) {
    Bucket& b = *bp;
    const Geometry& G = *b.geo;
    const size_t nv = b.verts.size();
    const int z = G.sig.z;
    end_call(b);
    if (!b.d_tab[0]) {
      for (int w = 0; w < 3; ++w) {
        b.d_tab[w] = (unsigned short*)itn_dev_alloc(ctx, std::max<size_t>(G.pass[w].table.size(), 8) * sizeof(unsigned short));
        CUDA_CHECK(cudaMemcpyAsync(b.d_tab[w], G.pass[w].table.data(), G.pass[w].table.size() * sizeof(unsigned short),
                                   cudaMemcpyHostToDevice, ctx->stream));
      }
    }
    size_t part_v = 0;  // doubles of partial results per vertex
    std::vector<size_t> part_off(z);
    for (int k = 0; k < z; ++k) {
      part_off[k] = part_v;
      const int nblk = k < G.h ? G.pass[2].desc.nblk : G.pass[1].desc.nblk;
      part_v += (size_t)nblk * PLN * G.sig.chi[k] * G.sig.chi[k];
    }
    b.scratch = new DevBuf(ctx, nv * 2 * (size_t)G.nelem * PLN * sizeof(double));
    b.part = new DevBuf(ctx, nv * part_v * sizeof(double));
    std::vector<BlkVertex> hv(3 * nv);
    std::vector<BlkRedJob> hr;
    for (size_t i = 0; i < nv; ++i) {
      const int v = b.verts[i];
      double* Pbuf = b.scratch->as<double>() + (2 * i) * (size_t)G.nelem * PLN;
      double* Sbuf = Pbuf + (size_t)G.nelem * PLN;
      for (int w = 0; w < 3; ++w) {
        BlkVertex& V = hv[w * nv + i];
        memset(&V, 0, sizeof(V));
        V.X = net->T[v].p;
        V.Pin = w == 1 ? Pbuf : (w == 2 ? Sbuf : nullptr);
        V.Wout = w == 0 ? Pbuf : (w == 1 ? Sbuf : nullptr);
        for (int k = 0; k < z; ++k) {
          const DevTensor& m = net->M[net->msg_into(v, net->inc[v][k])];
          ITN_REQUIRE(m.p != nullptr, ITN_EINVAL, "an incoming message is not set");
          V.msg[k] = m.p;
          V.part[k] = b.part->as<double>() + i * part_v + part_off[k];
        }
      }
      for (int k = 0; k < z; ++k) {
        ITN_REQUIRE(st[v][k] != nullptr, ITN_EINVAL, "block sweep: an outgoing message has no staging buffer");
        const int nblk = k < G.h ? G.pass[2].desc.nblk : G.pass[1].desc.nblk;
        hr.push_back({b.part->as<double>() + i * part_v + part_off[k], st[v][k], nblk, PLN * G.sig.chi[k] * G.sig.chi[k]});
      }
    }
    b.dverts = new DevBuf(ctx, hv.size() * sizeof(BlkVertex));
    b.dred = new DevBuf(ctx, hr.size() * sizeof(BlkRedJob));
    b.nred = hr.size();
    itn_upload(ctx, hv, *b.dverts);
    itn_upload(ctx, hr, *b.dred);
  }
}

// all_side: the caller is about to enqueue other work of the same sweep on the context stream (the tile path): every
// bucket goes to a side stream.  The context stream picks the results up in itn_block_bp_join.
void itn_block_bp_run(itn_net* net, bool all_side) {
  BlockCache* bc = (BlockCache*)net->block;
  if (!bc) return;
  itn_ctx* ctx = net->ctx;
  static const bool serial = getenv("ITN_BLOCK_SERIAL") != nullptr;  // experiments: every bucket on the context stream
  // largest bucket (by tensor elements) on the context stream, the others on side streams
  std::vector<Bucket*> order(bc->active.begin(), bc->active.end());
  std::stable_sort(order.begin(), order.end(), [](const Bucket* a, const Bucket* b) {
    return (double)a->verts.size() * (double)a->geo->nelem > (double)b->verts.size() * (double)b->geo->nelem;
  });
  if (all_side && !serial) order.insert(order.begin(), nullptr);  // nobody on the context stream
  const size_t nside = (serial || order.size() < 2) ? 0 : order.size() - 1;
  bc->pending_joins = 0;
  if (nside) {
    if (!bc->fork) CUDA_CHECK(cudaEventCreateWithFlags(&bc->fork, cudaEventDisableTiming));
    while (bc->side.size() < nside) {
      cudaStream_t st;
      cudaEvent_t ev;
      CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      bc->side.push_back(st);
      bc->join.push_back(ev);
    }
    CUDA_CHECK(cudaEventRecord(bc->fork, ctx->stream));
  }
  for (size_t bi = 0; bi < order.size(); ++bi) {
    if (!order[bi]) continue;
    Bucket& b = *order[bi];
    const Geometry& G = *b.geo;
    const size_t nv = b.verts.size();
    ITN_REQUIRE(b.dverts != nullptr, ITN_EINVAL, "block sweep is not prepared");
    cudaStream_t st = (bi == 0 || nside == 0) ? ctx->stream : bc->side[bi - 1];
    if (st != ctx->stream) CUDA_CHECK(cudaStreamWaitEvent(st, bc->fork, 0));
    for (int w = 0; w < 3; ++w)
      launch_pass(ctx, st, net->cplx, G.KS, (unsigned)(nv * G.pass[w].desc.nblk), G.pass[w].smem, &G.pass[w].desc,
                  b.dverts->as<BlkVertex>() + w * nv, b.d_tab[w]);
    int chimax = 0;
    for (int k = 0; k < G.sig.z; ++k) chimax = std::max(chimax, G.sig.chi[k]);
    const unsigned ry = (unsigned)std::max(1, std::min(16, (net->planes() * chimax * chimax + 127) / 128));
    k_block_reduce<<<dim3((unsigned)b.nred, ry), 128, 0, st>>>(b.dred->as<BlkRedJob>());
    ITN_LAUNCH_CHECK(ctx);
    if (st != ctx->stream) CUDA_CHECK(cudaEventRecord(bc->join[bi - 1], st));
  }
  bc->pending_joins = nside;
}

// the context stream waits for the side streams of the last itn_block_bp_run
void itn_block_bp_join(itn_net* net) {
  BlockCache* bc = (BlockCache*)net->block;
  if (!bc) return;
  for (size_t i = 0; i < bc->pending_joins; ++i) CUDA_CHECK(cudaStreamWaitEvent(net->ctx->stream, bc->join[i], 0));
  bc->pending_joins = 0;
}

void itn_block_bp_end(itn_net* net) {
  BlockCache* bc = (BlockCache*)net->block;
  if (!bc) return;
  for (Bucket* b : bc->active) end_call(*b);
}

// Instrumentation (host only, no device needed): the geometry the block path would use for a vertex signature, as a flat
// int32 stream, so that tests can replay the operation lists and fibre tables on the host (tests/block_emulator.py):
//   KS, MT, h, then per pass: nblk, nrows, rowlen, grow, gblk, nlev, lev_n[4], lev_s[4], PL, bufsz, nbuf, bulk, load_p, smem
//   bytes, excess wavefronts, wl, nmodes, {chi, S, ntile, tab, slot} x nmodes, nops, {type, src, dst, mode} x nops, tab_len, table...
extern "C" int itn_block_plan_export(int dtype, int d, int z, const int32_t* chi, int nverts, int32_t* out, int cap, int32_t* nout) {
  try {
    if (!chi || !nout || z < 1 || z > 8 || (dtype != ITN_F64 && dtype != ITN_C128)) {
      itn_set_error("itn_block_plan_export: bad arguments");
      return ITN_EINVAL;
    }
    Signature s;
    s.d = d;
    s.z = z;
    s.chi.fill(0);
    for (int k = 0; k < z; ++k) s.chi[k] = chi[k];
    std::vector<int32_t> v;
    bool supported = z >= 2 && d >= 1 && d <= 8;
    for (int k = 0; k < z; ++k) supported = supported && chi[k] >= 1 && chi[k] <= 32;
    const Geometry* g = supported ? geometry_for(s, dtype == ITN_C128, (size_t)std::max(nverts, 1)) : nullptr;
    if (!g || g->KS == 0) {
      *nout = 0;
      return ITN_OK;
    }
    v.push_back(g->KS);
    v.push_back(g->MT);
    v.push_back(g->h);
    for (int w = 0; w < 3; ++w) {
      const BlkPass& P = g->pass[w].desc;
      for (long long x : {(long long)P.nblk, (long long)P.nrows, (long long)P.rowlen, P.grow, P.gblk, (long long)P.nlev,
                          (long long)P.lev_n[0], (long long)P.lev_n[1], (long long)P.lev_n[2], (long long)P.lev_n[3],
                          (long long)P.lev_s[0], (long long)P.lev_s[1], (long long)P.lev_s[2], (long long)P.lev_s[3], (long long)P.PL,
                          (long long)P.bufsz, (long long)P.nbuf, (long long)P.bulk, (long long)P.load_p, (long long)g->pass[w].smem,
                          (long long)g->pass[w].excess, (long long)P.wl, (long long)P.nmodes})
        v.push_back((int32_t)x);
      for (int k = 0; k < P.nmodes; ++k)
        for (int x : {P.modes[k].chi, P.modes[k].S, P.modes[k].ntile, P.modes[k].tab, P.modes[k].slot}) v.push_back(x);  // (exact is derived)
      v.push_back(P.nops);
      for (int o = 0; o < P.nops; ++o)
        for (int x : {(int)P.ops[o].type, (int)P.ops[o].src, (int)P.ops[o].dst, (int)P.ops[o].mode}) v.push_back(x);
      v.push_back((int32_t)g->pass[w].table.size());
      for (unsigned short t : g->pass[w].table) v.push_back((int32_t)t);
    }
    *nout = (int32_t)v.size();
    if (out && cap >= (int)v.size()) memcpy(out, v.data(), v.size() * sizeof(int32_t));
    return ITN_OK;
  } catch (const std::exception& e) {
    itn_set_error(e.what());
    return ITN_EINVAL;
  }
}
