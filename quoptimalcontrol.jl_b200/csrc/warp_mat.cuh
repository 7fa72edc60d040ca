// Warp-resident complex-FP64 block matrices on DMMA (mma.sync.m8n8k4.f64, the native FP64 tensor shape
// on sm_100a: every f64 mma shape lowers to DMMA.8x8x4 SASS).
//
// A D x D complex matrix (D = 8*NB) is held by ONE warp in registers in the accumulator layout "CM":
//   lane (g = lane>>2, q = lane&3) holds X[8i+g][8j+2q+e], e = 0,1, for every 8x8 block (i, j).
// The MMA's k index is assigned as k = 2q + h (h = which of the two k-halves), so that
//   * CM(X) IS the left-operand fragment of X          (A[m=g][k=2q+h] = X[g][2q+h])
//   * CM(Y) IS the right-operand fragment of Y^T        (B[k=2q+h][n=g] = Y[g][2q+h] = (Y^T)[2q+h][g])
// i.e. the tensor pipe natively computes the "NT" product  C = A * B^T  (or A * B^H) on row-distributed
// registers and delivers C in the same layout: chains of products need NO fragment conversion.  The only data
// movement left is an explicit transpose, done through a padded per-warp shared-memory tile
// (2 STS.128 + 4 LDS.64 per complex 8x8 block, no selects).  Row stride 10 doubles keeps the four column-wise LDS.64
// conflict-free; the two row-wise STS.128 then cost one extra wavefront per quarter-warp (rows g and g+1 are 20 banks
// apart, so their 16-bank footprints overlap by 4): ~17 M conflict wavefronts per cfg4 launch in ncu, <1 % of the LSU
// traffic.  No stride serves both (stores want stride = 8 mod 16, which makes the loads 4-way conflicting).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace qoc {

constexpr unsigned FULL_MASK = 0xffffffffu;

template <int NB> struct CM { double re[NB][NB][2]; double im[NB][NB][2]; };

struct Lane {
  int lane, g, q;
  __device__ __forceinline__ explicit Lane(int l) : lane(l), g(l >> 2), q(l & 3) {}
};

__device__ __forceinline__ double dneg(double x) {   // sign flip on the integer pipe (keeps the FP64 pipe free)
  return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define QOC_FOR_CM(NB) \
  _Pragma("unroll") for (int i = 0; i < NB; i++) \
  _Pragma("unroll") for (int j = 0; j < NB; j++) \
  _Pragma("unroll") for (int e = 0; e < 2; e++)

// ---- products: C (+)= op(A) * op(B)^T,  op = identity or elementwise conjugate ---------------------------
//   CONJA: conj(A);  CONJB: conj(B)  (so CONJB alone gives A * B^H)
// QOC_3M = 1 (default): Gauss's three-multiplication form of the complex product.  With a = ar + i ai, b = br + i bi
//   K1 = (ar + ai) br,  K2 = ar (bi - br),  K3 = ai (br + bi):   re = K1 - K3,  im = K1 + K2
// every 8x8x4 step costs 3 DMMA instead of 4 (K1 is accumulated once and seeds both the real and the imaginary accumulator);
// the operand sums are one DADD per operand element (6 per 8x8 product) on the same FP64 pipe: 96 + 13 instead of 128 pipe
// clocks per complex 8x8x8 product.  Normwise as accurate as the four-multiplication form (the imaginary part loses the
// componentwise bound); observed deviation from the CPU restatement of the reference: 1e-14 level (tolerances 1e-10 / 1e-8).
#ifndef QOC_3M
#define QOC_3M 1
#endif
template <int NB, bool CONJA, bool CONJB, bool ACC>
__device__ __forceinline__ void mul_nt_impl(const CM<NB>& a, const CM<NB>& b, CM<NB>& c) {
#if QOC_3M
  double sA[NB][NB][2], sB[NB][NB][2], dB[NB][NB][2];
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int l = 0; l < NB; l++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        sA[i][l][h] = CONJA ? a.re[i][l][h] - a.im[i][l][h] : a.re[i][l][h] + a.im[i][l][h];     // ar + ai'
        sB[i][l][h] = CONJB ? b.re[i][l][h] - b.im[i][l][h] : b.re[i][l][h] + b.im[i][l][h];     // br + bi'
        dB[i][l][h] = CONJB ? -b.im[i][l][h] - b.re[i][l][h] : b.im[i][l][h] - b.re[i][l][h];    // bi' - br
      }
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double k0 = 0, k1 = 0;
      if (ACC) { k0 = c.re[i][j][0]; k1 = c.re[i][j][1]; }
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) dmma(k0, k1, sA[i][l][h], b.re[j][l][h]);                   // (c_re +) K1
      double r0 = k0, r1 = k1, m0 = k0, m1 = k1;
      if (ACC) { m0 += c.im[i][j][0] - c.re[i][j][0]; m1 += c.im[i][j][1] - c.re[i][j][1]; }
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double nai = CONJA ? a.im[i][l][h] : dneg(a.im[i][l][h]);                        // -ai'
          dmma(r0, r1, nai, sB[j][l][h]);                                                        // - K3
          dmma(m0, m1, a.re[i][l][h], dB[j][l][h]);                                              // + K2
        }
      c.re[i][j][0] = r0; c.re[i][j][1] = r1; c.im[i][j][0] = m0; c.im[i][j][1] = m1;
    }
#else
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double r0 = 0, r1 = 0, m0 = 0, m1 = 0;
      if (ACC) { r0 = c.re[i][j][0]; r1 = c.re[i][j][1]; m0 = c.im[i][j][0]; m1 = c.im[i][j][1]; }
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double ar = a.re[i][l][h];
          const double ai = CONJA ? dneg(a.im[i][l][h]) : a.im[i][l][h];
          const double br = b.re[j][l][h];
          const double bi = CONJB ? dneg(b.im[j][l][h]) : b.im[j][l][h];
          dmma(r0, r1, ar, br);
          dmma(m0, m1, ar, bi);
          dmma(r0, r1, ai, dneg(bi));
          dmma(m0, m1, ai, br);
        }
      c.re[i][j][0] = r0; c.re[i][j][1] = r1; c.im[i][j][0] = m0; c.im[i][j][1] = m1;
    }
#endif
}
template <int NB, bool CONJA = false, bool CONJB = false>
__device__ __forceinline__ CM<NB> mul_nt(const CM<NB>& a, const CM<NB>& b) {
  CM<NB> c; mul_nt_impl<NB, CONJA, CONJB, false>(a, b, c); return c;
}
template <int NB, bool CONJA = false, bool CONJB = false>
__device__ __forceinline__ void mul_nt_acc(const CM<NB>& a, const CM<NB>& b, CM<NB>& c) {
  mul_nt_impl<NB, CONJA, CONJB, true>(a, b, c);
}

// ---- transpose through the per-warp shared-memory tile (NB*NB*2*TB_PLANE doubles) ------------------------
template <int NB> __device__ __forceinline__ CM<NB> transpose(const Lane& L, const CM<NB>& x, double* tb) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double* pr = tb + ((i * NB + j) * 2) * TB_PLANE + L.g * 10 + 2 * L.q;
      *reinterpret_cast<double2*>(pr) = make_double2(x.re[i][j][0], x.re[i][j][1]);
      *reinterpret_cast<double2*>(pr + TB_PLANE) = make_double2(x.im[i][j][0], x.im[i][j][1]);
    }
  __syncwarp();
  CM<NB> t;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const double* pr = tb + ((j * NB + i) * 2) * TB_PLANE + (2 * L.q) * 10 + L.g;   // block (j,i) of x, transposed
      t.re[i][j][0] = pr[0]; t.re[i][j][1] = pr[10];
      t.im[i][j][0] = pr[TB_PLANE]; t.im[i][j][1] = pr[TB_PLANE + 10];
    }
  __syncwarp();
  return t;
}

// ---- elementwise helpers -------------------------------------------------------------------------------
template <int NB> __device__ __forceinline__ CM<NB> cm_conj(const CM<NB>& x) {
  CM<NB> y; QOC_FOR_CM(NB) { y.re[i][j][e] = x.re[i][j][e]; y.im[i][j][e] = dneg(x.im[i][j][e]); } return y;
}
template <int NB> __device__ __forceinline__ CM<NB> cm_negconj(const CM<NB>& x) {   // -conj(x)
  CM<NB> y; QOC_FOR_CM(NB) { y.re[i][j][e] = dneg(x.re[i][j][e]); y.im[i][j][e] = x.im[i][j][e]; } return y;
}
template <int NB> __device__ __forceinline__ CM<NB> cm_neg(const CM<NB>& x) {
  CM<NB> y; QOC_FOR_CM(NB) { y.re[i][j][e] = dneg(x.re[i][j][e]); y.im[i][j][e] = dneg(x.im[i][j][e]); } return y;
}
template <int NB> __device__ __forceinline__ CM<NB> cm_scale(const CM<NB>& x, double a) {
  CM<NB> y; QOC_FOR_CM(NB) { y.re[i][j][e] = a * x.re[i][j][e]; y.im[i][j][e] = a * x.im[i][j][e]; } return y;
}
template <int NB> __device__ __forceinline__ CM<NB> cm_cscale(const CM<NB>& x, double ar, double ai) {
  CM<NB> y;
  QOC_FOR_CM(NB) {
    y.re[i][j][e] = fma(ar, x.re[i][j][e], -ai * x.im[i][j][e]);
    y.im[i][j][e] = fma(ar, x.im[i][j][e], ai * x.re[i][j][e]);
  }
  return y;
}
template <int NB> __device__ __forceinline__ void cm_axpy(CM<NB>& y, double a, const CM<NB>& x) {
  QOC_FOR_CM(NB) { y.re[i][j][e] = fma(a, x.re[i][j][e], y.re[i][j][e]); y.im[i][j][e] = fma(a, x.im[i][j][e], y.im[i][j][e]); }
}
template <int NB> __device__ __forceinline__ void cm_add(CM<NB>& y, const CM<NB>& x) {
  QOC_FOR_CM(NB) { y.re[i][j][e] += x.re[i][j][e]; y.im[i][j][e] += x.im[i][j][e]; }
}
template <int NB> __device__ __forceinline__ void cm_caxpy(CM<NB>& y, double ar, double ai, const CM<NB>& x) {
  QOC_FOR_CM(NB) {
    y.re[i][j][e] = fma(ar, x.re[i][j][e], fma(-ai, x.im[i][j][e], y.re[i][j][e]));
    y.im[i][j][e] = fma(ar, x.im[i][j][e], fma(ai, x.re[i][j][e], y.im[i][j][e]));
  }
}
// y += a*I  (diagonal element of block (i,i): row g == column 2q+e)
template <int NB> __device__ __forceinline__ void cm_add_identity(const Lane& L, CM<NB>& y, double a) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int e = 0; e < 2; e++)
      if (L.g == 2 * L.q + e) y.re[i][i][e] += a;
}
// Packed storage: [NB*NB][2 (re,im)][32 lanes] double2; every access is one coalesced 512-byte warp transaction.
template <int NB> __device__ __forceinline__ void cm_store(const Lane& L, double2* p, const CM<NB>& x) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      p[((i * NB + j) * 2 + 0) * 32 + L.lane] = make_double2(x.re[i][j][0], x.re[i][j][1]);
      p[((i * NB + j) * 2 + 1) * 32 + L.lane] = make_double2(x.im[i][j][0], x.im[i][j][1]);
    }
}
template <int NB> __device__ __forceinline__ CM<NB> cm_load(const Lane& L, const double2* p) {
  CM<NB> x;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = p[((i * NB + j) * 2 + 0) * 32 + L.lane];
      double2 m = p[((i * NB + j) * 2 + 1) * 32 + L.lane];
      x.re[i][j][0] = r.x; x.re[i][j][1] = r.y; x.im[i][j][0] = m.x; x.im[i][j][1] = m.y;
    }
  return x;
}
// Packed system matrices live in shared memory (SH = true: explicit ld.shared, short scoreboard) or global.
template <bool SH> __device__ __forceinline__ double2 ld_packed(const double2* p) {
  if (SH) {
    double2 v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
  }
  return __ldg(p);
}
template <int NB, bool SH> __device__ __forceinline__ CM<NB> cm_load_sys(const Lane& L, const double2* p) {
  CM<NB> x;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = ld_packed<SH>(p + ((i * NB + j) * 2 + 0) * 32 + L.lane);
      double2 m = ld_packed<SH>(p + ((i * NB + j) * 2 + 1) * 32 + L.lane);
      x.re[i][j][0] = r.x; x.re[i][j][1] = r.y; x.im[i][j][0] = m.x; x.im[i][j][1] = m.y;
    }
  return x;
}
// y += a * (packed matrix at p)
template <int NB, bool SH> __device__ __forceinline__ void cm_axpy_packed(const Lane& L, CM<NB>& y, double a, const double2* p) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = ld_packed<SH>(p + ((i * NB + j) * 2 + 0) * 32 + L.lane);
      double2 m = ld_packed<SH>(p + ((i * NB + j) * 2 + 1) * 32 + L.lane);
      y.re[i][j][0] = fma(a, r.x, y.re[i][j][0]); y.re[i][j][1] = fma(a, r.y, y.re[i][j][1]);
      y.im[i][j][0] = fma(a, m.x, y.im[i][j][0]); y.im[i][j][1] = fma(a, m.y, y.im[i][j][1]);
    }
}
// per-lane partial of sum_ab conj(X[a][b]) * Y[a][b]
template <int NB> __device__ __forceinline__ void cm_dotc_partial(const CM<NB>& x, const CM<NB>& y, double& pr, double& pi) {
  pr = 0; pi = 0;
  QOC_FOR_CM(NB) {
    pr = fma(x.re[i][j][e], y.re[i][j][e], fma(x.im[i][j][e], y.im[i][j][e], pr));
    pi = fma(x.re[i][j][e], y.im[i][j][e], fma(-x.im[i][j][e], y.re[i][j][e], pi));
  }
}
// per-lane partial of Re sum_ab M[a][b] * X[a][b] with M in packed storage
// per-lane partial of Re sum_ab conj(M[a][b]) * X[a][b]
template <int NB, bool SH> __device__ __forceinline__ double cm_redotc_partial(const Lane& L, const double2* m, const CM<NB>& x) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = ld_packed<SH>(m + ((i * NB + j) * 2 + 0) * 32 + L.lane);
      double2 c = ld_packed<SH>(m + ((i * NB + j) * 2 + 1) * 32 + L.lane);
      s = fma(r.x, x.re[i][j][0], s); s = fma(c.x, x.im[i][j][0], s);
      s = fma(r.y, x.re[i][j][1], s); s = fma(c.y, x.im[i][j][1], s);
    }
  return s;
}
template <int NB, bool SH> __device__ __forceinline__ double cm_redot_partial(const Lane& L, const double2* m, const CM<NB>& x) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = ld_packed<SH>(m + ((i * NB + j) * 2 + 0) * 32 + L.lane);
      double2 c = ld_packed<SH>(m + ((i * NB + j) * 2 + 1) * 32 + L.lane);
      s = fma(r.x, x.re[i][j][0], s); s = fma(-c.x, x.im[i][j][0], s);
      s = fma(r.y, x.re[i][j][1], s); s = fma(-c.y, x.im[i][j][1], s);
    }
  return s;
}

// Conjugation step  P W P'  ( = nt(P, X) with X = nt(conj P, W) = (W P')^T ), the body of the closed-system recursion and of
// the density-type sweeps.  P is the left operand of both products, so in the three-multiplication form the variant with the
// operand sums on the LEFT side (K1 = ar (br + bi), K2 = (ai - ar) br, K3 = (ar + ai) bi) needs only s+ = pr + pi and
// s- = pr - pi for BOTH products (signs flip on the integer pipe): 8 instead of 12 DADD per element pair.
template <int NB> __device__ __forceinline__ CM<NB> conj_by(const CM<NB>& P, const CM<NB>& W) {
#if QOC_3M
  double nsp[NB][NB][2], nsm[NB][NB][2];                      // -(pr + pi),  pi - pr
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int l = 0; l < NB; l++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        nsp[i][l][h] = dneg(P.re[i][l][h] + P.im[i][l][h]);
        nsm[i][l][h] = P.im[i][l][h] - P.re[i][l][h];
      }
  CM<NB> X, R;
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    // pass 0: X = conj(P) W^T  (ai' = -pi: ai' - ar = -s+, ar + ai' = s-);  pass 1: R = P X^T  (ai - ar = -s-, ar + ai = s+)
    const CM<NB>& B = pass == 0 ? W : X;
    CM<NB>& C = pass == 0 ? X : R;
    double sB[NB][NB][2];
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) sB[j][l][h] = B.re[j][l][h] + B.im[j][l][h];
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) {
        double k0 = 0, k1 = 0;
#pragma unroll
        for (int l = 0; l < NB; l++)
#pragma unroll
          for (int h = 0; h < 2; h++) dmma(k0, k1, P.re[i][l][h], sB[j][l][h]);                 // K1 = ar (br + bi)
        double r0 = k0, r1 = k1, m0 = k0, m1 = k1;
#pragma unroll
        for (int l = 0; l < NB; l++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            dmma(r0, r1, pass == 0 ? nsm[i][l][h] : nsp[i][l][h], B.im[j][l][h]);               // - K3 = -(ar + ai') bi
            dmma(m0, m1, pass == 0 ? nsp[i][l][h] : nsm[i][l][h], B.re[j][l][h]);               // + K2 = (ai' - ar) br
          }
        C.re[i][j][0] = r0; C.re[i][j][1] = r1; C.im[i][j][0] = m0; C.im[i][j][1] = m1;
      }
  }
  return R;
#else
  const CM<NB> X = mul_nt<NB, true, false>(P, W);
  return mul_nt<NB>(P, X);
#endif
}
// G * G for an anti-Hermitian G, given Gt = G^T = -conj(G): the right operand's sums are the left operand's up to signs
// (br = -gr, bi = gi: br + bi = gi - gr, bi - br = gr + gi), 4 instead of 6 DADD per element pair.
template <int NB> __device__ __forceinline__ CM<NB> square_antiherm(const CM<NB>& G) {
#if QOC_3M
  double sp[NB][NB][2], dm[NB][NB][2];                        // gr + gi,  gi - gr
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int l = 0; l < NB; l++)
#pragma unroll
      for (int h = 0; h < 2; h++) { sp[i][l][h] = G.re[i][l][h] + G.im[i][l][h]; dm[i][l][h] = G.im[i][l][h] - G.re[i][l][h]; }
  CM<NB> C;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double k0 = 0, k1 = 0;
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) dmma(k0, k1, sp[i][l][h], dneg(G.re[j][l][h]));              // K1 = (ar + ai) br,  br = -gr
      double r0 = k0, r1 = k1, m0 = k0, m1 = k1;
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          dmma(r0, r1, dneg(G.im[i][l][h]), dm[j][l][h]);                                       // - K3 = -ai (br + bi)
          dmma(m0, m1, G.re[i][l][h], sp[j][l][h]);                                             // + K2 = ar (bi - br)
        }
      C.re[i][j][0] = r0; C.re[i][j][1] = r1; C.im[i][j][0] = m0; C.im[i][j][1] = m1;
    }
  return C;
#else
  return mul_nt<NB>(G, cm_negconj<NB>(G));
#endif
}

// ---- warp reductions ----------------------------------------------------------------------------------
template <int GS> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = GS / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// Sum EIGHT values over aligned groups of GS lanes (GS = 8, 16, 32) with the halving butterfly
// (7 / 8 / 9 shuffles instead of 24 / 32 / 40).  On return lane l holds in v[0] the group sum of value
// index ((l % GS) * 8) / GS.
template <int GS> __device__ __forceinline__ void group_sum8(int lane, double (&v)[8]) {
  int o = GS / 2;
  {
    bool hi = lane & o;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double keep = hi ? v[4 + i] : v[i], send = hi ? v[i] : v[4 + i];
      v[i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    }
    o >>= 1;
  }
  {
    bool hi = lane & o;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      double keep = hi ? v[2 + i] : v[i], send = hi ? v[i] : v[2 + i];
      v[i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    }
    o >>= 1;
  }
  {
    bool hi = lane & o;
    double keep = hi ? v[1] : v[0], send = hi ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    o >>= 1;
  }
  for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(FULL_MASK, v[0], o);
}

}  // namespace qoc
