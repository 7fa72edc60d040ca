// Warp-resident complex-FP64 block matrices on DMMA (mma.sync.m8n8k4.f64, the native FP64 tensor shape
// on sm_100a: every f64 mma shape lowers to DMMA.8x8x4 SASS).
//
// A D x D complex matrix (D = 8*NB) is held by ONE warp in registers.  Three register layouts exist, all
// per 8x8 block (i = block row, j = block column); g = lane>>2, q = lane&3:
//   CM  "accumulator layout": lane holds X[8i+g][8j+2q+e],   e = 0,1
//   FA  "left-operand layout": lane holds X[8i+g][8j+4h+q],  h = 0,1   (A fragment of block (i,j), k-half h)
//   FB  "right-operand layout": lane holds X[8i+4h+q][8j+g], h = 0,1   (B fragment of block (i,j), k-half h)
// A product needs its left operand in FA and its right operand in FB and delivers CM; CM -> FA / FB
// conversion costs two 64-bit shuffles per real 8x8 block.  Useful identities (used instead of data movement):
//   FA(X^T) = FB(X)   FB(X^T) = FA(X)   FA(X') = conj FB(X)   FB(X') = conj FA(X)
#pragma once
#include <cuda_runtime.h>

namespace qoc {

constexpr unsigned FULL_MASK = 0xffffffffu;

template <int NB> struct CM { double re[NB][NB][2]; double im[NB][NB][2]; };
template <int NB> struct FA { double re[NB][NB][2]; double im[NB][NB][2]; };
template <int NB> struct FB { double re[NB][NB][2]; double im[NB][NB][2]; };

struct Lane {
  int lane, g, q;
  int srcA1, srcA2, srcB1, srcB2;
  bool qlo, qodd, glo, godd;
  __device__ __forceinline__ explicit Lane(int l) {
    lane = l; g = l >> 2; q = l & 3;
    qlo = q < 2; qodd = q & 1; glo = g < 4; godd = g & 1;
    srcA1 = g * 4 + (q >> 1) + (qodd ? 2 : 0);
    srcA2 = g * 4 + (q >> 1) + (qodd ? 0 : 2);
    int lo = q * 4 + (g >> 1), hi = (4 + q) * 4 + (g >> 1);
    srcB1 = godd ? hi : lo;
    srcB2 = godd ? lo : hi;
  }
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---- layout conversions -------------------------------------------------------------------------------
__device__ __forceinline__ void cvtA(const Lane& L, double c0, double c1, double& a0, double& a1) {
  double s1 = L.qlo ? c0 : c1, s2 = L.qlo ? c1 : c0;
  double r1 = __shfl_sync(FULL_MASK, s1, L.srcA1);
  double r2 = __shfl_sync(FULL_MASK, s2, L.srcA2);
  a0 = L.qodd ? r2 : r1;
  a1 = L.qodd ? r1 : r2;
}
__device__ __forceinline__ void cvtB(const Lane& L, double c0, double c1, double& b0, double& b1) {
  double s1 = L.glo ? c0 : c1, s2 = L.glo ? c1 : c0;
  double r1 = __shfl_sync(FULL_MASK, s1, L.srcB1);
  double r2 = __shfl_sync(FULL_MASK, s2, L.srcB2);
  b0 = L.godd ? r2 : r1;
  b1 = L.godd ? r1 : r2;
}
template <int NB> __device__ __forceinline__ FA<NB> to_A(const Lane& L, const CM<NB>& c) {
  FA<NB> a;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      cvtA(L, c.re[i][j][0], c.re[i][j][1], a.re[i][j][0], a.re[i][j][1]);
      cvtA(L, c.im[i][j][0], c.im[i][j][1], a.im[i][j][0], a.im[i][j][1]);
    }
  return a;
}
template <int NB> __device__ __forceinline__ FB<NB> to_B(const Lane& L, const CM<NB>& c) {
  FB<NB> b;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      cvtB(L, c.re[i][j][0], c.re[i][j][1], b.re[i][j][0], b.re[i][j][1]);
      cvtB(L, c.im[i][j][0], c.im[i][j][1], b.im[i][j][0], b.im[i][j][1]);
    }
  return b;
}
// FA(X') = conj(FB(X)) with block indices swapped; FB(X') = conj(FA(X)) likewise.
template <int NB> __device__ __forceinline__ FA<NB> adjA(const FB<NB>& b) {
  FA<NB> a;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) { a.re[i][j][h] = b.re[j][i][h]; a.im[i][j][h] = -b.im[j][i][h]; }
  return a;
}
template <int NB> __device__ __forceinline__ FB<NB> adjB(const FA<NB>& a) {
  FB<NB> b;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) { b.re[i][j][h] = a.re[j][i][h]; b.im[i][j][h] = -a.im[j][i][h]; }
  return b;
}
// plain transposes (no conjugation): FA(X^T) = FB(X), FB(X^T) = FA(X)
template <int NB> __device__ __forceinline__ FA<NB> trA(const FB<NB>& b) {
  FA<NB> a;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) { a.re[i][j][h] = b.re[j][i][h]; a.im[i][j][h] = b.im[j][i][h]; }
  return a;
}
template <int NB> __device__ __forceinline__ FB<NB> trB(const FA<NB>& a) {
  FB<NB> b;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) { b.re[i][j][h] = a.re[j][i][h]; b.im[i][j][h] = a.im[j][i][h]; }
  return b;
}
template <int NB> __device__ __forceinline__ FA<NB> conjF(const FA<NB>& x) {
  FA<NB> y = x;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) y.im[i][j][h] = -x.im[i][j][h];
  return y;
}
template <int NB> __device__ __forceinline__ FB<NB> conjF(const FB<NB>& x) {
  FB<NB> y = x;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) y.im[i][j][h] = -x.im[i][j][h];
  return y;
}

// ---- products -----------------------------------------------------------------------------------------
// c (+)= a*b; four independent DMMA accumulator chains per output block keep the FP64 pipe busy
// (DMMA latency ~26 clk, issue 1 per 16 clk per SM sub-partition; see profiles/r01_fp64_pipes_microbench.txt).
template <int NB, bool ACC>
__device__ __forceinline__ void mul_impl(const FA<NB>& a, const FB<NB>& b, CM<NB>& c) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double rr0 = 0, rr1 = 0, ii0 = 0, ii1 = 0, ri0 = 0, ri1 = 0, ir0 = 0, ir1 = 0;
      if (ACC) { rr0 = c.re[i][j][0]; rr1 = c.re[i][j][1]; ri0 = c.im[i][j][0]; ri1 = c.im[i][j][1]; }
#pragma unroll
      for (int l = 0; l < NB; l++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          dmma(rr0, rr1, a.re[i][l][h], b.re[l][j][h]);
          dmma(ii0, ii1, a.im[i][l][h], b.im[l][j][h]);
          dmma(ri0, ri1, a.re[i][l][h], b.im[l][j][h]);
          dmma(ir0, ir1, a.im[i][l][h], b.re[l][j][h]);
        }
      c.re[i][j][0] = rr0 - ii0; c.re[i][j][1] = rr1 - ii1;
      c.im[i][j][0] = ri0 + ir0; c.im[i][j][1] = ri1 + ir1;
    }
}
template <int NB> __device__ __forceinline__ CM<NB> mul(const FA<NB>& a, const FB<NB>& b) {
  CM<NB> c; mul_impl<NB, false>(a, b, c); return c;
}
template <int NB> __device__ __forceinline__ void mul_acc(const FA<NB>& a, const FB<NB>& b, CM<NB>& c) {
  mul_impl<NB, true>(a, b, c);
}

// ---- elementwise helpers on CM ------------------------------------------------------------------------
#define QOC_FOR_CM(NB) \
  _Pragma("unroll") for (int i = 0; i < NB; i++) \
  _Pragma("unroll") for (int j = 0; j < NB; j++) \
  _Pragma("unroll") for (int e = 0; e < 2; e++)

template <int NB> __device__ __forceinline__ CM<NB> cm_zero() {
  CM<NB> z; QOC_FOR_CM(NB) { z.re[i][j][e] = 0; z.im[i][j][e] = 0; } return z;
}
// y = a*x (real a)
template <int NB> __device__ __forceinline__ CM<NB> cm_scale(const CM<NB>& x, double a) {
  CM<NB> y; QOC_FOR_CM(NB) { y.re[i][j][e] = a * x.re[i][j][e]; y.im[i][j][e] = a * x.im[i][j][e]; } return y;
}
// y = (ar + i ai) * x
template <int NB> __device__ __forceinline__ CM<NB> cm_cscale(const CM<NB>& x, double ar, double ai) {
  CM<NB> y;
  QOC_FOR_CM(NB) {
    y.re[i][j][e] = ar * x.re[i][j][e] - ai * x.im[i][j][e];
    y.im[i][j][e] = ar * x.im[i][j][e] + ai * x.re[i][j][e];
  }
  return y;
}
// y += a*x
template <int NB> __device__ __forceinline__ void cm_axpy(CM<NB>& y, double a, const CM<NB>& x) {
  QOC_FOR_CM(NB) { y.re[i][j][e] = fma(a, x.re[i][j][e], y.re[i][j][e]); y.im[i][j][e] = fma(a, x.im[i][j][e], y.im[i][j][e]); }
}
// y += (ar + i ai)*x
template <int NB> __device__ __forceinline__ void cm_caxpy(CM<NB>& y, double ar, double ai, const CM<NB>& x) {
  QOC_FOR_CM(NB) {
    y.re[i][j][e] += ar * x.re[i][j][e] - ai * x.im[i][j][e];
    y.im[i][j][e] += ar * x.im[i][j][e] + ai * x.re[i][j][e];
  }
}
template <int NB> __device__ __forceinline__ void cm_sub(CM<NB>& y, const CM<NB>& x) {
  QOC_FOR_CM(NB) { y.re[i][j][e] -= x.re[i][j][e]; y.im[i][j][e] -= x.im[i][j][e]; }
}
// y += a*I  (diagonal element of block (i,i): row g == column 2q+e)
template <int NB> __device__ __forceinline__ void cm_add_identity(const Lane& L, CM<NB>& y, double a) {
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int e = 0; e < 2; e++)
      if (L.g == 2 * L.q + e) y.re[i][i][e] += a;
}
// Packed storage of a CM / any 4*NB*NB-doubles-per-lane object: [NB*NB][2 (re,im)][32 lanes] double2.
// Every access is one fully coalesced 512-byte warp transaction.
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
// per-lane partial of sum_ab conj(X[a][b]) * Y[a][b]
template <int NB> __device__ __forceinline__ void cm_dotc_partial(const CM<NB>& x, const CM<NB>& y, double& pr, double& pi) {
  pr = 0; pi = 0;
  QOC_FOR_CM(NB) {
    pr += x.re[i][j][e] * y.re[i][j][e] + x.im[i][j][e] * y.im[i][j][e];
    pi += x.re[i][j][e] * y.im[i][j][e] - x.im[i][j][e] * y.re[i][j][e];
  }
}
// per-lane partial of Re sum_ab M[a][b] * X[a][b] with M in packed storage
template <int NB> __device__ __forceinline__ double cm_redot_partial(const Lane& L, const double2* m, const CM<NB>& x) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      double2 r = m[((i * NB + j) * 2 + 0) * 32 + L.lane];
      double2 c = m[((i * NB + j) * 2 + 1) * 32 + L.lane];
      s += r.x * x.re[i][j][0] - c.x * x.im[i][j][0] + r.y * x.re[i][j][1] - c.y * x.im[i][j][1];
    }
  return s;
}

// ---- warp reductions ----------------------------------------------------------------------------------
// Sum over aligned groups of GS lanes; result in every lane of the group.
template <int GS> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = GS / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
template <int GS> __device__ __forceinline__ float group_maxf(float v) {
#pragma unroll
  for (int o = GS / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
// Sum EIGHT values over aligned groups of GS lanes (GS = 8, 16, 32) with the halving butterfly
// (9 / 8 / 7 shuffles instead of 40 / 32 / 24).  On return lane l holds in v[0] the group sum of value
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
