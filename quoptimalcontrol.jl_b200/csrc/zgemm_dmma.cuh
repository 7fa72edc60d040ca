// Batched complex-FP64 GEMM on the FP64 tensor pipe (DMMA.8x8x4), for the large-dimension GRAPE path.
//   C[b] = epilogue( op(A[b]) * op(B[b]) ),  op in {N, C = conjugate transpose},  square D x D, D % 64 == 0,
//   interleaved (re, im) column-major matrices (the caller's own layout: Julia ComplexF64 arrays).
// CTA tile 64 x 64 x 8 (or 32 x 32 x 8), 256 threads = 8 warps of 32 x 16 (16 x 8) complex outputs, 4-stage cp.async pipeline.
// Operand tiles are staged in the orientation they have in global memory (so every cp.async moves a contiguous
// run) with row strides chosen so that the 16-byte fragment loads are bank-conflict free:
//   "KM" tile [k][mn], stride 66 (== 2 mod 8)   when the operand is contiguous along m / n
//   "MK" tile [mn][k], stride 12 (== 4 mod 8)   when it is contiguous along k
// One LDS.128 fetches (re, im) of a fragment element; the complex product is 4 real DMMAs with the imaginary
// left fragment negated once (4M, not Gauss 3M, to stay inside the 1e-10 parity budget without analysis).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qoc {

// Tile geometry: 8 warps as 2 (m) x 4 (n); each warp owns MI x NJ 8x8 blocks.  MI = 4, NJ = 2: CTA tile 64 x 64 (the
// workhorse); MI = 2, NJ = 1: CTA tile 32 x 32, used when D pads tighter to a multiple of 32 (e.g. 5 qubits, D = 32).
constexpr int GB_K = 8, GB_STAGES = 4, GB_THREADS = 256;
constexpr int GB_LD_MK = 12;
template <int MI> struct GemmGeom {
  static constexpr int BT = 16 * MI;                 // CTA tile edge (square tiles: BN = 32 * NJ with NJ = MI / 2)
  static constexpr int LD_KM = BT + 2;               // == 2 (mod 8)
  static constexpr int TILE_ELEMS = (GB_K * LD_KM > BT * GB_LD_MK) ? GB_K * LD_KM : BT * GB_LD_MK;
  static constexpr int SMEM_BYTES = GB_STAGES * 2 * TILE_ELEMS * 16;
  static constexpr int LOADS = BT * GB_K / GB_THREADS;   // cp.async per thread per operand per stage (2 or 1)
};

// A batched matrix argument: matrix b lives at ptr + index(b) * stride, index(b) = table ? table[b] + offset : b.
// A negative table entry means "this batch entry is idle in this launch" (ragged chunks).
struct BatchedMat {
  const double2* ptr;
  long stride;          // in double2
  const int* table;
  int offset;
};
// cz: 0 = real coefficient; 1 = additionally multiplied by the complex device scalar *cz_ptr; 2 = by its conjugate
struct EpiAux { BatchedMat m; double coef; int pow2; int cz; };   // contributes coef * 2^(-s*pow2) * [cz] * M
struct EpiOut {
  BatchedMat m;         // destination (ptr is written through a const_cast)
  double alpha; int alpha_pow2; int alpha_cz;   // alpha * 2^(-s*alpha_pow2) * [cz] * acc
  double ident; int naux;
  EpiAux aux[3];
};
struct GemmParams {
  int D, batch;
  BatchedMat A, B;
  EpiOut out[2];
  int nout;
  const int* s_ptr;     // device scalar: scaling power of the exponential (may be null -> 0)
  const double* cz_ptr; // device complex scalar (re, im) for the cz coefficient modes (may be null)
};

__device__ __forceinline__ const double2* bm_ptr(const BatchedMat& m, int b, bool& idle) {
  long idx = b;
  if (m.table) { int t = m.table[b]; if (t < 0) { idle = true; return nullptr; } idx = (long)t + m.offset; }
  return m.ptr + idx * m.stride;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int OPA, int OPB, int MI>
__global__ void __launch_bounds__(GB_THREADS, MI == 2 ? 4 : 2) zgemm_dmma_kernel(const GemmParams p) {   // 32x32 tiles: 4 CTAs (49 KB each) per SM
  extern __shared__ double2 gsm[];
  constexpr int NJ = MI / 2;
  constexpr int GB_M = GemmGeom<MI>::BT, GB_N = GemmGeom<MI>::BT, GB_LD_KM = GemmGeom<MI>::LD_KM;
  constexpr int GB_TILE_ELEMS = GemmGeom<MI>::TILE_ELEMS;
  const int b = blockIdx.z * gridDim.y + blockIdx.y;     // batches beyond 65535 spill into grid.z
  if (b >= p.batch) return;
  bool idle = false;
  const double2* Ag = bm_ptr(p.A, b, idle);
  const double2* Bg = bm_ptr(p.B, b, idle);
  if (idle) return;
  const int D = p.D;
  const int tiles_n = D / GB_N;
  const int m0 = (blockIdx.x / tiles_n) * GB_M, n0 = (blockIdx.x % tiles_n) * GB_N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm0 = (warp >> 2) * (8 * MI), wn0 = (warp & 3) * (8 * NJ);

  // acc[i][j][c]: block row i (4), block col j (2), c = re0, re1, im0, im1
  double acc[MI][NJ][4];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[i][j][c] = 0.0;

  auto load_tile = [&](int stage, int k0) {
    double2* As = gsm + (size_t)stage * 2 * GB_TILE_ELEMS;
    double2* Bs = As + GB_TILE_ELEMS;
#pragma unroll
    for (int r = 0; r < GemmGeom<MI>::LOADS; r++) {
      int e = tid + r * GB_THREADS;
      if (OPA == 0) { int k = e / GB_M, m = e % GB_M; cp_async16(As + k * GB_LD_KM + m, Ag + (size_t)(k0 + k) * D + m0 + m); }
      else          { int k = e & 7, m = e >> 3;      cp_async16(As + m * GB_LD_MK + k, Ag + (size_t)(m0 + m) * D + k0 + k); }
      if (OPB == 0) { int k = e & 7, n = e >> 3;      cp_async16(Bs + n * GB_LD_MK + k, Bg + (size_t)(n0 + n) * D + k0 + k); }
      else          { int k = e / GB_N, n = e % GB_N; cp_async16(Bs + k * GB_LD_KM + n, Bg + (size_t)(k0 + k) * D + n0 + n); }
    }
  };

  const int KT = D / GB_K;
#pragma unroll
  for (int s = 0; s < GB_STAGES - 1; s++) {
    if (s < KT) load_tile(s, s * GB_K);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<GB_STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + GB_STAGES - 1;
      if (nk < KT) load_tile(nk % GB_STAGES, nk * GB_K);
      cp_async_commit();
    }
    const double2* As = gsm + (size_t)(kt % GB_STAGES) * 2 * GB_TILE_ELEMS;
    const double2* Bs = As + GB_TILE_ELEMS;
#pragma unroll
    for (int kk = 0; kk < GB_K; kk += 4) {
      double2 a[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        a[i] = (OPA == 0) ? As[(kk + q) * GB_LD_KM + wm0 + 8 * i + g] : As[(wm0 + 8 * i + g) * GB_LD_MK + kk + q];
        if (OPA == 1) a[i].y = -a[i].y;
      }
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        bf[j] = (OPB == 0) ? Bs[(wn0 + 8 * j + g) * GB_LD_MK + kk + q] : Bs[(kk + q) * GB_LD_KM + wn0 + 8 * j + g];
        if (OPB == 1) bf[j].y = -bf[j].y;
      }
#pragma unroll
      for (int i = 0; i < MI; i++) {
        const double nai = -a[i].y;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          dmma884(acc[i][j][0], acc[i][j][1], a[i].x, bf[j].x);
          dmma884(acc[i][j][2], acc[i][j][3], a[i].x, bf[j].y);
          dmma884(acc[i][j][0], acc[i][j][1], nai, bf[j].y);
          dmma884(acc[i][j][2], acc[i][j][3], a[i].y, bf[j].x);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: out_o = alpha_o * acc + sum_i coef_i * aux_i + ident_o * I ----
  const int s = p.s_ptr ? *p.s_ptr : 0;
  for (int o = 0; o < p.nout; o++) {
    const EpiOut& eo = p.out[o];
    bool idl = false;
    double2* Cg = const_cast<double2*>(bm_ptr(eo.m, b, idl));
    if (idl) continue;
    const double czr = p.cz_ptr ? p.cz_ptr[0] : 1.0, czi = p.cz_ptr ? p.cz_ptr[1] : 0.0;
    auto ccoef = [&](double c, int pw, int mode, double& cr, double& ci) {
      const double v = c * scalbn(1.0, -s * pw);
      cr = mode ? v * czr : v; ci = mode == 1 ? v * czi : (mode == 2 ? -v * czi : 0.0);
    };
    double ar, ai;
    ccoef(eo.alpha, eo.alpha_pow2, eo.alpha_cz, ar, ai);
    const double2* auxp[3]; double auxr[3], auxi[3];
    for (int x = 0; x < eo.naux; x++) {
      bool dummy = false;
      auxp[x] = bm_ptr(eo.aux[x].m, b, dummy);
      ccoef(eo.aux[x].coef, eo.aux[x].pow2, eo.aux[x].cz, auxr[x], auxi[x]);
    }
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int row = m0 + wm0 + 8 * i + g, col = n0 + wn0 + 8 * j + 2 * q + e;
          const size_t off = (size_t)col * D + row;
          double re = ar * acc[i][j][e] - ai * acc[i][j][2 + e], im = ar * acc[i][j][2 + e] + ai * acc[i][j][e];
          for (int x = 0; x < eo.naux; x++) {
            double2 v = auxp[x][off];
            re = fma(auxr[x], v.x, fma(-auxi[x], v.y, re)); im = fma(auxr[x], v.y, fma(auxi[x], v.x, im));
          }
          if (row == col) re += eo.ident;
          Cg[off] = make_double2(re, im);
        }
  }
}

}  // namespace qoc
