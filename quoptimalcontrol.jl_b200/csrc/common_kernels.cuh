// Non-template kernels of the small-dimension path: matrix packing and the deterministic ensemble reduction.
// Included by exactly one translation unit (k_small_fused.cu); everyone else uses the launch_* wrappers of params.h.
#pragma once
#include "params.h"

namespace qoc {

// Pack caller matrices (column-major complex D x D) into the warp layout.  One thread per packed double2.
//   dst[(og*nmat_dst + mat_dst)*E + ((i*NB+j)*2+ri)*32 + lane] ; slot s of group og reads source matrix
//   src + src_index(og, s)*src_stride, transposed if asked; everything outside the chain's D x D block is 0.
__global__ void pack_kernel(const PackParams p) {
  const int E = p.NB * p.NB * 2 * 32;
  long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long)p.n_og * E) return;
  int og = (int)(tid / E), rem = (int)(tid - (long)og * E);
  int lane = rem & 31, ri = (rem >> 5) & 1, blk = rem >> 6;
  int i = blk / p.NB, j = blk - i * p.NB;
  int g = lane >> 2, q = lane & 3;
  int DPc = 8 * p.NB / p.CPW;
  double v[2];
  for (int e = 0; e < 2; e++) {
    int row = 8 * i + g, col = 8 * j + 2 * q + e;
    int s = row / DPc;
    int rr = row - s * DPc, cc = col - s * DPc;
    double val = 0.0;
    if (cc >= 0 && cc < DPc && rr < p.D && cc < p.D) {
      long member = p.pack_mode == 0 ? (long)og * p.CPW + s : og;
      if (member > p.n_src - 1) member = p.n_src - 1;
      const double2* m = p.src + member * p.src_stride;
      double2 z = p.transpose ? m[(size_t)rr * p.D + cc] : m[(size_t)cc * p.D + rr];
      val = ri ? (p.scale_re * z.y + p.scale_im * z.x) : (p.scale_re * z.x - p.scale_im * z.y);
    }
    v[e] = val;
  }
  p.dst[((size_t)og * p.nmat_dst + p.mat_dst) * E + rem] = make_double2(v[0], v[1]);
}

// Deterministic weighted ensemble reduction  F[r] = sum_k w_k fom[r,k],  G[r,:] = sum_k w_k grad[r,k,:]
// (/root/reference/src/solve.jl:171-191), two fixed-order passes: pass 1 folds `chunk` consecutive members (k ascending)
// into one of at most RED_MAX_CHUNKS partial rows, pass 2 folds the partial rows (8 interleaved lanes per element, then a
// fixed-order combine).  With a single chunk pass 1 writes the result itself (part == out, no pass 2).
__global__ void reduce_members_pass1(const double* __restrict__ gradc, const double* __restrict__ fomc,
                                     const double* __restrict__ wts, double* __restrict__ part,
                                     int M, int NK, int chunk, int nchunks, int ch0) {
  // grid: (ceil((NK+1)/256) * R, rows of this launch); part[r][chunk][NK+1] (entry 0 = fom); the launch covers the partial
  // rows ch0 .. ch0 + gridDim.y - 1 (a chain range of the chunk-parallel closed-system mode reduces its own members)
  const int bpr = (NK + 1 + blockDim.x - 1) / blockDim.x;
  int r = blockIdx.x / bpr;
  int e = (blockIdx.x - r * bpr) * blockDim.x + threadIdx.x;
  if (e > NK) return;
  int ch = ch0 + blockIdx.y;
  int k0 = ch * chunk, k1 = min(M, k0 + chunk);
  double s = 0.0;
  if (e == 0) { for (int k = k0; k < k1; k++) s += wts[k] * fomc[(size_t)r * M + k]; }
  else if (gradc) {
    const double* g = gradc + (size_t)r * M * NK + (e - 1);
#pragma unroll 4
    for (int k = k0; k < k1; k++) s += wts[k] * __ldg(g + (size_t)k * NK);
  }
  part[((size_t)r * nchunks + ch) * (NK + 1) + e] = s;
}
__global__ void __launch_bounds__(32 * RED_LANES) reduce_members_pass2(const double* __restrict__ part, double* __restrict__ out, int NK, int nchunks) {
  // block: 32 elements x RED_LANES partial-row lanes; grid: ceil((NK+1)/32) * R
  __shared__ double sm[RED_LANES][33];
  const int bpr = (NK + 1 + 31) / 32;
  const int r = blockIdx.x / bpr;
  const int e = (blockIdx.x - r * bpr) * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
  double s = 0.0;
  if (e <= NK)
    for (int ch = lane; ch < nchunks; ch += RED_LANES) s += part[((size_t)r * nchunks + ch) * (NK + 1) + e];
  sm[lane][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane == 0 && e <= NK) {
    double t = sm[0][threadIdx.x];
#pragma unroll
    for (int l = 1; l < RED_LANES; l++) t += sm[l][threadIdx.x];
    out[(size_t)r * (NK + 1) + e] = t;
  }
}

}  // namespace qoc
