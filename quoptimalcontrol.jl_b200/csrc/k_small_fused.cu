// Translation unit of the fused warp-per-chain kernels (small_d.cuh) and the non-template helpers
// (pack_kernel, reduce_members_pass1/2).  qocgrape.cu reaches them through the pickers / wrappers of params.h.
#include "small_d.cuh"
#include "common_kernels.cuh"

namespace qoc {

template <int NB, int CPW> static chain_fn pick_chain2(int sys, int grad) {
  if (sys == SYS_UNITARY) {
    if (grad == GRAD_NONE) return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_NONE>;
    if (grad == GRAD_FIRST) return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_FIRST>;
    return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_EXACT>;
  }
  if (grad == GRAD_NONE) return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_NONE>;
  if (grad == GRAD_FIRST) return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_FIRST>;
  return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_EXACT>;
}
chain_fn pick_chain(int NB, int CPW, int sys, int grad) {
  if (NB == 2) return pick_chain2<2, 1>(sys, grad);
  if (CPW == 4) return pick_chain2<1, 4>(sys, grad);
  if (CPW == 2) return pick_chain2<1, 2>(sys, grad);
  return pick_chain2<1, 1>(sys, grad);
}
chain_fn pick_chain_unitary(int NB, int CPW, int sys) {
  const bool u = sys == SYS_UNITARY;
  if (NB == 2) return u ? chain_unitary_kernel<2, 1, SYS_UNITARY> : chain_unitary_kernel<2, 1, SYS_DENSITY>;
  if (CPW == 4) return u ? chain_unitary_kernel<1, 4, SYS_UNITARY> : chain_unitary_kernel<1, 4, SYS_DENSITY>;
  if (CPW == 2) return u ? chain_unitary_kernel<1, 2, SYS_UNITARY> : chain_unitary_kernel<1, 2, SYS_DENSITY>;
  return u ? chain_unitary_kernel<1, 1, SYS_UNITARY> : chain_unitary_kernel<1, 1, SYS_DENSITY>;
}
slice_fn pick_slices(int NB, int CPW) {
  if (NB == 2) return expm_slices_kernel<2, 1>;
  if (CPW == 4) return expm_slices_kernel<1, 4>;
  if (CPW == 2) return expm_slices_kernel<1, 2>;
  return expm_slices_kernel<1, 1>;
}

cudaError_t launch_pack(const PackParams& pp, long total, cudaStream_t st) {
  pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_reduce_pass1(const double* gradc, const double* fomc, const double* wts, double* part, int M, int NK, int R,
                                int chunk, int nchunks, cudaStream_t st, int ch0, int nrows) {
  if (nrows < 0) nrows = nchunks - ch0;
  dim3 g1((unsigned)(((NK + 1 + 255) / 256) * (long)R), nrows);
  reduce_members_pass1<<<g1, 256, 0, st>>>(gradc, fomc, wts, part, M, NK, chunk, nchunks, ch0);
  return cudaGetLastError();
}
cudaError_t launch_reduce_pass2(const double* part, double* out, int NK, int R, int nchunks, cudaStream_t st) {
  dim3 g2((unsigned)(((NK + 1 + 31) / 32) * (long)R));
  reduce_members_pass2<<<g2, 32 * RED_LANES, 0, st>>>(part, out, NK, nchunks);
  return cudaGetLastError();
}

}  // namespace qoc
