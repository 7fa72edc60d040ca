// Microbenchmark: achievable DMMA rate for the GRAPE small-D instruction mix on B200.
//   mode 0: chain of complex 8x8 NT products, 2 accumulator chains of 4 dependent DMMAs (what chain_kernel does)
//   mode 1: same with 4 accumulator chains of 2 (+4 DADD)
//   mode 2: mode 0 + ~16 dependent-ish DFMA per product (the scalar FP64 share of the real kernel)
//   mode 3: mode 2 + one smem transpose per product
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../quoptimalcontrol.jl_b200/csrc -o product_mix product_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "warp_mat.cuh"
using namespace qoc;

template <int MODE>
__global__ void k(double* out, int iters) {
  __shared__ double tbs[32 * 160];
  const Lane L(threadIdx.x & 31);
  double* tb = tbs + (threadIdx.x >> 5) * 160;
  CM<1> A, B;
  QOC_FOR_CM(1) { A.re[i][j][e] = 1e-3 * (L.lane + e); A.im[i][j][e] = 2e-3 * (L.lane - e); B.re[i][j][e] = 1e-3 * e; B.im[i][j][e] = 1e-3; }
  for (int it = 0; it < iters; it++) {
    CM<1> C;
    if (MODE == 1) {
      double rr0 = 0, rr1 = 0, ii0 = 0, ii1 = 0, ri0 = 0, ri1 = 0, ir0 = 0, ir1 = 0;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        dmma(rr0, rr1, A.re[0][0][h], B.re[0][0][h]); dmma(ii0, ii1, A.im[0][0][h], B.im[0][0][h]);
        dmma(ri0, ri1, A.re[0][0][h], B.im[0][0][h]); dmma(ir0, ir1, A.im[0][0][h], B.re[0][0][h]);
      }
      C.re[0][0][0] = rr0 - ii0; C.re[0][0][1] = rr1 - ii1; C.im[0][0][0] = ri0 + ir0; C.im[0][0][1] = ri1 + ir1;
    } else {
      C = mul_nt<1>(A, B);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int r = 0; r < 4; r++) { cm_axpy<1>(C, 1e-3, B); }   // 16 DFMA, 4-deep chains
    }
    if (MODE >= 3) B = transpose<1>(L, C, tb);
    A = C;
  }
  double s = 0; QOC_FOR_CM(1) s += A.re[i][j][e] + A.im[i][j][e] + B.re[i][j][e];
  if (s == 123.456) out[0] = s;
}
template <int MODE> void run(int sms, double* out) {
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 20, 24, 32}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dm = (double)sms * warps * iters * 8;                     // DMMAs
    double cyc_per_prod = ms * 1e-3 * 1.965e9 / iters / (warps / 4.0);   // SMSP cycles per product per warp-slot
    printf("mode %d warps/SM %2d  %.3f ms  DMMA TFLOP/s %.2f  cycles/product/SMSP %.1f (128 = DMMA peak)\n", MODE, warps, ms,
           dm * 512 / (ms * 1e-3) / 1e12, cyc_per_prod);
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  run<0>(p.multiProcessorCount, out); run<1>(p.multiProcessorCount, out); run<2>(p.multiProcessorCount, out); run<3>(p.multiProcessorCount, out);
  return 0;
}
