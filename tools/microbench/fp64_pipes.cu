// Microbenchmark: FP64 pipe throughput/latency on B200 (sm_100a).
// Measures DFMA, DMMA (mma.sync f64 shapes m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16), mixed DFMA+DMMA,
// SHFL and LDS.64 rates, so that kernel designs and the FP64 roofline denominator rest on measurements.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* d, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template<int NACC>
__global__ void k_mma884(double* out, int iters) {
  double d[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { d[i][0] = i; d[i][1] = threadIdx.x; }
  double a = threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}
template<int NACC, int SHAPE>  // SHAPE 0: m16n8k4, 1: m16n8k8, 2: m16n8k16
__global__ void k_mma16(double* out, int iters) {
  double d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { d[i][0] = i; d[i][1] = threadIdx.x; d[i][2] = 1; d[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-9 + i * 1e-10;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1e-9 * (i + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (SHAPE == 0) mma1684(d[i], a, b[0]);
      else if (SHAPE == 1) mma1688(d[i], a, b);
      else mma16816(d[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}

// mixed: even warps DFMA (8 acc), odd warps DMMA m8n8k4 (8 acc)
__global__ void k_mixed(double* out, int iters, double a, double b) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i][0] = i; d[i][1] = threadIdx.x; }
    double fa = threadIdx.x * 1e-9, fb = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) mma884(d[i][0], d[i][1], fa, fb);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1];
  } else {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters * 8; it++) {   // 8 DFMA ~ 1 DMMA in MACs
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
  }
  if (s == 123.456) out[0] = s;
}

template<int NCH>
__global__ void k_shfl(double* out, int iters) {
  int v[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) v[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 3)) + 1;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += v[i];
  if (s == 123456789) out[0] = s;
}

// LDS.64 / LDS.128 rate: each lane reads distinct consecutive addresses
template<int W>  // W = 1: 64-bit, 2: 128-bit
__global__ void k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  int base = (threadIdx.x & 31) * W + (threadIdx.x >> 5) * 64;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int idx = (base + u * 64 * W + it) & 4095 & ~(W - 1);
      if (W == 1) s += sm[idx];
      else { double2 t = *reinterpret_cast<double2*>(&sm[idx]); s += t.x + t.y; }
    }
  }
  if (s == 123.456) out[0] = s;
}

struct Res { float ms; };
template<class F> float timeit(F launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clockRate %d kHz\n", p.name, sms, clk_khz);
  double* out; CK(cudaMalloc(&out, 1024));
  const int iters = 20000;
  printf("%-28s %6s %6s %10s %12s %14s\n", "test", "warps", "nacc", "ms", "TFLOP/s", "MAC/ns/SM");
  int wlist[] = {1, 4, 8, 16, 32};
  for (int wi = 0; wi < 5; wi++) {
    int warps = wlist[wi]; int threads = warps * 32;
    dim3 g(sms), b(threads);
    auto rep = [&](const char* name, int nacc, float ms, double macs_per_thread_or_warp, bool perwarp) {
      double macs = (double)sms * (perwarp ? warps : threads) * macs_per_thread_or_warp;
      printf("%-28s %6d %6d %10.4f %12.3f %14.2f\n", name, warps, nacc, ms, 2 * macs / (ms * 1e-3) / 1e12, macs / (ms * 1e6) / sms);
    };
    { float ms = timeit([&]{ k_dfma<1><<<g,b>>>(out, iters, 1.0000001, 1e-9); }); rep("dfma", 1, ms, (double)iters * 1, false); }
    { float ms = timeit([&]{ k_dfma<4><<<g,b>>>(out, iters, 1.0000001, 1e-9); }); rep("dfma", 4, ms, (double)iters * 4, false); }
    { float ms = timeit([&]{ k_dfma<16><<<g,b>>>(out, iters, 1.0000001, 1e-9); }); rep("dfma", 16, ms, (double)iters * 16, false); }
    { float ms = timeit([&]{ k_mma884<1><<<g,b>>>(out, iters); }); rep("dmma.m8n8k4", 1, ms, (double)iters * 1 * 256, true); }
    { float ms = timeit([&]{ k_mma884<4><<<g,b>>>(out, iters); }); rep("dmma.m8n8k4", 4, ms, (double)iters * 4 * 256, true); }
    { float ms = timeit([&]{ k_mma884<16><<<g,b>>>(out, iters); }); rep("dmma.m8n8k4", 16, ms, (double)iters * 16 * 256, true); }
    { float ms = timeit([&]{ k_mma16<1,0><<<g,b>>>(out, iters); }); rep("dmma.m16n8k4", 1, ms, (double)iters * 1 * 512, true); }
    { float ms = timeit([&]{ k_mma16<8,0><<<g,b>>>(out, iters); }); rep("dmma.m16n8k4", 8, ms, (double)iters * 8 * 512, true); }
    { float ms = timeit([&]{ k_mma16<1,1><<<g,b>>>(out, iters); }); rep("dmma.m16n8k8", 1, ms, (double)iters * 1 * 1024, true); }
    { float ms = timeit([&]{ k_mma16<8,1><<<g,b>>>(out, iters); }); rep("dmma.m16n8k8", 8, ms, (double)iters * 8 * 1024, true); }
    { float ms = timeit([&]{ k_mma16<1,2><<<g,b>>>(out, iters); }); rep("dmma.m16n8k16", 1, ms, (double)iters * 1 * 2048, true); }
    { float ms = timeit([&]{ k_mma16<8,2><<<g,b>>>(out, iters); }); rep("dmma.m16n8k16", 8, ms, (double)iters * 8 * 2048, true); }
    if (warps >= 4) {
      float ms = timeit([&]{ k_mixed<<<g,b>>>(out, iters, 1.0000001, 1e-9); });
      // half warps DMMA: iters*8*256 MACs per warp; half DFMA: iters*8*8*32 MACs per warp = iters*2048
      double macs = (double)sms * (warps / 2) * ((double)iters * 8 * 256 + (double)iters * 8 * 8 * 32);
      printf("%-28s %6d %6d %10.4f %12.3f %14.2f\n", "mixed dfma+dmma884", warps, 8, ms, 2 * macs / (ms * 1e-3) / 1e12, macs / (ms * 1e6) / sms);
    }
    { float ms = timeit([&]{ k_shfl<8><<<g,b>>>(out, iters); });
      printf("%-28s %6d %6d %10.4f   %10.2f warp-shfl/ns/SM\n", "shfl32 (8 chains)", warps, 8, ms, (double)warps * iters * 8 / (ms * 1e6)); }
    { float ms = timeit([&]{ k_lds<1><<<g,b,32768>>>(out, iters); });
      printf("%-28s %6d %6d %10.4f   %10.2f B/ns/SM\n", "lds.64", warps, 8, ms, (double)threads * iters * 8 * 8 / (ms * 1e6)); }
    { float ms = timeit([&]{ k_lds<2><<<g,b,32768>>>(out, iters); });
      printf("%-28s %6d %6d %10.4f   %10.2f B/ns/SM\n", "lds.128", warps, 8, ms, (double)threads * iters * 8 * 16 / (ms * 1e6)); }
  }
  // long sustained DFMA to see clocks under load (about 2 s)
  {
    dim3 g(sms * 2), b(512);
    float ms = timeit([&]{ k_dfma<16><<<g,b>>>(out, 400000, 1.0000001, 1e-9); });
    double macs = (double)sms * 2 * 512 * 400000.0 * 16;
    printf("sustained dfma 2x512thr: %.2f ms  %.3f TFLOP/s\n", ms, 2 * macs / (ms * 1e-3) / 1e12);
    ms = timeit([&]{ k_mma884<16><<<g,b>>>(out, 400000); });
    macs = (double)sms * 2 * 16 * 400000.0 * 16 * 256;
    printf("sustained dmma884 2x16warps: %.2f ms  %.3f TFLOP/s\n", ms, 2 * macs / (ms * 1e-3) / 1e12);
  }
  return 0;
}
