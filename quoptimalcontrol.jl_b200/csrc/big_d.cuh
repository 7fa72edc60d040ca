// Large-dimension GRAPE path (D > 16): every product is a batched DMMA complex GEMM (zgemm_dmma.cuh).
// The slice dependence of the forward/backward sweeps is broken into Cn chunks that advance in lock step
// (batch = Cn GEMMs per launch), after the chunk boundary states have been obtained from chunk-total propagators:
//   phase 1  P_t = exp(-i dt H_t) for all slices  (assemble, 1-norm -> scaling s, T8 in 3 fused GEMMs, s squarings)
//            -- pw_prop_save!, /root/reference/src/timeevolution.jl:98-110
//   phase 2  T_c = product of the chunk's propagators (Cn independent chains of L-1 products)
//   phase 3  boundary states S[start_c], costates C[start_c] through the T_c (short sequential chains)
//   phase 4  forward / backward sweeps inside all chunks concurrently, storing S_t and C_t
//            -- evolve_func!, /root/reference/src/GRAPE.jl:53-75, 216-251
//   phase 5  W_t = S_t C_t' (- C_t' S_t for density types) for all slices, then K trace-dots per slice over the
//            non-zeros of B_c  -- grad_func!, /root/reference/src/GRAPE.jl:261-287;  fom_func, cost_functions.jl:99-111
// Exact gradient (GRAD_EXACT): phases 1-4 as above, then per slice the direction matrix Y_t, the Frechet derivative
// L(G_t, Y_t) of the Taylor-8 scheme (8 batched GEMM launches + 2 per squaring) and the same trace-dots.
#pragma once
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "../../include/qocgrape.h"
#include "zgemm_dmma.cuh"
#include "pure_state.cuh"
#include "params.h"       // launch_reduce_pass1/2
#include "big_api.h"

namespace qoc {

constexpr double BT8_X1 = 0.10836465678522780852, BT8_X2 = 0.027091164196306952131, BT8_X3 = 0.66666666666666666667,
                 BT8_X4 = 0.54676145797072405251, BT8_X5 = 0.16112557339541759283, BT8_X6 = 0.014090917158378207731,
                 BT8_X7 = 0.033792797010870504141, BT8_Y2 = 0.13549236135285063166;

// ---------------------------------------------------------------------------------------------- small kernels
// out_t[e] = f(A[e] + sum_j x[t][j] B_j[e]); mode 0/2: -i*dt*H (generator), mode 1: H.  One thread per element,
// looping over a range of slices so A and B_j are read once.
// blockIdx.z = chain q of the batch: member[q] selects the system, pulse[q] the pulse; chain q owns the N+1 slots
// [q*(N+1), (q+1)*(N+1)) of `out` (slot N is zeroed: exp(0) = I keeps the batched launches well defined).
__global__ void big_assemble_kernel(const double2* __restrict__ A_all, const double2* __restrict__ B_all, const double* __restrict__ x_all,
                                    double2* __restrict__ out_all, int DD, int K, int N, int slices_per_block, double dt, int mode,
                                    const int* __restrict__ member, const int* __restrict__ pulse) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= DD) return;
  const int q = blockIdx.z;
  const double2* A = A_all + (size_t)member[q] * DD;
  const double2* B = B_all + (size_t)member[q] * (K > 0 ? K : 1) * DD;
  const double* x = x_all + (size_t)pulse[q] * N * K;
  double2* out = out_all + (size_t)q * (N + 1) * DD;
  int t0 = blockIdx.y * slices_per_block, t1 = min(N, t0 + slices_per_block);
  if (t1 == N) out[(size_t)N * DD + e] = make_double2(0.0, 0.0);
  double2 a = A[e];
  for (int t = t0; t < t1; t++) {
    double hr = a.x, hi = a.y;
    for (int j = 0; j < K; j++) {
      double xj = __ldg(x + (size_t)t * K + j);
      double2 bj = __ldg(B + (size_t)j * DD + e);
      hr = fma(xj, bj.x, hr); hi = fma(xj, bj.y, hi);
    }
    out[(size_t)t * DD + e] = mode == 1 ? make_double2(hr, hi) : make_double2(dt * hi, -dt * hr);
  }
}
// per-slice 1-norm bound (column sums of |re|+|im|); one block per slice, one warp per column round-robin.
__global__ void big_norm_kernel(const double2* __restrict__ G, int D, float* __restrict__ norms) {
  __shared__ float wmax[32];
  const double2* g = G + (size_t)blockIdx.x * D * D;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float best = 0.f;
  for (int col = warp; col < D; col += nw) {
    float s = 0.f;
    for (int r = lane; r < D; r += 32) { double2 v = g[(size_t)col * D + r]; s += (float)(fabs(v.x) + fabs(v.y)); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    best = fmaxf(best, s);
  }
  if (lane == 0) wmax[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) { float m = 0.f; for (int i = 0; i < nw; i++) m = fmaxf(m, wmax[i]); norms[blockIdx.x] = m * 1.000001f; }
}
__global__ void big_scale_power_kernel(const float* __restrict__ norms, int N, float theta, int* __restrict__ s_out) {
  __shared__ float sm[256];
  float m = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, norms[i]);
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) { int s = 0; if (sm[0] > theta) { s = ilogbf(sm[0] / theta) + 1; if (s > 60) s = 60; } *s_out = s; }
}
// tau = sum conj(X) .* Y over D*D elements (one block, fixed-order tree); then the figure of merit.
//   unitary (reference): X = S_N, Y = Xt, fom = Re(tau^2);   density: X = Xt, Y = S_N, fom = 1 - |tau|^2/D^2
// blockIdx.x = chain of the batch; X, Y advance by xs, ys elements per chain; tau_fom has 4 doubles per chain.
__global__ void big_fom_kernel(const double2* __restrict__ X_all, long xs, const double2* __restrict__ Y_all, long ys, int DD, int unitary,
                               double invD2, double* __restrict__ tau_fom_all /* [chain][4]: tau_re, tau_im, fom, - */) {
  __shared__ double sr[256], si[256];
  const double2* X = X_all + (size_t)blockIdx.x * xs;
  const double2* Y = Y_all + (size_t)blockIdx.x * ys;
  double* tau_fom = tau_fom_all + (size_t)blockIdx.x * 4;
  double pr = 0, pi = 0;
  for (int e = threadIdx.x; e < DD; e += blockDim.x) {
    double2 a = X[e], b = Y[e];
    pr += a.x * b.x + a.y * b.y; pi += a.x * b.y - a.y * b.x;
  }
  sr[threadIdx.x] = pr; si[threadIdx.x] = pi;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) { if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; } __syncthreads(); }
  if (threadIdx.x == 0) {
    tau_fom[0] = sr[0]; tau_fom[1] = si[0];
    tau_fom[2] = unitary ? sr[0] * sr[0] - si[0] * si[0] : 1.0 - (sr[0] * sr[0] + si[0] * si[0]) * invD2;
  }
}
// g[t][c] = Re( f * sum_nz B_c[a][b] * W_t[b][a] ), one block per (slice, control) over the control's non-zeros.
//   f = i*dt (density) or 2*(+-i dt)*tau (unitary, tau read from tau_fom)
// blockIdx.z = chain q of the batch: W slots q*(N+1)+t, COO lists of the chain's member (coo_off[q]), tau and g per chain.
struct BigTraceParams {
  const double2* W; int D, K, N; const int* coo_ptr_all; const int* coo_off; const int2* coo_idx; const double2* coo_val;
  const double* tau_fom; int unitary; double dt; int sign_static; double* g /* [chain][N][K] */; int exact; double invD2;
};
__global__ void big_trace_kernel(const BigTraceParams p) {
  __shared__ double sr[128];
  int t = blockIdx.x, c = blockIdx.y, q = blockIdx.z;
  const double* tau_fom = p.tau_fom + (size_t)q * 4;
  const int* coo_ptr = p.coo_ptr_all + p.coo_off[q];
  double fr, fi;
  if (p.exact) {   // -(2/D^2) (conj(tau)) (-i dt): tau is already folded into Y for the density types
    const double k = 2.0 * p.dt * p.invD2;
    if (p.unitary) { fr = k * tau_fom[1]; fi = k * tau_fom[0]; } else { fr = 0.0; fi = k; }
  } else if (p.unitary) { double sg = 2.0 * p.sign_static * p.dt; fr = -sg * tau_fom[1]; fi = sg * tau_fom[0]; }
  else { fr = 0.0; fi = p.dt; }
  const double2* W = p.W + ((size_t)q * (p.N + 1) + t) * p.D * p.D;
  double acc = 0;
  for (int e = coo_ptr[c] + threadIdx.x; e < coo_ptr[c + 1]; e += blockDim.x) {
    int2 ab = p.coo_idx[e]; double2 bv = p.coo_val[e];
    double2 w = W[(size_t)ab.x * p.D + ab.y];            // W[b][a] at column a = ab.x, row b = ab.y
    double zr = bv.x * w.x - bv.y * w.y, zi = bv.x * w.y + bv.y * w.x;
    acc += fr * zr - fi * zi;
  }
  sr[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) { if (threadIdx.x < o) sr[threadIdx.x] += sr[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) p.g[((size_t)q * p.N + t) * p.K + c] = sr[0];
}
// FG[0] += w * fom; FG[1 + i] += w * g[i]   (stream order makes the member sum deterministic: k ascending)
// the nb chains of a batch are folded in ascending order (chain = pulse*M + member), member 0 starts its pulse's row
__global__ void big_accumulate_kernel(double* __restrict__ FG_all, const double* __restrict__ tau_fom, const double* __restrict__ g,
                                      const double* __restrict__ wts, const int* __restrict__ member, const int* __restrict__ pulse,
                                      int nb, int NK, int want_grad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > NK) return;
  if (i > 0 && !want_grad) return;
  for (int q = 0; q < nb; q++) {
    double* FG = FG_all + (size_t)pulse[q] * (NK + 1);
    double w = wts[member[q]];
    double v = i == 0 ? tau_fom[(size_t)q * 4 + 2] : g[(size_t)q * NK + i - 1];
    FG[i] = member[q] == 0 ? w * v : FG[i] + w * v;
  }
}

// ---------------------------------------------------------------------------------------------- host state
struct BigState {
  qoc_desc d{};
  int Dp = 0, BT = 64, Cn = 1, Lmax = 1, unitary = 0;   // BT: CTA tile edge of the GEMM kernel (64, or 32 when that pads D tighter)
  // Chains (pulse r, member k) are evaluated Bc at a time: chain q of a batch owns slots [q*(N+1), (q+1)*(N+1)) of every
  // per-slice buffer and chunks [q*Cn, (q+1)*Cn), so all launches simply get a Bc times larger batch dimension.
  int Bc = 1, tabw = 1, CnMax = 1;
  long long ws_tables = 0, ws_coo = 0;
  int batch_c0 = -1, batch_nb = 0;                      // batch currently described on the device (reset by set_system)
  int *chain_member = nullptr, *chain_pulse = nullptr, *chain_coo = nullptr;   // device [Bc]
  double2 *XiQ = nullptr, *XtQ = nullptr;                                     // per-chain copies of Xi, Xt for the batch
  size_t DD = 0;
  std::vector<int> start, len;
  cudaStream_t sA = nullptr, sB = nullptr;
  cudaEvent_t evFork = nullptr, evA = nullptr, evB = nullptr;
  double2 *A = nullptr, *B = nullptr, *Xi = nullptr, *Xt = nullptr;   // [M] padded systems
  double2* buf[7] = {};                                               // (N+1) matrices each
  double2* xbuf[11] = {};                                             // exact gradient only: S, C, Y, T1, T2, dG2, dY1, dL8, dR8, dP, dPb
  int exact = 0, s_last = 0;
  int herm = 0;      // drift and every control Hermitian (exact host check): propagators unitary, conjugation recursion usable
  double2* Pfinal = nullptr;
  std::vector<double2*> pchain;                                       // exact gradient: P_1, P_2, ... (intermediate squares), grown on demand
  double2 *Q = nullptr, *T = nullptr, *tmpF = nullptr, *tmpB = nullptr;   // 2*Cn, Cn, Cn, Cn matrices
  int *tab2A = nullptr, *tab2T = nullptr, *tab0 = nullptr, *tab4F = nullptr, *tab4B = nullptr, *s_dev = nullptr;
  float* norms = nullptr;
  double *tau_fom = nullptr, *gk = nullptr;
  std::vector<int> coo_ptr_h;     // [M][K+1] offsets into the member's COO arrays
  int* coo_ptr = nullptr; int2* coo_idx = nullptr; double2* coo_val = nullptr;
  std::vector<size_t> coo_member_off;
  long long ws = 0;
  bool attr_set = false;
  std::vector<int*> scan_tab;     // closed-system boundary scan: level l lists the chunk columns q*Cn + c with c >= 2^l
  int *bndW0 = nullptr, *bndU = nullptr, *bndOut = nullptr;   // chunk-boundary conjugations: slot of W_0, column of U_c, slot of W[start_{c+1}]
  int *bndS0 = nullptr, *bndEnd = nullptr, *bndCN = nullptr;  // general path, per (chain, chunk): slot 0, slot end_c, slot N
  bool have_props = false;        // Pfinal / T hold the propagators and chunk totals of the last qoc_total_propagator call (one chain)
  PureState pure;                 // vector fast path for pure-state transfers on sparse closed systems (pure_state.cuh)
};

#define BIG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); \
  return e_ == cudaErrorMemoryAllocation ? QOC_ENOMEM : QOC_ECUDA; } } while (0)

template <class T> static int big_alloc(BigState* s, T** p, size_t n, std::string& err) {
  if (n == 0) n = 1;
  BIG_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  s->ws += (long long)(n * sizeof(T));
  return QOC_OK;
}
long long big_workspace(BigState* s) { return s ? s->ws + s->pure.ws : 0; }

static inline void big_free_tables(BigState* s);
void big_destroy(BigState* s) {
  if (!s) return;
  void* ptrs[] = {s->chain_member, s->chain_pulse, s->chain_coo, s->XiQ, s->XtQ, s->A, s->B, s->Xi, s->Xt, s->Q, s->T, s->tmpF, s->tmpB,
                  s->s_dev, s->norms, s->tau_fom, s->gk, s->coo_ptr, s->coo_idx, s->coo_val};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& b : s->buf) if (b) cudaFree(b);
  for (auto& b : s->xbuf) if (b) cudaFree(b);
  for (auto& b : s->pchain) if (b) cudaFree(b);
  big_free_tables(s);
  if (s->sA) cudaStreamDestroy(s->sA);
  if (s->sB) cudaStreamDestroy(s->sB);
  if (s->evFork) cudaEventDestroy(s->evFork);
  if (s->evA) cudaEventDestroy(s->evA);
  if (s->evB) cudaEventDestroy(s->evB);
  pure_free(s->pure);
  delete s;
}

// Chunk count.  Chunk-boundary operators come from parallel prefix / suffix products of the chunk totals (log2 Cn batched
// launches, Cn (log2 Cn + 2) extra products per chain and direction): enough chunks to fill a wave of 2 CTAs/SM together
// with the Bc chains of a batch, otherwise enough to keep the lock-step loops at ~32 launches unless the prefix would cost
// more than a tenth of the per-slice products.  Measured best for single closed chains (N = 2000): 296 / 74 / 37-74 chunks
// at D = 64 / 128 / 256.
static inline int big_chunk_policy(const BigState* s) {
  const int tiles = (s->Dp / s->BT) * (s->Dp / s->BT), N = s->d.N;
  const int fill = (296 + tiles * s->Bc - 1) / (tiles * s->Bc);
  int cap = 1;
  while ((cap + 1) * (std::log2((double)(cap + 1)) + 2.0) <= 0.7 * N) cap++;
  int Cn = std::max(fill, std::min(N / 32, cap));
  if (const char* e = getenv("QOC_BIG_CHUNKS")) Cn = std::max(1, atoi(e));   // tuning override
  return std::max(1, std::min(Cn, std::max(1, N / 2)));
}

static inline void big_free_tables(BigState* s) {
  for (int** t : {&s->tab2A, &s->tab2T, &s->tab0, &s->tab4F, &s->tab4B, &s->bndW0, &s->bndU, &s->bndOut, &s->bndS0, &s->bndEnd, &s->bndCN}) { if (*t) cudaFree(*t); *t = nullptr; }
  for (auto& b : s->scan_tab) if (b) cudaFree(b);
  s->scan_tab.clear();
  s->ws -= s->ws_tables; s->ws_tables = 0;
}

// chunk geometry and the lock-step / scan index tables for Cn chunks per chain (Cn <= CnMax)
static int big_build_chunks(BigState* s, int Cn, std::string& err) {
  const qoc_desc& d = s->d;
  big_free_tables(s);
  const long long ws0 = s->ws;
  s->Cn = Cn;
  s->start.resize(Cn); s->len.resize(Cn);
  for (int c = 0; c < Cn; c++) { int lo = (int)((long)c * d.N / Cn), hi = (int)((long)(c + 1) * d.N / Cn); s->start[c] = lo; s->len[c] = hi - lo; }
  s->Lmax = *std::max_element(s->len.begin(), s->len.end());
  const int Bc = s->Bc;
  s->tabw = Bc * Cn;
  int rc;
  // lock-step index tables over the virtual chain: row j, column q*Cn + c; entries are virtual slot indices q*(N+1) + t
  const int L = s->Lmax, W_ = s->tabw, N1 = d.N + 1;
  std::vector<int> t2A((size_t)L * W_, -1), t2T((size_t)L * W_, -1), t0(W_), t4F((size_t)L * W_, -1), t4B((size_t)L * W_, -1);
  for (int q = 0; q < Bc; q++)
    for (int c = 0; c < Cn; c++) {
      const int col = q * Cn + c, base = q * N1;
      t0[col] = base + s->start[c];
      for (int j = 0; j < L; j++) {
        if (j >= 1 && j < s->len[c]) t2A[(size_t)j * W_ + col] = base + s->start[c] + j;
        if (j == s->len[c] - 1) t2T[(size_t)j * W_ + col] = col;
        if (j < s->len[c] - 1) { t4F[(size_t)j * W_ + col] = base + s->start[c] + j; t4B[(size_t)j * W_ + col] = base + s->start[c] + s->len[c] - 1 - j; }
      }
    }
  auto up = [&](int** dst, const std::vector<int>& v) -> int {
    int r = big_alloc(s, dst, v.size(), err); if (r) return r;
    BIG_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    return QOC_OK;
  };
  if ((rc = up(&s->tab2A, t2A)) || (rc = up(&s->tab2T, t2T)) || (rc = up(&s->tab0, t0)) || (rc = up(&s->tab4F, t4F)) || (rc = up(&s->tab4B, t4B))) return rc;
  // closed-system boundaries by a parallel prefix over the chunk totals (log2(Cn) batched launches instead of 3 Cn sequential ones)
  for (int dd = 1; dd < Cn; dd <<= 1) {
    std::vector<int> lv;
    for (int q = 0; q < Bc; q++) for (int c = dd; c < Cn; c++) lv.push_back(q * Cn + c);
    int* dev = nullptr;
    if ((rc = up(&dev, lv))) return rc;
    s->scan_tab.push_back(dev);
  }
  if (Cn > 1) {
    std::vector<int> w0, uc, wo;
    for (int q = 0; q < Bc; q++) for (int c = 0; c + 1 < Cn; c++) { w0.push_back(q * N1); uc.push_back(q * Cn + c); wo.push_back(q * N1 + s->start[c + 1]); }
    if ((rc = up(&s->bndW0, w0)) || (rc = up(&s->bndU, uc)) || (rc = up(&s->bndOut, wo))) return rc;
  }
  {
    std::vector<int> s0, se, cn;
    for (int q = 0; q < Bc; q++) for (int c = 0; c < Cn; c++) { s0.push_back(q * N1); se.push_back(q * N1 + s->start[c] + s->len[c]); cn.push_back(q * N1 + d.N); }
    if ((rc = up(&s->bndS0, s0)) || (rc = up(&s->bndEnd, se)) || (rc = up(&s->bndCN, cn))) return rc;
  }
  s->ws_tables = s->ws - ws0;
  return QOC_OK;
}

int big_create(BigState** out, const qoc_desc& d, std::string& err, long long& ws_total) {
  BigState* s = new BigState();
  *out = s;
  s->d = d;
  s->exact = d.gradient == QOC_GRAD_EXACT;
  { const int d32 = ((d.D + 31) / 32) * 32; if (d32 % 64 == 32) { s->BT = 32; s->Dp = d32; } else { s->BT = 64; s->Dp = ((d.D + 63) / 64) * 64; } }
  s->DD = (size_t)s->Dp * s->Dp;
  s->unitary = d.sys_type == QOC_UNITARY_GATE;
  int rc;
  BIG_CUDA(cudaStreamCreateWithFlags(&s->sA, cudaStreamNonBlocking));
  BIG_CUDA(cudaStreamCreateWithFlags(&s->sB, cudaStreamNonBlocking));
  BIG_CUDA(cudaEventCreateWithFlags(&s->evFork, cudaEventDisableTiming));
  BIG_CUDA(cudaEventCreateWithFlags(&s->evA, cudaEventDisableTiming));
  BIG_CUDA(cudaEventCreateWithFlags(&s->evB, cudaEventDisableTiming));
  const size_t DD = s->DD;
  {  // batch size: as many chains as fit in half of the free device memory (exact mode keeps its many intermediates per chain)
    size_t fre = 0, tot = 0;
    BIG_CUDA(cudaMemGetInfo(&fre, &tot));
    const double per_chain = (7.5 * (double)(d.N + 1) + 5.0 * std::min(296, std::max(1, d.N / 2))) * DD * sizeof(double2);
    long bc = (long)(0.5 * (double)fre / per_chain);
    bc = std::max(1L, std::min(std::min(bc, 16384L), (long)d.M * d.R));
    if (s->exact) bc = 1;
    if (const char* e = getenv("QOC_BIG_BATCH")) bc = std::max(1L, std::min((long)atoi(e), (long)d.M * d.R));
    s->Bc = (int)bc;
  }
  const int Bc = s->Bc;
  const int Cn = big_chunk_policy(s);
  s->CnMax = Cn;
  if ((rc = big_alloc(s, &s->A, (size_t)d.M * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->B, (size_t)d.M * std::max(d.K, 1) * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->Xi, (size_t)d.M * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->Xt, (size_t)d.M * DD, err))) return rc;
  for (auto& b : s->buf) if ((rc = big_alloc(s, &b, (size_t)Bc * (d.N + 1) * DD, err))) return rc;
  if (s->exact) for (auto& b : s->xbuf) if ((rc = big_alloc(s, &b, (size_t)(d.N + 1) * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->Q, (size_t)2 * Bc * Cn * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->T, (size_t)Bc * Cn * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->tmpF, (size_t)Bc * Cn * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->tmpB, (size_t)Bc * Cn * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->XiQ, (size_t)Bc * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->XtQ, (size_t)Bc * DD, err))) return rc;
  if ((rc = big_alloc(s, &s->chain_member, (size_t)Bc, err))) return rc;
  if ((rc = big_alloc(s, &s->chain_pulse, (size_t)Bc, err))) return rc;
  if ((rc = big_alloc(s, &s->chain_coo, (size_t)Bc, err))) return rc;
  if ((rc = big_alloc(s, &s->s_dev, 1, err))) return rc;
  if ((rc = big_alloc(s, &s->norms, (size_t)Bc * (d.N + 1), err))) return rc;
  if ((rc = big_alloc(s, &s->tau_fom, (size_t)Bc * 4, err))) return rc;
  if ((rc = big_alloc(s, &s->gk, (size_t)Bc * d.N * std::max(d.K, 1), err))) return rc;
  if ((rc = big_build_chunks(s, Cn, err))) return rc;
  ws_total += s->ws;
  return QOC_OK;
}

// upload [count] column-major D x D matrices into zero-padded Dp x Dp slots
static int big_upload_padded(BigState* s, double2* dst, const double* src, size_t count, std::string& err) {
  const int D = s->d.D, Dp = s->Dp;
  BIG_CUDA(cudaMemset(dst, 0, count * s->DD * sizeof(double2)));
  for (size_t m = 0; m < count; m++)
    BIG_CUDA(cudaMemcpy2D(dst + m * s->DD, (size_t)Dp * sizeof(double2), src + m * 2 * (size_t)D * D, (size_t)D * sizeof(double2),
                          (size_t)D * sizeof(double2), D, cudaMemcpyHostToDevice));
  return QOC_OK;
}

int big_set_system(BigState* s, const double* A, const double* B, const double* Xi, const double* Xt, int shared, std::string& err) {
  const qoc_desc& d = s->d;
  const int M = d.M, K = d.K, D = d.D;
  const size_t dd = (size_t)D * D;
  int rc;
  s->batch_c0 = -1; s->pure.batch_c0 = -1; s->have_props = false;
  auto rep = [&](double2* dst, const double* src, size_t per_member, bool sh) -> int {
    for (int k = 0; k < M; k++)
      if ((rc = big_upload_padded(s, dst + (size_t)k * per_member * s->DD, src + (sh ? 0 : (size_t)k * per_member * 2 * dd), per_member, err))) return rc;
    return QOC_OK;
  };
  if ((rc = rep(s->A, A, 1, shared & QOC_SHARED_A))) return rc;
  if (K > 0 && (rc = rep(s->B, B, K, shared & QOC_SHARED_B))) return rc;
  if ((rc = rep(s->Xi, Xi, 1, shared & QOC_SHARED_XI))) return rc;
  if ((rc = rep(s->Xt, Xt, 1, shared & QOC_SHARED_XT))) return rc;
  {  // Hermitian drift and controls => unitary propagators (exact elementwise test, like Julia's ishermitian)
    auto is_herm = [&](const double* Mx) {
      for (int c = 0; c < D; c++)
        for (int r = 0; r <= c; r++) {
          const double* a = Mx + 2 * ((size_t)c * D + r); const double* b = Mx + 2 * ((size_t)r * D + c);
          if (a[0] != b[0] || a[1] != -b[1]) return false;
        }
      return true;
    };
    bool h = true;
    const int nA = (shared & QOC_SHARED_A) ? 1 : M, nB = (shared & QOC_SHARED_B) ? 1 : M;
    for (int k = 0; k < nA && h; k++) h = is_herm(A + 2 * (size_t)k * dd);
    for (int k = 0; k < nB * K && h; k++) h = is_herm(B + 2 * (size_t)k * dd);
    s->herm = h ? 1 : 0;
    if (const char* e = getenv("QOC_BIG_HERM")) s->herm = s->herm && atoi(e) != 0;   // tuning / A-B testing override
  }
  // COO lists of the non-zeros of every control (indices in the padded matrix): tr(B W) = sum_nz B[a][b] W[b][a]
  std::vector<int> ptr; std::vector<int2> idx; std::vector<double2> val;
  s->coo_member_off.assign(M, 0);
  for (int k = 0; k < M; k++) {
    s->coo_member_off[k] = ptr.size();
    const double* Bk = B + ((shared & QOC_SHARED_B) ? 0 : (size_t)k * K * 2 * dd);
    for (int c = 0; c < K; c++) {
      ptr.push_back((int)idx.size());
      for (int col = 0; col < D; col++)
        for (int row = 0; row < D; row++) {
          double re = Bk[2 * ((size_t)c * dd + (size_t)col * D + row)], im = Bk[2 * ((size_t)c * dd + (size_t)col * D + row) + 1];
          if (re != 0.0 || im != 0.0) { idx.push_back(make_int2(row, col)); val.push_back(make_double2(re, im)); }   // (a = row, b = col)
        }
    }
    ptr.push_back((int)idx.size());
  }
  for (void* p : {(void*)s->coo_ptr, (void*)s->coo_idx, (void*)s->coo_val}) if (p) cudaFree(p);
  s->coo_ptr = nullptr; s->coo_idx = nullptr; s->coo_val = nullptr;
  s->ws -= s->ws_coo;                      // a repeated qoc_set_system replaces the lists: keep workspace_bytes honest
  const long long ws_before = s->ws;
  if ((rc = big_alloc(s, &s->coo_ptr, ptr.size(), err)) || (rc = big_alloc(s, &s->coo_idx, idx.size(), err)) || (rc = big_alloc(s, &s->coo_val, val.size(), err))) return rc;
  s->ws_coo = s->ws - ws_before;
  BIG_CUDA(cudaMemcpy(s->coo_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (!idx.empty()) {
    BIG_CUDA(cudaMemcpy(s->coo_idx, idx.data(), idx.size() * sizeof(int2), cudaMemcpyHostToDevice));
    BIG_CUDA(cudaMemcpy(s->coo_val, val.data(), val.size() * sizeof(double2), cudaMemcpyHostToDevice));
  }
  return pure_setup(s->pure, d, A, B, Xi, Xt, shared, s->herm != 0, err);
}

// Replace only the initial / target operators (slice-parallel multi-GPU: every evaluation brings new boundary operators,
// drift and controls stay).  Propagators left by big_total_propagator stay valid.
int big_set_states(BigState* s, const double* Xi, const double* Xt, int shared, std::string& err) {
  if (s->pure.active) { err = "qoc_set_states: the handle runs the pure-state vector path; create it with QOC_FLAG_NO_PURE_STATE"; return QOC_EUNSUPPORTED; }
  const int M = s->d.M;
  const size_t dd = (size_t)s->d.D * s->d.D;
  int rc;
  for (int k = 0; k < M; k++) {
    if ((rc = big_upload_padded(s, s->Xi + (size_t)k * s->DD, Xi + ((shared & QOC_SHARED_XI) ? 0 : (size_t)k * 2 * dd), 1, err))) return rc;
    if ((rc = big_upload_padded(s, s->Xt + (size_t)k * s->DD, Xt + ((shared & QOC_SHARED_XT) ? 0 : (size_t)k * 2 * dd), 1, err))) return rc;
  }
  s->batch_c0 = -1;       // the per-batch copies XiQ / XtQ are stale
  return QOC_OK;
}

// ---------------------------------------------------------------------------------------------- launches
static inline BatchedMat bmat(const double2* p, long stride, const int* table = nullptr, int offset = 0) { return BatchedMat{p, stride, table, offset}; }
static inline EpiOut eout(BatchedMat m, double alpha = 1.0, int apow = 0, double ident = 0.0, int acz = 0) { EpiOut e{}; e.m = m; e.alpha = alpha; e.alpha_pow2 = apow; e.alpha_cz = acz; e.ident = ident; e.naux = 0; return e; }
static inline void eaux(EpiOut& e, BatchedMat m, double coef, int pow2 = 0, int cz = 0) { e.aux[e.naux].m = m; e.aux[e.naux].coef = coef; e.aux[e.naux].pow2 = pow2; e.aux[e.naux].cz = cz; e.naux++; }

static int big_gemm(BigState* s, int opA, int opB, GemmParams& p, cudaStream_t st, std::string& err, qoc_stats& stats) {
  typedef void (*kfn)(const GemmParams);
  kfn fn;
  int smem;
  if (s->BT == 64) {
    fn = opA == 0 ? (opB == 0 ? zgemm_dmma_kernel<0, 0, 4> : zgemm_dmma_kernel<0, 1, 4>) : (opB == 0 ? zgemm_dmma_kernel<1, 0, 4> : zgemm_dmma_kernel<1, 1, 4>);
    smem = GemmGeom<4>::SMEM_BYTES;
  } else {
    fn = opA == 0 ? (opB == 0 ? zgemm_dmma_kernel<0, 0, 2> : zgemm_dmma_kernel<0, 1, 2>) : (opB == 0 ? zgemm_dmma_kernel<1, 0, 2> : zgemm_dmma_kernel<1, 1, 2>);
    smem = GemmGeom<2>::SMEM_BYTES;
  }
  if (!s->attr_set) {
    kfn all4[] = {zgemm_dmma_kernel<0, 0, 4>, zgemm_dmma_kernel<0, 1, 4>, zgemm_dmma_kernel<1, 0, 4>, zgemm_dmma_kernel<1, 1, 4>};
    kfn all2[] = {zgemm_dmma_kernel<0, 0, 2>, zgemm_dmma_kernel<0, 1, 2>, zgemm_dmma_kernel<1, 0, 2>, zgemm_dmma_kernel<1, 1, 2>};
    for (kfn f : all4) BIG_CUDA(cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmGeom<4>::SMEM_BYTES));
    for (kfn f : all2) BIG_CUDA(cudaFuncSetAttribute((const void*)f, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmGeom<2>::SMEM_BYTES));
    s->attr_set = true;
  }
  p.D = s->Dp;
  const int by = std::min(p.batch, 32768);
  dim3 grid((s->Dp / s->BT) * (s->Dp / s->BT), by, (p.batch + by - 1) / by);
  fn<<<grid, GB_THREADS, smem, st>>>(p);
  BIG_CUDA(cudaGetLastError());
  stats.n_launches++; stats.launches_last_eval++;
  return QOC_OK;
}
#define BIG_COUNT() do { BIG_CUDA(cudaGetLastError()); stats.n_launches++; stats.launches_last_eval++; } while (0)

// Describe the batch [c0, c0 + nb) of chains (chain = pulse * M + member) to the device and gather their Xi, Xt.
static int big_set_batch(BigState* s, int c0, int nb, cudaStream_t st, std::string& err) {
  const int M = s->d.M;
  if (c0 == s->batch_c0 && nb == s->batch_nb) return QOC_OK;
  std::vector<int> mem(nb), pul(nb), coo(nb);
  for (int q = 0; q < nb; q++) { mem[q] = (c0 + q) % M; pul[q] = (c0 + q) / M; coo[q] = (int)s->coo_member_off[mem[q]]; }
  BIG_CUDA(cudaMemcpyAsync(s->chain_member, mem.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(s->chain_pulse, pul.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(s->chain_coo, coo.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int q = 0; q < nb; q++) {
    BIG_CUDA(cudaMemcpyAsync(s->XiQ + (size_t)q * s->DD, s->Xi + (size_t)mem[q] * s->DD, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    BIG_CUDA(cudaMemcpyAsync(s->XtQ + (size_t)q * s->DD, s->Xt + (size_t)mem[q] * s->DD, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  }
  BIG_CUDA(cudaStreamSynchronize(st));     // the host vectors above go out of scope
  s->batch_c0 = c0; s->batch_nb = nb;
  return QOC_OK;
}

// phase 1: propagators of the nb chains of the current batch (x_all: [R][N][K] on the device) into s->Pfinal
static int big_propagators_phase(BigState* s, int nb, const double* x_all, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const qoc_desc& d = s->d;
  const size_t DD = s->DD; const long sd = (long)DD;
  const int N = d.N, VB = nb * (N + 1);                   // batch over all virtual slots (slot N of every chain holds G = 0)
  double2 *G = s->buf[0], *G2 = s->buf[1], *Y1 = s->buf[2], *L8 = s->buf[3], *R8 = s->buf[4], *P = s->buf[5], *P2 = s->buf[6];
  dim3 ga((unsigned)((DD + 255) / 256), (N + 63) / 64, nb);
  big_assemble_kernel<<<ga, 256, 0, st>>>(s->A, s->B, x_all, G, (int)DD, d.K, N, 64, d.T / N, 0, s->chain_member, s->chain_pulse);
  BIG_COUNT();
  big_norm_kernel<<<VB, 256, 0, st>>>(G, s->Dp, s->norms);
  BIG_COUNT();
  big_scale_power_kernel<<<1, 256, 0, st>>>(s->norms, VB, (float)(d.expm_theta > 0 ? d.expm_theta : 0.0694), s->s_dev);
  BIG_COUNT();
  int rc;
  GemmParams p{};
  p.batch = VB; p.s_ptr = s->s_dev;
  // G2 = (G/2^s)^2 ; Y1 = x1 G/2^s + x2 G2
  p.A = bmat(G, sd); p.B = bmat(G, sd); p.nout = 2;
  p.out[0] = eout(bmat(G2, sd), 1.0, 2);
  p.out[1] = eout(bmat(Y1, sd), BT8_X2, 2); eaux(p.out[1], bmat(G, sd), BT8_X1, 1);
  if ((rc = big_gemm(s, 0, 0, p, st, err, stats))) return rc;
  // G4 = G2 * Y1 ; L8 = x3 G2 + G4 ; R8 = x4 I + x5 G/2^s + x6 G2 + x7 G4
  p.A = bmat(G2, sd); p.B = bmat(Y1, sd);
  p.out[0] = eout(bmat(L8, sd), 1.0, 0); eaux(p.out[0], bmat(G2, sd), BT8_X3);
  p.out[1] = eout(bmat(R8, sd), BT8_X7, 0, BT8_X4); eaux(p.out[1], bmat(G, sd), BT8_X5, 1); eaux(p.out[1], bmat(G2, sd), BT8_X6);
  if ((rc = big_gemm(s, 0, 0, p, st, err, stats))) return rc;
  // P = I + G/2^s + y2 G2 + L8 * R8
  p.A = bmat(L8, sd); p.B = bmat(R8, sd); p.nout = 1;
  p.out[0] = eout(bmat(P, sd), 1.0, 0, 1.0); eaux(p.out[0], bmat(G, sd), 1.0, 1); eaux(p.out[0], bmat(G2, sd), BT8_Y2);
  if ((rc = big_gemm(s, 0, 0, p, st, err, stats))) return rc;
  // squarings: the count lives on the device; one small synchronising read per batch
  int s_host = 0;
  BIG_CUDA(cudaMemcpyAsync(&s_host, s->s_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  BIG_CUDA(cudaStreamSynchronize(st));
  s->s_last = s_host;
  if (s->exact) {   // keep every intermediate square for the Frechet chain rule (Bc = 1 in exact mode)
    if (s_host > 12) { err = "exact gradient for D > 16 supports at most 12 squarings (||dt*H||_1 <= 284): use more slices"; return QOC_EUNSUPPORTED; }
    while ((int)s->pchain.size() < s_host) {      // allocated on first use
      double2* nbuf = nullptr;
      if ((rc = big_alloc(s, &nbuf, (size_t)(N + 1) * DD, err))) return rc;
      s->pchain.push_back(nbuf);
    }
    std::vector<double2*> chain(1, s->buf[5]);
    chain.insert(chain.end(), s->pchain.begin(), s->pchain.end());
    for (int j = 0; j < s_host; j++) {
      GemmParams q{};
      q.batch = VB; q.A = bmat(chain[j], sd); q.B = bmat(chain[j], sd); q.nout = 1; q.out[0] = eout(bmat(chain[j + 1], sd));
      if ((rc = big_gemm(s, 0, 0, q, st, err, stats))) return rc;
    }
    s->Pfinal = chain[s_host];
    return QOC_OK;
  }
  for (int j = 0; j < s_host; j++) {
    GemmParams q{};
    q.batch = VB; q.A = bmat(P, sd); q.B = bmat(P, sd); q.nout = 1; q.out[0] = eout(bmat(P2, sd));
    if ((rc = big_gemm(s, 0, 0, q, st, err, stats))) return rc;
    std::swap(P, P2);
  }
  if (P != s->buf[5]) std::swap(s->buf[5], s->buf[6]);   // keep "buf[5] is P"
  s->Pfinal = s->buf[5];
  return QOC_OK;
}

// phase 2: chunk totals T[q*Cn + c] (batch nb*Cn, Lmax-1 lock steps)
static int big_chunk_totals(BigState* s, int nb, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const long sd = (long)s->DD; const int Cn = s->Cn, N1 = s->d.N + 1;
  const size_t half = (size_t)s->Bc * Cn * s->DD;
  double2* P = s->Pfinal;
  int rc;
  for (int q = 0; q < nb; q++)
    for (int c = 0; c < Cn; c++)
      if (s->len[c] == 1)
        BIG_CUDA(cudaMemcpyAsync(s->T + (size_t)(q * Cn + c) * s->DD, P + (size_t)(q * N1 + s->start[c]) * s->DD, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  for (int j = 1; j < s->Lmax; j++) {
    GemmParams p{};
    p.batch = nb * Cn;
    p.A = bmat(P, sd, s->tab2A + (size_t)j * s->tabw);
    p.B = j == 1 ? bmat(P, sd, s->tab0) : bmat(s->Q + (size_t)((j - 1) & 1) * half, sd);
    p.nout = 2;
    p.out[0] = eout(bmat(s->Q + (size_t)(j & 1) * half, sd));
    p.out[1] = eout(bmat(s->T, sd, s->tab2T + (size_t)j * s->tabw));
    if ((rc = big_gemm(s, 0, 0, p, st, err, stats))) return rc;
  }
  return QOC_OK;
}

// Inclusive prefix (dir 0: U_c = T_c ... T_0) or suffix (dir 1: V_c = T_{Cn-1} ... T_c) products of the chunk totals of nb
// chains, Kogge-Stone over the chunk columns: log2(Cn) batched launches.  `src` is only read; the levels ping-pong between
// a and b.  *result holds the products, *spare is the other buffer (free for the caller).
static int big_scan(BigState* s, int nb, int dir, const double2* src, double2* a, double2* b, const double2** result, double2** spare,
                    cudaStream_t st, std::string& err, qoc_stats& stats) {
  const size_t DD = s->DD; const long sd = (long)DD;
  const int Cn = s->Cn;
  const size_t pitch = (size_t)Cn * DD * sizeof(double2);
  const double2* in = src;
  double2* out = a;
  int rc, lvl = 0;
  for (int dd = 1; dd < Cn; dd <<= 1, lvl++) {
    const int* tab = s->scan_tab[lvl];                      // columns q*Cn + c, c >= dd
    GemmParams q{}; q.batch = nb * (Cn - dd); q.nout = 1;
    if (dir == 0) { q.A = bmat(in, sd, tab); q.B = bmat(in, sd, tab, -dd); q.out[0] = eout(bmat(out, sd, tab)); }            // X_c X_{c-dd}
    else          { q.A = bmat(in, sd, tab); q.B = bmat(in, sd, tab, -dd); q.out[0] = eout(bmat(out, sd, tab, -dd)); }       // X_{c+dd} X_c
    if ((rc = big_gemm(s, 0, 0, q, st, err, stats))) return rc;
    const size_t off = dir == 0 ? 0 : (size_t)(Cn - dd) * DD;                                                                // finished columns
    BIG_CUDA(cudaMemcpy2DAsync(out + off, pitch, in + off, pitch, (size_t)dd * DD * sizeof(double2), nb, cudaMemcpyDeviceToDevice, st));
    in = out;
    out = out == a ? b : a;
  }
  *result = in;
  *spare = out;
  return QOC_OK;
}

// Closed systems (Hermitian drift and controls): every P_t is unitary, hence V_t = U_N U_t' and
//   W_t = S_t C_t' (- C_t' S_t) = U_t W_0 U_t',   W_0 = Xi C_0' (- C_0' Xi),   C_0 = U_N' Xt (U_N).
// One forward conjugation recursion W_{t+1} = P_t W_t P_t' replaces the separate state and costate sweeps and the
// per-slice W products: 7 GEMMs per slice instead of 11.  Same figure of merit and gradient to rounding.
// All nb chains of the batch advance together (batch dimension nb, nb*Cn or nb*(N+1)).
static int big_eval_batch_unitary(BigState* s, int nb, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const qoc_desc& d = s->d;
  const size_t DD = s->DD; const long sd = (long)DD;
  const int N = d.N, Cn = s->Cn, U = s->unitary;
  const long cs = (long)(N + 1) * (long)DD;          // chain stride in the per-slice buffers
  const long ts = (long)Cn * (long)DD;               // chain stride in the chunk-total buffer
  int rc;
  double2 *P = s->Pfinal, *W = s->buf[3];
  GemmParams p{}; p.batch = nb; p.nout = 1;
  // inclusive prefix products U_c = T_c ... T_0 of every chain
  const double2* Us; double2* tmp;
  if ((rc = big_scan(s, nb, 0, s->T, s->Q, s->tmpF, &Us, &tmp, st, err, stats))) return rc;
  const double2* Un = Us + (size_t)(Cn - 1) * DD; const long us = ts;      // U_N = U_{Cn-1}
  // C_0 = U_N' Xt (U_N)
  double2* C0 = s->tmpB;
  if (U) { p.A = bmat(Un, us); p.B = bmat(s->XtQ, sd); p.out[0] = eout(bmat(C0, sd)); if ((rc = big_gemm(s, 1, 0, p, st, err, stats))) return rc; }
  else {
    p.A = bmat(s->XtQ, sd); p.B = bmat(Un, us); p.out[0] = eout(bmat(tmp, sd)); if ((rc = big_gemm(s, 0, 0, p, st, err, stats))) return rc;
    p.A = bmat(Un, us); p.B = bmat(tmp, sd); p.out[0] = eout(bmat(C0, sd)); if ((rc = big_gemm(s, 1, 0, p, st, err, stats))) return rc;
  }
  // figure of merit: unitary tau = tr(S_N' Xt) = tr(Xi' C_0); density tau = tr(Xt' S_N) = tr(C_0' Xi)
  const double invD2 = 1.0 / ((double)d.D * d.D);
  if (U) big_fom_kernel<<<nb, 256, 0, st>>>(s->XiQ, sd, C0, sd, (int)DD, 1, invD2, s->tau_fom);
  else big_fom_kernel<<<nb, 256, 0, st>>>(C0, sd, s->XiQ, sd, (int)DD, 0, invD2, s->tau_fom);
  BIG_COUNT();
  // W_0 = Xi C_0' (- C_0' Xi) into slot 0 of every chain
  p.A = bmat(s->XiQ, sd); p.B = bmat(C0, sd); p.out[0] = eout(bmat(W, cs)); if ((rc = big_gemm(s, 0, 1, p, st, err, stats))) return rc;
  if (!U) {
    p.A = bmat(C0, sd); p.B = bmat(s->XiQ, sd); p.out[0] = eout(bmat(W, cs), -1.0); eaux(p.out[0], bmat(W, cs), 1.0);
    if ((rc = big_gemm(s, 1, 0, p, st, err, stats))) return rc;
  }
  // chunk-boundary operators W[start_{c+1}] = U_c W_0 U_c' for all chunks of all chains at once
  if (Cn > 1) {
    GemmParams q{}; q.batch = nb * (Cn - 1); q.nout = 1;
    q.A = bmat(W, sd, s->bndW0); q.B = bmat(Us, sd, s->bndU); q.out[0] = eout(bmat(tmp, sd));
    if ((rc = big_gemm(s, 0, 1, q, st, err, stats))) return rc;
    q.A = bmat(Us, sd, s->bndU); q.B = bmat(tmp, sd); q.out[0] = eout(bmat(W, sd, s->bndOut));
    if ((rc = big_gemm(s, 0, 0, q, st, err, stats))) return rc;
  }
  // lock-step conjugation sweeps inside all chunks of all chains: W[t+1] = P_t W[t] P_t'
  for (int j = 0; j < s->Lmax - 1; j++) {
    const int* tF = s->tab4F + (size_t)j * s->tabw;
    GemmParams q{}; q.batch = nb * Cn; q.nout = 1;
    q.A = bmat(W, sd, tF); q.B = bmat(P, sd, tF); q.out[0] = eout(bmat(s->tmpF, sd));
    if ((rc = big_gemm(s, 0, 1, q, st, err, stats))) return rc;
    q.A = bmat(P, sd, tF); q.B = bmat(s->tmpF, sd); q.out[0] = eout(bmat(W, sd, tF, 1));
    if ((rc = big_gemm(s, 0, 0, q, st, err, stats))) return rc;
  }
  if (d.K > 0) {
    BigTraceParams tp;
    tp.W = W; tp.D = s->Dp; tp.K = d.K; tp.N = N; tp.coo_ptr_all = s->coo_ptr; tp.coo_off = s->chain_coo; tp.coo_idx = s->coo_idx; tp.coo_val = s->coo_val;
    tp.tau_fom = s->tau_fom; tp.unitary = U; tp.dt = d.T / N; tp.sign_static = d.convention == QOC_REF_STATIC ? -1 : 1; tp.g = s->gk; tp.exact = 0; tp.invD2 = 0.0;
    big_trace_kernel<<<dim3(N, d.K, nb), 128, 0, st>>>(tp);
    BIG_COUNT();
  }
  return QOC_OK;
}

// General path (Liouvillians, exact gradient, value-only): ONE chain, in slot 0 of the batch buffers.
static int big_eval_member(BigState* s, int want_grad, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const qoc_desc& d = s->d;
  const size_t DD = s->DD; const long sd = (long)DD;
  const int N = d.N, Cn = s->Cn, U = s->unitary;
  const size_t tw = (size_t)s->tabw;
  int rc;
  double2 *P = s->Pfinal, *S = s->exact ? s->xbuf[0] : s->buf[1], *C = s->exact ? s->xbuf[1] : s->buf[2], *W = s->buf[3];
  // ---- phase 3: boundary states (stream sA) and boundary costates (stream sB), short sequential chains ----
  BIG_CUDA(cudaMemcpyAsync(S, s->XiQ, DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(C + (size_t)N * DD, s->XtQ, DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  BIG_CUDA(cudaEventRecord(s->evFork, st));
  BIG_CUDA(cudaStreamWaitEvent(s->sA, s->evFork, 0));
  BIG_CUDA(cudaStreamWaitEvent(s->sB, s->evFork, 0));
  {                                        // S[end_c] = U_c S_0 (U_c'),  U_c = T_c ... T_0, all chunks at once
    const double2* Up; double2* tmp;
    if ((rc = big_scan(s, 1, 0, s->T, s->Q, s->tmpF, &Up, &tmp, s->sA, err, stats))) return rc;
    GemmParams p{}; p.batch = Cn; p.nout = 1;
    if (U) { p.A = bmat(Up, sd); p.B = bmat(S, sd, s->bndS0); p.out[0] = eout(bmat(S, sd, s->bndEnd)); if ((rc = big_gemm(s, 0, 0, p, s->sA, err, stats))) return rc; }
    else {
      p.A = bmat(S, sd, s->bndS0); p.B = bmat(Up, sd); p.out[0] = eout(bmat(tmp, sd)); if ((rc = big_gemm(s, 0, 1, p, s->sA, err, stats))) return rc;
      p.A = bmat(Up, sd); p.B = bmat(tmp, sd); p.out[0] = eout(bmat(S, sd, s->bndEnd)); if ((rc = big_gemm(s, 0, 0, p, s->sA, err, stats))) return rc;
    }
  }
  if (want_grad) {
    {                                      // C[start_c] = V_c' C_N (V_c),  V_c = T_{Cn-1} ... T_c
      const double2* Vp; double2* tmp;
      if ((rc = big_scan(s, 1, 1, s->T, s->Q + (size_t)s->Bc * Cn * DD, s->tmpB, &Vp, &tmp, s->sB, err, stats))) return rc;
      GemmParams p{}; p.batch = Cn; p.nout = 1;
      if (U) { p.A = bmat(Vp, sd); p.B = bmat(C, sd, s->bndCN); p.out[0] = eout(bmat(C, sd, s->tab0)); if ((rc = big_gemm(s, 1, 0, p, s->sB, err, stats))) return rc; }
      else {
        p.A = bmat(C, sd, s->bndCN); p.B = bmat(Vp, sd); p.out[0] = eout(bmat(tmp, sd)); if ((rc = big_gemm(s, 0, 0, p, s->sB, err, stats))) return rc;
        p.A = bmat(Vp, sd); p.B = bmat(tmp, sd); p.out[0] = eout(bmat(C, sd, s->tab0)); if ((rc = big_gemm(s, 1, 0, p, s->sB, err, stats))) return rc;
      }
    }
    // ---- phase 4: lock-step sweeps inside the chunks, forward on sA, backward on sB ----
    for (int j = 0; j < s->Lmax - 1; j++) {
      const int* tF = s->tab4F + (size_t)j * tw;
      const int* tB = s->tab4B + (size_t)j * tw;
      GemmParams p{}; p.batch = Cn; p.nout = 1;
      if (U) {
        p.A = bmat(P, sd, tF); p.B = bmat(S, sd, tF); p.out[0] = eout(bmat(S, sd, tF, 1));                   // S[t+1] = P_t S_t
        if ((rc = big_gemm(s, 0, 0, p, s->sA, err, stats))) return rc;
        p.A = bmat(P, sd, tB); p.B = bmat(C, sd, tB, 1); p.out[0] = eout(bmat(C, sd, tB));                   // C[t] = P_t' C[t+1]
        if ((rc = big_gemm(s, 1, 0, p, s->sB, err, stats))) return rc;
      } else {
        p.A = bmat(S, sd, tF); p.B = bmat(P, sd, tF); p.out[0] = eout(bmat(s->tmpF, sd));                    // S_t P_t'
        if ((rc = big_gemm(s, 0, 1, p, s->sA, err, stats))) return rc;
        p.A = bmat(P, sd, tF); p.B = bmat(s->tmpF, sd); p.out[0] = eout(bmat(S, sd, tF, 1));                 // P_t (S_t P_t')
        if ((rc = big_gemm(s, 0, 0, p, s->sA, err, stats))) return rc;
        p.A = bmat(C, sd, tB, 1); p.B = bmat(P, sd, tB); p.out[0] = eout(bmat(s->tmpB, sd));                 // C[t+1] P_t
        if ((rc = big_gemm(s, 0, 0, p, s->sB, err, stats))) return rc;
        p.A = bmat(P, sd, tB); p.B = bmat(s->tmpB, sd); p.out[0] = eout(bmat(C, sd, tB));                    // P_t' (C[t+1] P_t)
        if ((rc = big_gemm(s, 1, 0, p, s->sB, err, stats))) return rc;
      }
    }
  }
  BIG_CUDA(cudaEventRecord(s->evA, s->sA));
  BIG_CUDA(cudaEventRecord(s->evB, s->sB));
  BIG_CUDA(cudaStreamWaitEvent(st, s->evA, 0));
  BIG_CUDA(cudaStreamWaitEvent(st, s->evB, 0));
  // ---- figure of merit from S[N] and Xt ----
  const double invD2 = 1.0 / ((double)d.D * d.D);
  if (U && !s->exact) big_fom_kernel<<<1, 256, 0, st>>>(S + (size_t)N * DD, 0, s->XtQ, 0, (int)DD, 1, invD2, s->tau_fom);
  else big_fom_kernel<<<1, 256, 0, st>>>(s->XtQ, 0, S + (size_t)N * DD, 0, (int)DD, 0, invD2, s->tau_fom);
  BIG_COUNT();
  if (!want_grad) return QOC_OK;
  BigTraceParams tp;
  tp.D = s->Dp; tp.K = d.K; tp.N = N; tp.coo_ptr_all = s->coo_ptr; tp.coo_off = s->chain_coo; tp.coo_idx = s->coo_idx; tp.coo_val = s->coo_val;
  tp.tau_fom = s->tau_fom; tp.unitary = U; tp.dt = d.T / N; tp.g = s->gk;
  if (s->exact) {
    // ---- exact gradient: Y_t, Frechet derivative Lam_t = L(G_t, Y_t), trace-dots over Lam_t ----
    double2 *G = s->buf[0], *G2 = s->buf[1], *Y1 = s->buf[2], *L8 = s->buf[3], *R8 = s->buf[4];
    double2 *Y = s->xbuf[2], *T1 = s->xbuf[3], *T2 = s->xbuf[4], *dG2 = s->xbuf[5], *dY1 = s->xbuf[6], *dL8 = s->xbuf[7],
            *dR8 = s->xbuf[8], *dP = s->xbuf[9], *dPb = s->xbuf[10];
    const double2* C1 = C + DD;                      // costate after slice t
    GemmParams p{}; p.batch = N; p.s_ptr = s->s_dev; p.cz_ptr = s->tau_fom;
    auto run = [&](int opA, int opB) { return big_gemm(s, opA, opB, p, st, err, stats); };
    if (U) {
      p.A = bmat(S, sd); p.B = bmat(C1, sd); p.nout = 1; p.out[0] = eout(bmat(Y, sd));                        // Y = S_t C_{t+1}'
      if ((rc = run(0, 1))) return rc;
    } else {
      p.A = bmat(P, sd); p.B = bmat(C1, sd); p.nout = 1; p.out[0] = eout(bmat(T1, sd)); if ((rc = run(1, 1))) return rc;   // P' C'
      p.out[0] = eout(bmat(T2, sd)); if ((rc = run(1, 0))) return rc;                                                      // P' C
      p.A = bmat(S, sd); p.B = bmat(T1, sd); p.out[0] = eout(bmat(Y, sd)); if ((rc = run(0, 0))) return rc;                // S P' C'
      p.A = bmat(S, sd); p.B = bmat(T2, sd);
      p.out[0] = eout(bmat(Y, sd), 1.0, 0, 0.0, 1); eaux(p.out[0], bmat(Y, sd), 1.0, 0, 2);                               // tau S' P' C + conj(tau) Y
      if ((rc = run(1, 0))) return rc;
    }
    // dG2 = (Y G + G Y)/4^s ; dY1 = x1 Y/2^s + x2 dG2
    p.A = bmat(Y, sd); p.B = bmat(G, sd); p.nout = 1; p.out[0] = eout(bmat(dG2, sd), 1.0, 2); if ((rc = run(0, 0))) return rc;
    p.A = bmat(G, sd); p.B = bmat(Y, sd); p.nout = 2;
    p.out[0] = eout(bmat(dY1, sd), BT8_X2, 2); eaux(p.out[0], bmat(dG2, sd), BT8_X2); eaux(p.out[0], bmat(Y, sd), BT8_X1, 1);
    p.out[1] = eout(bmat(dG2, sd), 1.0, 2); eaux(p.out[1], bmat(dG2, sd), 1.0);
    if ((rc = run(0, 0))) return rc;
    // dG4 = dG2 Y1 + G2 dY1 ; dR8 = x5 Y/2^s + x6 dG2 + x7 dG4 ; dL8 = x3 dG2 + dG4
    p.A = bmat(dG2, sd); p.B = bmat(Y1, sd); p.nout = 1; p.out[0] = eout(bmat(dL8, sd)); if ((rc = run(0, 0))) return rc;
    p.A = bmat(G2, sd); p.B = bmat(dY1, sd); p.nout = 2;
    p.out[0] = eout(bmat(dR8, sd), BT8_X7); eaux(p.out[0], bmat(dL8, sd), BT8_X7); eaux(p.out[0], bmat(Y, sd), BT8_X5, 1); eaux(p.out[0], bmat(dG2, sd), BT8_X6);
    p.out[1] = eout(bmat(dL8, sd), 1.0); eaux(p.out[1], bmat(dL8, sd), 1.0); eaux(p.out[1], bmat(dG2, sd), BT8_X3);
    if ((rc = run(0, 0))) return rc;
    // dP = dL8 R8 + L8 dR8 + Y/2^s + y2 dG2
    p.A = bmat(dL8, sd); p.B = bmat(R8, sd); p.nout = 1;
    p.out[0] = eout(bmat(dP, sd), 1.0); eaux(p.out[0], bmat(Y, sd), 1.0, 1); eaux(p.out[0], bmat(dG2, sd), BT8_Y2);
    if ((rc = run(0, 0))) return rc;
    p.A = bmat(L8, sd); p.B = bmat(dR8, sd); p.out[0] = eout(bmat(dP, sd), 1.0); eaux(p.out[0], bmat(dP, sd), 1.0);
    if ((rc = run(0, 0))) return rc;
    // squarings: dP <- dP P_j + P_j dP
    std::vector<double2*> chain(1, s->buf[5]);
    chain.insert(chain.end(), s->pchain.begin(), s->pchain.end());
    for (int j = 0; j < s->s_last; j++) {
      p.A = bmat(dP, sd); p.B = bmat(chain[j], sd); p.out[0] = eout(bmat(dPb, sd)); if ((rc = run(0, 0))) return rc;
      p.A = bmat(chain[j], sd); p.B = bmat(dP, sd); p.out[0] = eout(bmat(dPb, sd), 1.0); eaux(p.out[0], bmat(dPb, sd), 1.0);
      if ((rc = run(0, 0))) return rc;
      std::swap(dP, dPb);
    }
    if (d.K > 0) {
      tp.W = dP; tp.sign_static = 1; tp.exact = 1; tp.invD2 = invD2;
      big_trace_kernel<<<dim3(N, d.K, 1), 128, 0, st>>>(tp);
      BIG_COUNT();
    }
    return QOC_OK;
  }
  // ---- phase 5: W_t for all slices, then trace-dots ----
  {
    GemmParams p{}; p.batch = N; p.nout = 1;
    p.A = bmat(S, sd); p.B = bmat(C, sd); p.out[0] = eout(bmat(W, sd));                                      // S_t C_t'
    if ((rc = big_gemm(s, 0, 1, p, st, err, stats))) return rc;
    if (!U) {
      p.A = bmat(C, sd); p.B = bmat(S, sd); p.out[0] = eout(bmat(W, sd), -1.0); eaux(p.out[0], bmat(W, sd), 1.0);   // W - C_t' S_t
      if ((rc = big_gemm(s, 1, 0, p, st, err, stats))) return rc;
    }
  }
  if (d.K > 0) {
    tp.W = W; tp.sign_static = d.convention == QOC_REF_STATIC ? -1 : 1; tp.exact = 0; tp.invD2 = 0.0;
    big_trace_kernel<<<dim3(N, d.K, 1), 128, 0, st>>>(tp);
    BIG_COUNT();
  }
  return QOC_OK;
}

// Pure-state fast path: one sweep kernel (forward and backward CTAs of every chain), one gradient kernel, one fold.
static int pure_eval(BigState* s, const double* x_dev, double* FG_dev, int want_grad, const double* wts_dev, cudaStream_t st,
                     std::string& err, qoc_stats& stats) {
  PureState& ps = s->pure;
  const qoc_desc& d = s->d;
  const int NK = d.N * d.K, total = d.M * d.R, M = d.M;
  for (int c0 = 0; c0 < total; c0 += ps.cap) {
    const int nb = std::min(ps.cap, total - c0);
    if (c0 != ps.batch_c0 || nb != ps.batch_nb) {
      std::vector<int> mem(nb), pul(nb), coo(nb);
      for (int q = 0; q < nb; q++) { mem[q] = (c0 + q) % M; pul[q] = (c0 + q) / M; coo[q] = (int)s->coo_member_off[mem[q]]; }
      BIG_CUDA(cudaMemcpyAsync(ps.member, mem.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
      BIG_CUDA(cudaMemcpyAsync(ps.pulse, pul.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
      BIG_CUDA(cudaMemcpyAsync(ps.coo_off, coo.data(), nb * sizeof(int), cudaMemcpyHostToDevice, st));
      BIG_CUDA(cudaStreamSynchronize(st));
      ps.batch_c0 = c0; ps.batch_nb = nb;
    }
    PureParams pp{};
    pp.D = d.D; pp.K = d.K; pp.N = d.N; pp.tpr_log2 = ps.tpr_log2; pp.nthreads = ps.nthreads; pp.dt = d.T / d.N;
    pp.ucol = ps.ucol; pp.cj = ps.cj; pp.cval = ps.cval; pp.mstruct = ps.mstruct; pp.psi0 = ps.psi0; pp.phi0 = ps.phi0;
    pp.x = x_dev; pp.member = ps.member; pp.pulse = ps.pulse; pp.psi = ps.psi; pp.chi = ps.chi;
    ps.kernel<<<dim3(2, nb), ps.nthreads, ps.smem, st>>>(pp);
    BIG_COUNT();
    PureGradParams gp{};
    gp.D = d.D; gp.K = d.K; gp.N = d.N; gp.dt = d.T / d.N; gp.invD2 = 1.0 / ((double)d.D * d.D);
    gp.psi = ps.psi; gp.chi = ps.chi; gp.coo_ptr_all = s->coo_ptr; gp.coo_off = ps.coo_off; gp.coo_idx = s->coo_idx; gp.coo_val = s->coo_val;
    gp.g = ps.g; gp.tau_fom = ps.tau_fom; gp.fomc = ps.fomc; gp.want_grad = want_grad && d.K > 0; gp.t0 = gp.want_grad ? 0 : d.N - 1;
    pure_grad_kernel<<<dim3(gp.want_grad ? d.N : 1, nb), 256, (size_t)2 * d.D * sizeof(double2), st>>>(gp);
    BIG_COUNT();
    if (nb == total) {       // every chain in one batch (chain = r * M + k): deterministic two-pass weighted member reduction
      BIG_CUDA(launch_reduce_pass1(want_grad ? ps.g : nullptr, ps.fomc, wts_dev, ps.red_nchunks == 1 ? FG_dev : ps.part, M, NK, d.R,
                                   ps.red_chunk, ps.red_nchunks, st));
      stats.n_launches++; stats.launches_last_eval++;
      if (ps.red_nchunks > 1) {
        BIG_CUDA(launch_reduce_pass2(ps.part, FG_dev, NK, d.R, ps.red_nchunks, st));
        stats.n_launches++; stats.launches_last_eval++;
      }
    } else {
      big_accumulate_kernel<<<(NK + 1 + 255) / 256, 256, 0, st>>>(FG_dev, ps.tau_fom, ps.g, wts_dev, ps.member, ps.pulse, nb, NK, want_grad);
      BIG_COUNT();
    }
  }
  return QOC_OK;
}

// reuse: skip the propagator and chunk-total phases and continue from what big_total_propagator left (same pulse, one chain)
int big_eval(BigState* s, const double* x_dev, double* FG_dev, int want_grad, const double* wts_dev, cudaStream_t st,
                           std::string& err, qoc_stats& stats, bool reuse) {
  if (reuse) {
    if (s->pure.active || !s->have_props || s->d.M * s->d.R != 1) {
      err = "qoc_eval_continue: needs a dense-path handle with M = R = 1 and an immediately preceding qoc_total_propagator call";
      return QOC_EINVAL;
    }
  } else s->have_props = false;
  if (s->pure.active) return pure_eval(s, x_dev, FG_dev, want_grad, wts_dev, st, err, stats);
  const qoc_desc& d = s->d;
  const int NK = d.N * d.K, total = d.M * d.R;
  const bool batched = s->herm && !s->exact && want_grad;     // closed-system recursion: Bc chains per pass
  const int step = batched ? s->Bc : 1;
  int rc;
  for (int c0 = 0; c0 < total; c0 += step) {
    const int nb = std::min(step, total - c0);
    if ((rc = big_set_batch(s, c0, nb, st, err))) return rc;
    if (!reuse) {
      if ((rc = big_propagators_phase(s, nb, x_dev, st, err, stats))) return rc;
      if ((rc = big_chunk_totals(s, nb, st, err, stats))) return rc;
    }
    if (batched) { if ((rc = big_eval_batch_unitary(s, nb, st, err, stats))) return rc; }
    else if ((rc = big_eval_member(s, want_grad, st, err, stats))) return rc;
    big_accumulate_kernel<<<(NK + 1 + 255) / 256, 256, 0, st>>>(FG_dev, s->tau_fom, s->gk, wts_dev, s->chain_member, s->chain_pulse, nb, NK, want_grad);
    BIG_COUNT();
  }
  return QOC_OK;
}

// copy [count] padded matrices into a dense D x D output
static int big_unpad(BigState* s, double2* dst, const double2* src, size_t count, cudaStream_t st, std::string& err) {
  const int D = s->d.D, Dp = s->Dp;
  if (D == Dp) { BIG_CUDA(cudaMemcpyAsync(dst, src, count * s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st)); return QOC_OK; }
  for (size_t m = 0; m < count; m++)
    BIG_CUDA(cudaMemcpy2DAsync(dst + m * (size_t)D * D, (size_t)D * sizeof(double2), src + m * s->DD, (size_t)Dp * sizeof(double2),
                               (size_t)D * sizeof(double2), D, cudaMemcpyDeviceToDevice, st));
  return QOC_OK;
}

int big_propagators(BigState* s, const double* x_dev, double2* out, int mode, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const qoc_desc& d = s->d;
  int rc;
  s->have_props = false;
  for (int c = 0; c < d.M * d.R; c++) {
    if ((rc = big_set_batch(s, c, 1, st, err))) return rc;
    const double2* src;
    if (mode == 0) { if ((rc = big_propagators_phase(s, 1, x_dev, st, err, stats))) return rc; src = s->Pfinal; }
    else {
      dim3 ga((unsigned)((s->DD + 255) / 256), (d.N + 63) / 64, 1);
      big_assemble_kernel<<<ga, 256, 0, st>>>(s->A, s->B, x_dev, s->buf[0], (int)s->DD, d.K, d.N, 64, d.T / d.N, mode, s->chain_member, s->chain_pulse);
      BIG_COUNT();
      src = s->buf[0];
    }
    if ((rc = big_unpad(s, out + (size_t)c * d.N * d.D * d.D, src, d.N, st, err))) return rc;
  }
  return QOC_OK;
}

int big_total_propagator(BigState* s, const double* x_dev, double2* out, cudaStream_t st, std::string& err, qoc_stats& stats) {
  const qoc_desc& d = s->d;
  const size_t DD = s->DD;
  int rc;
  s->have_props = false;
  for (int c = 0; c < d.M * d.R; c++) {
    if ((rc = big_set_batch(s, c, 1, st, err))) return rc;
    if ((rc = big_propagators_phase(s, 1, x_dev, st, err, stats))) return rc;
    if ((rc = big_chunk_totals(s, 1, st, err, stats))) return rc;
    const double2* Up; double2* spare;                // U = T_{Cn-1} ... T_0 = last inclusive prefix
    if ((rc = big_scan(s, 1, 0, s->T, s->Q, s->tmpF, &Up, &spare, st, err, stats))) return rc;
    const double2* cur = Up + (size_t)(s->Cn - 1) * DD;
    if ((rc = big_unpad(s, out + (size_t)c * d.D * d.D, cur, 1, st, err))) return rc;
  }
  s->have_props = d.M * d.R == 1;
  return QOC_OK;
}

// ---- device-pointer entry points for the slice-parallel evaluation (qocgrape.cu, qoc_eval_slice) ------------------------
int big_padded_dim(const BigState* s) { return s->Dp; }

// U (padded Dp x Dp) = product of the range's propagators for the pulse x_dev, left in `U_out` (may be an exchange buffer
// that peers read); the propagators and chunk totals stay on the device for big_eval(..., reuse = true).  M = R = 1.
int big_range_propagator_device(BigState* s, const double* x_dev, double2* U_out, cudaStream_t st, std::string& err, qoc_stats& stats) {
  if (s->d.M * s->d.R != 1 || s->pure.active) { err = "slice-parallel evaluation needs a dense-path handle with M = R = 1"; return QOC_EINVAL; }
  int rc;
  s->have_props = false;
  if ((rc = big_set_batch(s, 0, 1, st, err))) return rc;
  if ((rc = big_propagators_phase(s, 1, x_dev, st, err, stats))) return rc;
  if ((rc = big_chunk_totals(s, 1, st, err, stats))) return rc;
  const double2* Up; double2* spare;
  if ((rc = big_scan(s, 1, 0, s->T, s->Q, s->tmpF, &Up, &spare, st, err, stats))) return rc;
  BIG_CUDA(cudaMemcpyAsync(U_out, Up + (size_t)(s->Cn - 1) * s->DD, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  s->have_props = true;
  return QOC_OK;
}
// install padded device matrices as the (single) member's Xi / Xt, asynchronously on st
int big_set_states_device(BigState* s, const double2* Xi_pad, const double2* Xt_pad, cudaStream_t st, std::string& err) {
  BIG_CUDA(cudaMemcpyAsync(s->Xi, Xi_pad, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(s->Xt, Xt_pad, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(s->XiQ, Xi_pad, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  BIG_CUDA(cudaMemcpyAsync(s->XtQ, Xt_pad, s->DD * sizeof(double2), cudaMemcpyDeviceToDevice, st));
  return QOC_OK;
}
// C = op(A) op(B) for single padded Dp x Dp matrices (op: 0 = as is, 1 = conjugate transpose) on the DMMA GEMM kernel;
// A and B may live in a peer GPU's memory (NVLink loads through cp.async)
int big_matmul(BigState* s, int opA, int opB, const double2* A, const double2* B, double2* C, cudaStream_t st, std::string& err, qoc_stats& stats) {
  GemmParams p{};
  p.batch = 1; p.nout = 1;
  p.A = bmat(A, 0); p.B = bmat(B, 0); p.out[0] = eout(bmat(C, 0));
  return big_gemm(s, opA, opB, p, st, err, stats);
}
int big_upload_states_padded(BigState* s, double2* dst, const double* src, std::string& err) { return big_upload_padded(s, dst, src, 1, err); }

}  // namespace qoc
