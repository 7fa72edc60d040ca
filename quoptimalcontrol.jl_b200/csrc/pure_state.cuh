// Pure-state fast path of the large-dimension GRAPE evaluation.
//
// Applies when (StateTransfer, first-order gradient, D > 16) and, checked exactly on the host in qoc_set_system,
//   * the drift and every control are Hermitian (closed system: P_t = exp(-i dt H_t) is unitary),
//   * Xi = psi psi' and Xt = phi phi' are rank-1 (pure states, e.g. |0..0><0..0| -> |1..1><1..1|),
//   * the union sparsity pattern of (A, B_1..B_K) has few entries per row (spin-chain Hamiltonians).
// Then S_t = psi_t psi_t', C_t = chi_t chi_t' with psi_{t+1} = P_t psi_t, chi_t = P_t' chi_{t+1}, and everything the
// reference computes (src/GRAPE.jl:235-251, :275-287; src/cost_functions.jl:103-111) follows from the vectors:
//   o_t = chi_t' psi_t,  b_{c,t} = chi_t' B_c psi_t,
//   g[c,t] = Re tr(i dt C_t' [B_c, S_t]) = -2 dt Im(conj(o_t) b_{c,t}),     fom = 1 - |tr(C' S) / D|^2 = 1 - |o|^4 / D^2.
// P_t psi is applied as a Taylor series in the sparse H_t (O(nnz) per term instead of O(D^3) per slice).  The result is
// the same F and G to rounding; the arithmetic is NOT the 9 dense products per slice the roofline contract credits, so
// bench.py reports this path separately from the contract figure.
//
// One CTA per (chain, direction): the whole time sweep of a chain is one kernel.  Every thread keeps its row's share of the
// union sparsity pattern (columns, contributions, assembled entries of -+i dt H_t) in registers, shared memory holds the
// ping-pong vector, one __syncthreads per Taylor term.  Chains (members x pulses) are independent CTAs.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

namespace qoc {

struct PureParams {
  int D, K, N, tpr_log2, nthreads;
  double dt;
  // per-thread static data of structure s, slot i (< LPT), contribution c (< CW): index ((s * LPT + i) * nthreads + tid)
  const int* ucol;          // column of the thread's i-th union entry (padding: 0 with zero contributions)
  const int* cj;            // [..][CW] coefficient index (0 = drift, 1 + j = control j; padding 0 with value 0)
  const double2* cval;      // [..][CW]
  const int* mstruct;       // [M] structure index of member m
  const double2 *psi0, *phi0;   // [M][D]
  const double* x;          // [R][N][K]
  const int *member, *pulse;    // [chains of this launch]
  double2 *psi, *chi;       // [chain][N+1][D]
};

__constant__ double PURE_RK[48] = {
    0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15,
    1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30,
    1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45,
    1.0 / 46, 1.0 / 47};

// registers a thread needs for its LPT entries with CW contributions each (static data + assembled values) plus a margin
constexpr int pure_regs(int LPT, int CW) { return LPT * (5 * CW + 5) + 24; }
constexpr int pure_max_threads(int LPT, int CW) { return pure_regs(LPT, CW) <= 64 ? 1024 : pure_regs(LPT, CW) <= 128 ? 512 : pure_regs(LPT, CW) <= 255 ? 256 : 0; }

// Row r of dt*H_t is shared by TPR = 2^tpr_log2 adjacent lanes; each keeps LPT union entries (column, contributions and the
// assembled value) in registers.  Shared memory only holds the two ping-pong vectors, so the per-term cost is the gather
// vin[col] (16 B per non-zero through the shared-memory pipe) plus one __syncthreads.
// dynamic shared memory: v0[D] v1[D] double2 | red[32] double | xeff[K+1] double
template <int LPT, int CW>
__global__ void __launch_bounds__(pure_max_threads(LPT, CW)) pure_sweep_kernel(const PureParams p) {
  extern __shared__ double2 psm[];
  const int D = p.D, K = p.K, N = p.N, nt = p.nthreads;
  const int dir = blockIdx.x, q = blockIdx.y;
  const int m = p.member[q], ms = p.mstruct[m];
  double2* vin = psm;
  double2* vout = vin + D;
  double* red = reinterpret_cast<double*>(vout + D);
  double* xeff = red + 32;
  const int tid = threadIdx.x, TPR = 1 << p.tpr_log2;
  const int r = tid >> p.tpr_log2;
  const bool writer = (tid & (TPR - 1)) == 0 && r < D;

  int col[LPT], cjr[LPT][CW];
  double2 cv[LPT][CW];
#pragma unroll
  for (int i = 0; i < LPT; i++) {
    const size_t base = ((size_t)ms * LPT + i) * nt + tid;
    col[i] = p.ucol[base];
#pragma unroll
    for (int c = 0; c < CW; c++) { cjr[i][c] = p.cj[base * CW + c]; cv[i][c] = p.cval[base * CW + c]; }
  }
  const double2* start = (dir == 0 ? p.psi0 : p.phi0) + (size_t)m * D;
  double2* store = (dir == 0 ? p.psi : p.chi) + (size_t)q * (N + 1) * D;
  double2 acc = make_double2(0.0, 0.0);
  if (r < D) acc = start[r];
  if (writer) { vin[r] = acc; store[(size_t)(dir == 0 ? 0 : N) * D + r] = acc; }
  const double* x = p.x + (size_t)p.pulse[q] * N * K;
  if (tid == 0) xeff[0] = 1.0;
  const double sgn = dir == 0 ? 1.0 : -1.0;     // forward: exp(-i dt H) ; backward: P' = exp(+i dt H)
  const int nwarps = (blockDim.x + 31) >> 5;

  double xnext = tid < K ? __ldg(x + (size_t)(dir == 0 ? 0 : N - 1) * K + tid) : 0.0;
  for (int step = 0; step < N; step++) {
    const int t = dir == 0 ? step : N - 1 - step;
    if (tid < K) {
      xeff[1 + tid] = xnext;
      if (step + 1 < N) xnext = __ldg(x + (size_t)(dir == 0 ? t + 1 : t - 1) * K + tid);    // prefetch: hidden behind this slice
    }
    __syncthreads();
    // ---- assemble this thread's entries of dt * H_t; infinity-norm bound (= 1-norm: H is Hermitian) ----
    double2 h[LPT];
    double rs = 0.0;
#pragma unroll
    for (int i = 0; i < LPT; i++) {
      double hr = 0.0, hi = 0.0;
#pragma unroll
      for (int c = 0; c < CW; c++) { const double xe = xeff[cjr[i][c]]; hr = fma(xe, cv[i][c].x, hr); hi = fma(xe, cv[i][c].y, hi); }
      hr *= p.dt; hi *= p.dt;
      h[i] = make_double2(hr, hi);
      rs += fabs(hr) + fabs(hi);
    }
    for (int o = 1; o < TPR; o <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);        // row sum over the row's lanes
    for (int o = TPR; o < 32; o <<= 1) rs = fmax(rs, __shfl_xor_sync(0xffffffffu, rs, o));  // max over the warp's rows
    if ((tid & 31) == 0) red[tid >> 5] = rs;
    __syncthreads();
    double nb = 0.0;
    for (int w = 0; w < nwarps; w++) nb = fmax(nb, red[w]);     // every thread reads the same values: uniform
    // ---- exp(-+ i dt H_t) applied as nsub Taylor sub-steps with ||dt H / nsub|| <= 1 ----
    const int nsub = nb > 1.0 ? (int)ceil(nb) : 1;
    const double inv_nsub = nsub > 1 ? 1.0 / nsub : 1.0, theta = nb * inv_nsub;
    // generator entries G = (-+ i) dt H / nsub:  (-i)(a + ib) = b - ia
    const double gs = sgn * inv_nsub;
#pragma unroll
    for (int i = 0; i < LPT; i++) h[i] = make_double2(gs * h[i].y, -gs * h[i].x);
    for (int sub = 0; sub < nsub; sub++) {
      // u_k = G u_{k-1} is propagated unscaled (|u_k| <= theta^k); the 1/k! only enters the accumulation, which is off the
      // critical path LDS -> FMA chain -> STS -> barrier.  The stopping bound theta^k / k! is kept one term ahead.
      double bound = theta, ifact = 1.0;
      double rk1 = 1.0, rk2 = 0.5;                      // 1/k and 1/(k+1): the constant-bank load runs one term ahead of its use
      for (int k = 1; k < 46; k++) {
        // (G u) with even / odd entries on separate dependency chains (FP64 latency is ~20 cycles per dependent op)
        double a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0}, b1[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < LPT; i++) {
          const double2 v = vin[col[i]];
          a1[i & 1] = fma(h[i].x, v.x, a1[i & 1]); a2[i & 1] = fma(h[i].y, v.y, a2[i & 1]);
          b1[i & 1] = fma(h[i].x, v.y, b1[i & 1]); b2[i & 1] = fma(h[i].y, v.x, b2[i & 1]);
        }
        double wr = (a1[0] + a1[1]) - (a2[0] + a2[1]), wi = (b1[0] + b1[1]) + (b2[0] + b2[1]);
        for (int o = 1; o < TPR; o <<= 1) { wr += __shfl_xor_sync(0xffffffffu, wr, o); wi += __shfl_xor_sync(0xffffffffu, wi, o); }
        if (writer) vout[r] = make_double2(wr, wi);
        ifact *= rk1;
        acc.x = fma(wr, ifact, acc.x); acc.y = fma(wi, ifact, acc.y);
        __syncthreads();
        double2* tmp = vin; vin = vout; vout = tmp;
        if (bound < 1e-18) break;                       // uniform: every thread holds the same bound
        bound *= theta * rk2;
        rk1 = rk2; rk2 = PURE_RK[k + 2];
      }
      // next sub-step / slice starts from the accumulated vector
      if (writer) vin[r] = acc;
      if (sub + 1 < nsub) __syncthreads();
    }
    if (writer) store[(size_t)(dir == 0 ? t + 1 : t) * D + r] = acc;
    // the __syncthreads after the xeff load of the next slice orders vin for the next term
  }
}

typedef void (*pure_kfn)(const PureParams);
template <int CW> static inline pure_kfn pure_kernel_lpt(int LPT) {
  switch (LPT) {
    case 1: return pure_sweep_kernel<1, CW>;
    case 2: return pure_sweep_kernel<2, CW>;
    case 3: return pure_sweep_kernel<3, CW>;
    case 4: return pure_sweep_kernel<4, CW>;
    case 5: return pure_sweep_kernel<5, CW>;
    case 6: return pure_sweep_kernel<6, CW>;
    case 7: return pure_sweep_kernel<7, CW>;
    case 8: return pure_sweep_kernel<8, CW>;
    case 9: return pure_sweep_kernel<9, CW>;
    case 10: return pure_sweep_kernel<10, CW>;
    case 12: return pure_sweep_kernel<12, CW>;
    default: return nullptr;
  }
}
static inline pure_kfn pure_kernel_for(int LPT, int CW) {
  return CW == 1 ? pure_kernel_lpt<1>(LPT) : CW == 2 ? pure_kernel_lpt<2>(LPT) : CW == 4 ? pure_kernel_lpt<4>(LPT) : nullptr;
}

// g[q][t][c] = -2 dt Im(conj(o_t) * chi_t' B_c psi_t), one block per (slice, chain), one warp per control (round-robin).
// The block of slice N-1 also writes the figure of merit 1 - |o|^4 / D^2 (reference index: src/GRAPE.jl:77, :94).
struct PureGradParams {
  int D, K, N; double dt, invD2;
  const double2 *psi, *chi;
  const int* coo_ptr_all; const int* coo_off; const int2* coo_idx; const double2* coo_val;
  double* g; double* tau_fom; double* fomc /* [chain], may be null */; int want_grad; int t0;
};
__global__ void __launch_bounds__(256) pure_grad_kernel(const PureGradParams p) {
  extern __shared__ double2 gsm2[];
  __shared__ double rr[8], ri[8];
  const int t = p.t0 + blockIdx.x, q = blockIdx.y, D = p.D;
  double2* ps = gsm2;
  double2* ch = ps + D;
  const double2* psi = p.psi + ((size_t)q * (p.N + 1) + t) * D;
  const double2* chi = p.chi + ((size_t)q * (p.N + 1) + t) * D;
  double pr = 0.0, pi = 0.0;
  for (int r = threadIdx.x; r < D; r += blockDim.x) {
    const double2 a = psi[r], b = chi[r];
    ps[r] = a; ch[r] = b;
    pr += b.x * a.x + b.y * a.y; pi += b.x * a.y - b.y * a.x;        // conj(chi) * psi
  }
  for (int o = 16; o; o >>= 1) { pr += __shfl_xor_sync(0xffffffffu, pr, o); pi += __shfl_xor_sync(0xffffffffu, pi, o); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { rr[warp] = pr; ri[warp] = pi; }
  __syncthreads();
  double orr = 0.0, oi = 0.0;
  for (int w = 0; w < nw; w++) { orr += rr[w]; oi += ri[w]; }
  if (t == p.N - 1 && threadIdx.x == 0) {
    const double o2 = orr * orr + oi * oi;
    double* tf = p.tau_fom + (size_t)q * 4;
    tf[0] = o2; tf[1] = 0.0; tf[2] = 1.0 - o2 * o2 * p.invD2;
    if (p.fomc) p.fomc[q] = tf[2];
  }
  if (!p.want_grad) return;
  const int* coo_ptr = p.coo_ptr_all + p.coo_off[q];
  for (int c = warp; c < p.K; c += nw) {
    double br = 0.0, bi = 0.0;
    for (int e = coo_ptr[c] + lane; e < coo_ptr[c + 1]; e += 32) {
      const int2 ab = p.coo_idx[e]; const double2 bv = p.coo_val[e];     // B_c[a][b], a = row, b = col
      const double2 xa = ch[ab.x], yb = ps[ab.y];
      const double zr = bv.x * yb.x - bv.y * yb.y, zi = bv.x * yb.y + bv.y * yb.x;    // B[a][b] psi[b]
      br += xa.x * zr + xa.y * zi; bi += xa.x * zi - xa.y * zr;                       // conj(chi[a]) * (.)
    }
    for (int o = 16; o; o >>= 1) { br += __shfl_xor_sync(0xffffffffu, br, o); bi += __shfl_xor_sync(0xffffffffu, bi, o); }
    if (lane == 0) p.g[((size_t)q * p.N + t) * p.K + c] = -2.0 * p.dt * (orr * bi - oi * br);   // Im(conj(o) b)
  }
}

// ---------------------------------------------------------------------------------------------- host side
struct PureState {
  bool active = false;
  int D = 0, K = 0, N = 0, M = 0, Lu = 0, Cw = 0, cap = 0, nthreads = 0, tpr_log2 = 0, LPT = 1, CWt = 1;
  size_t smem = 0;
  pure_kfn kernel = nullptr;
  int *ucol = nullptr, *cj = nullptr, *mstruct = nullptr, *member = nullptr, *pulse = nullptr, *coo_off = nullptr;
  double2 *cval = nullptr, *psi0 = nullptr, *phi0 = nullptr, *psi = nullptr, *chi = nullptr;
  double *g = nullptr, *tau_fom = nullptr, *fomc = nullptr, *part = nullptr;   // part: partial rows of the member reduction
  int red_chunk = 1, red_nchunks = 1;
  long long ws = 0;
  int batch_c0 = -1, batch_nb = 0;
};

static inline void pure_free(PureState& ps) {
  void* ptrs[] = {ps.ucol, ps.cj, ps.mstruct, ps.member, ps.pulse, ps.coo_off, ps.cval, ps.psi0, ps.phi0, ps.psi, ps.chi, ps.g, ps.tau_fom, ps.fomc, ps.part};
  for (void* p : ptrs) if (p) cudaFree(p);
  ps = PureState();
}

// Xi == v v' exactly up to rounding of the products?  Returns v (scaled so that v v' reproduces Xi) or false.
static inline bool pure_rank1(const double* X, int D, std::vector<double>& v) {
  int j = 0; double best = 0.0, amax = 0.0;
  for (int c = 0; c < D; c++) {
    const double d = X[2 * ((size_t)c * D + c)];
    if (d > best) { best = d; j = c; }
  }
  if (!(best > 0.0)) return false;
  const double inv = 1.0 / std::sqrt(best);
  v.resize(2 * (size_t)D);
  for (int r = 0; r < D; r++) { v[2 * r] = X[2 * ((size_t)j * D + r)] * inv; v[2 * r + 1] = X[2 * ((size_t)j * D + r) + 1] * inv; }
  for (size_t e = 0; e < (size_t)D * D; e++) amax = std::max(amax, std::max(std::fabs(X[2 * e]), std::fabs(X[2 * e + 1])));
  const double tol = 8.0 * 2.220446049250313e-16 * amax;
  for (int c = 0; c < D; c++)
    for (int r = 0; r < D; r++) {
      // (v v')[r][c] = v[r] conj(v[c])
      const double re = v[2 * r] * v[2 * c] + v[2 * r + 1] * v[2 * c + 1], im = v[2 * r + 1] * v[2 * c] - v[2 * r] * v[2 * c + 1];
      if (std::fabs(re - X[2 * ((size_t)c * D + r)]) > tol || std::fabs(im - X[2 * ((size_t)c * D + r) + 1]) > tol) return false;
    }
  return true;
}

#define PURE_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); \
  return e_ == cudaErrorMemoryAllocation ? QOC_ENOMEM : QOC_ECUDA; } } while (0)
template <class T> static int pure_alloc(PureState& ps, T** p, size_t n, std::string& err) {
  if (n == 0) n = 1;
  PURE_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  ps.ws += (long long)(n * sizeof(T));
  return QOC_OK;
}

// Decide whether the fast path applies and build its device data.  Host arrays as passed to qoc_set_system
// (column-major D x D, interleaved complex).  `herm` is the exact Hermitian test already made by the caller.
static inline int pure_setup(PureState& ps, const qoc_desc& d, const double* A, const double* B, const double* Xi, const double* Xt,
                             int shared, bool herm, std::string& err) {
  pure_free(ps);
  if (!herm || d.sys_type != QOC_STATE_TRANSFER || d.gradient != QOC_GRAD_FIRST_ORDER || (d.flags & QOC_FLAG_NO_PURE_STATE)) return QOC_OK;
  bool forced = false;
  if (const char* e = getenv("QOC_PURE_STATE")) { if (atoi(e) == 0) return QOC_OK; forced = true; }
  const int D = d.D, K = d.K, M = d.M, N = d.N;
  // The sweep is latency-bound (~12 dependent Taylor terms per slice, ~1.9 us per slice whatever D is) while chains are
  // independent CTAs; the dense path needs 0.33 / 1.0 / 5 / 32 us per slice at D = 32 / 64 / 128 / 256 for one chain and
  // grows with the number of chains.  Measured crossovers: always at D >= 128, from 3 chains at D = 64, from 16 below.
  const long chains = (long)M * d.R;
  if (!forced && !(D >= 128 || (D >= 64 && chains >= 3) || chains >= 16)) return QOC_OK;
  const size_t dd = (size_t)D * D;
  // pure states?
  const int nXi = (shared & QOC_SHARED_XI) ? 1 : M, nXt = (shared & QOC_SHARED_XT) ? 1 : M;
  std::vector<double> psi0((size_t)M * 2 * D), phi0((size_t)M * 2 * D), v;
  for (int k = 0; k < nXi; k++) { if (!pure_rank1(Xi + 2 * (size_t)k * dd, D, v)) return QOC_OK; std::copy(v.begin(), v.end(), psi0.begin() + (size_t)k * 2 * D); }
  for (int k = 0; k < nXt; k++) { if (!pure_rank1(Xt + 2 * (size_t)k * dd, D, v)) return QOC_OK; std::copy(v.begin(), v.end(), phi0.begin() + (size_t)k * 2 * D); }
  for (int k = 1; k < M; k++) {
    if (nXi == 1) std::copy(psi0.begin(), psi0.begin() + 2 * D, psi0.begin() + (size_t)k * 2 * D);
    if (nXt == 1) std::copy(phi0.begin(), phi0.begin() + 2 * D, phi0.begin() + (size_t)k * 2 * D);
  }
  // union sparsity pattern per structure (one structure if drift and controls are shared by all members)
  const bool shA = shared & QOC_SHARED_A, shB = shared & QOC_SHARED_B;
  const int nstruct = (shA && shB) ? 1 : M;
  struct Contrib { int j; double re, im; };
  std::vector<std::vector<std::vector<std::pair<int, std::vector<Contrib>>>>> rows(nstruct);   // [struct][row] -> list of (col, contribs)
  int Lu = 1, Cw = 1;
  for (int sidx = 0; sidx < nstruct; sidx++) {
    const double* Ak = A + (shA ? 0 : 2 * (size_t)sidx * dd);
    const double* Bk = B + (shB ? 0 : 2 * (size_t)sidx * K * dd);
    rows[sidx].resize(D);
    for (int r = 0; r < D; r++) {
      auto& lst = rows[sidx][r];
      for (int c = 0; c < D; c++) {
        std::vector<Contrib> cs;
        const double* a = Ak + 2 * ((size_t)c * D + r);
        if (a[0] != 0.0 || a[1] != 0.0) cs.push_back({0, a[0], a[1]});
        for (int j = 0; j < K; j++) {
          const double* b = Bk + 2 * ((size_t)j * dd + (size_t)c * D + r);
          if (b[0] != 0.0 || b[1] != 0.0) cs.push_back({1 + j, b[0], b[1]});
        }
        if (!cs.empty()) { Cw = std::max(Cw, (int)cs.size()); lst.emplace_back(c, std::move(cs)); }
      }
      std::stable_sort(lst.begin(), lst.end(), [r](const std::pair<int, std::vector<Contrib>>& a, const std::pair<int, std::vector<Contrib>>& b) {
        return (a.first ^ r) < (b.first ^ r); });
      Lu = std::max(Lu, (int)lst.size());
      if ((size_t)Lu * 4 > (size_t)D) return QOC_OK;             // not sparse enough: the dense GEMM path is the better tool
    }
  }
  // thread geometry: TPR lanes per row (as many as fit), LPT entries per lane, bounded by the register budget
  const int CWt = Cw <= 1 ? 1 : Cw <= 2 ? 2 : Cw <= 4 ? 4 : 0;
  if (!CWt || D > 1024) return QOC_OK;
  int tpr_log2 = -1, LPT = 0, nthreads = 0;
  for (int tl = 0; tl <= 5 && tpr_log2 < 0; tl++) {      // fewest lanes per row first: no shuffles, conflict-free gathers
    const int TPR = 1 << tl, threads = ((D * TPR + 31) / 32) * 32;
    if (threads > 1024) break;
    int lpt = (Lu + TPR - 1) / TPR;
    for (int cand : {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12}) if (cand >= lpt) { lpt = cand; break; }
    if (lpt > 12 || !pure_kernel_for(lpt, CWt)) continue;
    if (threads <= pure_max_threads(lpt, CWt)) { tpr_log2 = tl; LPT = lpt; nthreads = threads; }
  }
  if (tpr_log2 < 0 || K > nthreads) return QOC_OK;      // the sweep stages the K amplitudes of a slice with one thread each
  const int TPR = 1 << tpr_log2;
  ps.smem = 2 * (size_t)D * 16 + 32 * 8 + (size_t)(K + 1) * 8;
  ps.D = D; ps.K = K; ps.N = N; ps.M = M; ps.Lu = Lu; ps.Cw = Cw; ps.CWt = CWt; ps.LPT = LPT; ps.tpr_log2 = tpr_log2; ps.nthreads = nthreads;
  ps.kernel = pure_kernel_for(LPT, CWt);
  // per-thread static data: thread tid = r * TPR + lane keeps union entries l = lane + i * TPR, i < LPT
  const size_t per = (size_t)LPT * nthreads;
  std::vector<int> ucol((size_t)nstruct * per, 0), cj((size_t)nstruct * per * CWt, 0), mstruct(M);
  std::vector<double2> cval((size_t)nstruct * per * CWt, make_double2(0.0, 0.0));
  for (int sidx = 0; sidx < nstruct; sidx++)
    for (int tid = 0; tid < nthreads; tid++) {
      const int r = tid / TPR, lane = tid % TPR;
      if (r >= D) continue;
      const auto& lst = rows[sidx][r];
      for (int i = 0; i < LPT; i++) {
        const int l = lane + i * TPR;
        const size_t base = ((size_t)sidx * LPT + i) * nthreads + tid;
        if (l >= (int)lst.size()) { ucol[base] = r; continue; }
        ucol[base] = lst[l].first;
        for (size_t c = 0; c < lst[l].second.size(); c++) {
          cj[base * CWt + c] = lst[l].second[c].j;
          cval[base * CWt + c] = make_double2(lst[l].second[c].re, lst[l].second[c].im);
        }
      }
    }
  for (int k = 0; k < M; k++) mstruct[k] = nstruct == 1 ? 0 : k;
  // capacity: chains per launch, bounded by a tenth of the free memory for the two vector stores
  size_t fre = 0, tot = 0;
  PURE_CUDA(cudaMemGetInfo(&fre, &tot));
  const double per_chain = 2.0 * (double)(N + 1) * D * 16 + (double)N * std::max(K, 1) * 8;
  long cap = (long)(0.1 * (double)fre / per_chain);
  cap = std::max(1L, std::min(std::min(cap, 16384L), (long)M * d.R));
  ps.cap = (int)cap;
  int rc;
  if ((rc = pure_alloc(ps, &ps.ucol, ucol.size(), err)) || (rc = pure_alloc(ps, &ps.cj, cj.size(), err)) || (rc = pure_alloc(ps, &ps.cval, cval.size(), err)) ||
      (rc = pure_alloc(ps, &ps.mstruct, (size_t)M, err)) || (rc = pure_alloc(ps, &ps.member, (size_t)cap, err)) || (rc = pure_alloc(ps, &ps.pulse, (size_t)cap, err)) ||
      (rc = pure_alloc(ps, &ps.coo_off, (size_t)cap, err)) || (rc = pure_alloc(ps, &ps.psi0, (size_t)M * D, err)) || (rc = pure_alloc(ps, &ps.phi0, (size_t)M * D, err)) ||
      (rc = pure_alloc(ps, &ps.psi, (size_t)cap * (N + 1) * D, err)) || (rc = pure_alloc(ps, &ps.chi, (size_t)cap * (N + 1) * D, err)) ||
      (rc = pure_alloc(ps, &ps.g, (size_t)cap * N * std::max(K, 1), err)) || (rc = pure_alloc(ps, &ps.tau_fom, (size_t)cap * 4, err)) ||
      (rc = pure_alloc(ps, &ps.fomc, (size_t)cap, err))) return rc;
  ps.red_chunk = std::max(4, (M + 127) / 128);            // same two-pass member reduction as the small-D path (<= 128 partial rows)
  ps.red_nchunks = (M + ps.red_chunk - 1) / ps.red_chunk;
  if ((rc = pure_alloc(ps, &ps.part, (size_t)d.R * ps.red_nchunks * ((size_t)N * K + 1), err))) return rc;
  PURE_CUDA(cudaMemcpy(ps.ucol, ucol.data(), ucol.size() * sizeof(int), cudaMemcpyHostToDevice));
  PURE_CUDA(cudaMemcpy(ps.cj, cj.data(), cj.size() * sizeof(int), cudaMemcpyHostToDevice));
  PURE_CUDA(cudaMemcpy(ps.cval, cval.data(), cval.size() * sizeof(double2), cudaMemcpyHostToDevice));
  PURE_CUDA(cudaMemcpy(ps.mstruct, mstruct.data(), mstruct.size() * sizeof(int), cudaMemcpyHostToDevice));
  PURE_CUDA(cudaMemcpy(ps.psi0, psi0.data(), psi0.size() * sizeof(double), cudaMemcpyHostToDevice));
  PURE_CUDA(cudaMemcpy(ps.phi0, phi0.data(), phi0.size() * sizeof(double), cudaMemcpyHostToDevice));
  if ((size_t)2 * D * 16 > 48 * 1024) PURE_CUDA(cudaFuncSetAttribute((const void*)pure_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * D * 16));
  ps.active = true;
  return QOC_OK;
}

}  // namespace qoc
