// Small-dimension GRAPE kernels (per-chain dimension D <= 16): one warp owns one chain (or 8/DP packed
// chains for DP = 2, 4 as a block-diagonal 8x8 matrix) and keeps every matrix in registers; all products run
// on DMMA.  Replaces, for one (pulse, ensemble member) chain:
//   pw_prop_save!            /root/reference/src/timeevolution.jl:98-110   (assemble + expm_t8)
//   evolve_func! loops       /root/reference/src/GRAPE.jl:53-75, 216-251   (forward / backward sweeps)
//   grad_func! + fom_func    /root/reference/src/GRAPE.jl:79-94, 261-303; src/cost_functions.jl:99-111
//   exact gradient semantics /root/reference/src/solve.jl:268-290 + src/GRAPE.jl:14-18 (GRAD == 2)
#pragma once
#include "warp_mat.cuh"

namespace qoc {

// Degree-8 Taylor polynomial of exp in THREE matrix products (Bader, Blanes, Casas 2019):
//   A2 = A*A; A4 = A2*(x1 A + x2 A2); A8 = (x3 A2 + A4)*(x4 I + x5 A + x6 A2 + x7 A4);
//   T8 = I + A + y2 A2 + A8          (coefficients verified to reproduce 1/k!, k <= 8)
constexpr double T8_X1 = 0.10836465678522780852;
constexpr double T8_X2 = 0.027091164196306952131;
constexpr double T8_X3 = 0.66666666666666666667;
constexpr double T8_X4 = 0.54676145797072405251;
constexpr double T8_X5 = 0.16112557339541759283;
constexpr double T8_X6 = 0.014090917158378207731;
constexpr double T8_X7 = 0.033792797010870504141;
constexpr double T8_Y2 = 0.13549236135285063166;
// ||A||_1 <= theta  =>  ||A||^9/9! * e^||A|| <= 2^-53
constexpr double T8_THETA_DEFAULT = 0.0694;

enum { SYS_DENSITY = 0, SYS_UNITARY = 1 };
enum { GRAD_NONE = 0, GRAD_FIRST = 1, GRAD_EXACT = 2 };

struct SmallParams {
  int D, N, K, M, R;
  int pack_mode;       // 0: members packed in a warp (same pulse); 1: pulses packed (same member)
  int n_groups;        // warps of work
  int n_inner;         // pack_mode 0: ceil(M/CPW) member groups per pulse; pack_mode 1: M
  int nmat;            // packed system matrices per system group: 1 + K (+ K transposed controls if exact)
  int sys_in_smem;
  int have_P;          // propagators were precomputed into storeP by expm_slices_kernel
  int sign_static;     // first-order UnitaryGate: +1 grad_func! (in-place), -1 grad_func (static)
  int fom_exact;       // figure of merit of the exact (ADGRAPE / C1) functional even when no gradient is asked
  double dt, theta;
  const double2* sys;  // [n_sysgroups][nmat][NB*NB*2*32]
  const double2* xi;   // [n_sysgroups][NB*NB*2*32]
  const double2* xt;
  const double* x;     // [R][N][K]  (= the reference's K x N column-major control_array per pulse)
  double2* storeP;     // [n_groups][N][NB*NB*2*32]
  double2* storeS;     // [n_groups][N][NB*NB*2*32]
  double* fomc;        // [R][M]
  double* gradc;       // [R][M][N][K]
  double2* out_final;  // optional [R][M][D*D]: final forward state, column-major complex
};

template <int NB> __host__ __device__ constexpr int cm_elems() { return NB * NB * 2 * 32; }  // double2 per packed matrix

// 1-norm upper bound (column sums of |re|+|im|), warp-uniform max over all packed chains; float is enough
// for a scaling decision and is rounded up.
template <int NB> __device__ __forceinline__ float cm_norm1_bound(const CM<NB>& x) {
  float best = 0.f;
#pragma unroll
  for (int j = 0; j < NB; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NB; i++) s += (float)(fabs(x.re[i][j][e]) + fabs(x.im[i][j][e]));
      s += __shfl_xor_sync(FULL_MASK, s, 4);
      s += __shfl_xor_sync(FULL_MASK, s, 8);
      s += __shfl_xor_sync(FULL_MASK, s, 16);
      best = fmaxf(best, s);
    }
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 1));
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 2));
  return best * 1.000001f;
}
__device__ __forceinline__ int scaling_power(float nrm, float theta) {
  if (!(nrm > theta)) return 0;
  int s = ilogbf(nrm / theta) + 1;
  return s > 60 ? 60 : s;
}

// P = exp(G): scaling, T8 in 3 products, s squarings.
template <int NB> __device__ __forceinline__ CM<NB> expm_t8(const Lane& L, CM<NB> G, float theta) {
  int s = scaling_power(cm_norm1_bound<NB>(G), theta);
  if (s) G = cm_scale<NB>(G, scalbn(1.0, -s));
  FA<NB> Ga = to_A<NB>(L, G);
  FB<NB> Gb = to_B<NB>(L, G);
  CM<NB> G2 = mul<NB>(Ga, Gb);
  CM<NB> Y1 = cm_scale<NB>(G, T8_X1); cm_axpy<NB>(Y1, T8_X2, G2);
  CM<NB> G4 = mul<NB>(to_A<NB>(L, G2), to_B<NB>(L, Y1));
  CM<NB> L8 = G4; cm_axpy<NB>(L8, T8_X3, G2);
  CM<NB> R8 = cm_scale<NB>(G, T8_X5); cm_axpy<NB>(R8, T8_X6, G2); cm_axpy<NB>(R8, T8_X7, G4);
  cm_add_identity<NB>(L, R8, T8_X4);
  CM<NB> P = mul<NB>(to_A<NB>(L, L8), to_B<NB>(L, R8));
  cm_axpy<NB>(P, 1.0, G); cm_axpy<NB>(P, T8_Y2, G2); cm_add_identity<NB>(L, P, 1.0);
  for (int j = 0; j < s; j++) P = mul<NB>(to_A<NB>(L, P), to_B<NB>(L, P));
  return P;
}

// Frechet derivative L_exp(G, Y) of the same scheme (derivative of each of the three products and of every
// squaring).  Identity tr(W * L(G,E)) = tr(L(G,W) * E) lets ONE call per slice serve all K controls.
template <int NB> __device__ __forceinline__ CM<NB> frechet_t8(const Lane& L, CM<NB> G, CM<NB> Y, float theta) {
  int s = scaling_power(cm_norm1_bound<NB>(G), theta);
  if (s) { double sc = scalbn(1.0, -s); G = cm_scale<NB>(G, sc); Y = cm_scale<NB>(Y, sc); }
  FA<NB> Ga = to_A<NB>(L, G);
  FB<NB> Gb = to_B<NB>(L, G);
  CM<NB> G2 = mul<NB>(Ga, Gb);
  CM<NB> dG2 = mul<NB>(to_A<NB>(L, Y), Gb);
  mul_acc<NB>(Ga, to_B<NB>(L, Y), dG2);
  CM<NB> Y1 = cm_scale<NB>(G, T8_X1); cm_axpy<NB>(Y1, T8_X2, G2);
  FB<NB> Y1b = to_B<NB>(L, Y1);
  FA<NB> G2a = to_A<NB>(L, G2);
  CM<NB> G4 = mul<NB>(G2a, Y1b);
  CM<NB> dY1 = cm_scale<NB>(Y, T8_X1); cm_axpy<NB>(dY1, T8_X2, dG2);
  CM<NB> dG4 = mul<NB>(to_A<NB>(L, dG2), Y1b);
  mul_acc<NB>(G2a, to_B<NB>(L, dY1), dG4);
  CM<NB> L8 = G4; cm_axpy<NB>(L8, T8_X3, G2);
  CM<NB> R8 = cm_scale<NB>(G, T8_X5); cm_axpy<NB>(R8, T8_X6, G2); cm_axpy<NB>(R8, T8_X7, G4);
  cm_add_identity<NB>(L, R8, T8_X4);
  CM<NB> dL8 = dG4; cm_axpy<NB>(dL8, T8_X3, dG2);
  CM<NB> dR8 = cm_scale<NB>(Y, T8_X5); cm_axpy<NB>(dR8, T8_X6, dG2); cm_axpy<NB>(dR8, T8_X7, dG4);
  FA<NB> L8a = to_A<NB>(L, L8);
  FB<NB> R8b = to_B<NB>(L, R8);
  CM<NB> dP = mul<NB>(to_A<NB>(L, dL8), R8b);
  mul_acc<NB>(L8a, to_B<NB>(L, dR8), dP);
  cm_axpy<NB>(dP, 1.0, Y); cm_axpy<NB>(dP, T8_Y2, dG2);
  if (s) {
    CM<NB> P = mul<NB>(L8a, R8b);
    cm_axpy<NB>(P, 1.0, G); cm_axpy<NB>(P, T8_Y2, G2); cm_add_identity<NB>(L, P, 1.0);
    for (int j = 0; j < s; j++) {
      FA<NB> Pa = to_A<NB>(L, P);
      FB<NB> Pb = to_B<NB>(L, P);
      CM<NB> nd = mul<NB>(to_A<NB>(L, dP), Pb);
      mul_acc<NB>(Pa, to_B<NB>(L, dP), nd);
      dP = nd;
      if (j + 1 < s) P = mul<NB>(Pa, Pb);
    }
  }
  return dP;
}

// G_t = -i*dt*(A + sum_j x[j,t] B_j); packed system matrices: index 0 = A, 1..K = B_j.
template <int NB>
__device__ __forceinline__ CM<NB> assemble_generator(const Lane& L, const double2* sysw, const double* xt, int K, double dt) {
  CM<NB> H = cm_load<NB>(L, sysw);
  for (int j = 0; j < K; j++) {
    double xj = __ldg(xt + j);
    CM<NB> B = cm_load<NB>(L, sysw + (size_t)(j + 1) * cm_elems<NB>());
    cm_axpy<NB>(H, xj, B);
  }
  CM<NB> G;
  QOC_FOR_CM(NB) { G.re[i][j][e] = dt * H.im[i][j][e]; G.im[i][j][e] = -dt * H.re[i][j][e]; }
  return G;
}

// Which (pulse, member) does this lane's row block belong to.
template <int CPW> struct Slot {
  int r, k, sysgroup; bool valid;
  __device__ __forceinline__ Slot(const SmallParams& p, const Lane& L, int w) {
    int s = (CPW == 1) ? 0 : (L.g / (8 / CPW));
    int outer = w / p.n_inner, inner = w - outer * p.n_inner;
    if (p.pack_mode == 0) { r = outer; k = inner * CPW + s; sysgroup = inner; valid = k < p.M; if (!valid) k = p.M - 1; }
    else { k = inner; r = outer * CPW + s; sysgroup = inner; valid = r < p.R; if (!valid) r = p.R - 1; }
  }
};

// write the 8-chunked, group-reduced gradient values of slice t
template <int NB, int CPW>
__device__ __forceinline__ void emit_gradient(const SmallParams& p, const Lane& L, const Slot<CPW>& sl,
                                              const double2* mats, const CM<NB>& W, int t) {
  constexpr int GS = 32 / CPW;
  double* out = p.gradc + (((size_t)sl.r * p.M + sl.k) * p.N + t) * p.K;
  for (int c0 = 0; c0 < p.K; c0 += 8) {
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
      v[c] = (c0 + c < p.K) ? cm_redot_partial<NB>(L, mats + (size_t)(c0 + c) * cm_elems<NB>(), W) : 0.0;
    group_sum8<GS>(L.lane, v);
    int within = L.lane % GS;
    int idx = (within * 8) / GS;
    if ((within % (GS / 8)) == 0 && c0 + idx < p.K && sl.valid) out[c0 + idx] = v[0];
  }
}

// One warp = one packed group of chains; forward sweep (+ expm), figure of merit, backward sweep + gradient.
template <int NB, int CPW, int SYS, int GRAD>
__global__ void __launch_bounds__(128) chain_kernel(const SmallParams p) {
  extern __shared__ double2 smem[];
  const int warp_in_cta = threadIdx.x >> 5;
  const int w = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  if (w >= p.n_groups) return;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  const float theta = (float)p.theta;

  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  if (p.sys_in_smem) {
    double2* mine = smem + (size_t)warp_in_cta * p.nmat * E;
    for (int i = L.lane; i < p.nmat * E; i += 32) mine[i] = sysw[i];
    __syncwarp();
    sysw = mine;
  }
  const double* xr = p.x + (size_t)sl.r * p.N * p.K;
  double2* stP = p.storeP + (size_t)w * p.N * E;
  double2* stS = p.storeS + (size_t)w * p.N * E;
  const double invD2 = 1.0 / ((double)p.D * (double)p.D);

  // ---------------- forward sweep: S_{t+1} = P_t S_t  or  P_t S_t P_t' ----------------
  CM<NB> S = cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E);
  for (int t = 0; t < p.N; t++) {
    CM<NB> P;
    if (p.have_P) P = cm_load<NB>(L, stP + (size_t)t * E);
    else {
      P = expm_t8<NB>(L, assemble_generator<NB>(L, sysw, xr + (size_t)t * p.K, p.K, p.dt), theta);
      if (GRAD != GRAD_NONE) cm_store<NB>(L, stP + (size_t)t * E, P);
    }
    if (GRAD != GRAD_NONE) cm_store<NB>(L, stS + (size_t)t * E, S);
    FA<NB> Pa = to_A<NB>(L, P);
    if (SYS == SYS_UNITARY) {
      S = mul<NB>(Pa, to_B<NB>(L, S));
    } else {
      CM<NB> tmp = mul<NB>(to_A<NB>(L, S), adjB<NB>(Pa));      // S_t P_t'      (GRAPE.jl:245)
      S = mul<NB>(Pa, to_B<NB>(L, tmp));                          // P_t (S_t P_t') (GRAPE.jl:246)
    }
  }
  CM<NB> Xt = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);

  if (p.out_final) {   // final forward state in the caller's column-major layout (pw_evolve with U0 = Xi)
    constexpr int DPc = 8 * NB / CPW;
    int s = (CPW == 1) ? 0 : (L.g / DPc);
    double2* o = p.out_final + ((size_t)sl.r * p.M + sl.k) * p.D * p.D;
    QOC_FOR_CM(NB) {
      int row = 8 * i + L.g - s * DPc, col = 8 * j + 2 * L.q + e - s * DPc;
      if (sl.valid && row >= 0 && col >= 0 && row < p.D && col < p.D && col < DPc)
        o[(size_t)col * p.D + row] = make_double2(S.re[i][j][e], S.im[i][j][e]);
    }
  }

  // ---------------- figure of merit ----------------
  double tr_, ti_;
  const bool ref_unitary_fom = SYS == SYS_UNITARY && GRAD != GRAD_EXACT && !p.fom_exact;
  if (ref_unitary_fom) cm_dotc_partial<NB>(S, Xt, tr_, ti_);                            // tr(S' Xt)
  else cm_dotc_partial<NB>(Xt, S, tr_, ti_);                                            // tr(Xt' S)
  tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
  double fom;
  if (ref_unitary_fom) fom = tr_ * tr_ - ti_ * ti_;                                     // Re(tau*tau)
  else fom = 1.0 - (tr_ * tr_ + ti_ * ti_) * invD2;                                     // C1
  if (sl.valid && (L.lane % GS) == 0) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
  if (GRAD == GRAD_NONE) return;

  // ---------------- backward sweep + gradient ----------------
  const double2* Bmats = sysw + E;                        // B_1..B_K
  const double2* BTmats = sysw + (size_t)(1 + p.K) * E;   // transposed controls (exact mode only)
  CM<NB> C = Xt;
  if (GRAD == GRAD_FIRST) {
    double fr, fi;
    if (SYS == SYS_UNITARY) { double sg = 2.0 * p.sign_static * p.dt; fr = -sg * ti_; fi = sg * tr_; }   // 2(+-i dt) tau
    else { fr = 0.0; fi = p.dt; }                                                                       // i dt
    FB<NB> Cb = to_B<NB>(L, C);
    FA<NB> Ca = to_A<NB>(L, C);
    for (int t = p.N - 1; t >= 0; t--) {
      CM<NB> P = cm_load<NB>(L, stP + (size_t)t * E);
      CM<NB> St = cm_load<NB>(L, stS + (size_t)t * E);
      FB<NB> Pb = to_B<NB>(L, P);
      if (SYS == SYS_UNITARY) {
        C = mul<NB>(adjA<NB>(Pb), Cb);                          // C_t = P_t' C_{t+1}      (GRAPE.jl:228)
      } else {
        CM<NB> tmp = mul<NB>(Ca, Pb);                           // C_{t+1} P_t             (GRAPE.jl:248)
        C = mul<NB>(adjA<NB>(Pb), to_B<NB>(L, tmp));            // P_t' (C_{t+1} P_t)      (GRAPE.jl:249)
      }
      Ca = to_A<NB>(L, C);
      Cb = to_B<NB>(L, C);
      FA<NB> Sa = to_A<NB>(L, St);
      // W^T with W = S_t C_t' (unitary) or S_t C_t' - C_t' S_t (density): tr(B W) = sum B .* W^T
      CM<NB> WT = mul<NB>(conjF<NB>(Ca), trB<NB>(Sa));          // (S C')^T = conj(C) S^T
      if (SYS == SYS_DENSITY) {
        FB<NB> Sb = to_B<NB>(L, St);
        CM<NB> W2 = mul<NB>(trA<NB>(Sb), conjF<NB>(Cb));        // (C' S)^T = S^T conj(C)
        cm_sub<NB>(WT, W2);
      }
      CM<NB> Ws = cm_cscale<NB>(WT, fr, fi);
      emit_gradient<NB, CPW>(p, L, sl, Bmats, Ws, t);
    }
  } else {   // GRAD_EXACT: F = 1 - |tau|^2/D^2, tau = tr(Xt' S_N); dF/dx[c,t] = -(2/D^2) Re(conj(tau) dtau)
    const double kr = 2.0 * p.dt * invD2;
    for (int t = p.N - 1; t >= 0; t--) {
      CM<NB> P = cm_load<NB>(L, stP + (size_t)t * E);
      CM<NB> St = cm_load<NB>(L, stS + (size_t)t * E);
      FB<NB> Pb = to_B<NB>(L, P);
      FA<NB> Sa = to_A<NB>(L, St);
      FA<NB> Ca = to_A<NB>(L, C);
      FB<NB> Cb = to_B<NB>(L, C);
      CM<NB> Y, Cn;
      double cr, ci;
      if (SYS == SYS_UNITARY) {
        Y = mul<NB>(Sa, adjB<NB>(Ca));                          // Y = S_t C_{t+1}'
        Cn = mul<NB>(adjA<NB>(Pb), Cb);                         // C_t = P_t' C_{t+1}
        cr = kr * ti_; ci = kr * tr_;                           // -(2/D^2) conj(tau) (-i dt)
      } else {
        FA<NB> Pda = adjA<NB>(Pb);
        CM<NB> Q1 = mul<NB>(Pda, adjB<NB>(Ca));                 // P_t' C_{t+1}'
        CM<NB> Q2 = mul<NB>(Pda, Cb);                           // P_t' C_{t+1}
        FB<NB> Sb = to_B<NB>(L, St);
        CM<NB> Y1 = mul<NB>(Sa, to_B<NB>(L, Q1));               // S_t P_t' C_{t+1}'
        CM<NB> Y2 = mul<NB>(adjA<NB>(Sb), to_B<NB>(L, Q2));     // S_t' P_t' C_{t+1}
        Y = cm_cscale<NB>(Y1, tr_, -ti_);                       // conj(tau) Y1 + tau Y2
        cm_caxpy<NB>(Y, tr_, ti_, Y2);
        Cn = mul<NB>(to_A<NB>(L, Q2), Pb);                      // C_t = P_t' C_{t+1} P_t
        cr = 0.0; ci = kr;                                      // -(2/D^2)(-i dt)
      }
      CM<NB> G = assemble_generator<NB>(L, sysw, xr + (size_t)t * p.K, p.K, p.dt);
      CM<NB> Lam = frechet_t8<NB>(L, G, Y, theta);
      CM<NB> Ws = cm_cscale<NB>(Lam, cr, ci);
      emit_gradient<NB, CPW>(p, L, sl, BTmats, Ws, t);          // tr(Lam B_c) = sum Lam .* B_c^T
      C = Cn;
    }
  }
}

// Slice-parallel propagators: one warp per (system group / pulse, slice).  Writes the packed layout used by
// chain_kernel (storeP) and/or the caller's column-major layout (pw_prop_save!, timeevolution.jl:98-110).
struct SliceParams {
  int D, N, K, M, R, pack_mode, n_groups, n_inner, nmat;
  double dt, theta;
  const double2* sys;
  const double* x;
  double2* storeP;     // optional packed [n_groups][N][E]
  double2* out_user;   // optional [R][M][N][D*D] column-major complex
  int mode;            // 0: propagator exp(-i dt H); 1: Hamiltonian H (pw_ham_save!); 2: generator -i dt H (pw_gen_save!)
};
template <int NB, int CPW>
__global__ void __launch_bounds__(128) expm_slices_kernel(const SliceParams p) {
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long)p.n_groups * p.N) return;
  const int w = (int)(gw / p.N), t = (int)(gw - (long)w * p.N);
  const Lane L(threadIdx.x & 31);
  SmallParams sp; sp.M = p.M; sp.R = p.R; sp.pack_mode = p.pack_mode; sp.n_inner = p.n_inner;
  const Slot<CPW> sl(sp, L, w);
  constexpr int E = cm_elems<NB>();
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  const double* xt = p.x + ((size_t)sl.r * p.N + t) * p.K;
  CM<NB> out;
  if (p.mode == 0) out = expm_t8<NB>(L, assemble_generator<NB>(L, sysw, xt, p.K, p.dt), (float)p.theta);
  else if (p.mode == 2) out = assemble_generator<NB>(L, sysw, xt, p.K, p.dt);
  else {  // H = i/dt * G  (undo the -i dt factor exactly: re = -G.im/dt ... computed directly instead)
    out = cm_load<NB>(L, sysw);
    for (int j = 0; j < p.K; j++) cm_axpy<NB>(out, __ldg(xt + j), cm_load<NB>(L, sysw + (size_t)(j + 1) * E));
  }
  if (p.storeP) cm_store<NB>(L, p.storeP + ((size_t)w * p.N + t) * E, out);
  if (p.out_user) {
    constexpr int DPc = 8 * NB / CPW;
    int s = (CPW == 1) ? 0 : (L.g / DPc);
    double2* o = p.out_user + (((size_t)sl.r * p.M + sl.k) * p.N + t) * p.D * p.D;
    QOC_FOR_CM(NB) {
      int row = 8 * i + L.g - s * DPc, col = 8 * j + 2 * L.q + e - s * DPc;
      if (sl.valid && row >= 0 && col >= 0 && row < p.D && col < p.D && col < DPc)
        o[(size_t)col * p.D + row] = make_double2(out.re[i][j][e], out.im[i][j][e]);
    }
  }
}

// Pack caller matrices (column-major complex D x D) into the warp layout.  One thread per packed double2.
//   dst[(og*nmat_dst + mat_dst)*E + ((i*NB+j)*2+ri)*32 + lane] ; slot s of group og reads source matrix
//   src + src_index(og, s)*src_stride, transposed if asked; everything outside the chain's D x D block is 0.
struct PackParams {
  int D, NB, CPW, n_og, nmat_dst, mat_dst, transpose;
  int pack_mode;      // 0: slot s -> member og*CPW+s (clamped to n_src-1); 1: every slot -> member og
  int n_src;          // number of distinct source matrices (1 if shared)
  long src_stride;    // in double2 between consecutive source members (0 if shared)
  const double2* src;
  double2* dst;
};
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
      val = ri ? z.y : z.x;
    }
    v[e] = val;
  }
  p.dst[((size_t)og * p.nmat_dst + p.mat_dst) * E + rem] = make_double2(v[0], v[1]);
}

// Deterministic weighted ensemble reduction  F[r] = sum_k w_k fom[r,k],  G[r,:] = sum_k w_k grad[r,k,:]
// (/root/reference/src/solve.jl:171-191), two fixed-order passes.
__global__ void reduce_members_pass1(const double* __restrict__ gradc, const double* __restrict__ fomc,
                                     const double* __restrict__ wts, double* __restrict__ part,
                                     int M, int NK, int chunk, int nchunks) {
  // grid: (ceil((NK+1)/256), nchunks, R); part[r][chunk][NK+1] (entry 0 = fom)
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > NK) return;
  int ch = blockIdx.y, r = blockIdx.z;
  int k0 = ch * chunk, k1 = min(M, k0 + chunk);
  double s = 0.0;
  if (e == 0) { for (int k = k0; k < k1; k++) s += wts[k] * fomc[(size_t)r * M + k]; }
  else if (gradc) { for (int k = k0; k < k1; k++) s += wts[k] * gradc[((size_t)r * M + k) * NK + (e - 1)]; }
  part[((size_t)r * nchunks + ch) * (NK + 1) + e] = s;
}
__global__ void reduce_members_pass2(const double* __restrict__ part, double* __restrict__ out, int NK, int nchunks) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > NK) return;
  int r = blockIdx.y;
  double s = 0.0;
  for (int ch = 0; ch < nchunks; ch++) s += part[((size_t)r * nchunks + ch) * (NK + 1) + e];
  out[(size_t)r * (NK + 1) + e] = s;
}

}  // namespace qoc
