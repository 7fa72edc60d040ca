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




// Norm upper bound for the scaling decision (sums of |re|+|im|: rows = infinity norm, QOC_NORM_INF=0: columns = 1-norm),
// warp-uniform max over all packed chains; float is enough for a scaling decision and is rounded up.
// float upper bound of |x| from the high word of the double, on the integer pipe (the FP64 pipe is the
// bottleneck resource): exponent re-biased by 1023-127 = 896, 20 mantissa bits kept, rounded up.
__device__ __forceinline__ float abs_upper_f32(double x) {
  int hi = __double2hiint(x) & 0x7fffffff;
  int e = hi - 0x38000000;                       // (896 << 20)
  e = e < 0 ? 0 : (e > 0x0fdfffff ? 0x0fdfffff : e);
  return __int_as_float((e << 3) + 8);
}
#ifndef QOC_NORM_INF
#define QOC_NORM_INF 1
#endif
template <int NB> __device__ __forceinline__ float cm_norm1_bound(const CM<NB>& x) {
#if QOC_NORM_INF
  // infinity norm (row sums): any induced norm serves the Taylor bound, and in the register layout a row lives in the four
  // lanes of a quad, so a row sum needs 2 shuffles and the maximum over the rows 3 (column sums: 3 per column pair + 2)
  float best = 0.f;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) s += abs_upper_f32(x.re[i][j][e]) + abs_upper_f32(x.im[i][j][e]);
    s += __shfl_xor_sync(FULL_MASK, s, 1);
    s += __shfl_xor_sync(FULL_MASK, s, 2);
    best = fmaxf(best, s);
  }
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 4));
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 8));
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 16));
  return best * 1.000001f;
#else
  float best = 0.f;
#pragma unroll
  for (int j = 0; j < NB; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NB; i++) s += abs_upper_f32(x.re[i][j][e]) + abs_upper_f32(x.im[i][j][e]);
      s += __shfl_xor_sync(FULL_MASK, s, 4);
      s += __shfl_xor_sync(FULL_MASK, s, 8);
      s += __shfl_xor_sync(FULL_MASK, s, 16);
      best = fmaxf(best, s);
    }
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 1));
  best = fmaxf(best, __shfl_xor_sync(FULL_MASK, best, 2));
  return best * 1.000001f;
#endif
}
__device__ __forceinline__ int scaling_power(float nrm, float theta) {
  if (!(nrm > theta)) return 0;
  int s = ilogbf(nrm / theta) + 1;
  return s > 60 ? 60 : s;
}

// The exponential itself uses the same scheme with the scalar factors moved so that the elementwise FP64 work per matrix
// shrinks from 30 to 23 instructions (they share the pipe with the DMMAs):
//   Yh = A + (x2/x1) A2 + (x3/x1) I;   Lh = A2 Yh = (A4 + x3 A2) / x1;   A4 = x1 Lh - x3 A2
//   Rh = x1 (x4 I + x5 A + x6 A2 + x7 A4) = r4 I + r5 A + r6 A2 + r7 Lh;   T8 = I + A + y2 A2 + Lh Rh
// (constants from the x_i above in 40-digit arithmetic; x2/x1 = 1/4, r5 = 11/630, r7 = 1/2520)
constexpr double T8_YA = 0.25;
constexpr double T8_YB = 6.1520673478250353627;
constexpr double T8_R4 = 0.059249617736388271447;
constexpr double T8_R5 = 0.017460317460317460317;
constexpr double T8_R6 = -0.00091433916494050425595;
constexpr double T8_R7 = 0.00039682539682539682538;

// P = exp(G): scaling, T8 in 3 NT products, s squarings.  `herm`: G is anti-Hermitian (Hermitian drift and
// controls, checked on the host), so G^T = -conj(G) and (G^2)^T = conj(G^2) cost no data movement.
template <int NB> __device__ __forceinline__ CM<NB> expm_t8(const Lane& L, CM<NB> G, float theta, bool herm, double* tb) {
  const int s = scaling_power(cm_norm1_bound<NB>(G), theta);
  if (s) G = cm_scale<NB>(G, scalbn(1.0, -s));
  const CM<NB> Gt = herm ? cm_negconj<NB>(G) : transpose<NB>(L, G, tb);
  const CM<NB> G2 = herm ? square_antiherm<NB>(G) : mul_nt<NB>(G, Gt);
  const CM<NB> G2t = herm ? cm_conj<NB>(G2) : transpose<NB>(L, G2, tb);
  CM<NB> Yht = Gt; cm_axpy<NB>(Yht, T8_YA, G2t); cm_add_identity<NB>(L, Yht, T8_YB);
  const CM<NB> Lh = mul_nt<NB>(G2, Yht);
  const CM<NB> Lht = transpose<NB>(L, Lh, tb);
  CM<NB> Rht = cm_scale<NB>(Gt, T8_R5); cm_axpy<NB>(Rht, T8_R6, G2t); cm_axpy<NB>(Rht, T8_R7, Lht);
  cm_add_identity<NB>(L, Rht, T8_R4);
  CM<NB> P = G; cm_axpy<NB>(P, T8_Y2, G2); cm_add_identity<NB>(L, P, 1.0);
  mul_nt_acc<NB>(Lh, Rht, P);
  for (int j = 0; j < s; j++) { const CM<NB> Pt = transpose<NB>(L, P, tb); P = mul_nt<NB>(P, Pt); }
  return P;
}

// Frechet derivative L_exp(G, Y) of the same scheme (derivative of each of the three products and of every
// squaring).  Identity tr(W * L(G,E)) = tr(L(G,W) * E) lets ONE call per slice serve all K controls.
// Ordered for short live ranges: the right factor R8 = x4 I + x5 G + x6 G2 + x7 G4 and its derivative are accumulated
// term by term as soon as each (transposed) power exists, so G^T, Y^T, (G2)^T, (dG2)^T die after the second product instead
// of surviving until the third (D = 16, 32 registers per matrix: local-memory stack 1008 -> 840 B in chain_kernel, 448 ->
// 232 B in grad_slices_kernel).  Parking further matrices in shared memory explicitly brought that to 312 / 104 B but ran
// 3 % slower (spills already hit L1 at shared-memory speed, and the operand prefetch had to go): not kept, profiles/README.md.
template <int NB> __device__ __forceinline__ CM<NB> frechet_t8(const Lane& L, CM<NB> G, CM<NB> Y, float theta, bool herm, double* tb) {
  const int s = scaling_power(cm_norm1_bound<NB>(G), theta);
  if (s) { double sc = scalbn(1.0, -s); G = cm_scale<NB>(G, sc); Y = cm_scale<NB>(Y, sc); }
  CM<NB> R8t, dR8t, Y1t, dY1t, G2, dG2;
  {
    const CM<NB> Gt = herm ? cm_negconj<NB>(G) : transpose<NB>(L, G, tb);
    const CM<NB> Yt = transpose<NB>(L, Y, tb);
    G2 = mul_nt<NB>(G, Gt);
    dG2 = mul_nt<NB>(Y, Gt);
    mul_nt_acc<NB>(G, Yt, dG2);                                   // Y G + G Y
    R8t = cm_scale<NB>(Gt, T8_X5); cm_add_identity<NB>(L, R8t, T8_X4);
    dR8t = cm_scale<NB>(Yt, T8_X5);
    Y1t = cm_scale<NB>(Gt, T8_X1);
    dY1t = cm_scale<NB>(Yt, T8_X1);
  }
  {
    const CM<NB> G2t = herm ? cm_conj<NB>(G2) : transpose<NB>(L, G2, tb);
    cm_axpy<NB>(Y1t, T8_X2, G2t); cm_axpy<NB>(R8t, T8_X6, G2t);
    const CM<NB> dG2t = transpose<NB>(L, dG2, tb);
    cm_axpy<NB>(dY1t, T8_X2, dG2t); cm_axpy<NB>(dR8t, T8_X6, dG2t);
  }
  CM<NB> L8 = mul_nt<NB>(G2, Y1t);                                // G4
  CM<NB> dL8 = mul_nt<NB>(dG2, Y1t);
  mul_nt_acc<NB>(G2, dY1t, dL8);                                  // dG4 = dG2 Y1 + G2 dY1
  {
    const CM<NB> G4t = transpose<NB>(L, L8, tb);
    cm_axpy<NB>(R8t, T8_X7, G4t);
    const CM<NB> dG4t = transpose<NB>(L, dL8, tb);
    cm_axpy<NB>(dR8t, T8_X7, dG4t);
  }
  cm_axpy<NB>(L8, T8_X3, G2);                                     // L8 = G4 + x3 G2
  cm_axpy<NB>(dL8, T8_X3, dG2);
  CM<NB> dP = Y; cm_axpy<NB>(dP, T8_Y2, dG2);
  mul_nt_acc<NB>(dL8, R8t, dP);
  mul_nt_acc<NB>(L8, dR8t, dP);                                   // dL8 R8 + L8 dR8 + Y + y2 dG2
  if (s) {
    CM<NB> P = G; cm_axpy<NB>(P, T8_Y2, G2); cm_add_identity<NB>(L, P, 1.0);
    mul_nt_acc<NB>(L8, R8t, P);
    for (int j = 0; j < s; j++) {
      const CM<NB> Pt = transpose<NB>(L, P, tb);
      const CM<NB> dPt = transpose<NB>(L, dP, tb);
      CM<NB> nd = mul_nt<NB>(dP, Pt);
      mul_nt_acc<NB>(P, dPt, nd);                                 // dP P + P dP
      dP = nd;
      if (j + 1 < s) P = mul_nt<NB>(P, Pt);
    }
  }
  return dP;
}

// G_t = A~ + sum_j x[j,t] B~_j with the packed system matrices pre-multiplied by -i*dt on upload
// (index 0 = A~, 1..K = B~_j), so the generator needs no further scaling.
constexpr int XPF = 8;   // controls whose amplitudes are prefetched into registers one slice ahead
template <int NB, bool SH>
__device__ __forceinline__ CM<NB> assemble_generator(const Lane& L, const double2* sysw, const double* xt, int K,
                                                     const double* xpre = nullptr) {
  CM<NB> G = cm_load_sys<NB, SH>(L, sysw);
#pragma unroll
  for (int j = 0; j < XPF; j++)
    if (j < K) cm_axpy_packed<NB, SH>(L, G, xpre ? xpre[j] : __ldg(xt + j), sysw + (size_t)(j + 1) * cm_elems<NB>());
  for (int j = XPF; j < K; j++) cm_axpy_packed<NB, SH>(L, G, __ldg(xt + j), sysw + (size_t)(j + 1) * cm_elems<NB>());
  return G;
}

// Which (pulse, member) does this lane's row block belong to.
template <int CPW> struct Slot {
  int r, k, sysgroup; bool valid;
  __device__ __forceinline__ Slot(int pack_mode, int n_inner, int M, int R, const Lane& L, int w) {
    int s = (CPW == 1) ? 0 : (L.g / (8 / CPW));
    int outer = w / n_inner, inner = w - outer * n_inner;
    if (pack_mode == 0) { r = outer; k = inner * CPW + s; sysgroup = inner; valid = k < M; if (!valid) k = M - 1; }
    else { k = inner; r = outer * CPW + s; sysgroup = inner; valid = r < R; if (!valid) r = R - 1; }
  }
};

// write the 8-chunked, group-reduced gradient values of slice t: out[c] = Re sum mats_c .* W
template <int NB, int CPW, bool SH, bool CONJM = false>
__device__ __forceinline__ void emit_gradient(const SmallParams& p, const Lane& L, const Slot<CPW>& sl,
                                              const double2* mats, const CM<NB>& W, int t) {
  constexpr int GS = 32 / CPW;
  double* out = p.gradc + (((size_t)sl.r * p.M + sl.k) * p.N + t) * p.K;
  for (int c0 = 0; c0 < p.K; c0 += 8) {
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; c++)
      v[c] = (c0 + c < p.K) ? (CONJM ? cm_redotc_partial<NB, SH>(L, mats + (size_t)(c0 + c) * cm_elems<NB>(), W)
                                     : cm_redot_partial<NB, SH>(L, mats + (size_t)(c0 + c) * cm_elems<NB>(), W)) : 0.0;
    group_sum8<GS>(L.lane, v);
    int within = L.lane % GS;
    int idx = (within * 8) / GS;
    if ((within % (GS / 8)) == 0 && c0 + idx < p.K && sl.valid) out[c0 + idx] = v[0];
  }
}

// element (row, col) of the chain this lane's block belongs to, or row = -1 if the element is padding
template <int NB, int CPW> __device__ __forceinline__ void chain_coords(const Lane& L, int i, int j, int e, int D, int& row, int& col) {
  constexpr int DPc = 8 * NB / CPW;
  const int s = (CPW == 1) ? 0 : (L.g / DPc);
  row = 8 * i + L.g - s * DPc; col = 8 * j + 2 * L.q + e - s * DPc;
  if (row < 0 || col < 0 || row >= D || col >= D || col >= DPc) row = -1;
}

// One warp = one packed group of chains; forward sweep (+ expm), figure of merit, backward sweep + gradient.
// Stored per slice: the TRANSPOSED propagator Pt = P_t^T (what the backward recursions consume) and the state
// in the polarity its chain runs in (unitary: S_t^T, density: S_t).
//   unitary forward   S_{t+1}^T = S_t^T P_t^T                = nt(St, P)
//   density forward   S_{t+1}   = P (S P')                    = nt(P, X),  X = conj(P) S^T = nt(conj P, S)
//   unitary backward  C_t^T     = C_{t+1}^T conj(P)           = nt(Ct, conj Pt)
//   density backward  C_t       = P' (C P)                    = nt(conj Pt, Z),  Z = P^T C^T = nt(Pt, C)
// The scalar factor of the gradient formula is folded into the initial costate (everything is linear in C),
// next-slice operands (x, Pt, St) are prefetched one iteration ahead so HBM latency overlaps the DMMA work.
template <int NB, int CPW, int SYS, int GRAD, bool SH>
__device__ __forceinline__ void chain_body(const SmallParams& p, double2* smem) {
  const int warp_in_cta = threadIdx.x >> 5;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  const int Cn = p.Cn > 1 ? p.Cn : 1;
  if (gw >= p.n_groups * Cn) return;
  const int w = gw / Cn, ck = gw - w * Cn;
  const bool chunked = Cn > 1;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  constexpr int TBW = NB * NB * 2 * TB_PLANE;                     // doubles of transpose tile per warp
  const float theta = (float)p.theta;
  const bool herm = p.herm;
  const int K = p.K, N = p.N;
  const int t0 = (int)((long)ck * N / Cn), t1 = (int)((long)(ck + 1) * N / Cn);

  double* tb = reinterpret_cast<double*>(smem) + (size_t)warp_in_cta * TBW;
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  if (SH) {
    double2* mine = smem + (size_t)(blockDim.x >> 5) * TBW / 2 + (size_t)warp_in_cta * p.nmat * E;
    for (int i = L.lane; i < p.nmat * E; i += 32) mine[i] = sysw[i];
    __syncwarp();
    sysw = mine;
  }
  const double* xr = p.x + (size_t)sl.r * N * K;
  double2* stP = p.storeP + (size_t)w * N * E;
  double2* stS = p.storeS + (size_t)w * N * E;
  const double invD2 = 1.0 / ((double)p.D * (double)p.D);
  double xpre[XPF];
  auto prefetch_x = [&](int t) {
#pragma unroll
    for (int j = 0; j < XPF; j++) xpre[j] = (j < K) ? __ldg(xr + (size_t)t * K + j) : 0.0;
  };

  // ---------------- forward sweep ----------------
  // unitary: S holds S_t^T (xi packed transposed); density: S holds S_t
  CM<NB> S = chunked ? cm_load<NB>(L, p.bS + ((size_t)w * (Cn + 1) + ck) * E) : cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E);
  CM<NB> Pnext;
  if (p.have_P) Pnext = cm_load<NB>(L, stP + (size_t)t0 * E); else prefetch_x(t0);
  for (int t = t0; t < t1; t++) {
    CM<NB> P;
    if (p.have_P) {
      const CM<NB> Ptl = Pnext;
      if (t + 1 < t1) Pnext = cm_load<NB>(L, stP + (size_t)(t + 1) * E);
      P = transpose<NB>(L, Ptl, tb);
    } else {
      const CM<NB> G = assemble_generator<NB, SH>(L, sysw, xr + (size_t)t * K, K, xpre);
      if (t + 1 < t1) prefetch_x(t + 1);
      P = expm_t8<NB>(L, G, theta, herm, tb);
      if (GRAD != GRAD_NONE) cm_store<NB>(L, stP + (size_t)t * E, transpose<NB>(L, P, tb));
    }
    if (GRAD != GRAD_NONE) cm_store<NB>(L, stS + (size_t)t * E, S);
    if (SYS == SYS_UNITARY) {
      S = mul_nt<NB>(S, P);                                       // S^T P^T          (GRAPE.jl:226)
    } else {
      S = conj_by<NB>(P, S);                                      // P (S P'): X = conj(P) S^T = (S P')^T, then nt(P, X)   (GRAPE.jl:245-246)
    }
  }
  // xt packed transposed for unitary, so elementwise overlaps with S are consistent in both cases
  const CM<NB> Xt = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);

  if (p.out_final && !chunked) {   // final forward state in the caller's column-major layout (pw_evolve with U0 = Xi)
    double2* o = p.out_final + ((size_t)sl.r * p.M + sl.k) * p.D * p.D;
    QOC_FOR_CM(NB) {
      int row, col; chain_coords<NB, CPW>(L, i, j, e, p.D, row, col);
      if (sl.valid && row >= 0) {
        if (SYS == SYS_UNITARY) o[(size_t)row * p.D + col] = make_double2(S.re[i][j][e], S.im[i][j][e]);   // S holds S^T
        else o[(size_t)col * p.D + row] = make_double2(S.re[i][j][e], S.im[i][j][e]);
      }
    }
  }

  // ---------------- figure of merit ----------------
  double tr_, ti_;
  const bool ref_unitary_fom = SYS == SYS_UNITARY && GRAD != GRAD_EXACT && !p.fom_exact;
  if (ref_unitary_fom) cm_dotc_partial<NB>(S, Xt, tr_, ti_);                            // tr(S' Xt)
  else cm_dotc_partial<NB>(Xt, S, tr_, ti_);                                            // tr(Xt' S)
  tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
  if (chunked) {      // overlaps and figures of merit were produced by boundary2_kernel from the chunk totals
    const int slot = L.lane / GS;
    tr_ = p.tau_in[((size_t)w * CPW + slot) * 2 + 0]; ti_ = p.tau_in[((size_t)w * CPW + slot) * 2 + 1];
  } else {
    double fom;
    if (ref_unitary_fom) fom = tr_ * tr_ - ti_ * ti_;                                   // Re(tau*tau)
    else fom = 1.0 - (tr_ * tr_ + ti_ * ti_) * invD2;                                   // C1
    if (sl.valid && (L.lane % GS) == 0) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
  }
  if (GRAD == GRAD_NONE) return;

  // ---------------- backward sweep + gradient ----------------
  // Trace-dots run against the packed B~_c = -i dt B_c, i.e. B_c = (i/dt) B~_c.  With that factor folded in:
  //   first order, unitary:  f = 2(+-i dt) tau (i/dt) = -+2 tau          density:  f = (i dt)(i/dt) = -1
  //   exact, unitary:        f = -(2/D^2) conj(tau) (-i dt)(i/dt) = -(2/D^2) conj(tau)   density: f = -2/D^2
  // and f W(C) = W(conj(f) C): the costate chain starts from conj(f) Xt.
  const double2* Bmats = sysw + E;                        // B~_1..B~_K
  const double2* BTmats = sysw + (size_t)(1 + K) * E;     // transposed controls (exact mode only)
  CM<NB> C;                                               // unitary: C^T, density: C
  const CM<NB> C0 = chunked ? cm_load<NB>(L, p.bC + ((size_t)w * (Cn + 1) + ck + 1) * E) : Xt;   // costate after this chunk
  if (GRAD == GRAD_FIRST) {
    if (SYS == SYS_UNITARY) { const double sg = -2.0 * p.sign_static; C = cm_cscale<NB>(C0, sg * tr_, -sg * ti_); }
    else C = cm_neg<NB>(C0);
  } else {
    const double k2 = 2.0 * invD2;
    if (SYS == SYS_UNITARY) C = cm_cscale<NB>(C0, -k2 * tr_, -k2 * ti_);     // conj(-(2/D^2) conj(tau)) = -(2/D^2) tau
    else C = cm_scale<NB>(C0, -k2);
  }
  CM<NB> Pt_n = cm_load<NB>(L, stP + (size_t)(t1 - 1) * E);
  CM<NB> St_n = cm_load<NB>(L, stS + (size_t)(t1 - 1) * E);
  if (GRAD == GRAD_EXACT) prefetch_x(t1 - 1);
  for (int t = t1 - 1; t >= t0; t--) {
    const CM<NB> Pt = Pt_n;
    const CM<NB> St = St_n;
    if (t > t0) { Pt_n = cm_load<NB>(L, stP + (size_t)(t - 1) * E); St_n = cm_load<NB>(L, stS + (size_t)(t - 1) * E); }
    if (GRAD == GRAD_FIRST) {
      CM<NB> WT;
      if (SYS == SYS_UNITARY) {
        C = mul_nt<NB, false, true>(C, Pt);                       // C_t^T = C_{t+1}^T conj(P)   (GRAPE.jl:228)
        const CM<NB> Sn = transpose<NB>(L, St, tb);               // S_t
        const CM<NB> Cn = transpose<NB>(L, C, tb);                // C_t
        WT = mul_nt<NB, true, false>(Cn, Sn);                     // conj(C) S^T = (S C')^T
      } else {
        const CM<NB> Z = mul_nt<NB>(Pt, C);                       // P^T C^T = (C P)^T           (GRAPE.jl:248)
        C = mul_nt<NB, true, false>(Pt, Z);                       // P' (C P)                    (GRAPE.jl:249)
        const CM<NB> Stt = transpose<NB>(L, St, tb);              // S_t^T
        const CM<NB> Ctt = transpose<NB>(L, C, tb);               // C_t^T
        WT = mul_nt<NB, true, false>(C, St);                      // conj(C) S^T = (S C')^T
        mul_nt_acc<NB, false, true>(cm_neg<NB>(Stt), Ctt, WT);    // - S^T conj(C) = -(C' S)^T
      }
      emit_gradient<NB, CPW, SH>(p, L, sl, Bmats, WT, t);         // tr(B W) = sum B .* W^T
    } else {   // exact: F = 1 - |tau|^2/D^2, dF/dx[c,t] = -(2/D^2) Re(conj(tau) dtau), dtau = tr(Y L(G, -i dt B_c))
      CM<NB> Y;
      if (SYS == SYS_UNITARY) {
        const CM<NB> Sn = transpose<NB>(L, St, tb);               // S_t
        const CM<NB> Cn = transpose<NB>(L, C, tb);                // C_{t+1}
        Y = mul_nt<NB, false, true>(Sn, Cn);                      // Y = S_t C_{t+1}'
        C = mul_nt<NB, false, true>(C, Pt);                       // C_t^T
      } else {
        const CM<NB> Ctt = transpose<NB>(L, C, tb);               // C_{t+1}^T
        const CM<NB> Stt = transpose<NB>(L, St, tb);              // S_t^T
        const CM<NB> A1t = mul_nt<NB, true, true>(C, Pt);         // conj(C) conj(P) = (P' C')^T
        const CM<NB> A2t = mul_nt<NB, false, true>(Ctt, Pt);      // C^T conj(P)     = (P' C)^T
        const CM<NB> Y1 = mul_nt<NB>(St, A1t);                    // S P' C'
        const CM<NB> Y2 = mul_nt<NB, true, false>(Stt, A2t);      // S' P' C
        Y = cm_cscale<NB>(Y1, tr_, -ti_);                         // conj(tau) Y1 + tau Y2
        cm_caxpy<NB>(Y, tr_, ti_, Y2);
        const CM<NB> Z = mul_nt<NB>(Pt, C);                       // (C P)^T
        C = mul_nt<NB, true, false>(Pt, Z);                       // C_t = P' C P
      }
      const CM<NB> G = assemble_generator<NB, SH>(L, sysw, xr + (size_t)t * K, K, xpre);
      if (t > t0) prefetch_x(t - 1);
      const CM<NB> Lam = frechet_t8<NB>(L, G, Y, theta, herm, tb);
      emit_gradient<NB, CPW, SH>(p, L, sl, BTmats, Lam, t);       // tr(Lam B_c) = sum Lam .* B_c^T
    }
  }
}

// 5 CTAs/SM -> 96 registers, 20 warps/SM: best of {4,5,6,7} on cfg4 (6 and 7 spill; see profiles/README.md)
#ifndef QOC_CHAIN_MINB
#define QOC_CHAIN_MINB 5
#endif
template <int NB, int CPW, int SYS, int GRAD>
__global__ void __launch_bounds__(128, (NB == 1 && GRAD != GRAD_EXACT) ? QOC_CHAIN_MINB : 1) chain_kernel(const SmallParams p) {
  extern __shared__ double2 smem[];
  if (p.sys_in_smem) chain_body<NB, CPW, SYS, GRAD, true>(p, smem);
  else chain_body<NB, CPW, SYS, GRAD, false>(p, smem);
}

// From U_N^T: C_0 = U_N' Xt (U_N), the overlap tau, the figure of merit and the scaled, sign-prepared starting operator
//   W = -f W_0,  W_0 = Xi C_0' (- C_0' Xi),  f = gradient scalar with i/dt folded in,
// such that the gradient of slice t is  sum conj(B~_c) .* W_t  (for Hermitian B the transposed scaled control is -conj(B~)).
// XiP / XtP are the packed states (unitary: transposed, density: as is).
template <int NB, int CPW, int SYS>
__device__ __forceinline__ CM<NB> unitary_w0(const Lane& L, const CM<NB>& Ut, const CM<NB>& XiP, const CM<NB>& XtP,
                                             int sign_static, double invD2, double* tb, double& fom) {
  constexpr int GS = 32 / CPW;
  double tr_, ti_;
  if (SYS == SYS_UNITARY) {
    const CM<NB> C0 = mul_nt<NB, true, false>(Ut, XtP);         // conj(U^T) Xt = U' Xt
    const CM<NB> Xi = transpose<NB>(L, XiP, tb);
    cm_dotc_partial<NB>(Xi, C0, tr_, ti_);                      // tau = tr(S_N' Xt) = tr(Xi' C_0)
    tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
    fom = tr_ * tr_ - ti_ * ti_;                                // Re(tau*tau)          (cost_functions.jl:99-101)
    const CM<NB> W0 = mul_nt<NB, false, true>(Xi, C0);          // Xi C_0'
    const double sg = 2.0 * sign_static;                        // -f,  f = 2(+-i dt) tau (i/dt) = -+2 tau
    return cm_cscale<NB>(W0, sg * tr_, sg * ti_);
  }
  const CM<NB> Zt = mul_nt<NB>(Ut, XtP);                        // U^T Xt^T = (Xt U)^T
  const CM<NB> C0 = mul_nt<NB, true, false>(Ut, Zt);            // U' Xt U
  cm_dotc_partial<NB>(C0, XiP, tr_, ti_);                       // tau = tr(Xt' S_N) = tr(C_0' Xi)
  tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
  fom = 1.0 - (tr_ * tr_ + ti_ * ti_) * invD2;                  // C1                    (cost_functions.jl:13-17)
  const CM<NB> C0t = transpose<NB>(L, C0, tb);
  const CM<NB> Xit = transpose<NB>(L, XiP, tb);
  CM<NB> W0 = mul_nt<NB, false, true>(XiP, C0);                 // Xi C_0'
  mul_nt_acc<NB, true, false>(cm_neg<NB>(C0t), Xit, W0);        // - C_0' Xi
  return W0;                                                    // -f W_0 with f = (i dt)(i/dt) = -1
}

// Closed systems (Hermitian drift and controls, first-order gradient): every P_t is unitary, so V_t = U_N U_t' and the
// gradient operator is a conjugation,  W_t = S_t C_t' (- C_t' S_t) = U_t W_0 U_t',  W_0 = Xi C_0' (- C_0' Xi),
// C_0 = U_N' Xt (U_N).  Pass 1 computes and stores only P_t while accumulating U_N^T; pass 2 runs W <- P W P' with the
// trace-dots.  Per slice: the same 6 products, ONE transpose (inside expm) instead of four, 2 KB of HBM traffic instead of
// 4 KB, and no state store.  For Hermitian B the transposed scaled control is -conj(B~), so no extra matrices are needed.
template <int NB, int CPW, int SYS, bool SH>
__device__ __forceinline__ void chain_body_unitary(const SmallParams& p, double2* smem) {
  const int warp_in_cta = threadIdx.x >> 5;
  const int w = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  if (w >= p.n_groups) return;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  constexpr int TBW = NB * NB * 2 * TB_PLANE;
  const float theta = (float)p.theta;
  const int K = p.K, N = p.N;
  double* tb = reinterpret_cast<double*>(smem) + (size_t)warp_in_cta * TBW;
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  if (SH) {
    double2* mine = smem + (size_t)(blockDim.x >> 5) * TBW / 2 + (size_t)warp_in_cta * p.nmat * E;
    for (int i = L.lane; i < p.nmat * E; i += 32) mine[i] = sysw[i];
    __syncwarp();
    sysw = mine;
  }
  const double* xr = p.x + (size_t)sl.r * N * K;
  double2* stP = p.storeP + (size_t)w * N * E;
  const double invD2 = 1.0 / ((double)p.D * (double)p.D);
  double xpre[XPF];
  auto prefetch_x = [&](int t) {
#pragma unroll
    for (int j = 0; j < XPF; j++) xpre[j] = (j < K) ? __ldg(xr + (size_t)t * K + j) : 0.0;
  };
  // ---------------- pass 1: propagators and U_N^T = P_0^T P_1^T ... ----------------
  CM<NB> Ut = cm_load<NB>(L, p.ident + (size_t)sl.sysgroup * E);
  prefetch_x(0);
  for (int t = 0; t < N; t++) {
    const CM<NB> G = assemble_generator<NB, SH>(L, sysw, xr + (size_t)t * K, K, xpre);
    if (t + 1 < N) prefetch_x(t + 1);
    const CM<NB> P = expm_t8<NB>(L, G, theta, true, tb);
    cm_store<NB>(L, stP + (size_t)t * E, P);
    Ut = mul_nt<NB>(Ut, P);                                     // U_{t+1}^T = U_t^T P_t^T
  }
  // ---------------- C_0, overlap, figure of merit, W_0 ----------------
  double fom;
  CM<NB> W = unitary_w0<NB, CPW, SYS>(L, Ut, cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E), cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E),
                                     p.sign_static, invD2, tb, fom);
  if (sl.valid && (L.lane % GS) == 0) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
  // ---------------- pass 2: conjugation recursion with the trace-dots ----------------
  const double2* Bmats = sysw + E;
  CM<NB> Pn = cm_load<NB>(L, stP);
  for (int t = 0; t < N; t++) {
    const CM<NB> P = Pn;
    if (t + 1 < N) Pn = cm_load<NB>(L, stP + (size_t)(t + 1) * E);
    emit_gradient<NB, CPW, SH, true>(p, L, sl, Bmats, W, t);
    if (t + 1 < N) {
      W = conj_by<NB>(P, W);                                    // P W P'
    }
  }
}
template <int NB, int CPW, int SYS>
__global__ void __launch_bounds__(128, NB == 1 ? QOC_CHAIN_MINB : 1) chain_unitary_kernel(const SmallParams p) {
  extern __shared__ double2 smem[];
  if (p.sys_in_smem) chain_body_unitary<NB, CPW, SYS, true>(p, smem);
  else chain_body_unitary<NB, CPW, SYS, false>(p, smem);
}

// Slice-parallel propagators: one warp per (system group / pulse, slice).  Writes the packed TRANSPOSED
// propagator used by chain_kernel (storeP) and/or the caller's column-major layout
// (pw_prop_save!, timeevolution.jl:98-110).
template <int NB, int CPW>
__global__ void __launch_bounds__(128) expm_slices_kernel(const SliceParams p) {
  extern __shared__ double2 smem[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long)p.n_groups * p.N) return;
  const int w = (int)(gw / p.N), t = (int)(gw - (long)w * p.N);
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int E = cm_elems<NB>();
  double* tb = reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x >> 5) * (NB * NB * 2 * TB_PLANE);
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  const double* xt = p.x + ((size_t)sl.r * p.N + t) * p.K;
  CM<NB> out = assemble_generator<NB, false>(L, sysw, xt, p.K);                // G = -i dt H
  if (p.mode == 0) out = expm_t8<NB>(L, out, (float)p.theta, p.herm, tb);
  else if (p.mode == 1) out = cm_cscale<NB>(out, 0.0, 1.0 / p.dt);            // H = (i/dt) G
  if (p.storeP) cm_store<NB>(L, p.storeP + ((size_t)w * p.N + t) * E, transpose<NB>(L, out, tb));
  if (p.storeP2) cm_store<NB>(L, p.storeP2 + ((size_t)w * p.N + t) * E, out);
  if (p.out_user) {
    double2* o = p.out_user + (((size_t)sl.r * p.M + sl.k) * p.N + t) * p.D * p.D;
    QOC_FOR_CM(NB) {
      int row, col; chain_coords<NB, CPW>(L, i, j, e, p.D, row, col);
      if (sl.valid && row >= 0) o[(size_t)col * p.D + row] = make_double2(out.re[i][j][e], out.im[i][j][e]);
    }
  }
}

}  // namespace qoc
