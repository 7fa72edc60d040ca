// Low-batch ("phased") small-D pipeline: when there are too few (pulse, member) chains to fill 148 SMs with one
// warp per chain, the slice axis supplies the parallelism.  A chunked three-level prefix scan over slices replaces
// the sequential sweeps of /root/reference/src/GRAPE.jl:53-75:
//   A  expm_slices_kernel      P_t for every (group, slice)                         [slice-parallel]
//   B1 chunk_totals_kernel     T_c = P_{t1-1} ... P_{t0} for every (group, chunk)   [chunk-parallel, L-1 products]
//   B2 boundary_kernel         S[start_c], C[start_c] through the T_c               [Cn sequential steps per group]
//   B3 sweep_kernel            forward and backward sweeps inside every chunk, concurrently, storing S_t, C_t
//   C  grad_slices_kernel      W_t (or the Frechet derivative for the exact gradient) + K trace-dots per slice
//                              [slice-parallel]  -- grad_func!, /root/reference/src/GRAPE.jl:261-287
// With Cn = 1 (B1/B2 skipped) this is plain slice-parallel expm/gradient around two sequential sweep warps per chain.
#pragma once
#include "small_d.cuh"

// resident CTAs (4 warps each) per SM requested for the D = 8 chunk kernels: 5 -> 96 registers per thread, 6 -> 80.
// Measured on cfg4 (expm, sweep): (5,5) 2.437 ms, (5,6) 2.408, (6,5) 2.489, (6,6) 2.459, (7,7) 2.515 -- the exponential spills at 80.
#ifndef QOC_EXPM_MINB
#define QOC_EXPM_MINB 5
#endif
// A/B switches of the exponential work item (profiles/README.md, r02s): deferred chunk-total update (measured slower:
// 1.934 vs 1.902 ms at 4096 chains, off), pulse prefetch one batch ahead (neutral, on)
#ifndef QOC_E_DEFER
#define QOC_E_DEFER 0
#endif
#ifndef QOC_E_XPRE
#define QOC_E_XPRE 1
#endif
#ifndef QOC_SWEEP_MINB
#define QOC_SWEEP_MINB 6
#endif

namespace qoc {



// B1: chunk-total propagators.  T <- T * P_t for t descending: nt(T, Pt) = T * (P_t^T)^T, no transposes needed.
template <int NB, int CPW>
__global__ void __launch_bounds__(128) chunk_totals_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= p.n_groups * p.Cn) return;
  const int w = gw / p.Cn, c = gw - w * p.Cn;
  const Lane L(threadIdx.x & 31);
  constexpr int E = cm_elems<NB>();
  double* tb = reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x >> 5) * (NB * NB * 2 * TB_PLANE);
  const int t0 = chunk_lo(c, p.N, p.Cn), t1 = chunk_lo(c + 1, p.N, p.Cn);
  const double2* Pt = p.storePt + (size_t)w * p.N * E;
  CM<NB> T = cm_load<NB>(L, p.storeP + ((size_t)w * p.N + (t1 - 1)) * E);
  CM<NB> nxt;
  if (t1 - 2 >= t0) nxt = cm_load<NB>(L, Pt + (size_t)(t1 - 2) * E);
  for (int t = t1 - 2; t >= t0; t--) {
    const CM<NB> cur = nxt;
    if (t - 1 >= t0) nxt = cm_load<NB>(L, Pt + (size_t)(t - 1) * E);
    T = mul_nt<NB>(T, cur);
  }
  cm_store<NB>(L, p.totT + ((size_t)w * p.Cn + c) * E, T);
  cm_store<NB>(L, p.totTt + ((size_t)w * p.Cn + c) * E, transpose<NB>(L, T, tb));
}

// B2: boundary states / costates.  Warp 2w: forward over chunks; warp 2w+1: backward.
template <int NB, int CPW, int SYS>
__global__ void __launch_bounds__(128) boundary_kernel(const PhasedParams p) {
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= p.n_groups * 2) return;
  const int w = gw >> 1, dir = gw & 1;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int E = cm_elems<NB>();
  const int N = p.N, Cn = p.Cn;
  if (dir == 0) {
    CM<NB> S = cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E);
    double2* st = p.stS + (size_t)w * (N + 1) * E;
    cm_store<NB>(L, st, S);
    for (int c = 0; c < Cn - 1; c++) {       // the last boundary (S_N) is produced by the sweep itself
      const CM<NB> T = cm_load<NB>(L, p.totT + ((size_t)w * Cn + c) * E);
      if (SYS == SYS_UNITARY) S = mul_nt<NB>(S, T);
      else S = conj_by<NB>(T, S);
      cm_store<NB>(L, st + (size_t)chunk_lo(c + 1, N, Cn) * E, S);
    }
  } else {
    CM<NB> C = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);
    double2* st = p.stC + (size_t)w * (N + 1) * E;
    cm_store<NB>(L, st + (size_t)N * E, C);
    for (int c = Cn - 1; c >= 1; c--) {
      const CM<NB> Tt = cm_load<NB>(L, p.totTt + ((size_t)w * Cn + c) * E);
      if (SYS == SYS_UNITARY) C = mul_nt<NB, false, true>(C, Tt);
      else { const CM<NB> Z = mul_nt<NB>(Tt, C); C = mul_nt<NB, true, false>(Tt, Z); }
      cm_store<NB>(L, st + (size_t)chunk_lo(c, N, Cn) * E, C);
    }
  }
}

// B3: within-chunk sweeps.  Warp index = ((w * Cn + c) * 2 + dir).  Forward reads P, backward reads P^T; both
// prefetch two slices ahead (an iteration is shorter than DRAM latency).
template <int NB, int CPW, int SYS>
__global__ void __launch_bounds__(128) sweep_kernel(const PhasedParams p) {
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= p.n_groups * p.Cn * 2) return;
  const int dir = gw & 1, rest = gw >> 1;
  const int w = rest / p.Cn, c = rest - w * p.Cn;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  const int N = p.N, Cn = p.Cn;
  const int t0 = chunk_lo(c, N, Cn), t1 = chunk_lo(c + 1, N, Cn);
  constexpr int PF = 4;     // prefetch ring depth: a sweep step is far shorter than DRAM latency
  if (dir == 0) {
    const double2* Pm = p.storeP + (size_t)w * N * E;
    double2* st = p.stS + (size_t)w * (N + 1) * E;
    CM<NB> S = cm_load<NB>(L, st + (size_t)t0 * E);
    CM<NB> ring[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) if (t0 + i < t1) ring[i] = cm_load<NB>(L, Pm + (size_t)(t0 + i) * E);
    for (int tb0 = t0; tb0 < t1; tb0 += PF) {
#pragma unroll
      for (int i = 0; i < PF; i++) {
        const int t = tb0 + i;
        if (t < t1) {
          const CM<NB> P = ring[i];
          if (t + PF < t1) ring[i] = cm_load<NB>(L, Pm + (size_t)(t + PF) * E);
          if (SYS == SYS_UNITARY) S = mul_nt<NB>(S, P);
          else S = conj_by<NB>(P, S);
          if (t + 1 < t1 || c == Cn - 1) cm_store<NB>(L, st + (size_t)(t + 1) * E, S);
        }
      }
    }
    if (c == Cn - 1) {      // S holds the final state: overlap and figure of merit (cost_functions.jl:99-111)
      const CM<NB> Xt = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);
      double tr_, ti_;
      const bool ref_unitary_fom = SYS == SYS_UNITARY && !p.fom_exact;
      if (ref_unitary_fom) cm_dotc_partial<NB>(S, Xt, tr_, ti_); else cm_dotc_partial<NB>(Xt, S, tr_, ti_);
      tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
      const double invD2 = 1.0 / ((double)p.D * (double)p.D);
      const double fom = ref_unitary_fom ? tr_ * tr_ - ti_ * ti_ : 1.0 - (tr_ * tr_ + ti_ * ti_) * invD2;
      if ((L.lane % GS) == 0) {
        const int s = L.lane / GS;
        p.tau[((size_t)w * CPW + s) * 2 + 0] = tr_; p.tau[((size_t)w * CPW + s) * 2 + 1] = ti_;
        if (sl.valid) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
      }
    }
  } else {
    const double2* Pm = p.storePt + (size_t)w * N * E;
    double2* st = p.stC + (size_t)w * (N + 1) * E;
    CM<NB> C = cm_load<NB>(L, st + (size_t)t1 * E);
    CM<NB> ring[PF];
#pragma unroll
    for (int i = 0; i < PF; i++) if (t1 - 1 - i >= t0) ring[i] = cm_load<NB>(L, Pm + (size_t)(t1 - 1 - i) * E);
    for (int tb0 = t1 - 1; tb0 >= t0; tb0 -= PF) {
#pragma unroll
      for (int i = 0; i < PF; i++) {
        const int t = tb0 - i;
        if (t >= t0) {
          const CM<NB> Pt = ring[i];
          if (t - PF >= t0) ring[i] = cm_load<NB>(L, Pm + (size_t)(t - PF) * E);
          if (SYS == SYS_UNITARY) C = mul_nt<NB, false, true>(C, Pt);
          else { const CM<NB> Z = mul_nt<NB>(Pt, C); C = mul_nt<NB, true, false>(Pt, Z); }
          if (t > t0 || c == 0) cm_store<NB>(L, st + (size_t)t * E, C);
        }
      }
    }
  }
}

// C: gradient of every (group, slice).  Same formulas and factors as chain_body's backward loop.
template <int NB, int CPW, int SYS, int GRAD>
__global__ void __launch_bounds__(128) grad_slices_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long)p.n_groups * p.N) return;
  const int w = (int)(gw / p.N), t = (int)(gw - (long)w * p.N);
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  double* tb = reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x >> 5) * (NB * NB * 2 * TB_PLANE);
  const int N = p.N, K = p.K;
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  const double2* Bmats = sysw + E;
  const double2* BTmats = sysw + (size_t)(1 + K) * E;
  const int slot = L.lane / GS;
  const double tr_ = p.tau[((size_t)w * CPW + slot) * 2 + 0], ti_ = p.tau[((size_t)w * CPW + slot) * 2 + 1];
  const double invD2 = 1.0 / ((double)p.D * (double)p.D);
  const double2* stS = p.stS + (size_t)w * (N + 1) * E;
  const double2* stC = p.stC + (size_t)w * (N + 1) * E;
  SmallParams sp;   // only the fields emit_gradient reads
  sp.M = p.M; sp.N = N; sp.K = K; sp.gradc = p.gradc;
  if (GRAD == GRAD_FIRST) {
    const CM<NB> St = cm_load<NB>(L, stS + (size_t)t * E);
    const CM<NB> Ct = cm_load<NB>(L, stC + (size_t)t * E);
    CM<NB> WT;
    double fr, fi;
    if (SYS == SYS_UNITARY) {
      const CM<NB> Sn = transpose<NB>(L, St, tb), Cn = transpose<NB>(L, Ct, tb);
      WT = mul_nt<NB, true, false>(Cn, Sn);
      const double sg = -2.0 * p.sign_static; fr = sg * tr_; fi = sg * ti_;
    } else {
      const CM<NB> Stt = transpose<NB>(L, St, tb), Ctt = transpose<NB>(L, Ct, tb);
      WT = mul_nt<NB, true, false>(Ct, St);
      mul_nt_acc<NB, false, true>(cm_neg<NB>(Stt), Ctt, WT);
      fr = -1.0; fi = 0.0;
    }
    emit_gradient<NB, CPW, false>(sp, L, sl, Bmats, cm_cscale<NB>(WT, fr, fi), t);
  } else {
    const double k2 = 2.0 * invD2;
    const CM<NB> St = cm_load<NB>(L, stS + (size_t)t * E);
    const CM<NB> C1 = cm_load<NB>(L, stC + (size_t)(t + 1) * E);      // costate after slice t
    CM<NB> Y;
    double cr, ci;
    if (SYS == SYS_UNITARY) {
      const CM<NB> Sn = transpose<NB>(L, St, tb), Cn = transpose<NB>(L, C1, tb);
      Y = mul_nt<NB, false, true>(Sn, Cn);
      cr = -k2 * tr_; ci = k2 * ti_;
    } else {
      const CM<NB> Pt = cm_load<NB>(L, p.storePt + ((size_t)w * N + t) * E);
      const CM<NB> Ctt = transpose<NB>(L, C1, tb), Stt = transpose<NB>(L, St, tb);
      const CM<NB> A1t = mul_nt<NB, true, true>(C1, Pt);
      const CM<NB> A2t = mul_nt<NB, false, true>(Ctt, Pt);
      const CM<NB> Y1 = mul_nt<NB>(St, A1t);
      const CM<NB> Y2 = mul_nt<NB, true, false>(Stt, A2t);
      Y = cm_cscale<NB>(Y1, tr_, -ti_);
      cm_caxpy<NB>(Y, tr_, ti_, Y2);
      cr = -k2; ci = 0.0;
    }
    const double* xr = p.x + ((size_t)sl.r * N + t) * K;
    const CM<NB> G = assemble_generator<NB, false>(L, sysw, xr, K);
    const CM<NB> Lam = frechet_t8<NB>(L, G, Y, (float)p.theta, p.herm, tb);
    emit_gradient<NB, CPW, false>(sp, L, sl, BTmats, cm_cscale<NB>(Lam, cr, ci), t);
  }
}

// ---- chunk-parallel fused mode (150..1500 chains) ------------------------------------------------------
// K1: warp (w, c): exponentials of the chunk's slices (stored transposed for the later sweeps) and the running
//     chunk total, T_new^T = T_old^T P_t^T = nt(Tt, P).
template <int NB, int CPW, bool SH>
__device__ __forceinline__ void chunk_expm_body(const PhasedParams& p, double2* smem) {
  // grid: one CTA per (chain, group of 4 chunks) of the chains [w_off, w_off + w_cnt): its warps share the chain's system
  // matrices, staged ONCE per CTA (was once per warp: 4x the L2 traffic and shared memory at every CTA start)
  const int warp_in_cta = threadIdx.x >> 5;
  const int Cg = (p.Cn + 3) >> 2;
  const int wl = blockIdx.x / Cg, cg = blockIdx.x - wl * Cg;
  const int w = p.w_off + wl, c = cg * 4 + warp_in_cta;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int E = cm_elems<NB>();
  constexpr int TBW = NB * NB * 2 * TB_PLANE;
  const int N = p.N, K = p.K;
  double* tb = reinterpret_cast<double*>(smem) + (size_t)warp_in_cta * TBW;
  const double2* sysw = p.sys + (size_t)sl.sysgroup * p.nmat * E;
  if (SH) {
    double2* shared_sys = smem + (size_t)(blockDim.x >> 5) * TBW / 2;
    for (int i = threadIdx.x; i < (1 + K) * E; i += blockDim.x) shared_sys[i] = sysw[i];
    __syncthreads();
    sysw = shared_sys;
  }
  if (c >= p.Cn) return;
  const int t0 = chunk_lo(c, N, p.Cn), t1 = chunk_lo(c + 1, N, p.Cn);
  const double* xr = p.x + (size_t)sl.r * N * K;
  double2* stP = p.storePt + (size_t)w * N * E;
  double xpre[XPF];
#pragma unroll
  for (int j = 0; j < XPF; j++) xpre[j] = (j < K) ? __ldg(xr + (size_t)t0 * K + j) : 0.0;
  CM<NB> Tt;
  for (int t = t0; t < t1; t++) {
    const CM<NB> G = assemble_generator<NB, SH>(L, sysw, xr + (size_t)t * K, K, xpre);
    if (t + 1 < t1) {
#pragma unroll
      for (int j = 0; j < XPF; j++) xpre[j] = (j < K) ? __ldg(xr + (size_t)(t + 1) * K + j) : 0.0;
    }
    const CM<NB> P = expm_t8<NB>(L, G, (float)p.theta, p.herm, tb);
    if (p.store_plain) {                                       // closed-system mode: P is stored, P^T only seeds the total
      cm_store<NB>(L, stP + (size_t)t * E, P);
      if (t == t0) Tt = transpose<NB>(L, P, tb); else Tt = mul_nt<NB>(Tt, P);
    } else {
      const CM<NB> Pt = transpose<NB>(L, P, tb);
      cm_store<NB>(L, stP + (size_t)t * E, Pt);
      Tt = (t == t0) ? Pt : mul_nt<NB>(Tt, P);
    }
  }
  if (!p.store_plain) cm_store<NB>(L, p.totTt + ((size_t)w * p.Cn + c) * E, Tt);
  cm_store<NB>(L, p.totT + ((size_t)w * p.Cn + c) * E, transpose<NB>(L, Tt, tb));
}
template <int NB, int CPW>
__global__ void __launch_bounds__(128, NB == 1 ? QOC_EXPM_MINB : 1) chunk_expm_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  if (p.sys_in_smem) chunk_expm_body<NB, CPW, true>(p, smem); else chunk_expm_body<NB, CPW, false>(p, smem);
}

// K2: per group, forward warp: boundary states bS[0..Cn] and the overlap / figure of merit; backward warp: boundary
// costates bC[Cn..1] (unscaled: chain_body folds the gradient factor in when it loads them).
template <int NB, int CPW, int SYS>
__global__ void __launch_bounds__(128) boundary2_kernel(const PhasedParams p) {
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= p.n_groups * 2) return;
  const int w = gw >> 1, dir = gw & 1;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  const int Cn = p.Cn;
  if (dir == 0) {
    CM<NB> S = cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E);
    double2* st = p.bS + (size_t)w * (Cn + 1) * E;
    cm_store<NB>(L, st, S);
    for (int c = 0; c < Cn; c++) {
      const CM<NB> T = cm_load<NB>(L, p.totT + ((size_t)w * Cn + c) * E);
      if (SYS == SYS_UNITARY) S = mul_nt<NB>(S, T);
      else S = conj_by<NB>(T, S);
      cm_store<NB>(L, st + (size_t)(c + 1) * E, S);
    }
    const CM<NB> Xt = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);
    double tr_, ti_;
    const bool ref_unitary_fom = SYS == SYS_UNITARY && !p.fom_exact;
    if (ref_unitary_fom) cm_dotc_partial<NB>(S, Xt, tr_, ti_); else cm_dotc_partial<NB>(Xt, S, tr_, ti_);
    tr_ = group_sum<GS>(tr_); ti_ = group_sum<GS>(ti_);
    const double invD2 = 1.0 / ((double)p.D * (double)p.D);
    const double fom = ref_unitary_fom ? tr_ * tr_ - ti_ * ti_ : 1.0 - (tr_ * tr_ + ti_ * ti_) * invD2;
    if ((L.lane % GS) == 0) {
      const int s = L.lane / GS;
      p.tau[((size_t)w * CPW + s) * 2 + 0] = tr_; p.tau[((size_t)w * CPW + s) * 2 + 1] = ti_;
      if (sl.valid) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
    }
  } else {
    CM<NB> C = cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E);
    double2* st = p.bC + (size_t)w * (Cn + 1) * E;
    cm_store<NB>(L, st + (size_t)Cn * E, C);
    for (int c = Cn - 1; c >= 1; c--) {
      const CM<NB> Tt = cm_load<NB>(L, p.totTt + ((size_t)w * Cn + c) * E);
      if (SYS == SYS_UNITARY) C = mul_nt<NB, false, true>(C, Tt);
      else { const CM<NB> Z = mul_nt<NB>(Tt, C); C = mul_nt<NB, true, false>(Tt, Z); }
      cm_store<NB>(L, st + (size_t)c * E, C);
    }
  }
}

// ---- chunk-parallel closed-system mode: K1 = chunk_expm (stores P, not P^T), K2u, K3u ---------------------------
// K2u: one warp per group: U_N^T from the chunk totals, W_0 (unitary_w0), then the chunk-boundary operators
//      bW[c+1] = T_c bW[c] T_c'  (stored in the bS buffer).
// CG: the chunk totals were written by other CTAs of the SAME launch (fused use below): read them through L2.
template <int NB, bool CG> __device__ __forceinline__ CM<NB> cm_load_tot(const Lane& L, const double2* p) {
  if (!CG) return cm_load<NB>(L, p);
  CM<NB> x;
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const double2 r = __ldcg(p + ((i * NB + j) * 2 + 0) * 32 + L.lane);
      const double2 m = __ldcg(p + ((i * NB + j) * 2 + 1) * 32 + L.lane);
      x.re[i][j][0] = r.x; x.re[i][j][1] = r.y; x.im[i][j][0] = m.x; x.im[i][j][1] = m.y;
    }
  return x;
}
template <int NB, int CPW, int SYS, bool CG>
__device__ __forceinline__ void boundary_unitary_chain(const PhasedParams& p, const Lane& L, int w, double* tb) {
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int GS = 32 / CPW;
  constexpr int E = cm_elems<NB>();
  const int Cn = p.Cn;
  const double2* T = p.totT + (size_t)w * Cn * E;
  CM<NB> Tn = cm_load_tot<NB, CG>(L, T + (size_t)(Cn > 1 ? 1 : 0) * E);         // loads run one chunk ahead of the products
  CM<NB> Ut = transpose<NB>(L, cm_load_tot<NB, CG>(L, T), tb);                  // U^T after chunk 0
  for (int c = 1; c < Cn; c++) {
    const CM<NB> Tc = Tn;
    if (c + 1 < Cn) Tn = cm_load_tot<NB, CG>(L, T + (size_t)(c + 1) * E);
    Ut = mul_nt<NB>(Ut, Tc);                                                  // U^T T_c^T
  }
  double fom;
  CM<NB> W = unitary_w0<NB, CPW, SYS>(L, Ut, cm_load<NB>(L, p.xi + (size_t)sl.sysgroup * E), cm_load<NB>(L, p.xt + (size_t)sl.sysgroup * E),
                                     p.sign_static, 1.0 / ((double)p.D * (double)p.D), tb, fom);
  if (sl.valid && (L.lane % GS) == 0) p.fomc[(size_t)sl.r * p.M + sl.k] = fom;
  double2* bW = p.bS + (size_t)w * (Cn + 1) * E;
  cm_store<NB>(L, bW, W);
  Tn = cm_load_tot<NB, CG>(L, T);
  for (int c = 0; c + 1 < Cn; c++) {
    const CM<NB> Tc = Tn;
    if (c + 2 < Cn) Tn = cm_load_tot<NB, CG>(L, T + (size_t)(c + 1) * E);
    W = conj_by<NB>(Tc, W);
    cm_store<NB>(L, bW + (size_t)(c + 1) * E, W);
  }
}
template <int NB, int CPW, int SYS>
__global__ void __launch_bounds__(128) boundary_unitary_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  const int wl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wl >= p.w_cnt) return;
  const Lane L(threadIdx.x & 31);
  double* tb = reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x >> 5) * (NB * NB * 2 * TB_PLANE);
  boundary_unitary_chain<NB, CPW, SYS, false>(p, L, p.w_off + wl, tb);
}
// K3u: warp (w, c): conjugation recursion over the chunk's slices with the trace-dots.
template <int NB, int CPW, bool SH>
__device__ __forceinline__ void sweep_unitary_body(const PhasedParams& p, double2* smem) {
  const int warp_in_cta = threadIdx.x >> 5;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  if (gw >= p.w_cnt * p.Cn) return;
  const int w = p.w_off + gw / p.Cn, c = gw % p.Cn;
  const Lane L(threadIdx.x & 31);
  const Slot<CPW> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  constexpr int E = cm_elems<NB>();
  const int N = p.N, K = p.K;
  const int t0 = chunk_lo(c, N, p.Cn), t1 = chunk_lo(c + 1, N, p.Cn);
  const double2* Bmats = p.sys + (size_t)sl.sysgroup * p.nmat * E + E;
  if (SH) {
    double2* mine = smem + (size_t)warp_in_cta * K * E;
    for (int i = L.lane; i < K * E; i += 32) mine[i] = Bmats[i];
    __syncwarp();
    Bmats = mine;
  }
  SmallParams sp; sp.M = p.M; sp.N = N; sp.K = K; sp.gradc = p.gradc;
  const double2* stP = p.storePt + (size_t)w * N * E;      // holds P (not P^T) in this mode
  CM<NB> W = cm_load<NB>(L, p.bS + ((size_t)w * (p.Cn + 1) + c) * E);
  CM<NB> Pn = cm_load<NB>(L, stP + (size_t)t0 * E);
  for (int t = t0; t < t1; t++) {
    const CM<NB> P = Pn;
    if (t + 1 < t1) Pn = cm_load<NB>(L, stP + (size_t)(t + 1) * E);
    emit_gradient<NB, CPW, SH, true>(sp, L, sl, Bmats, W, t);
    if (t + 1 < t1) W = conj_by<NB>(P, W);
  }
}
template <int NB, int CPW>
__global__ void __launch_bounds__(128, NB == 1 ? QOC_SWEEP_MINB : 1) sweep_unitary_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  if (p.sys_in_smem) sweep_unitary_body<NB, CPW, true>(p, smem); else sweep_unitary_body<NB, CPW, false>(p, smem);
}

// K3u with the trace-dots on the tensor pipe (D = 5..8, one chain per warp, K <= 8).
// The scalar form of emit_gradient costs 24 DFMA + a 9-stage shuffle / DADD butterfly per slice, every one of them queueing
// for the FP64 pipe behind the other warps' DMMAs (ncu r02e: 75 % of this kernel's warp time, top stall math-pipe throttle).
// Here W_t of 8 consecutive slices is staged in a per-warp shared-memory tile and all K x 8 dots
//   g[c][t] = sum_f Bflat[c][f] * Wflat[t][f],   f over the 128 real entries (re plane, then im plane),
// are one accumulated chain of 32 DMMA.8x8x4 (m = control, n = slice, k = f): the contraction over k is the cross-lane
// reduction, the result lands as g[c = lane / 4][t = 2 (lane % 4) + {0, 1}].  Same FP64-pipe time (4 DMMA per slice), but
// ~110 instead of ~250 instructions per slice and no dependent shuffle chain.
//   Bd [32 k-steps][32 lanes]  control matrices in fragment order (lane (c, q) <-> Bflat[c][4 ks + q]), shared by the CTA:
//                              the four warps of a CTA work on four chunks of the SAME chain
//   Wb [8 slices][132]         per warp; row stride 132 doubles makes the fragment loads conflict-free
constexpr int DOT_LD = 132;
// One CTA-sized work item: chunks 4 cg .. 4 cg + 3 of chain w.  `Wstride`: doubles between the per-warp tiles.
// CG: the boundary operator was written by another CTA of the SAME launch (persistent kernel): read it through L2.
template <bool CG>
__device__ __forceinline__ void sweep_unitary_dmma_item(const PhasedParams& p, double* Bd, int Wstride, int w, int cg) {
  constexpr int NB = 1;
  constexpr int E = cm_elems<NB>();
  const int warp = threadIdx.x >> 5;
  double* Wb = Bd + 1024 + warp * Wstride;
  const int c = cg * 4 + warp;
  const Lane L(threadIdx.x & 31);
  const Slot<1> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  const int N = p.N, K = p.K;
  const int nks = p.dot_nks;                                  // > 0: only the union of the controls' non-zero entries is contracted
  int o_r0 = 0, o_r1 = 0, o_i0 = 0, o_i1 = 0;                 // compact slots of this lane's four W values (-1: not needed)
  {
    const double* Bm = reinterpret_cast<const double*>(p.sys + (size_t)sl.sysgroup * p.nmat * E + E);
    if (nks) {
      for (int i = threadIdx.x; i < nks * 32; i += blockDim.x) {
        const int ks = i >> 5, ln = i & 31, cc = ln >> 2, q = ln & 3;
        const int f = __ldg(p.dot_tab + 128 + 4 * ks + q);
        Bd[i] = (cc < K && f >= 0) ? Bm[(size_t)cc * (2 * E) + f] : 0.0;
      }
      o_r0 = __ldg(p.dot_tab + 2 * L.lane); o_r1 = __ldg(p.dot_tab + 2 * L.lane + 1);
      o_i0 = __ldg(p.dot_tab + 64 + 2 * L.lane); o_i1 = __ldg(p.dot_tab + 64 + 2 * L.lane + 1);
      for (int r = 0; r < 8; r++)                             // padding slots are never stored to: they must hold finite values
        for (int sidx = L.lane; sidx < 4 * nks; sidx += 32) Wb[r * DOT_LD + sidx] = 0.0;
    } else {
      for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int ks = i >> 5, ln = i & 31, cc = ln >> 2, q = ln & 3;
        Bd[i] = cc < K ? Bm[(size_t)cc * (2 * E) + 4 * ks + q] : 0.0;
      }
    }
  }
  __syncthreads();
  if (c >= p.Cn) return;
  const int t0 = chunk_lo(c, N, p.Cn), t1 = chunk_lo(c + 1, N, p.Cn);
  const double2* stP = p.storePt + (size_t)w * N * E;      // holds P (not P^T) in this mode
  double* out = p.gradc + ((size_t)sl.r * p.M + sl.k) * N * K;
  CM<NB> W = cm_load_tot<NB, CG>(L, p.bS + ((size_t)w * (p.Cn + 1) + c) * E);
  CM<NB> Pn = cm_load<NB>(L, stP + (size_t)t0 * E);
  const int cc = L.g, tq = 2 * L.q;
  for (int tb = t0; tb < t1; tb += 8) {
    const int nb = min(8, t1 - tb);
    for (int i = 0; i < nb; i++) {
      const int t = tb + i;
      const CM<NB> P = Pn;
      if (t + 1 < t1) Pn = cm_load<NB>(L, stP + (size_t)(t + 1) * E);
      if (nks) {
        double* row = Wb + i * DOT_LD;
        if (o_r0 >= 0) row[o_r0] = W.re[0][0][0];
        if (o_r1 >= 0) row[o_r1] = W.re[0][0][1];
        if (o_i0 >= 0) row[o_i0] = W.im[0][0][0];
        if (o_i1 >= 0) row[o_i1] = W.im[0][0][1];
      } else {
        double* row = Wb + i * DOT_LD + 2 * L.lane;
        *reinterpret_cast<double2*>(row) = make_double2(W.re[0][0][0], W.re[0][0][1]);
        *reinterpret_cast<double2*>(row + 64) = make_double2(W.im[0][0][0], W.im[0][0][1]);
      }
      if (t + 1 < t1) W = conj_by<NB>(P, W);
    }
    __syncwarp();
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0;      // four accumulator chains
    const double* bp = Bd + L.lane;
    const double* wp = Wb + L.g * DOT_LD + L.q;
    if (nks) {
      for (int ks = 0; ks < nks; ks += 4) {
        dmma(a0, a1, bp[(ks + 0) * 32], wp[4 * (ks + 0)]);
        dmma(b0, b1, bp[(ks + 1) * 32], wp[4 * (ks + 1)]);
        dmma(c0, c1, bp[(ks + 2) * 32], wp[4 * (ks + 2)]);
        dmma(d0, d1, bp[(ks + 3) * 32], wp[4 * (ks + 3)]);
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < 32; ks += 4) {
        dmma(a0, a1, bp[(ks + 0) * 32], wp[4 * (ks + 0)]);
        dmma(b0, b1, bp[(ks + 1) * 32], wp[4 * (ks + 1)]);
        dmma(c0, c1, bp[(ks + 2) * 32], wp[4 * (ks + 2)]);
        dmma(d0, d1, bp[(ks + 3) * 32], wp[4 * (ks + 3)]);
      }
    }
    const double g0 = (a0 + b0) + (c0 + d0), g1 = (a1 + b1) + (c1 + d1);
    if (cc < K && sl.valid) {
      if (tq < nb) out[(size_t)(tb + tq) * K + cc] = g0;
      if (tq + 1 < nb) out[(size_t)(tb + tq + 1) * K + cc] = g1;
    }
    __syncwarp();
  }
}
__global__ void __launch_bounds__(128, 5) sweep_unitary_dmma_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  const int Cg = (p.Cn + 3) >> 2;
  const int wl = blockIdx.x / Cg;
  sweep_unitary_dmma_item<false>(p, reinterpret_cast<double*>(smem), 8 * DOT_LD, p.w_off + wl, blockIdx.x - wl * Cg);
}

// K1 with the generator assembly on the tensor pipe (D = 5..8, one chain per warp, K <= 7).
// G_t = A~ + sum_j x[j,t] B~_j costs 24 dependent DFMA + 14 LDS.128 per slice in scalar form; like the trace-dots of the sweep
// these scalar FP64 instructions queue behind the other warps' DMMAs.  Here the generators of 8 consecutive slices are
// one batched product  Gflat[f][t] = sum_j' Coef[f][j'] * X[j'][t]  (m = flat matrix entry f, 16 blocks of 8; n = slice;
// k = coefficient index j' = 0 (drift, X = 1), 1..K (controls)): 32 DMMA.8x8x4 per 8 slices, results staged in a per-warp
// shared-memory tile from which every slice's G is read back in the register layout with two LDS.128.
//   Ad [16 m-blocks][2 k-steps][32 lanes]   coefficient matrices in fragment order, shared by the CTA (same chain)
//   per warp: [28 spare][8 rows x DOT_LD]   the transpose tile of warp_mat.cuh (160 doubles) aliases the spare doubles + row 0,
//                                           which is dead as soon as slice 0 of the batch sits in registers
constexpr int ASM_WARP_DOUBLES = 28 + 8 * DOT_LD;
// One CTA-sized work item: chunks 4 cg .. 4 cg + 3 of chain w.
__device__ __forceinline__ void chunk_expm_dmma_item(const PhasedParams& p, double* Ad, int w, int cg) {
  constexpr int NB = 1;
  constexpr int E = cm_elems<NB>();
  const int warp = threadIdx.x >> 5;
  double* tile = Ad + 1024 + warp * ASM_WARP_DOUBLES;      // transpose tile (2 * TB_PLANE doubles) ...
  double* Gb = tile + 28;                                  // ... overlapping the first staging row
  const int c = cg * 4 + warp;
  const Lane L(threadIdx.x & 31);
  const Slot<1> sl(p.pack_mode, p.n_inner, p.M, p.R, L, w);
  const int N = p.N, K = p.K;
  const bool sparse = p.asm_sparse != 0;
  const int jr = (int)((p.asm_lr >> (8 * L.q)) & 0xffu), ji = (int)((p.asm_li >> (8 * L.q)) & 0xffu);   // this lane's k index -> coefficient
  {
    const double* sysd = reinterpret_cast<const double*>(p.sys + (size_t)sl.sysgroup * p.nmat * E);
    if (sparse && p.asm_nblk) {    // compact blocks over the union of non-zero entries; position table behind the coefficients
      int* posS = reinterpret_cast<int*>(Ad + 512);
      for (int i = threadIdx.x; i < p.asm_nblk * 32; i += blockDim.x) {
        const int ln = i & 31, b = i >> 5, g = ln >> 2, q = ln & 3;
        const int f = __ldg(p.asm_pos + 8 * b + g);
        const int j = (int)(((b < p.asm_nblk_re ? p.asm_lr : p.asm_li) >> (8 * q)) & 0xffu);
        Ad[i] = (f >= 0 && j <= K) ? sysd[(size_t)j * (2 * E) + f] : 0.0;
      }
      for (int i = threadIdx.x; i < 128; i += blockDim.x) posS[i] = i < 8 * p.asm_nblk ? __ldg(p.asm_pos + i) : -1;
      for (int i = L.lane; i < 8 * DOT_LD; i += 32) Gb[i] = 0.0;      // entries outside the union stay zero in rows 1..7
    } else if (sparse) {       // Ad [16 m-blocks][32 lanes]: blocks 0..7 = real plane (list asm_lr), 8..15 = imaginary plane (asm_li)
      for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int ln = i & 31, mb = i >> 5, g = ln >> 2, q = ln & 3;
        const int j = (int)(((mb < 8 ? p.asm_lr : p.asm_li) >> (8 * q)) & 0xffu);
        Ad[i] = j <= K ? sysd[(size_t)j * (2 * E) + 8 * mb + g] : 0.0;
      }
    } else {
      for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int ln = i & 31, ks = (i >> 5) & 1, mb = i >> 6, g = ln >> 2, q = ln & 3, j = 4 * ks + q;
        Ad[i] = j <= K ? sysd[(size_t)j * (2 * E) + 8 * mb + g] : 0.0;
      }
    }
  }
  __syncthreads();
  if (c >= p.Cn) return;
  const int t0 = chunk_lo(c, N, p.Cn), t1 = chunk_lo(c + 1, N, p.Cn);
  const double* xr = p.x + (size_t)sl.r * N * K;
  double2* stP = p.storePt + (size_t)w * N * E;
  CM<NB> Tt;
#if QOC_E_DEFER
  CM<NB> Pprev;                     // the chunk-total update of slice t is issued at the start of slice t + 1 (see below)
#endif
  // X fragments of a batch of 8 slices: lane (g = slice, q): dense: k-step 0 -> (1, x_1, x_2, x_3)[q], k-step 1 -> x_{4+q};
  // plane-wise: real-plane list entry q, imaginary-plane list entry q.  Loaded one batch ahead (an L2 round trip per batch
  // would otherwise sit on every warp's critical path).
  auto load_x = [&](int tb, double& f0, double& f1) {
    const int ts = min(tb + L.g, t1 - 1);
    const double* xs = xr + (size_t)ts * K;
    if (sparse) {
      f0 = jr == 0 ? 1.0 : (jr <= K ? __ldg(xs + jr - 1) : 0.0);
      f1 = ji == 0 ? 1.0 : (ji <= K ? __ldg(xs + ji - 1) : 0.0);
    } else {
      f0 = L.q == 0 ? 1.0 : (L.q <= K ? __ldg(xs + L.q - 1) : 0.0);
      f1 = 4 + L.q <= K ? __ldg(xs + 3 + L.q) : 0.0;
    }
  };
  double xn0, xn1;
  load_x(t0, xn0, xn1);
  for (int tb = t0; tb < t1; tb += 8) {
    const int nb = min(8, t1 - tb);
    {
      const double b0 = xn0, b1 = xn1;
#if QOC_E_XPRE
      if (tb + 8 < t1) load_x(tb + 8, xn0, xn1);
#endif
      const double* ap = Ad + L.lane;
      double* gp = Gb + (2 * L.q) * DOT_LD + L.g;
      if (sparse && p.asm_nblk) {
        // row 0 doubles as the transpose tile: clear it again, then scatter the blocks' results to their flat positions
        *reinterpret_cast<double2*>(Gb + 2 * L.lane) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(Gb + 64 + 2 * L.lane) = make_double2(0.0, 0.0);
        __syncwarp();
        const int* posS = reinterpret_cast<const int*>(Ad + 512) + L.g;
        double* gq = Gb + (2 * L.q) * DOT_LD;
        const int nblk = p.asm_nblk, nre = p.asm_nblk_re;
        for (int b = 0; b < nblk; b++) {
          double d0 = 0.0, d1 = 0.0;
          dmma(d0, d1, ap[b * 32], b < nre ? b0 : b1);
          const int f = posS[8 * b];
          if (f >= 0) { gq[f] = d0; gq[f + DOT_LD] = d1; }
        }
      } else if (sparse) {
#pragma unroll
        for (int mb = 0; mb < 16; mb++) {
          double d0 = 0.0, d1 = 0.0;
          dmma(d0, d1, ap[mb * 32], mb < 8 ? b0 : b1);
          gp[8 * mb] = d0; gp[8 * mb + DOT_LD] = d1;
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < 16; mb++) {
          double d0 = 0.0, d1 = 0.0;
          dmma(d0, d1, ap[(2 * mb) * 32], b0);
          dmma(d0, d1, ap[(2 * mb + 1) * 32], b1);
          gp[8 * mb] = d0; gp[8 * mb + DOT_LD] = d1;
        }
      }
#if !QOC_E_XPRE
      if (tb + 8 < t1) load_x(tb + 8, xn0, xn1);
#endif
    }
    __syncwarp();
    for (int i = 0; i < nb; i++) {
      const int t = tb + i;
      CM<NB> G;
      {
        const double2 r = *reinterpret_cast<const double2*>(Gb + i * DOT_LD + 2 * L.lane);
        const double2 m = *reinterpret_cast<const double2*>(Gb + i * DOT_LD + 64 + 2 * L.lane);
        G.re[0][0][0] = r.x; G.re[0][0][1] = r.y; G.im[0][0][0] = m.x; G.im[0][0][1] = m.y;
      }
      __syncwarp();                                           // row 0 may now be overwritten by the transposes
#if QOC_E_DEFER
      // closed-system mode: the chunk-total product of the PREVIOUS slice is independent of this slice's norm estimate and
      // first product, so its DMMAs fill the shuffle latency of the estimate (in-order issue: a warp has nothing else to issue)
      if (p.store_plain && t > t0 + 1) Tt = mul_nt<NB>(Tt, Pprev);
#endif
      const CM<NB> P = expm_t8<NB>(L, G, (float)p.theta, p.herm, tile);
      if (p.store_plain) {
        cm_store<NB>(L, stP + (size_t)t * E, P);
#if QOC_E_DEFER
        if (t == t0) Tt = transpose<NB>(L, P, tile);
        Pprev = P;
#else
        if (t == t0) Tt = transpose<NB>(L, P, tile); else Tt = mul_nt<NB>(Tt, P);
#endif
      } else {
        const CM<NB> Pt = transpose<NB>(L, P, tile);
        cm_store<NB>(L, stP + (size_t)t * E, Pt);
        Tt = (t == t0) ? Pt : mul_nt<NB>(Tt, P);
      }
    }
    __syncwarp();
  }
#if QOC_E_DEFER
  if (p.store_plain && t1 - 1 > t0) Tt = mul_nt<NB>(Tt, Pprev);
#endif
  if (!p.store_plain) cm_store<NB>(L, p.totTt + ((size_t)w * p.Cn + c) * E, Tt);
  cm_store<NB>(L, p.totT + ((size_t)w * p.Cn + c) * E, transpose<NB>(L, Tt, tile));
}
__global__ void __launch_bounds__(128, 5) chunk_expm_dmma_kernel(const PhasedParams p) {
  extern __shared__ double2 smem[];
  const int Cg = (p.Cn + 3) >> 2;
  const int wl = blockIdx.x / Cg;
  chunk_expm_dmma_item(p, reinterpret_cast<double*>(smem), p.w_off + wl, blockIdx.x - wl * Cg);
}

// ---- persistent closed-system kernel: K1 + K2u + K3u of ALL chains in ONE launch -----------------------------------------
// The three-launch form leaves SM slots idle at every kernel boundary: the exponential CTAs of a chain range live ~70 us, the
// 10 us boundary kernel behind them waits for free slots, and the sweeps cannot start before it (profiles/r02p_shard_timeline.txt:
// a 512-chain shard runs 0.297 ms against 0.257 ms of work).  Here 5 CTAs per SM stay resident and pull CTA-sized work items
// from device-memory queues until everything is done:
//   E (chain, 4 chunks)  exponentials + chunk totals (chunk_expm_dmma_item).  The CTA that finishes the LAST E item of a chain
//                        (per-chain counter) runs the chain's boundary stage -- U_N, figure of merit, W_0, chunk-boundary
//                        operators (boundary_unitary_chain) -- and appends the chain's S items to the S list
//   S (chain, 4 chunks)  conjugation recursion + trace-dots (sweep_unitary_dmma_item)
// E items go first (long items first: the short S items fill the tail), except that S items are taken early while more than
// `reserve` of them are waiting (E and S work mixed on an SM use the FP64 pipe better than either alone, and a chain's
// propagators are re-read while they still sit in L2).  Both queues are ticket
// counters (one atomicAdd per item; a compare-and-swap claim serialises: one winner per L2 round trip, measured 0.7 us per
// item).  An S ticket may be drawn before its entry is published; the holder waits for it (bounded to ~4 s: a producer that
// never arrives yields NaN, not a hung GPU).  That wait cannot deadlock: S
// tickets are drawn only once every E item has been claimed, and every unpublished entry depends only on E items that
// resident CTAs are executing.  A list entry is (id + 1); the whole control block is zeroed by ONE memset node before the
// launch.  Data written by other CTAs of the launch (chunk totals, boundary operators) is read with ld.global.cg after a
// fence; P_t is read once per launch, so no stale L1 line can exist.
struct PersistCtl { int e_next, s_head, s_tail, pad[5]; };
__device__ __forceinline__ int ld_vol(const int* q) { return *reinterpret_cast<const volatile int*>(q); }
__device__ __forceinline__ void st_vol(int* q, int v) { *reinterpret_cast<volatile int*>(q) = v; }
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
inline int persist_ctl_ints(int n_groups, int Cn) { return 8 + n_groups + n_groups * ((Cn + 3) / 4); }

template <int SYS>
__global__ void __launch_bounds__(128, 5) closed_persistent_kernel(const PhasedParams p, int* ctl_raw, int reserve, unsigned long long* trace) {
  extern __shared__ double2 smem[];
  __shared__ int sh_item[2];
  double* smd = reinterpret_cast<double*>(smem);
  const int Cg = (p.Cn + 3) >> 2;
  const int nE = p.n_groups * Cg;
  PersistCtl* ctl = reinterpret_cast<PersistCtl*>(ctl_raw);
  int* chain_cnt = ctl_raw + 8;
  int* s_ready = chain_cnt + p.n_groups;
  bool e_done = false;                                       // meaningful in thread 0 only
  for (;;) {
    if (threadIdx.x == 0) {
      int a = -1;                                            // >= 0: E item a;  < -1: S item -(a + 2);  -1: exit
      bool take_s = false;
      if (!e_done) {
        // mixing: an S item is taken early when more than `reserve` of them are waiting (reserve >= gridDim.x, so the ticket
        // drawn below is certainly one of the published ones); the reserve itself stays for the tail
        take_s = ld_vol(&ctl->s_tail) - ld_vol(&ctl->s_head) > reserve;
        if (!take_s) {
          a = atomicAdd(&ctl->e_next, 1);
          if (a >= nE) { e_done = true; a = -1; }
        }
      }
      if (e_done || take_s) {
        const int j = atomicAdd(&ctl->s_head, 1);
        if (j < nE) {
          int v;
          unsigned backoff = 32;
          unsigned long long t_wait = 0;
          while ((v = ld_vol(s_ready + j)) == 0) {
            __nanosleep(backoff);
            if (backoff < 256) backoff <<= 1;
            const unsigned long long now = global_ns();
            if (t_wait == 0) t_wait = now;
            else if (now - t_wait > 4000000000ULL) break;      // ~4 s: the producer never arrived -- poison the result, do not hang the GPU
          }
          __threadfence();
          if (v > 0) a = -(v - 1) - 2;
          else p.fomc[0] = __longlong_as_double(0x7ff8000000000000LL);      // a stays -1: this CTA leaves
        }
      }
      sh_item[0] = a;
    }
    __syncthreads();
    const int a = sh_item[0];
    if (a == -1) return;
    unsigned long long t_begin = 0;
    if (trace && threadIdx.x == 0) t_begin = global_ns();
    if (a >= 0) {
      const int w = a / Cg;
      chunk_expm_dmma_item(p, smd, w, a - w * Cg);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) sh_item[1] = atomicAdd(&chain_cnt[w], 1) == Cg - 1;      // this CTA finished the chain's last E item
      __syncthreads();
      if (sh_item[1] && threadIdx.x < 32) {
        const Lane L(threadIdx.x);
        __threadfence();
        boundary_unitary_chain<1, 1, SYS, true>(p, L, w, smd + 1024);
        __threadfence();
        __syncwarp();
        if (L.lane == 0) {
          const int slot = atomicAdd(&ctl->s_tail, Cg);
          for (int i = 0; i < Cg; i++) st_vol(s_ready + slot + i, w * Cg + i + 1);
        }
      }
    } else {
      const int item = -(a + 2);
      const int w = item / Cg;
      sweep_unitary_dmma_item<true>(p, smd, ASM_WARP_DOUBLES, w, item - w * Cg);
    }
    __syncthreads();                                          // the shared tiles and sh_item are reused by the next item
    if (trace && threadIdx.x == 0) {                          // QOC_PERSIST_TRACE: (kind, CTA, SM), item, begin, end [ns]
      const unsigned long long t_end = global_ns();
      const unsigned long long slot = atomicAdd(trace, 1ULL);
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* r = trace + 4 + 4 * slot;
      r[0] = (a >= 0 ? 1ULL : 3ULL) | ((unsigned long long)blockIdx.x << 8) | ((unsigned long long)smid << 32);
      r[1] = (unsigned long long)(a >= 0 ? a : -(a + 2));
      r[2] = t_begin; r[3] = t_end;
    }
  }
}

}  // namespace qoc
