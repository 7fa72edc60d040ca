// Plain parameter blocks, constants and the cross-translation-unit entry points of the small-dimension kernels.
// The kernels are compiled in separate translation units (k_small_fused.cu, k_small_phased.cu, k_big.cu) so that a
// change to one kernel family does not recompile the others; the host orchestration (qocgrape.cu) reaches them through
// the pick_* functions (kernel function pointers) and launch_* wrappers declared at the end of this file.
#pragma once
#include <cuda_runtime.h>

namespace qoc {

constexpr int TB_PLANE = 80;   // doubles per real 8x8 plane in the transpose tile: 8 rows x stride 10

enum { SYS_DENSITY = 0, SYS_UNITARY = 1 };
enum { GRAD_NONE = 0, GRAD_FIRST = 1, GRAD_EXACT = 2 };

// ||A||_1 <= theta  =>  ||A||^9/9! * e^||A|| <= 2^-53
constexpr double T8_THETA_DEFAULT = 0.0694;

template <int NB> __host__ __device__ constexpr int cm_elems() { return NB * NB * 2 * 32; }  // double2 per packed matrix

constexpr int RED_MAX_CHUNKS = 128, RED_LANES = 8;

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
  int herm;            // drift and all controls are Hermitian (host-checked): generator is anti-Hermitian
  double dt, theta;
  const double2* sys;  // [n_sysgroups][nmat][NB*NB*2*32], pre-multiplied by -i*dt
  const double2* xi;   // [n_sysgroups][NB*NB*2*32]; UnitaryGate: packed transposed (the chain runs on S^T)
  const double2* xt;
  const double* x;     // [R][N][K]  (= the reference's K x N column-major control_array per pulse)
  double2* storeP;     // [n_groups][N][NB*NB*2*32]
  double2* storeS;     // [n_groups][N][NB*NB*2*32]
  double* fomc;        // [R][M]
  double* gradc;       // [R][M][N][K]
  double2* out_final;  // optional [R][M][D*D]: final forward state, column-major complex
  // chunk-parallel fused mode (Cn > 1): warp (w, c) handles slices [c*N/Cn, (c+1)*N/Cn) of group w, starting from
  // the boundary state bS[w][c] and boundary costate bC[w][c+1]; overlaps come from tau_in (boundary2_kernel)
  int Cn;
  const double2* bS;   // [n_groups][Cn+1][E]
  const double2* bC;   // [n_groups][Cn+1][E]
  const double* tau_in;  // [n_groups][CPW][2]
  const double2* ident;  // packed identity (closed-system kernel)
};

struct SliceParams {
  int D, N, K, M, R, pack_mode, n_groups, n_inner, nmat, herm;
  double dt, theta;
  const double2* sys;
  const double* x;
  double2* storeP;     // optional packed [n_groups][N][E]: TRANSPOSED result
  double2* storeP2;    // optional packed [n_groups][N][E]: result as is
  double2* out_user;   // optional [R][M][N][D*D] column-major complex
  int mode;            // 0: propagator exp(-i dt H); 1: Hamiltonian H (pw_ham_save!); 2: generator -i dt H (pw_gen_save!)
};

struct PackParams {
  int D, NB, CPW, n_og, nmat_dst, mat_dst, transpose;
  int pack_mode;      // 0: slot s -> member og*CPW+s (clamped to n_src-1); 1: every slot -> member og
  int n_src;          // number of distinct source matrices (1 if shared)
  long src_stride;    // in double2 between consecutive source members (0 if shared)
  double scale_re, scale_im;   // every element is multiplied by this complex factor (-i*dt for A, B)
  const double2* src;
  double2* dst;
};

struct PhasedParams {
  int D, N, K, M, R, pack_mode, n_groups, n_inner, nmat, herm, Cn;
  int sign_static, fom_exact;
  int w_off, w_cnt;        // chunk-parallel kernels: the launch covers the chains (groups) [w_off, w_off + w_cnt)
  double theta;
  const double2* sys;      // packed, pre-multiplied by -i dt
  const double2* xi;       // packed (unitary: transposed)
  const double2* xt;
  const double* x;
  double2* storePt;        // [n_groups][N][E]   P_t^T
  double2* storeP;         // [n_groups][N][E]   P_t
  double2* stS;            // [n_groups][N+1][E] unitary: S_t^T, density: S_t
  double2* stC;            // [n_groups][N+1][E] unitary: C_t^T, density: C_t   (C_N = Xt)
  double2* totT;           // [n_groups][Cn][E]  T_c
  double2* totTt;          // [n_groups][Cn][E]  T_c^T
  double* tau;             // [n_groups][CPW][2] overlap per chain (written by the forward sweep)
  double2* bS;             // [n_groups][Cn+1][E] chunk-boundary states   (chunk-parallel fused mode)
  double2* bC;             // [n_groups][Cn+1][E] chunk-boundary costates
  int sys_in_smem;
  int store_plain;         // chunk_expm_kernel stores P_t instead of P_t^T (closed-system mode)
  // chunk_expm_dmma_item: when, over all members, at most 4 of the 1 + K scaled matrices (-i dt A, -i dt B_j) have a non-zero
  // real plane and at most 4 a non-zero imaginary plane (Pauli-type controls: each is purely real or purely imaginary), every
  // plane is assembled from its own coefficient list in ONE k-step: 16 instead of 32 DMMA per 8 slices.  Lists as 4 packed
  // bytes (coefficient index 0 = drift, j = control j; 0xff = unused).
  int asm_sparse;
  unsigned asm_lr, asm_li;
  // sweep_unitary_dmma_item: trace-dots over the union of the controls' non-zero entries only.  dot_tab[0..127] = compact slot
  // of flat entry f (re plane 0..63, im plane 64..127; -1: no control has a non-zero there), dot_tab[128..255] = flat entry of
  // slot s (-1: padding); dot_nks = k-steps of 4 slots (a multiple of 4; 0 = dense form, all 32 k-steps).
  const int* dot_tab;
  int dot_nks;
  // plane-wise assembly over the union of the matrices' non-zero entries only: asm_pos[8 b + g] = flat entry (row g of compact
  // block b; -1 = padding), blocks [0, asm_nblk_re) belong to the real plane, [asm_nblk_re, asm_nblk) to the imaginary plane;
  // asm_nblk = 0: all 16 blocks in natural order.  Entries outside the union are zero in every generator.
  const int* asm_pos;
  int asm_nblk_re, asm_nblk;
  double* fomc;
  double* gradc;
};

__host__ __device__ __forceinline__ int chunk_lo(int c, int N, int Cn) { return (int)((long)c * N / Cn); }

// ---- kernel pickers (defined next to the kernels they instantiate) ---------------------------------------------------
typedef void (*chain_fn)(const SmallParams);
typedef void (*slice_fn)(const SliceParams);
typedef void (*phased_fn)(const PhasedParams);
chain_fn pick_chain(int NB, int CPW, int sys, int grad);            // chain_kernel<NB, CPW, SYS, GRAD>
chain_fn pick_chain_unitary(int NB, int CPW, int sys);              // chain_unitary_kernel<NB, CPW, SYS>
slice_fn pick_slices(int NB, int CPW);                              // expm_slices_kernel<NB, CPW>
void pick_phased(int NB, int CPW, int sys, int grad, phased_fn& tot, phased_fn& bnd, phased_fn& swp, phased_fn& grd);
phased_fn pick_chunk_expm(int NB, int CPW);
phased_fn pick_boundary2(int NB, int CPW, int sys);
phased_fn pick_boundary_unitary(int NB, int CPW, int sys);
phased_fn pick_sweep_unitary(int NB, int CPW);
phased_fn pick_chunk_expm_dmma();         // D = 5..8 (NB = 1, one chain per warp), K <= 7: generator assembly on the tensor pipe
int chunk_expm_dmma_smem();
phased_fn pick_sweep_unitary_dmma();     // D = 5..8 (NB = 1, one chain per warp), K <= 8: trace-dots on the tensor pipe
int sweep_unitary_dmma_smem();           // its dynamic shared memory per CTA; grid = chains x ceil(Cn / 4)
// persistent closed-system kernel (D = 5..8, 1 <= K <= 7): exponentials, boundary stage and sweeps of all chains in one launch,
// CTA-sized work items pulled from device-memory queues; ctl = closed_persistent_ctl_ints() ints, zeroed before every launch
typedef void (*persist_fn)(const PhasedParams, int*, int, unsigned long long*);
persist_fn pick_closed_persistent(int sys);
int closed_persistent_smem();
int closed_persistent_ctl_ints(int n_groups, int Cn);
// ---- launch wrappers of the non-template kernels (defined in k_small_fused.cu); return cudaGetLastError() ---------------
cudaError_t launch_pack(const PackParams& pp, long total, cudaStream_t st);
cudaError_t launch_reduce_pass1(const double* gradc, const double* fomc, const double* wts, double* part, int M, int NK, int R,
                                int chunk, int nchunks, cudaStream_t st, int ch0 = 0, int nrows = -1);   // partial rows [ch0, ch0 + nrows)
cudaError_t launch_reduce_pass2(const double* part, double* out, int NK, int R, int nchunks, cudaStream_t st);

}  // namespace qoc
