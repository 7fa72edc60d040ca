/* qocgrape.h — C ABI of libqocgrape.so: B200-native (sm_100a) GRAPE fidelity + gradient evaluation.
 *
 * This is the drop-in boundary for ONE hot path of QuOptimalControl.jl.  The reference has no FFI of its
 * own (it is pure Julia); each entry point below states the reference routine it replaces
 * (paths relative to /root/reference/).  The Julia glue that binds these with `ccall`
 * (struct GPUGRAPE + solve methods) is shown in INTEGRATION.md; the same ABI is driven from Python ctypes by
 * quoptimalcontrol.jl_b200/_lib.py.
 *
 * Conventions (identical to the reference's in-memory data, so no transposes on either side):
 *   - complex numbers are interleaved (re, im) doubles == Julia ComplexF64 == C `double _Complex`;
 *   - matrices are column-major D x D;
 *   - a pulse x is Julia's K x N column-major `control_array[j, i]` (control j, slice i): x[j + i*K];
 *     the gradient has the same shape;
 *   - every function returns a status (0 = QOC_OK); no exception crosses the boundary;
 *   - the library copies inputs during the call and never retains host pointers;
 *   - a handle is used by one thread at a time; different handles are independent;
 *   - a handle created with n_devices > 1 shards the ensemble members over several GPUs inside ONE process
 *     (single-process multi-device: one qoc_eval call drives all of them, see qoc_desc.n_devices);
 *   - there is NO CPU fallback: without a CUDA device or for unsupported shapes an error is returned;
 *   - for D > 16 qoc_eval_device synchronises its stream once per evaluation (a 4-byte read of the scaling power).
 */
#ifndef QOCGRAPE_H
#define QOCGRAPE_H

#ifdef __cplusplus
extern "C" {
#endif

#define QOC_OK 0
#define QOC_EINVAL 1        /* bad argument / shape */
#define QOC_ECUDA 2         /* CUDA runtime error (see qoc_last_error) */
#define QOC_ENOMEM 3        /* device workspace does not fit */
#define QOC_EUNSUPPORTED 4  /* shape or mode not implemented by the CUDA path */

/* src/problems.jl:8-10 — StateTransfer / UnitaryGate / CoherenceTransfer (README's ClosedStateTransfer /
 * UnitarySynthesis / OpenSystemCoherenceTransfer) */
#define QOC_STATE_TRANSFER 0
#define QOC_UNITARY_GATE 1
#define QOC_COHERENCE_TRANSFER 2

/* gradient: first order = grad_func!/grad_func (src/GRAPE.jl:261-303); exact = real(Zygote.gradient) of the
 * ADGRAPE functional (src/GRAPE.jl:14-18, src/solve.jl:268-290), computed by Frechet derivatives. */
#define QOC_GRAD_FIRST_ORDER 0
#define QOC_GRAD_EXACT 1

/* first-order sign convention for UnitaryGate: in-place grad_func! (+i dt, src/GRAPE.jl:272) or static
 * grad_func (-i dt, src/GRAPE.jl:290).  Ignored otherwise. */
#define QOC_REF_INPLACE 0
#define QOC_REF_STATIC 1

/* qoc_desc.flags */
#define QOC_FLAG_NO_PURE_STATE 1 /* D > 16 StateTransfer between pure states on a sparse closed system is evaluated with
                                    state vectors (same F, G to rounding, O(nnz) per slice); set to force the dense path */

/* shared_flags for qoc_set_system: the argument holds ONE matrix (set) used by all M members */
#define QOC_SHARED_A 1
#define QOC_SHARED_B 2
#define QOC_SHARED_XI 4
#define QOC_SHARED_XT 8

#define QOC_MAX_DEVICES 16
#define QOC_IPC_HANDLE_BYTES 64   /* sizeof(cudaIpcMemHandle_t): what the ranks exchange for the peer-memory entry points */
#define QOC_MAX_RANKS 16

typedef struct qoc_handle qoc_handle;

typedef struct qoc_desc {
  int sys_type;      /* QOC_STATE_TRANSFER | QOC_UNITARY_GATE | QOC_COHERENCE_TRANSFER */
  int D;             /* matrix dimension (Hilbert or Liouville) */
  int K;             /* n_controls (Problem.n_controls, src/problems.jl:25) */
  int N;             /* n_slices (Piecewise.n_slices, src/timeevolution.jl:11-14) */
  int M;             /* ensemble members (EnsembleProblem.n_ens, src/problems.jl:35); 1 for a plain Problem */
  int R;             /* independent pulses evaluated per call (multi-start batch); 1 for solve() */
  double T;          /* pulse duration (Problem.T) ; dt = T/N (src/GRAPE.jl:42) */
  int gradient;      /* QOC_GRAD_FIRST_ORDER | QOC_GRAD_EXACT */
  int convention;    /* QOC_REF_INPLACE | QOC_REF_STATIC */
  int device;        /* CUDA device ordinal */
  double expm_theta; /* scaling threshold of the degree-8 Taylor exponential; <= 0 selects the default
                        (0.0694: truncation error below 2^-53) */
  int flags;         /* QOC_FLAG_* bits, 0 = defaults */
  /* Single-process multi-device ensembles (SURVEY.md 8b "Threading", 8e): with n_devices > 1 the M members are
   * block-partitioned over device_ids[0 .. n_devices) (member k -> device k * n_devices / M, like the serial member loop
   * src/solve.jl:166 cut into contiguous blocks); qoc_set_system uploads each block to its device, qoc_eval launches all
   * blocks concurrently from the calling thread (one CUDA graph spanning the devices) and device_ids[0] sums the weighted
   * partial [F|G] rows in fixed device order over NVLink peer memory before the single D2H copy.  `device` is ignored
   * then.  An ordinal may be repeated (several shards on one device: useful for testing on a single-GPU box).  n_devices <= 1 keeps the single-device behaviour (and a zero-initialised tail keeps old callers valid). */
  int n_devices;
  int device_ids[QOC_MAX_DEVICES];
} qoc_desc;

typedef struct qoc_stats {
  long long n_evals;          /* calls of qoc_eval / qoc_eval_device */
  long long n_launches;       /* kernels launched by this handle so far */
  int launches_last_eval;     /* kernels launched by the most recent evaluation */
  float gpu_ms_last_eval;     /* device time of the most recent qoc_eval (CUDA events), 0 for eval_device */
  long long workspace_bytes;  /* device memory held */
  int path;                   /* 1 = warp-resident DMMA (D <= 16), 2 = tiled DMMA GEMM (D > 16),
                                 3 = pure-state vector sweep (D > 16, see QOC_FLAG_NO_PURE_STATE) */
  float main_kernel_ms_avg;   /* mean device time of the dominant kernel over the evaluations since the previous
                                 qoc_get_stats call (CUDA events on the launching stream, at most 64 samples) */
  int main_kernel_samples;
} qoc_stats;

const char* qoc_version(void);

/* Allocates device workspace for the shape in `desc`.  Replaces init_GRAPE (src/grape_tools.jl:4-16). */
int qoc_create(qoc_handle** out, const qoc_desc* desc);
int qoc_destroy(qoc_handle* h);

/* Uploads the problem(s): Problem.A/B/Xi/Xt (src/problems.jl:19-28), or for an ensemble the M member problems
 * produced by init_ensemble (src/tools.jl:42-53) and EnsembleProblem.wts.
 *   A  [M][D*D]      (or [D*D] with QOC_SHARED_A)
 *   B  [M][K][D*D]   (or [K][D*D] with QOC_SHARED_B)
 *   Xi [M][D*D], Xt [M][D*D] (or single with the SHARED flags)
 *   wts [M] or NULL (all 1.0; a plain Problem is M = 1, weight 1). */
int qoc_set_system(qoc_handle* h, const double* A, const double* B, const double* Xi, const double* Xt,
                   const double* wts, int shared_flags);

/* One fidelity+gradient evaluation per pulse: the body of the Optim.only_fg! closure
 * (src/solve.jl:75-100 single problem, :164-196 ensemble) = _fom_and_gradient_GRAPE! (src/GRAPE.jl:25-96)
 * over all members, weighted and summed.
 *   x [R][N*K] host;  F [R] or NULL;  G [R][N*K] or NULL (value-only evaluation skips the backward sweep).
 * With G == NULL and R > 1 this is the batched fidelity-only call for gradient-free callers: one call evaluates a
 * whole Nelder-Mead simplex / dCRAB candidate set (src/dCRAB.jl:52-70 evaluates them one user_func call at a time). */
int qoc_eval(qoc_handle* h, const double* x, double* F, double* G);

/* Optional control-amplitude / control-variation penalties added to every evaluation (the PenaltyFunctionals C3 and C4
 * of src/cost_functions.jl:29-39, weighted):  F += w_amp * sum(x.^2) + w_var * sum(diff(x, dims = 2).^2) and G gets the
 * matching derivative; applied once per pulse after the ensemble reduction (and after the all-reduce / device sum).
 * Both weights 0 (the default) disables the extra kernel. */
int qoc_set_penalty(qoc_handle* h, double w_amp, double w_var);

/* Same with DEVICE pointers, asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the CUDA default
 * stream, as everywhere in the CUDA runtime): x_dev [R][N*K];  FG_dev [R][1 + N*K] with F first, then G.  Lets one-process-per-GPU callers
 * all-reduce FG_dev (NCCL) without a host round trip.  With want_gradient == 0 the G columns of FG_dev are written as 0. */
int qoc_eval_device(qoc_handle* h, const double* x_dev, double* FG_dev, int want_gradient, void* stream);

/* pw_evolve (src/timeevolution.jl:28-39) with U0 = I:  U [R][M][D*D] = P_N ... P_1. */
int qoc_total_propagator(qoc_handle* h, const double* x, double* U);

/* pw_prop_save! (src/timeevolution.jl:98-110) for mode 0, pw_ham_save! (:64-75) for mode 1,
 * pw_gen_save! (:80-92) for mode 2:  out [R][M][N][D*D]. */
int qoc_propagators(qoc_handle* h, const double* x, double* out, int mode);

/* ---- slice-parallel evaluation of ONE large instance over several GPUs (SURVEY.md 8e "config 5", 8f rank 4) -------
 * Each rank owns a contiguous range of slices and a handle created with that range's N and T (D > 16, M = R = 1,
 * QOC_FLAG_NO_PURE_STATE).  Per evaluation: (1) qoc_total_propagator gives the range's propagator product U_r;
 * (2) the ranks exchange the U_r (D x D each: the path's only exchange) and form their boundary operators, the state
 * before and the costate after their range (pw_evolve semantics, src/GRAPE.jl:216-251 applied to whole ranges);
 * (3) qoc_set_states installs them as Xi / Xt; (4) qoc_eval_continue finishes the evaluation from the propagators that
 * step (1) left on the device: F is the full figure of merit, G [N_r][K] the gradient entries of the rank's slices. */
int qoc_set_states(qoc_handle* h, const double* Xi, const double* Xt, int shared_flags);
int qoc_eval_continue(qoc_handle* h, double* F, double* G);

/* The same four steps inside the library, one process per GPU, no host staging and no library GEMMs: the range propagators
 * are exchanged through CUDA-IPC peer buffers (flag signalling as in the all-reduce below) and the boundary operators are
 * formed by the library's own DMMA GEMM kernel reading the peers' propagators straight over NVLink.
 *   qoc_slice_export   allocate this rank's exchange buffer, return its IPC handle (handle for the range: D > 16, M = R = 1,
 *                      QOC_FLAG_NO_PURE_STATE; its qoc_set_system Xi / Xt are placeholders)
 *   qoc_slice_connect  open the peers' buffers (handles [world][QOC_IPC_HANDLE_BYTES]); Xi, Xt: the GLOBAL initial / target
 *                      operators [D*D] of the whole problem (src/problems.jl:19-28)
 *   qoc_eval_slice     x [N_r*K] = this rank's slices of the pulse; F = the full figure of merit (identical on every rank),
 *                      G [N_r*K] = the gradient entries of this rank's slices (or NULL).  Every rank must call it, in the
 *                      same order.  qoc_set_penalty does not apply here (C4 couples adjacent ranges). */
int qoc_slice_export(qoc_handle* h, unsigned char* handle /* [QOC_IPC_HANDLE_BYTES] */);
int qoc_slice_connect(qoc_handle* h, int world, int rank, const unsigned char* handles, const double* Xi, const double* Xt);
int qoc_eval_slice(qoc_handle* h, const double* x, double* F, double* G);

/* ---- multi-GPU, one process per GPU: fused one-shot all-reduce of [F|G] over NVLink peer memory -------------------
 * Replaces the cross-shard part of the ensemble reduction `sum(gradient .* wts, dims = 1)` (src/solve.jl:171-191) when
 * the members are sharded over GPUs.  Each rank publishes its weighted partial in a CUDA-IPC shared exchange buffer,
 * raises an epoch flag in every peer, and sums all peers' partials in fixed rank order (deterministic, bit-identical on
 * all ranks).  The host side only has to carry the 64-byte IPC handles between processes (e.g. torch.distributed
 * all_gather_object, MPI, a file).
 *   qoc_comm_export   allocate this rank's exchange buffer and return its IPC handle
 *   qoc_comm_connect  open the peers' buffers; handles = [world][QOC_IPC_HANDLE_BYTES], own entry included
 *   qoc_eval_allreduce_device   like qoc_eval_device, but FG_dev receives the sum over all ranks
 *   qoc_eval_allreduce          like qoc_eval (HOST buffers; every rank passes the same x and receives the summed F, G):
 *                               H2D, kernels, the all-reduce and D2H are one CUDA-graph launch per call
 * The member reduction's second pass is folded into the all-reduce kernel, which keeps its epoch counter in device memory
 * (so the whole sequence is graph-replayable).  All calls of a handle must be issued in the same order on every rank;
 * qoc_comm_connect may be called once per handle. */
int qoc_comm_export(qoc_handle* h, unsigned char* handle /* [QOC_IPC_HANDLE_BYTES] */);
int qoc_comm_connect(qoc_handle* h, int world, int rank, const unsigned char* handles);
int qoc_eval_allreduce_device(qoc_handle* h, const double* x_dev, double* FG_dev, int want_gradient, void* stream);
int qoc_eval_allreduce(qoc_handle* h, const double* x, double* F, double* G);

/* ---- the caller of the path: L-BFGS inside the library ---------------------------------------
 * Replaces `Optim.optimize(Optim.only_fg!(topt), guess, LBFGS(), optim_options)` (src/solve.jl:138, :244) for a
 * single pulse (R = 1): two-loop recursion with `history` pairs (initial inverse Hessian scaled by s'y / y'y like Optim's
 * scaleinvH0) and the Hager-Zhang line search with approximate Wolfe conditions (Optim.LBFGS's default, restated from the
 * published algorithm with LineSearches.jl's default constants; initial trial step 1 = InitialStatic) on the host, every
 * trial being one qoc_eval (a CUDA-graph replay; also on multi-device handles).  Stops on ||g||_inf <= g_tol (Optim's default criterion,
 * 1e-8), on a relative decrease below f_tol, or after max_iters iterations.  x0 / x_out: [N*K] like qoc_eval. */
#define QOC_LS_HAGER_ZHANG 0
#define QOC_LS_BACKTRACKING 1
typedef struct qoc_lbfgs_options {
  int max_iters;        /* <= 0: 1000 (Optim.Options default) */
  int history;          /* <= 0: 10 (Optim.LBFGS default m) */
  double g_tol;         /* <= 0: 1e-8 */
  double f_tol;         /* < 0: 0 (disabled, Optim default) */
  int max_linesearch;   /* evaluations per line search; <= 0: 50 (Hager-Zhang, LineSearches' linesearchmax) / 30 (backtracking) */
  int linesearch;       /* QOC_LS_HAGER_ZHANG (0, default: what Optim.LBFGS() uses) | QOC_LS_BACKTRACKING */
} qoc_lbfgs_options;
typedef struct qoc_lbfgs_result {
  double minimum;       /* res.minimum */
  double g_norm;        /* ||g||_inf at the minimizer */
  int iterations, f_calls, converged;
} qoc_lbfgs_result;
int qoc_minimize_lbfgs(qoc_handle* h, const double* x0, const qoc_lbfgs_options* opt, double* x_out, qoc_lbfgs_result* res);

int qoc_get_stats(qoc_handle* h, qoc_stats* out);
/* Diagnostics (pure host code, no CUDA call, usable without a GPU): the exact elementwise structure qoc_set_system derives
   from the caller's matrices to select kernel forms -- Hermiticity (closed-system recursion; the reference has no such
   notion: /root/reference/src/GRAPE.jl:53-75 always runs the general sweeps), which of -i dt A, -i dt B_j have a real /
   imaginary plane, and the unions of non-zero entries for the compact generator assembly and the compact trace-dots
   (D <= 8).  A [M or 1][D*D], B [M or 1][K][D*D] as in qoc_set_system.
   out (392 ints): [0] herm, [1] asm_sparse, [2] asm_lr, [3] asm_li (4 packed bytes each: 0 = drift, j = control j, 0xff =
   unused), [4] asm_nblk_re, [5] asm_nblk, [6] dot_nks, [7] 0, [8..135] asm_pos, [136..391] dot_tab. */
int qoc_analyze_structure(int D, int K, int M, const double* A, const double* B, int shared_flags, int* out);
/* Message of the last error on this handle (or of the last failed qoc_create when h == NULL). */
const char* qoc_last_error(qoc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* QOCGRAPE_H */
