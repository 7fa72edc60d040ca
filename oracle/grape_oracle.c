/* grape_oracle.c — plain-C restatement of the reference's GRAPE fidelity+gradient CPU path.
 *
 * TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline); never linked into libqocgrape.so.
 * PARITY UNPINNED: the reference ships no golden vectors and Julia is absent; see oracle/grape_oracle.py.
 * This file is cross-checked against the numpy oracle in tests/test_oracle_c.py.
 *
 * It keeps the reference's loop order AND operation counts (so that timing it is a fair stand-in for the Julia
 * code, which cannot run here): per slice one Pade scaling-and-squaring exponential, 1 (UnitaryGate) or 2
 * (density) GEMMs per direction, and 3 GEMMs per (control, slice) in the gradient double loop.
 *   pw_prop_save!              /root/reference/src/timeevolution.jl:98-110
 *   _fom_and_gradient_GRAPE!   /root/reference/src/GRAPE.jl:25-96
 *   evolve_func!               /root/reference/src/GRAPE.jl:216-251
 *   grad_func! / grad_func     /root/reference/src/GRAPE.jl:261-303
 *   fom_func / C1              /root/reference/src/cost_functions.jl:99-111, 13-17
 *   ensemble closure           /root/reference/src/solve.jl:164-196   (members serial in the reference; here
 *                              optionally spread over OpenMP threads, which flatters the reference)
 * exp: restates the published algorithm of Julia's LinearAlgebra.exp! (stdlib dense.jl; Higham 2005): Pade
 * degree 3/5/7/9 for ||A||_1 <= 0.015/0.25/0.95/2.1, else degree 13 with scaling and squaring; (V-U)\(V+U) by
 * LU with partial pivoting (gesv).  gebal balancing is omitted (it only permutes/scales; result identical up
 * to rounding).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;
#define IDX(r, c, D) ((size_t)(c) * (D) + (r))   /* column-major */

enum { STATE_TRANSFER = 0, UNITARY_GATE = 1, COHERENCE_TRANSFER = 2 };

/* C = op(A) * op(B); opA/opB: 0 = N, 1 = conjugate transpose */
static void gemm(int D, const cplx* A, int opA, const cplx* B, int opB, cplx* C) {
  for (int j = 0; j < D; j++) {
    for (int i = 0; i < D; i++) C[IDX(i, j, D)] = 0;
    for (int k = 0; k < D; k++) {
      cplx b = opB ? conj(B[IDX(j, k, D)]) : B[IDX(k, j, D)];
      if (!opA) {
        const cplx* a = A + (size_t)k * D;
        cplx* c = C + (size_t)j * D;
        for (int i = 0; i < D; i++) c[i] += a[i] * b;
      } else {
        for (int i = 0; i < D; i++) C[IDX(i, j, D)] += conj(A[IDX(k, i, D)]) * b;
      }
    }
  }
}
static cplx trace(int D, const cplx* A) { cplx t = 0; for (int i = 0; i < D; i++) t += A[IDX(i, i, D)]; return t; }
static double norm1(int D, const cplx* A) {
  double best = 0;
  for (int j = 0; j < D; j++) { double s = 0; for (int i = 0; i < D; i++) s += cabs(A[IDX(i, j, D)]); if (s > best) best = s; }
  return best;
}
/* solve A X = B in place (B <- X), LU with partial pivoting; A is destroyed */
static void gesv(int D, cplx* A, cplx* B) {
  for (int k = 0; k < D; k++) {
    int p = k; double best = cabs(A[IDX(k, k, D)]);
    for (int i = k + 1; i < D; i++) { double v = cabs(A[IDX(i, k, D)]); if (v > best) { best = v; p = i; } }
    if (p != k) for (int j = 0; j < D; j++) {
      cplx t = A[IDX(k, j, D)]; A[IDX(k, j, D)] = A[IDX(p, j, D)]; A[IDX(p, j, D)] = t;
      t = B[IDX(k, j, D)]; B[IDX(k, j, D)] = B[IDX(p, j, D)]; B[IDX(p, j, D)] = t;
    }
    cplx inv = 1.0 / A[IDX(k, k, D)];
    for (int i = k + 1; i < D; i++) {
      cplx l = A[IDX(i, k, D)] * inv;
      if (l == 0) continue;
      for (int j = k + 1; j < D; j++) A[IDX(i, j, D)] -= l * A[IDX(k, j, D)];
      for (int j = 0; j < D; j++) B[IDX(i, j, D)] -= l * B[IDX(k, j, D)];
    }
  }
  for (int j = 0; j < D; j++)
    for (int i = D - 1; i >= 0; i--) {
      cplx s = B[IDX(i, j, D)];
      for (int k = i + 1; k < D; k++) s -= A[IDX(i, k, D)] * B[IDX(k, j, D)];
      B[IDX(i, j, D)] = s / A[IDX(i, i, D)];
    }
}

/* X <- exp(X); work: 6*D*D.  Returns the number of matrix products executed (excluding the solve). */
static int expm_pade(int D, cplx* X, cplx* work) {
  static const double C3[] = {120., 60., 12., 1.};
  static const double C5[] = {30240., 15120., 3360., 420., 30., 1.};
  static const double C7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
  static const double C9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
  static const double CC[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                              129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                              40840800., 960960., 16380., 182., 1.};
  const size_t n = (size_t)D * D;
  cplx *A2 = work, *P = work + n, *U = work + 2 * n, *V = work + 3 * n, *T1 = work + 4 * n, *T2 = work + 5 * n;
  double nA = norm1(D, X);
  int prods = 0;
  if (nA <= 2.1) {
    const double* C; int nc;
    if (nA > 0.95) { C = C9; nc = 10; } else if (nA > 0.25) { C = C7; nc = 8; } else if (nA > 0.015) { C = C5; nc = 6; } else { C = C3; nc = 4; }
    gemm(D, X, 0, X, 0, A2); prods++;
    memcpy(P, A2, n * sizeof(cplx));
    for (size_t i = 0; i < n; i++) { U[i] = C[3] * P[i]; V[i] = C[2] * P[i]; }
    for (int i = 0; i < D; i++) { U[IDX(i, i, D)] += C[1]; V[IDX(i, i, D)] += C[0]; }
    for (int k = 2; k <= nc / 2 - 1; k++) {
      gemm(D, P, 0, A2, 0, T1); prods++;
      memcpy(P, T1, n * sizeof(cplx));
      for (size_t i = 0; i < n; i++) { U[i] += C[2 * k + 1] * P[i]; V[i] += C[2 * k] * P[i]; }
    }
    gemm(D, X, 0, U, 0, T1); prods++;
    for (size_t i = 0; i < n; i++) { T2[i] = V[i] - T1[i]; X[i] = V[i] + T1[i]; }
    gesv(D, T2, X);
  } else {
    double s = log2(nA / 5.4);
    int si = 0;
    if (s > 0) { si = (int)ceil(s); double sc = ldexp(1.0, -si); for (size_t i = 0; i < n; i++) X[i] *= sc; }
    cplx *A4 = P, *A6 = T2;
    gemm(D, X, 0, X, 0, A2); gemm(D, A2, 0, A2, 0, A4); gemm(D, A2, 0, A4, 0, A6); prods += 3;
    for (size_t i = 0; i < n; i++) T1[i] = CC[13] * A6[i] + CC[11] * A4[i] + CC[9] * A2[i];
    gemm(D, A6, 0, T1, 0, U); prods++;
    for (size_t i = 0; i < n; i++) U[i] += CC[7] * A6[i] + CC[5] * A4[i] + CC[3] * A2[i];
    for (int i = 0; i < D; i++) U[IDX(i, i, D)] += CC[1];
    gemm(D, X, 0, U, 0, T1); prods++;                        /* T1 = U (odd part) */
    for (size_t i = 0; i < n; i++) U[i] = CC[12] * A6[i] + CC[10] * A4[i] + CC[8] * A2[i];
    gemm(D, A6, 0, U, 0, V); prods++;
    for (size_t i = 0; i < n; i++) V[i] += CC[6] * A6[i] + CC[4] * A4[i] + CC[2] * A2[i];
    for (int i = 0; i < D; i++) V[IDX(i, i, D)] += CC[0];
    for (size_t i = 0; i < n; i++) { A2[i] = V[i] - T1[i]; X[i] = V[i] + T1[i]; }
    gesv(D, A2, X);
    for (int t = 0; t < si; t++) { gemm(D, X, 0, X, 0, T1); memcpy(X, T1, n * sizeof(cplx)); prods++; }
  }
  return prods;
}

/* One member: _fom_and_gradient_GRAPE! / _fom_and_gradient_sGRAPE.  x: [N][K] (= Julia K x N column-major).
 * g: [N][K].  variant 0 = in-place (grad_func!), 1 = static (grad_func).  Returns fom. */
static double eval_member(int sys, int D, int K, int N, double T, const cplx* A, const cplx* B, const cplx* Xi,
                          const cplx* Xt, const double* x, int variant, double* g, cplx* ws) {
  const size_t n = (size_t)D * D;
  const double dt = T / N;
  cplx *P = ws, *S = P + (size_t)N * n, *C = S + (size_t)(N + 1) * n, *H = C + (size_t)(N + 1) * n;
  cplx *store = H + n, *t1 = store + n, *t2 = t1 + n, *ework = t2 + n;   /* ework: 6n */
  memcpy(S, Xi, n * sizeof(cplx));                                         /* GRAPE.jl:44 */
  memcpy(C + (size_t)N * n, Xt, n * sizeof(cplx));                         /* GRAPE.jl:45 */
  for (int i = 0; i < N; i++) {                                            /* timeevolution.jl:102-109 */
    for (size_t e = 0; e < n; e++) H[e] = 0;
    for (int j = 0; j < K; j++) { double xj = x[(size_t)i * K + j]; const cplx* Bj = B + (size_t)j * n; for (size_t e = 0; e < n; e++) H[e] += Bj[e] * xj; }
    cplx* Pi = P + (size_t)i * n;
    for (size_t e = 0; e < n; e++) Pi[e] = (-I * dt) * (H[e] + A[e]);
    expm_pade(D, Pi, ework);
  }
  const int unitary = sys == UNITARY_GATE;
  for (int t = 0; t < N; t++) {                                            /* GRAPE.jl:53-63 */
    cplx *Pt = P + (size_t)t * n, *St = S + (size_t)t * n, *Sn = S + (size_t)(t + 1) * n;
    if (unitary) gemm(D, Pt, 0, St, 0, Sn);                                /* :226 */
    else { gemm(D, St, 0, Pt, 1, store); gemm(D, Pt, 0, store, 0, Sn); }   /* :245-246 */
  }
  for (int t = N - 1; t >= 0; t--) {                                       /* GRAPE.jl:65-75 */
    cplx *Pt = P + (size_t)t * n, *Cn = C + (size_t)(t + 1) * n, *Ct = C + (size_t)t * n;
    if (unitary) gemm(D, Pt, 1, Cn, 0, Ct);                                /* :228 */
    else { gemm(D, Cn, 0, Pt, 0, store); gemm(D, Pt, 1, store, 0, Ct); }   /* :248-249 */
  }
  for (int c = 0; c < K; c++) {                                            /* GRAPE.jl:79-92 */
    const cplx* Bc = B + (size_t)c * n;
    for (int t = 0; t < N; t++) {
      cplx *St = S + (size_t)t * n, *Ct = C + (size_t)t * n;
      if (unitary) {
        gemm(D, St, 1, Ct, 0, store);                                      /* :271 / :290 */
        gemm(D, Ct, 1, Bc, 0, t1); gemm(D, t1, 0, St, 0, t2);
        cplx f = (variant == 0 ? I : -I) * dt;
        g[(size_t)t * K + c] = 2.0 * creal(f * trace(D, t2) * trace(D, store));
      } else {
        gemm(D, Bc, 0, St, 0, t1); gemm(D, St, 0, Bc, 0, t2);              /* commutator, tools.jl:17-19 */
        for (size_t e = 0; e < n; e++) t1[e] -= t2[e];
        gemm(D, Ct, 1, t1, 0, store);                                      /* :285 */
        g[(size_t)t * K + c] = creal(I * dt * trace(D, store));            /* :286 */
      }
    }
  }
  cplx *St = S + (size_t)(N - 1) * n, *Ct = C + (size_t)(N - 1) * n;        /* GRAPE.jl:77,94: t = n_timeslices */
  if (unitary) { gemm(D, St, 1, Ct, 0, store); cplx tau = trace(D, store); return creal(tau * tau); }   /* cost_functions.jl:99-101 */
  gemm(D, Ct, 1, St, 0, store);                                            /* C1, cost_functions.jl:13-17 */
  cplx tau = trace(D, store) / D;
  return 1.0 - creal(tau * conj(tau));
}

size_t qoc_oracle_workspace_elems(int D, int N) { return ((size_t)N + 2 * ((size_t)N + 1) + 4 + 6) * D * D; }

/* Ensemble evaluation.  A [M][D*D], B [M][K][D*D], Xi/Xt [M][D*D], wts [M], x [N][K]; F scalar out, G [N][K].
 * nthreads <= 1 reproduces the reference's serial member loop.  Returns 0, or 1 on allocation failure. */
int qoc_oracle_eval(int sys, int D, int K, int N, int M, double T, const double* A_, const double* B_,
                    const double* Xi_, const double* Xt_, const double* wts, const double* x, int variant,
                    int nthreads, double* F, double* G) {
  const cplx *A = (const cplx*)A_, *B = (const cplx*)B_, *Xi = (const cplx*)Xi_, *Xt = (const cplx*)Xt_;
  const size_t n = (size_t)D * D, NK = (size_t)N * K;
  double* foms = (double*)malloc(sizeof(double) * M);
  double* grads = (double*)malloc(sizeof(double) * M * NK);
  if (!foms || !grads) { free(foms); free(grads); return 1; }
  int fail = 0;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    cplx* ws = (cplx*)malloc(sizeof(cplx) * qoc_oracle_workspace_elems(D, N));
    if (!ws) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      fail = 1;
    } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
      for (int k = 0; k < M; k++)
        foms[k] = eval_member(sys, D, K, N, T, A + k * n, B + (size_t)k * K * n, Xi + k * n, Xt + k * n, x, variant,
                              grads + (size_t)k * NK, ws);
      free(ws);
    }
  }
  if (!fail) {
    double f = 0;                                                         /* solve.jl:171-186, k ascending */
    for (int k = 0; k < M; k++) f += foms[k] * wts[k];
    *F = f;
    if (G) for (size_t e = 0; e < NK; e++) { double s = 0; for (int k = 0; k < M; k++) s += grads[(size_t)k * NK + e] * wts[k]; G[e] = s; }   /* solve.jl:191 */
  }
  free(foms); free(grads);
  return fail;
}

/* single matrix exponential, for cross-checking the Pade restatement; returns the product count */
int qoc_oracle_expm(int D, double* X_) {
  cplx* work = (cplx*)malloc(sizeof(cplx) * 6 * (size_t)D * D);
  if (!work) return -1;
  int p = expm_pade(D, (cplx*)X_, work);
  free(work);
  return p;
}

int qoc_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
