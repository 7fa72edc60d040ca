// libqocgrape.so — host orchestration + C ABI (include/qocgrape.h).  No CPU fallback anywhere: every
// entry point either runs the sm_100a kernels or returns an error.
#include "../../include/qocgrape.h"
#include "small_d.cuh"
#include "small_phased.cuh"
#include "big_d.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

using namespace qoc;

static thread_local std::string g_create_error;

struct qoc_handle {
  qoc_desc d{};
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  static constexpr int KRING = 64;           // event pairs around the dominant kernel of recent evaluations
  cudaEvent_t ek0[KRING] = {}, ek1[KRING] = {};
  int kring_count = 0;
  // derived small-path geometry
  int path = 0, NB = 1, CPW = 1, pack_mode = 0, n_groups = 0, n_inner = 0, n_sysgroups = 0, nmat = 0;
  int have_P = 0, sys_in_smem = 0, smem_bytes = 0, tb_bytes = 0, herm = 0, phased = 0, Cn = 1;
  double2 *storeP2 = nullptr, *stS = nullptr, *stC = nullptr, *totT = nullptr, *totTt = nullptr;
  double* tau = nullptr;
  int chunked = 0;                       // chunk-parallel fused mode (150..1500 chains)
  bool big_reuse = false;                // qoc_eval_continue: big_eval continues from the propagators of qoc_total_propagator
  int chunked_closed = 0;                // >= 1500 chains: chunking (Cn = 4) only pays with the closed-system recursion
  int unitary_fast = 1;                  // closed-system conjugation kernel when the problem is Hermitian (QOC_UNITARY_FAST=0 disables)
  double2 *bS = nullptr, *bC = nullptr;
  int NK = 0, red_chunk = 0, red_nchunks = 0;
  bool system_set = false;
  // device buffers
  double2 *sys = nullptr, *xi = nullptr, *xt = nullptr, *ident = nullptr, *storeP = nullptr, *storeS = nullptr;
  double *wts = nullptr, *x = nullptr, *fomc = nullptr, *gradc = nullptr, *part = nullptr, *out = nullptr;
  double2* staging = nullptr;   // raw caller matrices on device (set_system) / outputs of propagator calls
  size_t staging_bytes = 0;
  double *hx = nullptr, *hout = nullptr;   // pinned host staging
  long long ws_bytes = 0;
  qoc_stats st{};
  BigState* big = nullptr;
  // CUDA-graph replay of the whole qoc_eval sequence (small path): [0] value only, [1] value + gradient
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int graph_launches[2] = {0, 0};
  bool use_graph = true, in_capture = false;
  // one-shot NVLink all-reduce (qoc_comm_*)
  char* comm_local = nullptr;            // this rank's exchange buffer: data[2][n] doubles, then flags[2][QOC_MAX_RANKS]
  char* comm_peer_host[QOC_MAX_RANKS] = {};
  char** comm_peers = nullptr;           // device array of the peers' base pointers
  int comm_world = 0, comm_rank = -1;
  unsigned long long comm_epoch = 0;
  size_t comm_n = 0;
};

#define QOC_CUDA(h, call)                                                                           \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
      return e_ == cudaErrorMemoryAllocation ? QOC_ENOMEM : QOC_ECUDA;                              \
    }                                                                                               \
  } while (0)

template <class T> static int dev_alloc(qoc_handle* h, T** p, size_t n) {
  if (n == 0) n = 1;
  QOC_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
  h->ws_bytes += (long long)(n * sizeof(T));
  return QOC_OK;
}

extern "C" const char* qoc_version(void) { return "qocgrape-b200 0.1 (sm_100a, DMMA)"; }

extern "C" const char* qoc_last_error(qoc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// ------------------------------------------------------------------------------------------------ dispatch
typedef void (*chain_fn)(const SmallParams);
template <int NB, int CPW> static chain_fn pick_chain2(int sys, int grad) {
  if (sys == SYS_UNITARY) {
    if (grad == GRAD_NONE) return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_NONE>;
    if (grad == GRAD_FIRST) return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_FIRST>;
    return chain_kernel<NB, CPW, SYS_UNITARY, GRAD_EXACT>;
  }
  if (grad == GRAD_NONE) return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_NONE>;
  if (grad == GRAD_FIRST) return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_FIRST>;
  return chain_kernel<NB, CPW, SYS_DENSITY, GRAD_EXACT>;
}
static chain_fn pick_chain(int NB, int CPW, int sys, int grad) {
  if (NB == 2) return pick_chain2<2, 1>(sys, grad);
  if (CPW == 4) return pick_chain2<1, 4>(sys, grad);
  if (CPW == 2) return pick_chain2<1, 2>(sys, grad);
  return pick_chain2<1, 1>(sys, grad);
}
typedef void (*slice_fn)(const SliceParams);
static slice_fn pick_slices(int NB, int CPW) {
  if (NB == 2) return expm_slices_kernel<2, 1>;
  if (CPW == 4) return expm_slices_kernel<1, 4>;
  if (CPW == 2) return expm_slices_kernel<1, 2>;
  return expm_slices_kernel<1, 1>;
}

static int launch_check(qoc_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string(what) + ": " + cudaGetErrorString(e); return QOC_ECUDA; }
  h->st.n_launches++; h->st.launches_last_eval++;
  return QOC_OK;
}

// ------------------------------------------------------------------------------------------------ create
extern "C" int qoc_create(qoc_handle** out, const qoc_desc* desc) {
  if (!out || !desc) { g_create_error = "qoc_create: null argument"; return QOC_EINVAL; }
  *out = nullptr;
  const qoc_desc& d = *desc;
  if (d.D < 1 || d.K < 0 || d.N < 1 || d.M < 1 || d.R < 1 || !(d.T == d.T) ||
      d.sys_type < 0 || d.sys_type > 2 || d.gradient < 0 || d.gradient > 1 || d.convention < 0 || d.convention > 1) {
    g_create_error = "qoc_create: invalid descriptor"; return QOC_EINVAL;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("qoc_create: no CUDA device (") + cudaGetErrorString(ce) + "); there is no CPU fallback";
    return QOC_ECUDA;
  }
  if (d.device < 0 || d.device >= ndev) { g_create_error = "qoc_create: bad device ordinal"; return QOC_EINVAL; }
  qoc_handle* h = new qoc_handle();
  h->d = d;
  if (h->d.expm_theta <= 0) h->d.expm_theta = T8_THETA_DEFAULT;
  h->NK = d.N * d.K;
  auto fail = [&](int rc) { g_create_error = h->err; qoc_destroy(h); return rc; };
#define CR(call) do { int rc_ = (call); if (rc_ != QOC_OK) return fail(rc_); } while (0)
#define CRC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return fail(e_ == cudaErrorMemoryAllocation ? QOC_ENOMEM : QOC_ECUDA); } } while (0)
  CRC(cudaSetDevice(d.device));
  CRC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CRC(cudaEventCreate(&h->ev0));
  CRC(cudaEventCreate(&h->ev1));
  for (int i = 0; i < qoc_handle::KRING; i++) { CRC(cudaEventCreate(&h->ek0[i])); CRC(cudaEventCreate(&h->ek1[i])); }

  const size_t DD = (size_t)d.D * d.D;
  if (d.D <= 16) {
    h->path = 1;
    if (d.D <= 2) { h->NB = 1; h->CPW = 4; }
    else if (d.D <= 4) { h->NB = 1; h->CPW = 2; }
    else if (d.D <= 8) { h->NB = 1; h->CPW = 1; }
    else { h->NB = 2; h->CPW = 1; }
    h->pack_mode = (d.M < h->CPW && d.R > 1) ? 1 : 0;
    if (h->pack_mode == 0) { h->n_inner = (d.M + h->CPW - 1) / h->CPW; h->n_groups = d.R * h->n_inner; h->n_sysgroups = h->n_inner; }
    else { h->n_inner = d.M; h->n_groups = ((d.R + h->CPW - 1) / h->CPW) * d.M; h->n_sysgroups = d.M; }
    h->nmat = 1 + d.K + (d.gradient == QOC_GRAD_EXACT ? d.K : 0);
    // Execution strategy (thresholds measured on cfg4 shards, profiles/README.md):
    //   >= 1500 chains        fused chain_kernel, one warp per chain
    //   150..1499 chains      chunk-parallel fused: each chain split into Cn chunks (full occupancy)
    //   < 150 chains          fully slice-parallel pipeline with a chunked prefix scan (small_phased.cuh)
    // Short pulses stay on the plain fused kernel.
    if (const char* e = getenv("QOC_UNITARY_FAST")) h->unitary_fast = atoi(e) != 0;
    if (const char* e = getenv("QOC_GRAPH")) h->use_graph = atoi(e) != 0;
    h->phased = h->n_groups < 150 && d.N >= 32;
    h->chunked = !h->phased && h->n_groups < 1500 && d.N >= 64;
    if (const char* e = getenv("QOC_PHASED")) { h->phased = atoi(e) != 0; if (h->phased) h->chunked = 0; }
    if (const char* e = getenv("QOC_CHUNKED")) { h->chunked = atoi(e) != 0; if (h->chunked) h->phased = 0; }
    h->have_P = h->phased;               // slice-parallel exponentials feed the value-only path as well
    if (const char* e = getenv("QOC_HAVE_P")) h->have_P = atoi(e) != 0;   // tuning override
    if (h->chunked) {
      h->have_P = 0;
      h->Cn = std::max(2, std::min((8192 + h->n_groups - 1) / h->n_groups, std::max(2, d.N / 16)));   // ~8192 warps measured best
      if (const char* e = getenv("QOC_CHUNKS")) h->Cn = std::max(2, std::min(atoi(e), d.N));
      h->Cn = std::min(h->Cn, d.N / 2);              // every chunk needs at least two slices
      if (h->Cn < 2) { h->chunked = 0; h->Cn = 1; }
    }
    // with the closed-system recursion (6 products per slice in either mode) 4 chunks per chain still win at full batch:
    // cfg4 4096 chains 2.44 ms vs 2.52 ms fused, 2048 chains 1.26 vs 1.50 ms
    if (!h->phased && !h->chunked && h->n_groups >= 1500 && d.N >= 64 && d.gradient == QOC_GRAD_FIRST_ORDER && !getenv("QOC_CHUNKED")) {
      h->chunked_closed = 1; h->Cn = std::min(8, std::max(2, d.N / 16));   // measured 2 / 3 / 4 / 6 / 8 chunks at 4096 chains: 2.452 / 2.421 / 2.408 / 2.400 / 2.398 ms
      if (const char* e = getenv("QOC_CHUNKS")) h->Cn = std::max(2, std::min(atoi(e), d.N / 2));   // tuning override
    }
    if (h->phased) {
      h->have_P = 1;
      int want = (592 + 2 * h->n_groups - 1) / (2 * h->n_groups);        // sweep warps >= one per SM sub-partition
      int cap = (int)std::sqrt((double)d.N);                              // depth L + Cn is minimal near sqrt(N)
      h->Cn = std::max(1, std::min(std::min(want, cap), d.N / 2 > 0 ? d.N / 2 : 1));
      if (const char* e = getenv("QOC_CHUNKS")) h->Cn = std::max(1, std::min(atoi(e), d.N));
    }
    const size_t E = (size_t)h->NB * h->NB * 64;
    h->tb_bytes = 4 * h->NB * h->NB * 2 * TB_PLANE * (int)sizeof(double);     // per-warp transpose tiles
    size_t smem = (size_t)4 * h->nmat * E * sizeof(double2);
    h->sys_in_smem = smem + h->tb_bytes <= 96 * 1024;
    h->smem_bytes = h->tb_bytes + (h->sys_in_smem ? (int)smem : 0);
    CR(dev_alloc(h, &h->sys, (size_t)h->n_sysgroups * h->nmat * E));
    CR(dev_alloc(h, &h->xi, (size_t)h->n_sysgroups * E));
    CR(dev_alloc(h, &h->xt, (size_t)h->n_sysgroups * E));
    CR(dev_alloc(h, &h->ident, (size_t)h->n_sysgroups * E));
    CR(dev_alloc(h, &h->storeP, (size_t)h->n_groups * d.N * E));
    if (!h->phased) CR(dev_alloc(h, &h->storeS, (size_t)h->n_groups * d.N * E));
    if (h->chunked || h->chunked_closed) {
      CR(dev_alloc(h, &h->totT, (size_t)h->n_groups * h->Cn * E));
      CR(dev_alloc(h, &h->totTt, (size_t)h->n_groups * h->Cn * E));
      CR(dev_alloc(h, &h->bS, (size_t)h->n_groups * (h->Cn + 1) * E));
      CR(dev_alloc(h, &h->bC, (size_t)h->n_groups * (h->Cn + 1) * E));
      CR(dev_alloc(h, &h->tau, (size_t)h->n_groups * h->CPW * 2));
    }
    if (h->phased) {
      CR(dev_alloc(h, &h->storeP2, (size_t)h->n_groups * d.N * E));
      CR(dev_alloc(h, &h->stS, (size_t)h->n_groups * (d.N + 1) * E));
      CR(dev_alloc(h, &h->stC, (size_t)h->n_groups * (d.N + 1) * E));
      CR(dev_alloc(h, &h->totT, (size_t)h->n_groups * h->Cn * E));
      CR(dev_alloc(h, &h->totTt, (size_t)h->n_groups * h->Cn * E));
      CR(dev_alloc(h, &h->tau, (size_t)h->n_groups * h->CPW * 2));
    }
    CR(dev_alloc(h, &h->fomc, (size_t)d.R * d.M));
    CR(dev_alloc(h, &h->gradc, (size_t)d.R * d.M * h->NK));
  } else {
    h->path = 2;
    int rc = big_create(&h->big, d, h->err, h->ws_bytes);
    if (rc != QOC_OK) return fail(rc);
  }
  h->red_chunk = std::max(4, (d.M + RED_MAX_CHUNKS - 1) / RED_MAX_CHUNKS);
  h->red_nchunks = (d.M + h->red_chunk - 1) / h->red_chunk;
  CR(dev_alloc(h, &h->wts, (size_t)d.M));
  CR(dev_alloc(h, &h->x, (size_t)d.R * h->NK));
  CR(dev_alloc(h, &h->part, (size_t)d.R * h->red_nchunks * (h->NK + 1)));
  CR(dev_alloc(h, &h->out, (size_t)d.R * (h->NK + 1)));
  h->staging_bytes = (size_t)d.M * (size_t)(d.K > 1 ? d.K : 1) * DD * sizeof(double2);
  CR(dev_alloc(h, (char**)&h->staging, h->staging_bytes));
  CRC(cudaMallocHost((void**)&h->hx, (size_t)d.R * (h->NK > 0 ? h->NK : 1) * sizeof(double)));
  CRC(cudaMallocHost((void**)&h->hout, (size_t)d.R * (h->NK + 1) * sizeof(double)));
#undef CR
#undef CRC
  h->st.path = h->path;
  h->st.workspace_bytes = h->ws_bytes;
  *out = h;
  return QOC_OK;
}

extern "C" int qoc_destroy(qoc_handle* h) {
  if (!h) return QOC_OK;
  cudaSetDevice(h->d.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->big) big_destroy(h->big);
  for (auto& g : h->graph_exec) if (g) cudaGraphExecDestroy(g);
  for (int r = 0; r < h->comm_world; r++) if (r != h->comm_rank && h->comm_peer_host[r]) cudaIpcCloseMemHandle(h->comm_peer_host[r]);
  if (h->comm_local) cudaFree(h->comm_local);
  if (h->comm_peers) cudaFree(h->comm_peers);
  void* bufs[] = {h->bS, h->bC, h->storeP2, h->stS, h->stC, h->totT, h->totTt, h->tau, h->sys, h->xi, h->xt, h->ident, h->storeP, h->storeS, h->wts, h->x, h->fomc, h->gradc, h->part, h->out, h->staging};
  for (void* b : bufs) if (b) cudaFree(b);
  if (h->hx) cudaFreeHost(h->hx);
  if (h->hout) cudaFreeHost(h->hout);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < qoc_handle::KRING; i++) { if (h->ek0[i]) cudaEventDestroy(h->ek0[i]); if (h->ek1[i]) cudaEventDestroy(h->ek1[i]); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QOC_OK;
}

// ------------------------------------------------------------------------------------------------ set_system
static int pack_one(qoc_handle* h, const double2* src_dev, int n_src, long src_stride, int transpose,
                    double2* dst, int nmat_dst, int mat_dst, int pack_mode, double sre = 1.0, double sim = 0.0) {
  PackParams pp;
  pp.scale_re = sre; pp.scale_im = sim;
  pp.D = h->d.D; pp.NB = h->NB; pp.CPW = h->CPW; pp.n_og = h->n_sysgroups; pp.nmat_dst = nmat_dst; pp.mat_dst = mat_dst;
  pp.transpose = transpose; pp.pack_mode = pack_mode; pp.n_src = n_src; pp.src_stride = src_stride; pp.src = src_dev; pp.dst = dst;
  long total = (long)h->n_sysgroups * h->NB * h->NB * 64;
  pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(pp);
  return launch_check(h, "pack_kernel");
}

extern "C" int qoc_set_system(qoc_handle* h, const double* A, const double* B, const double* Xi, const double* Xt,
                              const double* wts, int shared_flags) {
  if (!h) return QOC_EINVAL;
  if (!A || !Xi || !Xt || (h->d.K > 0 && !B)) { h->err = "qoc_set_system: null matrix argument"; return QOC_EINVAL; }
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  const qoc_desc& d = h->d;
  const size_t DD = (size_t)d.D * d.D;
  std::vector<double> w(d.M, 1.0);
  if (wts) memcpy(w.data(), wts, sizeof(double) * d.M);
  QOC_CUDA(h, cudaMemcpyAsync(h->wts, w.data(), sizeof(double) * d.M, cudaMemcpyHostToDevice, h->stream));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->path == 2) {
    int rc = big_set_system(h->big, A, B, Xi, Xt, shared_flags, h->err);
    if (rc == QOC_OK) { h->system_set = true; h->st.path = h->big->pure.active ? 3 : 2; }
    return rc;
  }
  // small path: stage raw matrices on the device, then pack into the warp layout
  const int member_mode = h->pack_mode;   // 0: slot -> member og*CPW+s ; 1: all slots -> member og
  const double dt = d.T / d.N;              // A, B are packed pre-multiplied by -i*dt
  const int tr_states = d.sys_type == QOC_UNITARY_GATE ? 1 : 0;   // the unitary chain runs on S^T, C^T
  {  // Hermitian drift and controls  =>  anti-Hermitian generators (exact elementwise test)
    auto is_herm = [&](const double* Mx) {
      for (int c = 0; c < d.D; c++)
        for (int r = 0; r <= c; r++) {
          const double* a = Mx + 2 * ((size_t)c * d.D + r); const double* b = Mx + 2 * ((size_t)r * d.D + c);
          if (a[0] != b[0] || a[1] != -b[1]) return false;
        }
      return true;
    };
    bool herm = true;
    const int nA = (shared_flags & QOC_SHARED_A) ? 1 : d.M, nB = (shared_flags & QOC_SHARED_B) ? 1 : d.M;
    for (int k = 0; k < nA && herm; k++) herm = is_herm(A + 2 * (size_t)k * DD);
    for (int k = 0; k < nB * d.K && herm; k++) herm = is_herm(B + 2 * (size_t)k * DD);
    h->herm = herm ? 1 : 0;
  }
  auto upload = [&](const double* src, size_t count) -> int {
    QOC_CUDA(h, cudaMemcpyAsync(h->staging, src, count * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
    return QOC_OK;
  };
  int rc;
  {  // A
    bool sh = shared_flags & QOC_SHARED_A;
    if ((rc = upload(A, (sh ? 1 : (size_t)d.M) * DD)) != QOC_OK) return rc;
    if ((rc = pack_one(h, h->staging, sh ? 1 : d.M, sh ? 0 : (long)DD, 0, h->sys, h->nmat, 0, member_mode, 0.0, -dt)) != QOC_OK) return rc;
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  if (d.K > 0) {  // B (and transposed copies for the exact gradient)
    bool sh = shared_flags & QOC_SHARED_B;
    if ((rc = upload(B, (sh ? 1 : (size_t)d.M) * d.K * DD)) != QOC_OK) return rc;
    for (int c = 0; c < d.K; c++) {
      if ((rc = pack_one(h, h->staging + (size_t)c * DD, sh ? 1 : d.M, sh ? 0 : (long)(d.K * DD), 0, h->sys, h->nmat, 1 + c, member_mode, 0.0, -dt)) != QOC_OK) return rc;
      if (d.gradient == QOC_GRAD_EXACT)
        if ((rc = pack_one(h, h->staging + (size_t)c * DD, sh ? 1 : d.M, sh ? 0 : (long)(d.K * DD), 1, h->sys, h->nmat, 1 + d.K + c, member_mode, 0.0, -dt)) != QOC_OK) return rc;
    }
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  {  // Xi, Xt
    bool sh = shared_flags & QOC_SHARED_XI;
    if ((rc = upload(Xi, (sh ? 1 : (size_t)d.M) * DD)) != QOC_OK) return rc;
    if ((rc = pack_one(h, h->staging, sh ? 1 : d.M, sh ? 0 : (long)DD, tr_states, h->xi, 1, 0, member_mode)) != QOC_OK) return rc;
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
    sh = shared_flags & QOC_SHARED_XT;
    if ((rc = upload(Xt, (sh ? 1 : (size_t)d.M) * DD)) != QOC_OK) return rc;
    if ((rc = pack_one(h, h->staging, sh ? 1 : d.M, sh ? 0 : (long)DD, tr_states, h->xt, 1, 0, member_mode)) != QOC_OK) return rc;
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  {  // identity (initial state of qoc_total_propagator)
    std::vector<double> I(2 * DD, 0.0);
    for (int i = 0; i < d.D; i++) I[2 * ((size_t)i * d.D + i)] = 1.0;
    if ((rc = upload(I.data(), DD)) != QOC_OK) return rc;
    if ((rc = pack_one(h, h->staging, 1, 0, 0, h->ident, 1, 0, member_mode)) != QOC_OK) return rc;
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->system_set = true;
  return QOC_OK;
}

// ------------------------------------------------------------------------------------------------ evaluation
static SmallParams small_params(qoc_handle* h, const double* x_dev) {
  const qoc_desc& d = h->d;
  SmallParams p;
  p.D = d.D; p.N = d.N; p.K = d.K; p.M = d.M; p.R = d.R;
  p.pack_mode = h->pack_mode; p.n_groups = h->n_groups; p.n_inner = h->n_inner; p.nmat = h->nmat;
  p.sys_in_smem = h->sys_in_smem; p.have_P = h->have_P;
  p.sign_static = d.convention == QOC_REF_STATIC ? -1 : 1;
  p.fom_exact = d.gradient == QOC_GRAD_EXACT;
  p.herm = h->herm;
  p.dt = d.T / d.N; p.theta = d.expm_theta;
  p.sys = h->sys; p.xi = h->xi; p.xt = h->xt; p.x = x_dev;
  p.storeP = h->storeP; p.storeS = h->storeS; p.fomc = h->fomc; p.gradc = h->gradc; p.out_final = nullptr;
  p.Cn = 1; p.bS = nullptr; p.bC = nullptr; p.tau_in = nullptr; p.ident = h->ident;
  return p;
}
static SliceParams slice_params(qoc_handle* h, const double* x_dev) {
  const qoc_desc& d = h->d;
  SliceParams s;
  s.D = d.D; s.N = d.N; s.K = d.K; s.M = d.M; s.R = d.R; s.pack_mode = h->pack_mode; s.n_groups = h->n_groups;
  s.n_inner = h->n_inner; s.nmat = h->nmat; s.herm = h->herm; s.dt = d.T / d.N; s.theta = d.expm_theta;
  s.sys = h->sys; s.x = x_dev; s.storeP = nullptr; s.storeP2 = nullptr; s.out_user = nullptr; s.mode = 0;
  return s;
}
static int launch_chain(qoc_handle* h, const SmallParams& p, int sys, int grad, cudaStream_t st) {
  chain_fn fn = pick_chain(h->NB, h->CPW, sys, grad);
  if (h->smem_bytes > 48 * 1024)
    QOC_CUDA(h, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  fn<<<(unsigned)(((long)h->n_groups * (p.Cn > 1 ? p.Cn : 1) + 3) / 4), 128, h->smem_bytes, st>>>(p);
  return launch_check(h, "chain_kernel");
}
// closed-system kernel (Hermitian drift and controls, first-order gradient, fused mode)
static int launch_chain_unitary(qoc_handle* h, const SmallParams& p, int sys, cudaStream_t st) {
  chain_fn fn;
  const bool u = sys == SYS_UNITARY;
  if (h->NB == 2) fn = u ? chain_unitary_kernel<2, 1, SYS_UNITARY> : chain_unitary_kernel<2, 1, SYS_DENSITY>;
  else if (h->CPW == 4) fn = u ? chain_unitary_kernel<1, 4, SYS_UNITARY> : chain_unitary_kernel<1, 4, SYS_DENSITY>;
  else if (h->CPW == 2) fn = u ? chain_unitary_kernel<1, 2, SYS_UNITARY> : chain_unitary_kernel<1, 2, SYS_DENSITY>;
  else fn = u ? chain_unitary_kernel<1, 1, SYS_UNITARY> : chain_unitary_kernel<1, 1, SYS_DENSITY>;
  if (h->smem_bytes > 48 * 1024)
    QOC_CUDA(h, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  fn<<<(unsigned)((h->n_groups + 3) / 4), 128, h->smem_bytes, st>>>(p);
  return launch_check(h, "chain_unitary_kernel");
}
static int launch_slices(qoc_handle* h, const SliceParams& s, cudaStream_t st) {
  long warps = (long)h->n_groups * h->d.N;
  pick_slices(h->NB, h->CPW)<<<(unsigned)((warps + 3) / 4), 128, h->tb_bytes, st>>>(s);
  return launch_check(h, "expm_slices_kernel");
}

static PhasedParams phased_params(qoc_handle* h, const double* x_dev) {
  const qoc_desc& d = h->d;
  PhasedParams p;
  p.D = d.D; p.N = d.N; p.K = d.K; p.M = d.M; p.R = d.R; p.pack_mode = h->pack_mode; p.n_groups = h->n_groups; p.n_inner = h->n_inner;
  p.nmat = h->nmat; p.herm = h->herm; p.Cn = h->Cn; p.sign_static = d.convention == QOC_REF_STATIC ? -1 : 1;
  p.fom_exact = d.gradient == QOC_GRAD_EXACT; p.theta = d.expm_theta;
  p.sys = h->sys; p.xi = h->xi; p.xt = h->xt; p.x = x_dev; p.storePt = h->storeP; p.storeP = h->storeP2; p.stS = h->stS; p.stC = h->stC;
  p.totT = h->totT; p.totTt = h->totTt; p.tau = h->tau; p.fomc = h->fomc; p.gradc = h->gradc;
  p.bS = h->bS; p.bC = h->bC; p.sys_in_smem = 0; p.store_plain = 0;
  return p;
}

// dispatch of the phased-pipeline kernels
typedef void (*phased_fn)(const PhasedParams);
template <int NB, int CPW> static void pick_phased(int sys, int grad, phased_fn& tot, phased_fn& bnd, phased_fn& swp, phased_fn& grd) {
  tot = chunk_totals_kernel<NB, CPW>;
  if (sys == SYS_UNITARY) {
    bnd = boundary_kernel<NB, CPW, SYS_UNITARY>; swp = sweep_kernel<NB, CPW, SYS_UNITARY>;
    grd = grad == GRAD_EXACT ? grad_slices_kernel<NB, CPW, SYS_UNITARY, GRAD_EXACT> : grad_slices_kernel<NB, CPW, SYS_UNITARY, GRAD_FIRST>;
  } else {
    bnd = boundary_kernel<NB, CPW, SYS_DENSITY>; swp = sweep_kernel<NB, CPW, SYS_DENSITY>;
    grd = grad == GRAD_EXACT ? grad_slices_kernel<NB, CPW, SYS_DENSITY, GRAD_EXACT> : grad_slices_kernel<NB, CPW, SYS_DENSITY, GRAD_FIRST>;
  }
}

static int eval_phased(qoc_handle* h, const double* x_dev, int sys, int grad, cudaStream_t st) {
  const qoc_desc& d = h->d;
  int rc;
  SliceParams s = slice_params(h, x_dev);
  s.storeP = h->storeP; s.storeP2 = h->storeP2;
  if ((rc = launch_slices(h, s, st)) != QOC_OK) return rc;
  PhasedParams p = phased_params(h, x_dev);
  phased_fn tot, bnd, swp, grd;
  if (h->NB == 2) pick_phased<2, 1>(sys, grad, tot, bnd, swp, grd);
  else if (h->CPW == 4) pick_phased<1, 4>(sys, grad, tot, bnd, swp, grd);
  else if (h->CPW == 2) pick_phased<1, 2>(sys, grad, tot, bnd, swp, grd);
  else pick_phased<1, 1>(sys, grad, tot, bnd, swp, grd);
  auto blocks = [](long warps) { return (unsigned)((warps + 3) / 4); };
  if (h->Cn > 1) {
    tot<<<blocks((long)h->n_groups * h->Cn), 128, h->tb_bytes, st>>>(p);
    if ((rc = launch_check(h, "chunk_totals_kernel")) != QOC_OK) return rc;
  }
  bnd<<<blocks((long)h->n_groups * 2), 128, 0, st>>>(p);
  if ((rc = launch_check(h, "boundary_kernel")) != QOC_OK) return rc;
  swp<<<blocks((long)h->n_groups * h->Cn * 2), 128, 0, st>>>(p);
  if ((rc = launch_check(h, "sweep_kernel")) != QOC_OK) return rc;
  grd<<<blocks((long)h->n_groups * d.N), 128, h->tb_bytes, st>>>(p);
  return launch_check(h, "grad_slices_kernel");
}

// chunk-parallel fused mode: exponentials + chunk totals, boundary states, then the fused body per chunk
static int eval_chunked(qoc_handle* h, SmallParams cp, const double* x_dev, int sys, int grad, cudaStream_t st) {
  int rc;
  PhasedParams p = phased_params(h, x_dev);
  const size_t E = (size_t)h->NB * h->NB * 64;
  const size_t sys_bytes = (size_t)4 * (1 + h->d.K) * E * sizeof(double2);
  p.sys_in_smem = sys_bytes + h->tb_bytes <= 96 * 1024;
  const int smem1 = h->tb_bytes + (p.sys_in_smem ? (int)sys_bytes : 0);
  typedef void (*kfn)(const PhasedParams);
  kfn k1, k2;
  if (h->NB == 2) { k1 = chunk_expm_kernel<2, 1>; k2 = sys == SYS_UNITARY ? boundary2_kernel<2, 1, SYS_UNITARY> : boundary2_kernel<2, 1, SYS_DENSITY>; }
  else if (h->CPW == 4) { k1 = chunk_expm_kernel<1, 4>; k2 = sys == SYS_UNITARY ? boundary2_kernel<1, 4, SYS_UNITARY> : boundary2_kernel<1, 4, SYS_DENSITY>; }
  else if (h->CPW == 2) { k1 = chunk_expm_kernel<1, 2>; k2 = sys == SYS_UNITARY ? boundary2_kernel<1, 2, SYS_UNITARY> : boundary2_kernel<1, 2, SYS_DENSITY>; }
  else { k1 = chunk_expm_kernel<1, 1>; k2 = sys == SYS_UNITARY ? boundary2_kernel<1, 1, SYS_UNITARY> : boundary2_kernel<1, 1, SYS_DENSITY>; }
  if (smem1 > 48 * 1024) QOC_CUDA(h, cudaFuncSetAttribute((const void*)k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
  const bool closed = grad == GRAD_FIRST && h->herm && h->unitary_fast;     // closed-system conjugation recursion
  p.store_plain = closed;
  k1<<<(unsigned)(((long)h->n_groups * h->Cn + 3) / 4), 128, smem1, st>>>(p);
  if ((rc = launch_check(h, "chunk_expm_kernel")) != QOC_OK) return rc;
  if (closed) {
    kfn kb, ks;
    const bool u = sys == SYS_UNITARY;
    if (h->NB == 2) { kb = u ? boundary_unitary_kernel<2, 1, SYS_UNITARY> : boundary_unitary_kernel<2, 1, SYS_DENSITY>; ks = sweep_unitary_kernel<2, 1>; }
    else if (h->CPW == 4) { kb = u ? boundary_unitary_kernel<1, 4, SYS_UNITARY> : boundary_unitary_kernel<1, 4, SYS_DENSITY>; ks = sweep_unitary_kernel<1, 4>; }
    else if (h->CPW == 2) { kb = u ? boundary_unitary_kernel<1, 2, SYS_UNITARY> : boundary_unitary_kernel<1, 2, SYS_DENSITY>; ks = sweep_unitary_kernel<1, 2>; }
    else { kb = u ? boundary_unitary_kernel<1, 1, SYS_UNITARY> : boundary_unitary_kernel<1, 1, SYS_DENSITY>; ks = sweep_unitary_kernel<1, 1>; }
    kb<<<(unsigned)((h->n_groups + 3) / 4), 128, h->tb_bytes, st>>>(p);
    if ((rc = launch_check(h, "boundary_unitary_kernel")) != QOC_OK) return rc;
    const size_t bbytes = (size_t)4 * h->d.K * E * sizeof(double2);
    p.sys_in_smem = bbytes <= 96 * 1024;
    const int smem3 = p.sys_in_smem ? (int)bbytes : 0;
    if (smem3 > 48 * 1024) QOC_CUDA(h, cudaFuncSetAttribute((const void*)ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    ks<<<(unsigned)(((long)h->n_groups * h->Cn + 3) / 4), 128, smem3, st>>>(p);
    return launch_check(h, "sweep_unitary_kernel");
  }
  k2<<<(unsigned)((h->n_groups * 2 + 3) / 4), 128, 0, st>>>(p);
  if ((rc = launch_check(h, "boundary2_kernel")) != QOC_OK) return rc;
  cp.Cn = h->Cn; cp.bS = h->bS; cp.bC = h->bC; cp.tau_in = h->tau; cp.have_P = 1;
  return launch_chain(h, cp, sys, grad, st);
}

static int eval_small(qoc_handle* h, const double* x_dev, double* fg_dev, int want_grad, cudaStream_t st) {
  const qoc_desc& d = h->d;
  int rc;
  SmallParams p = small_params(h, x_dev);
  const int sys = d.sys_type == QOC_UNITARY_GATE ? SYS_UNITARY : SYS_DENSITY;
  const int grad = !want_grad ? GRAD_NONE : (d.gradient == QOC_GRAD_EXACT ? GRAD_EXACT : GRAD_FIRST);
  const int slot = h->kring_count % qoc_handle::KRING;
  if (!h->in_capture) QOC_CUDA(h, cudaEventRecord(h->ek0[slot], st));
  if (h->phased && want_grad) {
    if ((rc = eval_phased(h, x_dev, sys, grad, st)) != QOC_OK) return rc;
  } else if ((h->chunked || (h->chunked_closed && h->herm && h->unitary_fast)) && want_grad) {
    if ((rc = eval_chunked(h, p, x_dev, sys, grad, st)) != QOC_OK) return rc;
  } else if (grad == GRAD_FIRST && h->herm && !h->have_P && h->unitary_fast) {
    if ((rc = launch_chain_unitary(h, p, sys, st)) != QOC_OK) return rc;
  } else {
    if (h->have_P) {
      SliceParams s = slice_params(h, x_dev);
      s.storeP = h->storeP;
      if ((rc = launch_slices(h, s, st)) != QOC_OK) return rc;
    }
    if ((rc = launch_chain(h, p, sys, grad, st)) != QOC_OK) return rc;
  }
  if (!h->in_capture) { QOC_CUDA(h, cudaEventRecord(h->ek1[slot], st)); h->kring_count++; }
  dim3 g1((unsigned)(((h->NK + 1 + 255) / 256) * (long)d.R), h->red_nchunks);
  reduce_members_pass1<<<g1, 256, 0, st>>>(want_grad ? h->gradc : nullptr, h->fomc, h->wts, h->red_nchunks == 1 ? fg_dev : h->part, d.M, h->NK,
                                           h->red_chunk, h->red_nchunks);
  if ((rc = launch_check(h, "reduce_members_pass1")) != QOC_OK) return rc;
  if (h->red_nchunks == 1) return QOC_OK;
  dim3 g2((unsigned)(((h->NK + 1 + 31) / 32) * (long)d.R));
  reduce_members_pass2<<<g2, 32 * RED_LANES, 0, st>>>(h->part, fg_dev, h->NK, h->red_nchunks);
  return launch_check(h, "reduce_members_pass2");
}

static int eval_device_on(qoc_handle* h, const double* x_dev, double* FG_dev, int want_gradient, cudaStream_t st) {
  h->st.launches_last_eval = 0;
  h->st.n_evals++;
  if (h->path == 2) return big_eval(h->big, x_dev, FG_dev, want_gradient, h->wts, st, h->err, h->st, h->big_reuse);
  return eval_small(h, x_dev, FG_dev, want_gradient, st);
}

extern "C" int qoc_eval_device(qoc_handle* h, const double* x_dev, double* FG_dev, int want_gradient, void* stream) {
  if (!h) return QOC_EINVAL;
  if (!x_dev || !FG_dev) { h->err = "qoc_eval_device: null pointer"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_eval_device: qoc_set_system has not been called"; return QOC_EINVAL; }
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  return eval_device_on(h, x_dev, FG_dev, want_gradient, (cudaStream_t)stream);   // NULL = CUDA default stream
}

// D2H of the result rows: everything when the gradient was asked for, else only the F column
static cudaError_t copy_result_async(qoc_handle* h, bool grad) {
  const size_t row = (size_t)h->NK + 1;
  if (grad) return cudaMemcpyAsync(h->hout, h->out, (size_t)h->d.R * row * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  return cudaMemcpy2DAsync(h->hout, row * sizeof(double), h->out, row * sizeof(double), sizeof(double), h->d.R, cudaMemcpyDeviceToHost, h->stream);
}

// Captures H2D(x) -> kernels -> D2H([F|G]) once per variant and replays it afterwards: one graph launch per
// evaluation.  Everything in the sequence is static (pinned staging buffers, device buffers, launch geometry).
static bool eval_via_graph(qoc_handle* h, bool grad) {
  const int gi = grad ? 1 : 0;
  const size_t nx = (size_t)h->d.R * h->NK;
  if (!h->graph_exec[gi]) {
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    h->in_capture = true;
    const long long before = h->st.n_launches;
    bool ok = cudaMemcpyAsync(h->x, h->hx, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
    ok = ok && eval_small(h, h->x, h->out, grad, h->stream) == QOC_OK;
    ok = ok && copy_result_async(h, grad) == cudaSuccess;
    h->in_capture = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    h->graph_launches[gi] = (int)(h->st.n_launches - before);
    h->st.n_launches = before;
    if (!ok || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return false; }
    e = cudaGraphInstantiate(&h->graph_exec[gi], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { h->graph_exec[gi] = nullptr; cudaGetLastError(); return false; }
  }
  if (cudaGraphLaunch(h->graph_exec[gi], h->stream) != cudaSuccess) { cudaGetLastError(); return false; }
  h->st.n_launches += h->graph_launches[gi];
  h->st.launches_last_eval = h->graph_launches[gi];
  return cudaStreamSynchronize(h->stream) == cudaSuccess;
}

extern "C" int qoc_eval(qoc_handle* h, const double* x, double* F, double* G) {
  if (!h) return QOC_EINVAL;
  if (!x) { h->err = "qoc_eval: null pulse"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_eval: qoc_set_system has not been called"; return QOC_EINVAL; }
  const qoc_desc& d = h->d;
  QOC_CUDA(h, cudaSetDevice(d.device));
  const size_t nx = (size_t)d.R * h->NK;
  memcpy(h->hx, x, nx * sizeof(double));
  const size_t row = (size_t)h->NK + 1;
  bool done = false;
  if (h->path == 1 && h->use_graph) {
    h->st.n_evals++;
    done = eval_via_graph(h, G != nullptr);
    if (!done) { h->use_graph = false; h->st.n_evals--; h->in_capture = false; }   // capture unavailable: plain launches from now on
    else h->st.gpu_ms_last_eval = 0.f;
  }
  if (!done) {
    QOC_CUDA(h, cudaMemcpyAsync(h->x, h->hx, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    QOC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    int rc = eval_device_on(h, h->x, h->out, G != nullptr, h->stream);
    if (rc != QOC_OK) return rc;
    QOC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    QOC_CUDA(h, copy_result_async(h, G != nullptr));
    QOC_CUDA(h, cudaStreamSynchronize(h->stream));
    QOC_CUDA(h, cudaEventElapsedTime(&h->st.gpu_ms_last_eval, h->ev0, h->ev1));
  }
  for (int r = 0; r < d.R; r++) {
    if (F) F[r] = h->hout[r * row];
    if (G) memcpy(G + (size_t)r * h->NK, h->hout + r * row + 1, sizeof(double) * h->NK);
  }
  return QOC_OK;
}

// Slice-parallel use (one rank per range of slices): new boundary operators per evaluation, drift and controls stay.
extern "C" int qoc_set_states(qoc_handle* h, const double* Xi, const double* Xt, int shared_flags) {
  if (!h) return QOC_EINVAL;
  if (!Xi || !Xt) { h->err = "qoc_set_states: null pointer"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_set_states: qoc_set_system has not been called"; return QOC_EINVAL; }
  if (h->path != 2) { h->err = "qoc_set_states: only implemented for D > 16 (tiled GEMM path)"; return QOC_EUNSUPPORTED; }
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  return big_set_states(h->big, Xi, Xt, shared_flags, h->err);
}

// F (and G) for the pulse of the immediately preceding qoc_total_propagator call, reusing its propagators and chunk totals.
extern "C" int qoc_eval_continue(qoc_handle* h, double* F, double* G) {
  if (!h) return QOC_EINVAL;
  if (!h->system_set) { h->err = "qoc_eval_continue: qoc_set_system has not been called"; return QOC_EINVAL; }
  if (h->path != 2) { h->err = "qoc_eval_continue: only implemented for D > 16 (tiled GEMM path)"; return QOC_EUNSUPPORTED; }
  const qoc_desc& d = h->d;
  QOC_CUDA(h, cudaSetDevice(d.device));
  QOC_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  h->big_reuse = true;
  int rc = eval_device_on(h, h->x, h->out, G != nullptr, h->stream);
  h->big_reuse = false;
  if (rc != QOC_OK) return rc;
  QOC_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  QOC_CUDA(h, copy_result_async(h, G != nullptr));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  QOC_CUDA(h, cudaEventElapsedTime(&h->st.gpu_ms_last_eval, h->ev0, h->ev1));
  const size_t row = (size_t)h->NK + 1;
  if (F) F[0] = h->hout[0];
  if (G) memcpy(G, h->hout + 1, sizeof(double) * h->NK);
  (void)row;
  return QOC_OK;
}

// ------------------------------------------------------------------------------------------------ propagators
static int ensure_staging(qoc_handle* h, size_t bytes) {
  if (bytes <= h->staging_bytes) return QOC_OK;
  if (h->staging) { cudaFree(h->staging); h->ws_bytes -= (long long)h->staging_bytes; h->staging = nullptr; h->staging_bytes = 0; }
  QOC_CUDA(h, cudaMalloc((void**)&h->staging, bytes));
  h->staging_bytes = bytes; h->ws_bytes += (long long)bytes; h->st.workspace_bytes = h->ws_bytes;
  return QOC_OK;
}
static int upload_x(qoc_handle* h, const double* x) {
  const size_t nx = (size_t)h->d.R * h->NK;
  memcpy(h->hx, x, nx * sizeof(double));
  QOC_CUDA(h, cudaMemcpyAsync(h->x, h->hx, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return QOC_OK;
}

extern "C" int qoc_total_propagator(qoc_handle* h, const double* x, double* U) {
  if (!h) return QOC_EINVAL;
  if (!x || !U) { h->err = "qoc_total_propagator: null pointer"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_total_propagator: qoc_set_system has not been called"; return QOC_EINVAL; }
  const qoc_desc& d = h->d;
  QOC_CUDA(h, cudaSetDevice(d.device));
  h->st.launches_last_eval = 0;
  int rc;
  if ((rc = upload_x(h, x)) != QOC_OK) return rc;
  const size_t bytes = (size_t)d.R * d.M * d.D * d.D * sizeof(double2);
  if ((rc = ensure_staging(h, bytes)) != QOC_OK) return rc;
  if (h->path == 2) {
    if ((rc = big_total_propagator(h->big, h->x, h->staging, h->stream, h->err, h->st)) != QOC_OK) return rc;
  } else {
    SmallParams p = small_params(h, h->x);
    p.xi = h->ident; p.out_final = h->staging; p.have_P = 0; p.fom_exact = 0;
    if ((rc = launch_chain(h, p, SYS_UNITARY, GRAD_NONE, h->stream)) != QOC_OK) return rc;
  }
  QOC_CUDA(h, cudaMemcpyAsync(U, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  return QOC_OK;
}

extern "C" int qoc_propagators(qoc_handle* h, const double* x, double* out, int mode) {
  if (!h) return QOC_EINVAL;
  if (!x || !out || mode < 0 || mode > 2) { h->err = "qoc_propagators: bad argument"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_propagators: qoc_set_system has not been called"; return QOC_EINVAL; }
  const qoc_desc& d = h->d;
  QOC_CUDA(h, cudaSetDevice(d.device));
  h->st.launches_last_eval = 0;
  int rc;
  if ((rc = upload_x(h, x)) != QOC_OK) return rc;
  const size_t bytes = (size_t)d.R * d.M * d.N * d.D * d.D * sizeof(double2);
  if ((rc = ensure_staging(h, bytes)) != QOC_OK) return rc;
  if (h->path == 2) {
    if ((rc = big_propagators(h->big, h->x, h->staging, mode, h->stream, h->err, h->st)) != QOC_OK) return rc;
  } else {
    SliceParams s = slice_params(h, h->x);
    s.out_user = h->staging; s.mode = mode;
    if ((rc = launch_slices(h, s, h->stream)) != QOC_OK) return rc;
  }
  QOC_CUDA(h, cudaMemcpyAsync(out, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
  QOC_CUDA(h, cudaStreamSynchronize(h->stream));
  return QOC_OK;
}

// ------------------------------------------------------------------------------------------------ one-shot all-reduce
// Exchange buffer layout per rank: double data[2][n]; unsigned long long flags[2][QOC_MAX_RANKS].
// Epoch e uses half b = e & 1.  A rank can only be one epoch ahead of its slowest peer (it needs that peer's flag to
// finish an epoch), so two halves suffice: nobody overwrites a half that a peer may still be reading.
__global__ void oneshot_allreduce_kernel(char* const* peers, int world, int rank, unsigned long long epoch, size_t n,
                                         double* __restrict__ out) {
  const int b = (int)(epoch & 1);
  const size_t flags_off = 2 * n * sizeof(double);
  if (blockIdx.x == 0 && threadIdx.x < world) {
    // this rank's partial was written by the preceding kernel on the same stream; publish it system-wide, then signal
    __threadfence_system();
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(peers[threadIdx.x] + flags_off) + b * QOC_MAX_RANKS + rank;
    *f = epoch;
  }
  // wait for every peer's signal in OUR flag array (local memory polls)
  const volatile unsigned long long* mine = reinterpret_cast<const volatile unsigned long long*>(peers[rank] + flags_off) + b * QOC_MAX_RANKS;
  if (threadIdx.x < world) { while (mine[threadIdx.x] < epoch) __nanosleep(64); }
  __syncthreads();
  __threadfence_system();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; r++) {
      const volatile double* src = reinterpret_cast<const volatile double*>(peers[r]) + (size_t)b * n;
      s += src[i];                                   // fixed rank order: bit-identical on every rank
    }
    out[i] = s;
  }
}

extern "C" int qoc_comm_export(qoc_handle* h, unsigned char* handle) {
  if (!h || !handle) return QOC_EINVAL;
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  if (!h->comm_local) {
    h->comm_n = (size_t)h->d.R * (h->NK + 1);
    const size_t bytes = 2 * h->comm_n * sizeof(double) + 2 * QOC_MAX_RANKS * sizeof(unsigned long long);
    QOC_CUDA(h, cudaMalloc((void**)&h->comm_local, bytes));
    QOC_CUDA(h, cudaMemset(h->comm_local, 0, bytes));
    QOC_CUDA(h, cudaDeviceSynchronize());
    h->ws_bytes += (long long)bytes;
  }
  cudaIpcMemHandle_t ih;
  QOC_CUDA(h, cudaIpcGetMemHandle(&ih, h->comm_local));
  static_assert(sizeof(cudaIpcMemHandle_t) == QOC_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle, &ih, QOC_IPC_HANDLE_BYTES);
  return QOC_OK;
}

extern "C" int qoc_comm_connect(qoc_handle* h, int world, int rank, const unsigned char* handles) {
  if (!h || !handles || world < 1 || world > QOC_MAX_RANKS || rank < 0 || rank >= world) { if (h) h->err = "qoc_comm_connect: bad argument"; return QOC_EINVAL; }
  if (!h->comm_local) { h->err = "qoc_comm_connect: call qoc_comm_export first"; return QOC_EINVAL; }
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  for (int r = 0; r < world; r++) {
    if (r == rank) { h->comm_peer_host[r] = h->comm_local; continue; }
    cudaIpcMemHandle_t ih;
    memcpy(&ih, handles + (size_t)r * QOC_IPC_HANDLE_BYTES, QOC_IPC_HANDLE_BYTES);
    void* ptr = nullptr;
    QOC_CUDA(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    h->comm_peer_host[r] = (char*)ptr;
  }
  if (!h->comm_peers) QOC_CUDA(h, cudaMalloc((void**)&h->comm_peers, QOC_MAX_RANKS * sizeof(char*)));
  QOC_CUDA(h, cudaMemcpy(h->comm_peers, h->comm_peer_host, QOC_MAX_RANKS * sizeof(char*), cudaMemcpyHostToDevice));
  h->comm_world = world; h->comm_rank = rank; h->comm_epoch = 0;
  return QOC_OK;
}

extern "C" int qoc_eval_allreduce_device(qoc_handle* h, const double* x_dev, double* FG_dev, int want_gradient, void* stream) {
  if (!h) return QOC_EINVAL;
  if (!x_dev || !FG_dev) { h->err = "qoc_eval_allreduce_device: null pointer"; return QOC_EINVAL; }
  if (!h->system_set) { h->err = "qoc_eval_allreduce_device: qoc_set_system has not been called"; return QOC_EINVAL; }
  if (h->comm_world < 1) { h->err = "qoc_eval_allreduce_device: qoc_comm_connect has not been called"; return QOC_EINVAL; }
  QOC_CUDA(h, cudaSetDevice(h->d.device));
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long epoch = ++h->comm_epoch;
  double* mine = reinterpret_cast<double*>(h->comm_local) + (size_t)(epoch & 1) * h->comm_n;
  if (!want_gradient) QOC_CUDA(h, cudaMemsetAsync(mine, 0, h->comm_n * sizeof(double), st));   // G part undefined otherwise
  int rc = eval_device_on(h, x_dev, mine, want_gradient, st);
  if (rc != QOC_OK) return rc;
  const int blocks = (int)std::min<size_t>((h->comm_n + 255) / 256, 64);
  oneshot_allreduce_kernel<<<blocks, 256, 0, st>>>(h->comm_peers, h->comm_world, h->comm_rank, epoch, h->comm_n, FG_dev);
  return launch_check(h, "oneshot_allreduce_kernel");
}

// ------------------------------------------------------------------------------------------------ L-BFGS
extern "C" int qoc_minimize_lbfgs(qoc_handle* h, const double* x0, const qoc_lbfgs_options* opt, double* x_out, qoc_lbfgs_result* res) {
  if (!h) return QOC_EINVAL;
  if (!x0 || !x_out || !res) { h->err = "qoc_minimize_lbfgs: null pointer"; return QOC_EINVAL; }
  if (h->d.R != 1) { h->err = "qoc_minimize_lbfgs: the handle must be created with R = 1"; return QOC_EINVAL; }
  const int n = h->NK;
  const int max_iters = opt && opt->max_iters > 0 ? opt->max_iters : 1000;
  const int m = opt && opt->history > 0 ? opt->history : 10;
  const double g_tol = opt && opt->g_tol > 0 ? opt->g_tol : 1e-8;
  const double f_tol = opt && opt->f_tol >= 0 ? opt->f_tol : 0.0;
  const int max_ls = opt && opt->max_linesearch > 0 ? opt->max_linesearch : 30;
  std::vector<double> x(x0, x0 + n), g(n), xn(n), gn(n), dir(n), alpha(m), rho(m);
  std::vector<std::vector<double>> S(m, std::vector<double>(n)), Y(m, std::vector<double>(n));
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; };
  auto ninf = [&](const std::vector<double>& a) { double s = 0; for (int i = 0; i < n; i++) s = std::max(s, std::fabs(a[i])); return s; };
  double f = 0, fnew = 0;
  int rc = qoc_eval(h, x.data(), &f, g.data());
  if (rc != QOC_OK) return rc;
  int f_calls = 1, stored = 0, head = 0, it = 0, converged = 0;
  for (; it < max_iters; it++) {
    if (ninf(g) <= g_tol) { converged = 1; break; }
    // two-loop recursion: dir = -H g
    for (int i = 0; i < n; i++) dir[i] = -g[i];
    for (int j = 0; j < stored; j++) {
      const int idx = (head - 1 - j + 2 * m) % m;
      alpha[idx] = rho[idx] * dot(S[idx], dir);
      for (int i = 0; i < n; i++) dir[i] -= alpha[idx] * Y[idx][i];
    }
    if (stored > 0) {
      const int last = (head - 1 + m) % m;
      const double gamma = dot(S[last], Y[last]) / dot(Y[last], Y[last]);
      for (int i = 0; i < n; i++) dir[i] *= gamma;
    }
    for (int j = stored - 1; j >= 0; j--) {
      const int idx = (head - 1 - j + 2 * m) % m;
      const double beta = rho[idx] * dot(Y[idx], dir);
      for (int i = 0; i < n; i++) dir[i] += (alpha[idx] - beta) * S[idx][i];
    }
    double slope = dot(g, dir);
    if (!(slope < 0)) { for (int i = 0; i < n; i++) dir[i] = -g[i]; slope = -dot(g, g); stored = 0; }   // not a descent direction: restart
    // backtracking Armijo line search (first iteration: scale the step to a unit-length move)
    double step = (stored == 0) ? 1.0 / std::max(1.0, std::sqrt(dot(g, g))) : 1.0;
    bool ok = false;
    for (int ls = 0; ls < max_ls; ls++) {
      for (int i = 0; i < n; i++) xn[i] = x[i] + step * dir[i];
      if ((rc = qoc_eval(h, xn.data(), &fnew, gn.data())) != QOC_OK) return rc;
      f_calls++;
      if (fnew <= f + 1e-4 * step * slope) { ok = true; break; }
      step *= 0.5;
    }
    if (!ok) break;                                  // no acceptable step: stop at the current point
    // expansion towards the (weak) Wolfe curvature condition: while the slope along dir is still steep, try doubling
    {
      std::vector<double> xe(n), ge(n);
      for (int ex = 0; ex < 6 && dot(gn, dir) < 0.9 * slope; ex++) {
        const double s2 = 2.0 * step;
        double fe = 0;
        for (int i = 0; i < n; i++) xe[i] = x[i] + s2 * dir[i];
        if ((rc = qoc_eval(h, xe.data(), &fe, ge.data())) != QOC_OK) return rc;
        f_calls++;
        if (!(fe <= f + 1e-4 * s2 * slope) || fe >= fnew) break;
        step = s2; fnew = fe; xn.swap(xe); gn.swap(ge);
      }
    }
    for (int i = 0; i < n; i++) { S[head][i] = xn[i] - x[i]; Y[head][i] = gn[i] - g[i]; }
    const double sy = dot(S[head], Y[head]);
    const double fprev = f;
    x.swap(xn); g.swap(gn); f = fnew;
    if (sy > 1e-12 * std::sqrt(dot(S[head], S[head]) * dot(Y[head], Y[head]))) { rho[head] = 1.0 / sy; head = (head + 1) % m; stored = std::min(stored + 1, m); }
    if (f_tol > 0 && std::fabs(fprev - f) <= f_tol * std::fabs(f)) { converged = 1; it++; break; }
  }
  if (!converged && ninf(g) <= g_tol) converged = 1;
  memcpy(x_out, x.data(), sizeof(double) * n);
  res->minimum = f; res->g_norm = ninf(g); res->iterations = it; res->f_calls = f_calls; res->converged = converged;
  return QOC_OK;
}

extern "C" int qoc_get_stats(qoc_handle* h, qoc_stats* out) {
  if (!h || !out) return QOC_EINVAL;
  h->st.workspace_bytes = h->ws_bytes + (h->big ? big_workspace(h->big) : 0);
  {  // average over the event pairs recorded since the previous qoc_get_stats (at most KRING)
    int n = h->kring_count < qoc_handle::KRING ? h->kring_count : qoc_handle::KRING;
    double sum = 0; int got = 0;
    for (int i = 0; i < n; i++) {
      float ms = 0;
      if (cudaEventSynchronize(h->ek1[i]) == cudaSuccess && cudaEventElapsedTime(&ms, h->ek0[i], h->ek1[i]) == cudaSuccess) { sum += ms; got++; }
    }
    h->st.main_kernel_ms_avg = got ? (float)(sum / got) : 0.f;
    h->st.main_kernel_samples = got;
    h->kring_count = 0;
  }
  *out = h->st;
  return QOC_OK;
}
