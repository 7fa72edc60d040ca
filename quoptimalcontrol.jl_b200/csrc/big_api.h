// Host entry points of the large-dimension path (D > 16), defined in k_big.cu (big_d.cuh).
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "../../include/qocgrape.h"

namespace qoc {

struct BigState;
int big_create(BigState** out, const qoc_desc& d, std::string& err, long long& ws_total);
void big_destroy(BigState* s);
long long big_workspace(BigState* s);
int big_set_system(BigState* s, const double* A, const double* B, const double* Xi, const double* Xt, int shared, std::string& err);
int big_set_states(BigState* s, const double* Xi, const double* Xt, int shared, std::string& err);
bool big_pure_active(const BigState* s);      // the handle runs the pure-state vector path (qoc_stats.path == 3)
// reuse: skip the propagator and chunk-total phases and continue from what big_total_propagator left (same pulse, one chain)
int big_eval(BigState* s, const double* x_dev, double* FG_dev, int want_grad, const double* wts_dev, cudaStream_t st,
             std::string& err, qoc_stats& stats, bool reuse = false);
int big_propagators(BigState* s, const double* x_dev, double2* out, int mode, cudaStream_t st, std::string& err, qoc_stats& stats);
int big_total_propagator(BigState* s, const double* x_dev, double2* out, cudaStream_t st, std::string& err, qoc_stats& stats);

// slice-parallel evaluation of one instance over several GPUs (device-pointer entry points)
int big_padded_dim(const BigState* s);
int big_range_propagator_device(BigState* s, const double* x_dev, double2* U_out, cudaStream_t st, std::string& err, qoc_stats& stats);
int big_set_states_device(BigState* s, const double2* Xi_pad, const double2* Xt_pad, cudaStream_t st, std::string& err);
int big_matmul(BigState* s, int opA, int opB, const double2* A, const double2* B, double2* C, cudaStream_t st, std::string& err, qoc_stats& stats);
int big_upload_states_padded(BigState* s, double2* dst, const double* src, std::string& err);

}  // namespace qoc
