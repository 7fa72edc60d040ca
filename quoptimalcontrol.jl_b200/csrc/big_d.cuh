// Large-dimension path (D > 16): tiled DMMA complex GEMM pipeline.  Placeholder until the tiled kernels land:
// every entry returns QOC_EUNSUPPORTED (there is no CPU fallback).
#pragma once
#include <string>
#include "../../include/qocgrape.h"

namespace qoc {
struct BigState { int dummy; };
static inline int big_create(BigState**, const qoc_desc& d, std::string& err, long long&) {
  err = "D = " + std::to_string(d.D) + " > 16 is not implemented by the CUDA path yet"; return QOC_EUNSUPPORTED;
}
static inline void big_destroy(BigState*) {}
static inline long long big_workspace(BigState*) { return 0; }
static inline int big_set_system(BigState*, const double*, const double*, const double*, const double*, int, std::string&) { return QOC_EUNSUPPORTED; }
static inline int big_eval(BigState*, const double*, double*, int, const double*, cudaStream_t, std::string&, qoc_stats&) { return QOC_EUNSUPPORTED; }
static inline int big_total_propagator(BigState*, const double*, double2*, cudaStream_t, std::string&, qoc_stats&) { return QOC_EUNSUPPORTED; }
static inline int big_propagators(BigState*, const double*, double2*, int, cudaStream_t, std::string&, qoc_stats&) { return QOC_EUNSUPPORTED; }
}  // namespace qoc
