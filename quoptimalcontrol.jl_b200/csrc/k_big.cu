// Translation unit of the large-dimension path: tiled DMMA GEMM pipeline + pure-state vector sweep.
#include "big_d.cuh"

namespace qoc {
bool big_pure_active(const BigState* s) { return s && s->pure.active; }
}  // namespace qoc
