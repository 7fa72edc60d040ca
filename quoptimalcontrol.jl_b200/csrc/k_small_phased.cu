// Translation unit of the slice-parallel ("phased") and chunk-parallel kernels (small_phased.cuh).
#include "small_phased.cuh"

namespace qoc {

template <int NB, int CPW> static void pick_phased2(int sys, int grad, phased_fn& tot, phased_fn& bnd, phased_fn& swp, phased_fn& grd) {
  tot = chunk_totals_kernel<NB, CPW>;
  if (sys == SYS_UNITARY) {
    bnd = boundary_kernel<NB, CPW, SYS_UNITARY>; swp = sweep_kernel<NB, CPW, SYS_UNITARY>;
    grd = grad == GRAD_EXACT ? grad_slices_kernel<NB, CPW, SYS_UNITARY, GRAD_EXACT> : grad_slices_kernel<NB, CPW, SYS_UNITARY, GRAD_FIRST>;
  } else {
    bnd = boundary_kernel<NB, CPW, SYS_DENSITY>; swp = sweep_kernel<NB, CPW, SYS_DENSITY>;
    grd = grad == GRAD_EXACT ? grad_slices_kernel<NB, CPW, SYS_DENSITY, GRAD_EXACT> : grad_slices_kernel<NB, CPW, SYS_DENSITY, GRAD_FIRST>;
  }
}
void pick_phased(int NB, int CPW, int sys, int grad, phased_fn& tot, phased_fn& bnd, phased_fn& swp, phased_fn& grd) {
  if (NB == 2) pick_phased2<2, 1>(sys, grad, tot, bnd, swp, grd);
  else if (CPW == 4) pick_phased2<1, 4>(sys, grad, tot, bnd, swp, grd);
  else if (CPW == 2) pick_phased2<1, 2>(sys, grad, tot, bnd, swp, grd);
  else pick_phased2<1, 1>(sys, grad, tot, bnd, swp, grd);
}
phased_fn pick_chunk_expm(int NB, int CPW) {
  if (NB == 2) return chunk_expm_kernel<2, 1>;
  if (CPW == 4) return chunk_expm_kernel<1, 4>;
  if (CPW == 2) return chunk_expm_kernel<1, 2>;
  return chunk_expm_kernel<1, 1>;
}
phased_fn pick_boundary2(int NB, int CPW, int sys) {
  const bool u = sys == SYS_UNITARY;
  if (NB == 2) return u ? boundary2_kernel<2, 1, SYS_UNITARY> : boundary2_kernel<2, 1, SYS_DENSITY>;
  if (CPW == 4) return u ? boundary2_kernel<1, 4, SYS_UNITARY> : boundary2_kernel<1, 4, SYS_DENSITY>;
  if (CPW == 2) return u ? boundary2_kernel<1, 2, SYS_UNITARY> : boundary2_kernel<1, 2, SYS_DENSITY>;
  return u ? boundary2_kernel<1, 1, SYS_UNITARY> : boundary2_kernel<1, 1, SYS_DENSITY>;
}
phased_fn pick_boundary_unitary(int NB, int CPW, int sys) {
  const bool u = sys == SYS_UNITARY;
  if (NB == 2) return u ? boundary_unitary_kernel<2, 1, SYS_UNITARY> : boundary_unitary_kernel<2, 1, SYS_DENSITY>;
  if (CPW == 4) return u ? boundary_unitary_kernel<1, 4, SYS_UNITARY> : boundary_unitary_kernel<1, 4, SYS_DENSITY>;
  if (CPW == 2) return u ? boundary_unitary_kernel<1, 2, SYS_UNITARY> : boundary_unitary_kernel<1, 2, SYS_DENSITY>;
  return u ? boundary_unitary_kernel<1, 1, SYS_UNITARY> : boundary_unitary_kernel<1, 1, SYS_DENSITY>;
}
phased_fn pick_sweep_unitary(int NB, int CPW) {
  if (NB == 2) return sweep_unitary_kernel<2, 1>;
  if (CPW == 4) return sweep_unitary_kernel<1, 4>;
  if (CPW == 2) return sweep_unitary_kernel<1, 2>;
  return sweep_unitary_kernel<1, 1>;
}

phased_fn pick_chunk_expm_dmma() { return chunk_expm_dmma_kernel; }
int chunk_expm_dmma_smem() { return (1024 + 4 * ASM_WARP_DOUBLES) * (int)sizeof(double); }
phased_fn pick_sweep_unitary_dmma() { return sweep_unitary_dmma_kernel; }
int sweep_unitary_dmma_smem() { return (1024 + 4 * 8 * DOT_LD) * (int)sizeof(double); }
persist_fn pick_closed_persistent(int sys) { return sys == SYS_UNITARY ? closed_persistent_kernel<SYS_UNITARY> : closed_persistent_kernel<SYS_DENSITY>; }
int closed_persistent_smem() { return (1024 + 4 * ASM_WARP_DOUBLES) * (int)sizeof(double); }
int closed_persistent_ctl_ints(int n_groups, int Cn) { return persist_ctl_ints(n_groups, Cn); }
}  // namespace qoc
