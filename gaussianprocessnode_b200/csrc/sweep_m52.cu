// Instantiates the fused sweep kernels of one kernel family (see sweep_kernel.cuh / sweep.cu).
#include "sweep_kernel.cuh"

namespace sgp_sweep {
int launch_m52(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad, int TM) {
    return launch_kind<SGP_KERNEL_MATERN52>(ctx, p, weighted, grid, dpad, TM);
}
}  // namespace sgp_sweep
