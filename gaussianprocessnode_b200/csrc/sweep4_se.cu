// Instantiates the generate-once sweep kernels of one kernel family (see sweep4_kernel.cuh / sweep.cu).
#include "sweep4_kernel.cuh"

namespace sgp_sweep4 {
int launch4_se(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad, int TM) {
    return launch4_kind<SGP_KERNEL_SE>(ctx, p, weighted, grid, dpad, TM);
}
}  // namespace sgp_sweep4
