// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's ema_grid_samples_nerf_api
// (/root/reference/extensions/ngp_raymarch/src/ema_grid_samples_nerf.cu:29-50), compiled for CPU.
#include "gen/ema_grid_samples_nerf.cu"
#include "harness_common.h"
extern "C" void ref_ema(const float *grid_tmp, int n_elements, float decay, float *grid) {
    auto g = T(grid, {n_elements});
    ema_grid_samples_nerf_api(T(grid_tmp, {n_elements}), n_elements, decay, g);
}
