// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's splat_grid_samples_nerf_max_nearest_neighbor_api
// (/root/reference/extensions/ngp_raymarch/src/splat_grid_samples_nerf_max_nearest_neighbor.cu:30-57), compiled for CPU.
#include "gen/splat_grid_samples_nerf_max_nearest_neighbor.cu"
#include "harness_common.h"
extern "C" void ref_splat(const float *mlp_out, const int32_t *indices, int padded_width, int n, float *grid_tmp) {
    auto g = T(grid_tmp, {0});
    splat_grid_samples_nerf_max_nearest_neighbor_api(T(mlp_out, {n, padded_width}), T(indices, {n}, at::ScalarType::Int), padded_width, n, g);
}
