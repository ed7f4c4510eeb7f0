// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's generate_grid_samples_nerf_nonuniform_api
// (/root/reference/extensions/ngp_raymarch/src/generate_grid_samples_nerf_nonuniform.cu:44-87), compiled for CPU.
#include "gen/generate_grid_samples_nerf_nonuniform.cu"
#include "harness_common.h"
REF_RNG_CONTROL(generate_grid_samples)
extern "C" void ref_generate_grid_samples(const float *grid, int ema_step, int n_elements, int max_cascade, float thresh, float aabb0,
                                          float aabb1, float *positions, int32_t *indices) {
    auto po = T(positions, {n_elements, 3}); auto io = T(indices, {n_elements}, at::ScalarType::Int);
    generate_grid_samples_nerf_nonuniform_api(T(grid, {0}), ema_step, n_elements, max_cascade, thresh, aabb0, aabb1, po, io);
}
