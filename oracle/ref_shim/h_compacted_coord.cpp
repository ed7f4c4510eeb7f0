// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's compacted_coord_api
// (/root/reference/extensions/ngp_raymarch/src/compacted_coord.cu:79-143), compiled for CPU.
#include "gen/compacted_coord.cu"
#include "harness_common.h"
extern "C" void ref_compacted_coord(const float *raw, const float *coords_in, const int32_t *numsteps, const float *bg3,
                                    int n_samples, int n_rays, int max_compacted, int rgb_act, int dens_act, float aabb0, float aabb1,
                                    float *coords_out, int32_t *numsteps_c, int32_t *ray_counter, int32_t *step_counter) {
    auto co = T(coords_out, {max_compacted, 7}); auto nc = T(numsteps_c, {n_rays, 2}, at::ScalarType::Int);
    auto rc = T(ray_counter, {1}, at::ScalarType::Int); auto sc = T(step_counter, {1}, at::ScalarType::Int);
    compacted_coord_api(T(raw, {n_samples, 4}), T(coords_in, {n_samples, 7}), T(numsteps, {n_rays, 2}, at::ScalarType::Int), T(bg3, {3}),
                        rgb_act, dens_act, aabb0, aabb1, co, nc, rc, sc);
}
