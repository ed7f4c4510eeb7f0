// TEST INFRASTRUCTURE ONLY: C-ABI doors onto the reference's calc_rgb_{forward,backward,influence}_api
// (/root/reference/extensions/ngp_raymarch/src/calc_rgb.cu:208-389), compiled for CPU.
#include "gen/calc_rgb.cu"
#include "harness_common.h"
extern "C" void ref_calc_rgb_forward(const float *raw, const float *coords, const int32_t *numsteps, const int32_t *numsteps_c,
                                     const float *bg, int n_samples, int n_rays, int rgb_act, int dens_act, float aabb0, float aabb1,
                                     float *rgb_out) {
    auto out = T(rgb_out, {n_rays, 3});
    calc_rgb_forward_api(T(raw, {n_samples, 4}), T(coords, {n_samples, 7}), T(numsteps, {n_rays, 2}, at::ScalarType::Int),
                         T(numsteps_c, {n_rays, 2}, at::ScalarType::Int), T(bg, {n_rays, 3}), rgb_act, dens_act, aabb0, aabb1, out);
}
extern "C" void ref_calc_rgb_backward(const float *raw, const int32_t *numsteps_c, const float *coords, const float *grad_rgb,
                                      const float *rgb, const float *grid_mean, int n_samples, int n_rays, int rgb_act, int dens_act,
                                      float aabb0, float aabb1, float *dl_draw) {
    auto out = T(dl_draw, {n_samples, 4});
    calc_rgb_backward_api(T(raw, {n_samples, 4}), T(numsteps_c, {n_rays, 2}, at::ScalarType::Int), T(coords, {n_samples, 7}),
                          T(grad_rgb, {n_rays, 3}), T(rgb, {n_rays, 3}), T(grid_mean, {1}), rgb_act, dens_act, aabb0, aabb1, out);
}
extern "C" void ref_calc_rgb_inference(const float *raw, const float *coords, const int32_t *numsteps, const float *bg3, int n_samples,
                                       int n_rays, int rgb_act, int dens_act, float aabb0, float aabb1, float *rgb_out, float *alpha_out) {
    auto o1 = T(rgb_out, {n_rays, 3}); auto o2 = T(alpha_out, {n_rays, 1});
    calc_rgb_influence_api(T(raw, {n_samples, 4}), T(coords, {n_samples, 7}), T(numsteps, {n_rays, 2}, at::ScalarType::Int), T(bg3, {3}),
                           rgb_act, dens_act, aabb0, aabb1, o1, o2);
}
