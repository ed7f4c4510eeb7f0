// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's rays_sampler_api
// (/root/reference/extensions/ngp_raymarch/src/ray_sampler.cu:118-200), compiled for CPU.
#include "gen/ray_sampler.cu"
#include "harness_common.h"
REF_RNG_CONTROL(ray_sampler)
extern "C" void ref_rays_sampler(const float *rays_o, const float *rays_d, const uint8_t *bitfield, const float *metadata,
                                 const int32_t *img_ids, const float *xforms, int n_rays, int n_img, int max_samples,
                                 float aabb0, float aabb1, float near_distance, float cone_angle, float *coords_out,
                                 int32_t *rays_index, int32_t *numsteps, int32_t *counters) {
    auto coords = T(coords_out, {max_samples, 7}); auto ridx = T(rays_index, {n_rays, 1}, at::ScalarType::Int);
    auto ns = T(numsteps, {n_rays, 2}, at::ScalarType::Int); auto cnt = T(counters, {2}, at::ScalarType::Int);
    rays_sampler_api(T(rays_o, {n_rays, 3}), T(rays_d, {n_rays, 3}), T(bitfield, {0}, at::ScalarType::Byte), T(metadata, {n_img, 11}),
                     T(img_ids, {n_rays}, at::ScalarType::Int), T(xforms, {n_img, 4, 3}), aabb0, aabb1, near_distance, cone_angle,
                     coords, ridx, ns, cnt);
}
