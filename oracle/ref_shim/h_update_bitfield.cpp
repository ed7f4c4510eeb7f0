// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's update_bitfield_api
// (/root/reference/extensions/ngp_raymarch/src/update_bitfield.cu:74-116), compiled for CPU.
#include "gen/update_bitfield.cu"
#include "harness_common.h"
extern "C" void ref_update_bitfield(const float *grid, float *mean, uint8_t *bitfield) {
    auto m = T(mean, {1}); auto b = T(bitfield, {0}, at::ScalarType::Byte);
    update_bitfield_api(T(grid, {0}), m, b);
}
