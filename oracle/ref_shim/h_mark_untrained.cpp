// TEST INFRASTRUCTURE ONLY: C-ABI door onto the reference's mark_untrained_density_grid_api
// (/root/reference/extensions/ngp_raymarch/src/mark_untrained_density_grid.cu:53-82), compiled for CPU.
// The caller pre-fills `grid` (the reference reads it uninitialised).
#include "gen/mark_untrained_density_grid.cu"
#include "harness_common.h"
extern "C" void ref_mark_untrained(const float *focal, const float *xforms, int n_elements, int n_images, int res0, int res1, float *grid) {
    auto g = T(grid, {n_elements});
    mark_untrained_density_grid_api(T(focal, {n_images, 2}), T(xforms, {n_images, 4, 3}), n_elements, n_images, res0, res1, g);
}
