// xrnerf_b200 — shared device helpers (sm_100a only).
//
// Index-path arithmetic is written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn):
// nvcc never contracts those into FMAs, which makes the per-ray sample sets bit-reproducible against the
// un-contracted IEEE fp32 semantics of the reference's C++ source. Reference citations are
// relative to /root/reference/extensions/ngp_raymarch/.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <float.h>
#include "../../include/xrnerf_b200.h"

namespace xrb {

constexpr uint32_t NERF_STEPS = 1024u;      // raymarch_shared.h:42
constexpr uint32_t NERF_CASCADES = 8u;      // :43
constexpr uint32_t NERF_GRIDSIZE = 128u;    // :48
constexpr uint32_t GRID_CELLS = 128u * 128u * 128u;
constexpr float NERF_MIN_OPTICAL_THICKNESS = 0.01f;  // :56
constexpr int NUM_SMS = 148;

// constants folded exactly as the reference's constexpr chain folds them (fp32, left to right)
__host__ __device__ constexpr float SQRT3() { return 1.73205080757f; }
__host__ __device__ constexpr float MIN_CONE_STEPSIZE() { return SQRT3() / NERF_STEPS; }
__host__ __device__ constexpr float MAX_CONE_STEPSIZE() { return (SQRT3() / NERF_STEPS) * (1 << (NERF_CASCADES - 1)) * NERF_STEPS / NERF_GRIDSIZE; }

void set_error(const char *msg);
int check_launch(const char *what);

#define XRB_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) { xrb::set_error(msg); return XRB_E_BADARG; } \
    } while (0)

// ---------------------------------------------------------------- pcg32 (include/op_include/pcg32/pcg32.h:41-165)
struct Pcg32 {
    uint64_t state, inc;
    __host__ __device__ uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    __host__ __device__ void seed(uint64_t initstate, uint64_t initseq) {
        state = 0; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint();
    }
    __host__ __device__ float next_float() {
#ifdef __CUDA_ARCH__
        return __fadd_rn(__uint_as_float((next_uint() >> 9) | 0x3f800000u), -1.0f);
#else
        union { uint32_t u; float f; } x; x.u = (next_uint() >> 9) | 0x3f800000u; return x.f - 1.0f;
#endif
    }
    __host__ __device__ void advance(uint64_t delta) {
        uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
            cur_plus = (cur_mult + 1) * cur_plus; cur_mult *= cur_mult; delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};
// host rng of one reference translation unit after n_prior_calls API calls (raymarch_shared.h:38, ray_sampler.cu:198)
inline Pcg32 host_rng(uint64_t seed, int64_t n_prior_calls) {
    Pcg32 r; r.seed(seed, 1u);
    r.advance((uint64_t)n_prior_calls << 32);   // k calls of advance(2^32) == one jump of k*2^32 (mod 2^64): O(64) on the host instead of O(64 k)
    return r;
}

// ---------------------------------------------------------------- Morton (raymarch_shared.h:122-131, :753-768)
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249; x = (x | (x >> 2)) & 0xc30c30c3; x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff; x = (x | (x >> 16)) & 0x0000ffff; return x;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- exact-rounding helpers
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }  // raymarch_shared.h:104-107
__device__ __forceinline__ float calc_dt(float t, float cone) { return clampf(mul_(t, cone), MIN_CONE_STEPSIZE(), MAX_CONE_STEPSIZE()); }  // ray_sampler_header.h:24-25

// floor() of |p| < 2^22 and its integer value on the FMA/ALU pipes (F2I / FRND / I2F are quarter-rate XU ops on sm_100):
// t = p + 1.5*2^23 rounds p to the nearest integer into the low mantissa bits; one compare turns round-to-nearest into floor. Exact.
__device__ __forceinline__ float floor_small(float p, int *ip) {
    float t = __fadd_rn(p, 12582912.0f);
    float r = __fadd_rn(t, -12582912.0f);
    int i = __float_as_int(t) - 0x4B400000;
    if (r > p) { r = __fadd_rn(r, -1.0f); i -= 1; }
    *ip = i;
    return r;
}
// expand_bits() of raymarch_shared.h:753-760 for 7-bit inputs via the identity spread3(v) = sum of bit b moved to 3b: same value, 2 steps
__device__ __forceinline__ uint32_t expand_bits7(uint32_t v) {
    v = (v | (v << 8)) & 0x0000F00Fu;   // bits 0-3 stay, bits 4-6 -> 12-14
    v = (v | (v << 4)) & 0x000C30C3u;   // pairs apart
    v = (v | (v << 2)) & 0x00249249u;   // every third bit
    return v;
}

// frexpf exponent for finite positive normal/zero inputs (the only ones the march produces); matches frexpf incl. 0 -> 0
__device__ __forceinline__ int frexp_exponent(float v) {
    // v >= 0 finite. Normal numbers: biased exponent - 126 (frexpf's convention 0.5 <= m < 1); zero -> 0; denormals take libm.
    uint32_t e = __float_as_uint(v) >> 23;
    if (e != 0) return (int)e - 126;
    if (v == 0.f) return 0;
    int r; frexpf(v, &r); return r;
}
// frexpf exponent e of m: m in [2^(e-1), 2^e). The march only needs clamp(e + 1, 0, 7):
//   m >= 0.25 (a normal number): e = biased exponent - 126 >= -1;   0 < m < 0.25 (incl. denormals): e <= -2 -> 0;   m == 0: frexpf yields e = 0 -> 1
__device__ __forceinline__ int mip_from_pos(float px, float py, float pz) {  // ray_sampler_header.h:37-43
    float m = fmaxf(fabsf(sub_(px, 0.5f)), fmaxf(fabsf(sub_(py, 0.5f)), fabsf(sub_(pz, 0.5f))));
    if (m >= 0.25f) return min((int)NERF_CASCADES - 1, (int)(__float_as_uint(m) >> 23) - 125);
    return m == 0.f ? 1 : 0;
}
__device__ __forceinline__ int mip_from_dt(float dt, float px, float py, float pz) {  // ray_sampler_header.h:45-54
    int mip = mip_from_pos(px, py, pz);
    dt = mul_(dt, (float)(2 * NERF_GRIDSIZE));
    if (dt < 1.f) return mip;
    int e = (int)(__float_as_uint(dt) >> 23) - 126;   // dt >= 1: a normal number, frexpf exponent = biased exponent - 126
    return min((int)NERF_CASCADES - 1, max(e, mip));
}
// expand_bits() of a 7-bit cell coordinate through a 128-entry shared-memory table (3 LDS + 2 LEA instead of ~27 ALU ops per tested position)
__device__ __forceinline__ void morton_lut_init(uint32_t *lut, int tid, int nthreads) {
    for (int v = tid; v < 128; v += nthreads) lut[v] = expand_bits7((uint32_t)v);
}
__device__ __forceinline__ uint32_t cascaded_grid_idx_at(float px, float py, float pz, uint32_t mip, const uint32_t *lut = nullptr) {  // ray_sampler_header.h:298-313
    float s = __uint_as_float((127u - mip) << 23);  // scalbnf(1, -mip), exact
    float q[3] = {px, py, pz}; uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = add_(mul_(sub_(q[k], 0.5f), s), 0.5f);
        // the reference truncates toward zero (cast<int>) then clamps to [0,127]; floor differs from trunc only for negative values,
        // which both clamp to 0 -> identical cell. floor_small keeps this off the XU pipe.
        int i; floor_small(mul_(v, (float)NERF_GRIDSIZE), &i);
        c[k] = (uint32_t)min(max(i, 0), (int)NERF_GRIDSIZE - 1);
    }
    if (lut) return lut[c[0]] + (lut[c[1]] << 1) + (lut[c[2]] << 2);   // disjoint bit sets: + == |
    return expand_bits7(c[0]) | (expand_bits7(c[1]) << 1) | (expand_bits7(c[2]) << 2);
}
__device__ __forceinline__ bool occupied_at(float px, float py, float pz, const uint8_t *__restrict__ bitfield, uint32_t mip, const uint32_t *lut = nullptr) {  // :315-319
    uint32_t idx = cascaded_grid_idx_at(px, py, pz, mip, lut);
    return __ldg(bitfield + (idx >> 3) + ((GRID_CELLS * mip) >> 3)) & (1u << (idx & 7u));
}
__device__ __forceinline__ float signf_(float x) { return copysignf(1.0f, x); }
__device__ __forceinline__ float distance_to_next_voxel(const float p[3], const float d[3], const float idir[3], uint32_t res) {  // :271-280
    float t3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float pr = mul_((float)res, p[k]);
        int unused; float fl = floor_small(add_(add_(pr, 0.5f), mul_(0.5f, signf_(d[k]))), &unused);   // == floorf, |arg| < 2^22
        t3[k] = mul_(sub_(fl, pr), idir[k]);
    }
    float t = fminf(fminf(t3[0], t3[1]), t3[2]);
    return fmaxf(mul_(t, __uint_as_float((254u - (__float_as_uint((float)res) >> 23)) << 23)), 0.0f);  // t / res, res = 2^k: exact reciprocal, same bits as the division
}
__device__ __forceinline__ float advance_to_next_voxel(float t, float cone, const float p[3], const float d[3], const float idir[3], uint32_t res) {  // :282-296
    float t_target = add_(t, distance_to_next_voxel(p, d, idir, res));
    do { t = add_(t, calc_dt(t, cone)); } while (t < t_target);
    return t;
}
// BoundingBox::ray_intersect (raymarch_shared.h:506-563): returns tmin (FLT_MAX on a miss)
__device__ __forceinline__ float aabb_ray_tmin(float lo, float hi, const float o[3], const float d[3]) {
    float tmin = div_(sub_(lo, o[0]), d[0]), tmax = div_(sub_(hi, o[0]), d[0]);
    if (tmin > tmax) { float c = tmin; tmin = tmax; tmax = c; }
    float tymin = div_(sub_(lo, o[1]), d[1]), tymax = div_(sub_(hi, o[1]), d[1]);
    if (tymin > tymax) { float c = tymin; tymin = tymax; tymax = c; }
    if (tmin > tymax || tymin > tmax) return FLT_MAX;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = div_(sub_(lo, o[2]), d[2]), tzmax = div_(sub_(hi, o[2]), d[2]);
    if (tzmin > tzmax) { float c = tzmin; tzmin = tzmax; tzmax = c; }
    if (tmin > tzmax || tzmin > tmax) return FLT_MAX;
    if (tzmin > tmin) tmin = tzmin;
    return tmin;
}
__device__ __forceinline__ bool aabb_contains(float lo, float hi, float x, float y, float z) {  // raymarch_shared.h:570-575
    return x >= lo && x <= hi && y >= lo && y <= hi && z >= lo && z <= hi;
}
__device__ __forceinline__ float warp_dt(float dt) {  // raymarch_shared.h:110-114
    const float max_stepsize = MIN_CONE_STEPSIZE() * (1 << (NERF_CASCADES - 1));
    return div_(sub_(dt, MIN_CONE_STEPSIZE()), max_stepsize - MIN_CONE_STEPSIZE());
}
__device__ __forceinline__ float unwarp_dt(float dt) {  // ray_sampler_header.h:388-392
    const float max_stepsize = MIN_CONE_STEPSIZE() * (1 << (NERF_CASCADES - 1));
    return add_(mul_(dt, max_stepsize - MIN_CONE_STEPSIZE()), MIN_CONE_STEPSIZE());
}
// start of the march for ray i (ray_sampler.cu:31,:43-51)
__device__ __forceinline__ float ray_start_t(Pcg32 rng, uint32_t i, float lo, float hi, const float o[3], const float d[3], float near_distance, float cone) {
    rng.advance((uint64_t)(i * 8u));
    float tmin = fmaxf(aabb_ray_tmin(lo, hi, o, d), near_distance);
    return add_(tmin, mul_(calc_dt(tmin, cone), rng.next_float()));
}

// activations (ray_sampler_header.h:440-456, :534-574; raymarch_shared.h:615-642). __expf as in the reference.
__device__ __forceinline__ float logistic(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float net_to_rgb(float v, int act) {
    switch (act) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return logistic(v); case 3: return __expf(clampf(v, -10.f, 10.f)); }
    return 0.f;
}
__device__ __forceinline__ float net_to_density(float v, int act) {
    switch (act) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return logistic(v); case 3: return __expf(v); }
    return 0.f;
}
__device__ __forceinline__ float net_to_rgb_deriv(float v, int act) {
    switch (act) { case 0: return 1.f; case 1: return v > 0.f ? 1.f : 0.f; case 2: { float s = logistic(v); return s * (1 - s); } case 3: return __expf(clampf(v, -10.f, 10.f)); }
    return 0.f;
}
__device__ __forceinline__ float net_to_density_deriv(float v, int act) {
    switch (act) { case 0: return 1.f; case 1: return v > 0.f ? 1.f : 0.f; case 2: { float s = logistic(v); return s * (1 - s); } case 3: return __expf(clampf(v, -15.f, 15.f)); }
    return 0.f;
}

// warp scans
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    return v;
}
#endif  // __CUDACC__

}  // namespace xrb
