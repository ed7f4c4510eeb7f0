/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle ("port") for the Instant-NGP ray-march /
 * compaction / compositing / occupancy-grid kernels of the reference's
 * extensions/ngp_raymarch. Plain C restatement; every function cites the
 * reference file:line it follows (paths relative to
 * /root/reference/extensions/ngp_raymarch/). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PINNING: tests/test_oracle_vs_ref.py checks every function here bit-for-bit
 * (index path) / to 1e-6 (compositing) against oracle/_ref/libraymarch_ref.so,
 * which is the reference's own .cu code compiled for CPU (see oracle/Makefile).
 *
 * Arithmetic contract: un-contracted IEEE fp32 (-ffp-contract=off); the fast
 * exponential `__expf` of the reference is expf here.
 *
 * Layout decisions: sample bases are assigned in RAY ORDER
 * (Q2: the reference's are atomic-arrival order; a serial run of the reference
 * gives exactly this layout), rng = (seed 9121, n_prior_calls) explicit (Q9).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

/* ---- constants: include/raymarch_shared.h:41-56 */
#define NERF_STEPS 1024u
#define NERF_CASCADES 8u
#define NERF_GRIDSIZE 128u
#define GRID_CELLS (128u * 128u * 128u)
static const float SQRT3 = 1.73205080757f;
static inline float MIN_CONE_STEPSIZE(void) { return SQRT3 / NERF_STEPS; }
static inline float MAX_CONE_STEPSIZE(void) { return (SQRT3 / NERF_STEPS) * (1 << (NERF_CASCADES - 1)) * NERF_STEPS / NERF_GRIDSIZE; }
#define NERF_MIN_OPTICAL_THICKNESS 0.01f
#define N_MAX_RANDOM_SAMPLES_PER_RAY 8u

/* ---- pcg32: include/op_include/pcg32/pcg32.h:41-58 (seed), :62-68 (next_uint), :103-112 (next_float), :145-165 (advance) */
typedef struct { uint64_t state, inc; } pcg32_t;
#define PCG32_MULT 0x5851f42d4c957f2dULL
static inline uint32_t pcg_next_uint(pcg32_t *r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
static inline void pcg_seed(pcg32_t *r, uint64_t initstate, uint64_t initseq) {
    r->state = 0; r->inc = (initseq << 1u) | 1u; pcg_next_uint(r); r->state += initstate; pcg_next_uint(r);
}
static inline float pcg_next_float(pcg32_t *r) {
    union { uint32_t u; float f; } x; x.u = (pcg_next_uint(r) >> 9) | 0x3f800000u; return x.f - 1.0f;
}
static inline void pcg_advance(pcg32_t *r, uint64_t delta) {
    uint64_t cur_mult = PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus; cur_mult *= cur_mult; delta /= 2;
    }
    r->state = acc_mult * r->state + acc_plus;
}
/* host rng of one translation unit after `n_prior_calls` API calls: raymarch_shared.h:38 + ray_sampler.cu:198 */
static pcg32_t host_rng(uint64_t seed, int64_t n_prior_calls) {
    pcg32_t r; pcg_seed(&r, seed, 1u);
    for (int64_t k = 0; k < n_prior_calls; ++k) pcg_advance(&r, 1ull << 32);
    return r;
}
void oracle_pcg32_floats(uint64_t seed, int64_t n_prior_calls, uint64_t advance, int n, float *out) {
    pcg32_t r = host_rng(seed, n_prior_calls); pcg_advance(&r, advance);
    for (int i = 0; i < n; ++i) out[i] = pcg_next_float(&r);
}

/* ---- small helpers */
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); } /* raymarch_shared.h:104-107 */
static inline float signf_(float x) { return copysignf(1.0f, x); }                                    /* raymarch_shared.h:166-169 */
static inline float calc_dt(float t, float cone) { return clampf(t * cone, MIN_CONE_STEPSIZE(), MAX_CONE_STEPSIZE()); } /* ray_sampler_header.h:24-25 */
static inline uint32_t expand_bits(uint32_t v) {                                                      /* raymarch_shared.h:753-760 */
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); } /* :762-768 */
static inline uint32_t morton3D_invert(uint32_t x) {                                                  /* raymarch_shared.h:122-131 */
    x = x & 0x49249249; x = (x | (x >> 2)) & 0xc30c30c3; x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff; x = (x | (x >> 16)) & 0x0000ffff; return x;
}
uint32_t oracle_morton3D(uint32_t x, uint32_t y, uint32_t z) { return morton3D(x, y, z); }
uint32_t oracle_morton3D_invert(uint32_t x) { return morton3D_invert(x); }

static inline int mip_from_pos(const float p[3]) {                                                    /* ray_sampler_header.h:37-43 */
    int e; float m = fmaxf(fmaxf(fabsf(p[0] - 0.5f), fabsf(p[1] - 0.5f)), fabsf(p[2] - 0.5f));
    frexpf(m, &e);
    int v = e + 1; if (v < 0) v = 0; if (v > (int)NERF_CASCADES - 1) v = NERF_CASCADES - 1; return v;
}
static inline int mip_from_dt(float dt, const float p[3]) {                                           /* ray_sampler_header.h:45-54 */
    int mip = mip_from_pos(p);
    dt *= 2 * NERF_GRIDSIZE;
    if (dt < 1.f) return mip;
    int e; frexpf(dt, &e);
    int v = e > mip ? e : mip; if (v > (int)NERF_CASCADES - 1) v = NERF_CASCADES - 1; return v;
}
static inline uint32_t cascaded_grid_idx_at(const float pos[3], uint32_t mip) {                       /* ray_sampler_header.h:298-313 */
    float s = scalbnf(1.0f, -(int)mip); int c[3];
    for (int k = 0; k < 3; ++k) {
        float q = pos[k] - 0.5f; q *= s; q += 0.5f;
        int i = (int)(q * NERF_GRIDSIZE);
        float cl = clampf((float)i, 0.f, (float)(NERF_GRIDSIZE - 1)); /* the reference's clamp() is the float overload */
        c[k] = (int)cl;
    }
    return morton3D((uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2]);
}
static inline int occupied_at(const float pos[3], const uint8_t *bitfield, uint32_t mip) {            /* ray_sampler_header.h:315-319 */
    uint32_t idx = cascaded_grid_idx_at(pos, mip);
    return bitfield[idx / 8 + (GRID_CELLS * mip) / 8] & (1 << (idx % 8));
}
static inline float distance_to_next_voxel(const float pos[3], const float d[3], const float idir[3], uint32_t res) { /* :271-280 */
    float t3[3];
    for (int k = 0; k < 3; ++k) { float p = res * pos[k]; t3[k] = (floorf(p + 0.5f + 0.5f * signf_(d[k])) - p) * idir[k]; }
    float t = fminf(fminf(t3[0], t3[1]), t3[2]);
    return fmaxf(t / res, 0.0f);
}
static inline float advance_to_next_voxel(float t, float cone, const float pos[3], const float d[3], const float idir[3], uint32_t res) { /* :282-296 */
    float t_target = t + distance_to_next_voxel(pos, d, idir, res);
    do { t += calc_dt(t, cone); } while (t < t_target);
    return t;
}
/* BoundingBox::ray_intersect, raymarch_shared.h:506-563 (division by dir, ordered swaps) */
static inline void aabb_ray_intersect(float lo, float hi, const float o[3], const float d[3], float *tmin_out, float *tmax_out) {
    float tmin = (lo - o[0]) / d[0], tmax = (hi - o[0]) / d[0];
    if (tmin > tmax) { float c = tmin; tmin = tmax; tmax = c; }
    float tymin = (lo - o[1]) / d[1], tymax = (hi - o[1]) / d[1];
    if (tymin > tymax) { float c = tymin; tymin = tymax; tymax = c; }
    if (tmin > tymax || tymin > tmax) { *tmin_out = FLT_MAX; *tmax_out = FLT_MAX; return; }
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (lo - o[2]) / d[2], tzmax = (hi - o[2]) / d[2];
    if (tzmin > tzmax) { float c = tzmin; tzmin = tzmax; tzmax = c; }
    if (tmin > tzmax || tzmin > tmax) { *tmin_out = FLT_MAX; *tmax_out = FLT_MAX; return; }
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tmin_out = tmin; *tmax_out = tmax;
}
static inline int aabb_contains(float lo, float hi, const float p[3]) {                               /* raymarch_shared.h:570-575 */
    return p[0] >= lo && p[0] <= hi && p[1] >= lo && p[1] <= hi && p[2] >= lo && p[2] <= hi;
}
static inline float warp_dt(float dt) {                                                               /* raymarch_shared.h:110-114 */
    float max_stepsize = MIN_CONE_STEPSIZE() * (1 << (NERF_CASCADES - 1));
    return (dt - MIN_CONE_STEPSIZE()) / (max_stepsize - MIN_CONE_STEPSIZE());
}
static inline float unwarp_dt(float dt) {                                                             /* ray_sampler_header.h:388-392 */
    float max_stepsize = MIN_CONE_STEPSIZE() * (1 << (NERF_CASCADES - 1));
    return dt * (max_stepsize - MIN_CONE_STEPSIZE()) + MIN_CONE_STEPSIZE();
}
static inline float logistic(float x) { return 1.0f / (1.0f + expf(-x)); }                            /* raymarch_shared.h:615-618 */
static inline float net_to_rgb(float v, int act) {                                                    /* ray_sampler_header.h:440-456 */
    switch (act) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return logistic(v); case 3: return expf(clampf(v, -10.f, 10.f)); }
    return 0.f;
}
static inline float net_to_density(float v, int act) {                                                /* raymarch_shared.h:626-642 */
    switch (act) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return logistic(v); case 3: return expf(v); }
    return 0.f;
}
static inline float net_to_rgb_deriv(float v, int act) {                                              /* ray_sampler_header.h:534-553 */
    switch (act) { case 0: return 1.f; case 1: return v > 0.f ? 1.f : 0.f; case 2: { float s = logistic(v); return s * (1 - s); } case 3: return expf(clampf(v, -10.f, 10.f)); }
    return 0.f;
}
static inline float net_to_density_deriv(float v, int act) {                                          /* ray_sampler_header.h:555-574 */
    switch (act) { case 0: return 1.f; case 1: return v > 0.f ? 1.f : 0.f; case 2: { float s = logistic(v); return s * (1 - s); } case 3: return expf(clampf(v, -15.f, 15.f)); }
    return 0.f;
}

/* one march of one ray; emit != 0 writes up to `limit` NerfCoordinate rows. src/ray_sampler.cu:58-72 (count) and :99-115 (emit) */
static uint32_t march_ray(const float o[3], const float d[3], float lo, float hi, float startt, float cone,
                          const uint8_t *bitfield, uint32_t limit, float *coords /* [limit,7] or NULL */) {
    float idir[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float wdir[3] = {(d[0] + 1.0f) * 0.5f, (d[1] + 1.0f) * 0.5f, (d[2] + 1.0f) * 0.5f}; /* warp_direction, ray_sampler_header.h:362-365 */
    float diag = hi - lo;
    uint32_t j = 0; float t = startt; float pos[3];
    for (;;) {
        for (int k = 0; k < 3; ++k) pos[k] = o[k] + t * d[k];
        if (!(aabb_contains(lo, hi, pos) && j < limit)) break;
        float dt = calc_dt(t, cone);
        uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
        if (occupied_at(pos, bitfield, mip)) {
            if (coords) {
                float *c = coords + 7 * (size_t)j;
                for (int k = 0; k < 3; ++k) c[k] = (pos[k] - lo) / diag;     /* warp_position = relative_pos, raymarch_shared.h:457-460,610-613 */
                c[3] = warp_dt(dt);
                for (int k = 0; k < 3; ++k) c[4 + k] = wdir[k];
            }
            ++j; t += dt;
        } else {
            uint32_t res = NERF_GRIDSIZE >> mip;
            t = advance_to_next_voxel(t, cone, pos, d, idir, res);
        }
    }
    return j;
}

/* rays_sampler_cuda + rays_sampler_api, src/ray_sampler.cu:5-116,118-200.
 * counters[0] = rays with >=0 slot reserved (ray_counter), counters[1] = total reserved samples. */
void oracle_rays_sampler(const float *rays_o, const float *rays_d, const uint8_t *bitfield, int n_rays, int max_samples,
                         float aabb0, float aabb1, float near_distance, float cone, uint64_t seed, int64_t n_prior_calls,
                         float *coords_out, int32_t *rays_index, int32_t *numsteps, int32_t *counters) {
    uint32_t ray_counter = (uint32_t)counters[0], step_counter = (uint32_t)counters[1];
    for (int i = 0; i < n_rays; ++i) {
        pcg32_t r = host_rng(seed, n_prior_calls);
        pcg_advance(&r, (uint64_t)((uint32_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY));   /* :31 (uint32 product) */
        const float *o = rays_o + 3 * (size_t)i, *d = rays_d + 3 * (size_t)i;
        float tmin, tmax; aabb_ray_intersect(aabb0, aabb1, o, d, &tmin, &tmax);     /* :43 */
        tmin = fmaxf(tmin, near_distance);                                          /* :47 */
        float startt = tmin; startt += calc_dt(startt, cone) * pcg_next_float(&r);  /* :49-51 */
        uint32_t n = march_ray(o, d, aabb0, aabb1, startt, cone, bitfield, NERF_STEPS, NULL);
        uint32_t base = step_counter; step_counter += n;                            /* :75 */
        if (base + n > (uint32_t)max_samples) { numsteps[2 * i] = 0; numsteps[2 * i + 1] = (int32_t)base; continue; } /* :76-82 */
        uint32_t ridx = ray_counter++; rays_index[i] = (int32_t)ridx;               /* :86-87 */
        numsteps[2 * i] = (int32_t)n; numsteps[2 * i + 1] = (int32_t)base;
        if (n == 0) { rays_index[i] = -1; continue; }                               /* :91-95 */
        march_ray(o, d, aabb0, aabb1, startt, cone, bitfield, n, coords_out + 7 * (size_t)base);
    }
    counters[0] = (int32_t)ray_counter; counters[1] = (int32_t)step_counter;
}

/* compacted_coord_cuda, src/compacted_coord.cu:5-77: truncating gather in ray order (the T loop is dead work, Q3). */
void oracle_compacted_coord(const float *coords_in, const int32_t *numsteps, int n_rays, int max_compacted,
                            float *coords_out, int32_t *numsteps_c, int32_t *ray_counter, int32_t *step_counter) {
    uint32_t counter = (uint32_t)*step_counter, rc = (uint32_t)*ray_counter, T = (uint32_t)max_compacted;
    for (int i = 0; i < n_rays; ++i) {
        uint32_t n = (uint32_t)numsteps[2 * i], base = (uint32_t)numsteps[2 * i + 1];
        uint32_t cbase = counter; counter += n;                                     /* :63 */
        uint32_t mb = T < cbase ? T : cbase; uint32_t room = T - mb; uint32_t nc = room < n ? room : n; /* :64 */
        numsteps_c[2 * i] = (int32_t)nc; numsteps_c[2 * i + 1] = (int32_t)cbase;
        if (nc == 0) continue;
        rc++;
        memcpy(coords_out + 7 * (size_t)cbase, coords_in + 7 * (size_t)base, sizeof(float) * 7 * nc);
    }
    *step_counter = (int32_t)counter; *ray_counter = (int32_t)rc;
}

/* compute_rgbs, src/calc_rgb.cu:5-67 */
void oracle_calc_rgb_forward(const float *raw, const float *coords, const int32_t *numsteps, const int32_t *numsteps_c, const float *bg,
                             int n_rays, int rgb_act, int dens_act, float *rgb_out) {
#pragma omp parallel for schedule(static, 64)
    for (int i = 0; i < n_rays; ++i) {
        uint32_t n = (uint32_t)numsteps_c[2 * i], base = (uint32_t)numsteps_c[2 * i + 1];
        const float *b = bg + 3 * (size_t)i; float *out = rgb_out + 3 * (size_t)i;
        if (n == 0) { out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; continue; }
        float T = 1.f, acc[3] = {0, 0, 0}; uint32_t k = 0;
        for (; k < n; ++k) {
            const float *r = raw + 4 * (size_t)(base + k); float dt = unwarp_dt(coords[7 * (size_t)(base + k) + 3]);
            float density = net_to_density(r[3], dens_act);
            float alpha = 1.f - expf(-density * dt), w = alpha * T;
            for (int c = 0; c < 3; ++c) acc[c] += w * net_to_rgb(r[c], rgb_act);
            T *= (1.f - alpha);
        }
        if (k == (uint32_t)numsteps[2 * i]) for (int c = 0; c < 3; ++c) acc[c] += T * b[c]; /* :61-64 */
        out[0] = acc[0]; out[1] = acc[1]; out[2] = acc[2];
    }
}

/* compute_rgbs_grad, src/calc_rgb.cu:70-140 */
void oracle_calc_rgb_backward(const float *raw, const int32_t *numsteps_c, const float *coords, const float *grad_rgb, const float *rgb,
                              const float *grid_mean, int n_rays, int rgb_act, int dens_act, float *dl_draw) {
    float loss_scale = 128; loss_scale /= n_rays;                                   /* :92-93 */
    const float l2 = rgb_act == 3 ? 1e-4f : 0.0f;                                   /* :103 */
    const float l1 = *grid_mean < NERF_MIN_OPTICAL_THICKNESS ? 1e-4f : 0.0f;        /* :104 */
#pragma omp parallel for schedule(static, 64)
    for (int i = 0; i < n_rays; ++i) {
        uint32_t n = (uint32_t)numsteps_c[2 * i], base = (uint32_t)numsteps_c[2 * i + 1];
        const float *g = grad_rgb + 3 * (size_t)i, *fin = rgb + 3 * (size_t)i;
        float T = 1.f, rgb2[3] = {0, 0, 0};
        for (uint32_t k = 0; k < n; ++k) {
            const float *r = raw + 4 * (size_t)(base + k); float *o = dl_draw + 4 * (size_t)(base + k);
            float c[3]; for (int q = 0; q < 3; ++q) c[q] = net_to_rgb(r[q], rgb_act);
            float dt = unwarp_dt(coords[7 * (size_t)(base + k) + 3]);
            float density = net_to_density(r[3], dens_act);
            float alpha = 1.f - expf(-density * dt), w = alpha * T;
            for (int q = 0; q < 3; ++q) rgb2[q] += w * c[q];
            T *= (1.f - alpha);
            float term3[3];
            for (int q = 0; q < 3; ++q) {
                float suffix = fin[q] - rgb2[q];
                float dl_drgb = w * g[q];
                o[q] = loss_scale * (dl_drgb * net_to_rgb_deriv(r[q], rgb_act) + fmaxf(0.0f, l2 * r[q]));
                term3[q] = g[q] * (T * c[q] - suffix);
            }
            /* Eigen's fixed-size-3 redux splits [0,1)+[1,3): p0 + (p1 + p2) (Eigen/src/Core/Redux.h redux_novec_unroller) */
            float dot = term3[0] + (term3[1] + term3[2]);
            float dd = net_to_density_deriv(r[3], dens_act);
            float dl_dmlp = dd * (dt * dot);
            o[3] = loss_scale * dl_dmlp + (r[3] < 0 ? -l1 : 0.0f);
        }
    }
}

/* compute_rgbs_inference, src/calc_rgb.cu:143-206 */
void oracle_calc_rgb_inference(const float *raw, const float *coords, const int32_t *numsteps, const float *bg3, int n_rays,
                               int rgb_act, int dens_act, float *rgb_out, float *alpha_out) {
#pragma omp parallel for schedule(static, 64)
    for (int i = 0; i < n_rays; ++i) {
        uint32_t n = (uint32_t)numsteps[2 * i], base = (uint32_t)numsteps[2 * i + 1];
        float *out = rgb_out + 3 * (size_t)i;
        if (n == 0) { out[0] = bg3[0]; out[1] = bg3[1]; out[2] = bg3[2]; alpha_out[i] = 0; continue; }
        float T = 1.f, acc[3] = {0, 0, 0};
        for (uint32_t k = 0; k < n; ++k) {
            const float *r = raw + 4 * (size_t)(base + k); float dt = unwarp_dt(coords[7 * (size_t)(base + k) + 3]);
            float density = net_to_density(r[3], dens_act);
            float alpha = 1.f - expf(-density * dt), w = alpha * T;
            for (int c = 0; c < 3; ++c) acc[c] += w * net_to_rgb(r[c], rgb_act);
            T *= (1.f - alpha);
        }
        for (int c = 0; c < 3; ++c) out[c] = acc[c] + T * bg3[c];
        alpha_out[i] = 1 - T;
    }
}

/* mark_untrained_density_grid_cuda, src/mark_untrained_density_grid.cu:5-51. `grid` must be pre-filled (Q1). */
void oracle_mark_untrained(const float *focal /*[I,2]*/, const float *xforms /*[I,4,3] = column-major 3x4*/, int n_elements, int n_images,
                           int res0, int res1, float *grid) {
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < n_elements; ++i) {
        uint32_t level = (uint32_t)i / GRID_CELLS, pos_idx = (uint32_t)i % GRID_CELLS;
        uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
        float half_resx = res0 * 0.5f, half_resy = res1 * 0.5f;
        float s = scalbnf(1.0f, (int)level);
        float pos[3] = {(((float)x + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f, (((float)y + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
                        (((float)z + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
        float voxel_radius = 0.5f * SQRT3 * s / NERF_GRIDSIZE;
        int count = 0;
        for (int j = 0; j < n_images; ++j) {
            const float *m = xforms + 12 * (size_t)j;          /* Eigen 3x4 column-major: col c = m[3c..3c+2] */
            float pl[3] = {pos[0] - m[9], pos[1] - m[10], pos[2] - m[11]};
            /* Eigen fixed-size-3 dot = p0 + (p1 + p2) */
            float cx = pl[0] * m[0] + (pl[1] * m[1] + pl[2] * m[2]);
            float cy = pl[0] * m[3] + (pl[1] * m[4] + pl[2] * m[5]);
            float cz = pl[0] * m[6] + (pl[1] * m[7] + pl[2] * m[8]);
            if (cz > 0.f) {
                const float *f = focal + 2 * (size_t)j;
                if (fabsf(cx) - voxel_radius < cz / f[0] * half_resx && fabsf(cy) - voxel_radius < cz / f[1] * half_resy) { count++; break; }
            }
        }
        if ((grid[i] < 0) != (count <= 0)) grid[i] = (count > 0) ? 0.f : -1.f;
    }
}

/* generate_grid_samples_nerf_nonuniform_cuda, src/generate_grid_samples_nerf_nonuniform.cu:6-42 */
void oracle_generate_grid_samples(const float *grid, int step, int n_elements, int max_cascade, float thresh, float aabb0, float aabb1,
                                  uint64_t seed, int64_t n_prior_calls, float *positions, int32_t *indices) {
    uint32_t n_cascades = (uint32_t)max_cascade + 1, n = (uint32_t)n_elements;
    float diag = aabb1 - aabb0;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t ii = 0; ii < n_elements; ++ii) {
        uint32_t i = (uint32_t)ii;
        pcg32_t r = host_rng(seed, n_prior_calls); pcg_advance(&r, (uint64_t)(i * 4u));
        uint32_t level = (uint32_t)(pcg_next_float(&r) * n_cascades) % n_cascades;
        uint32_t idx = 0;
        for (uint32_t j = 0; j < 10; ++j) {
            idx = ((i + (uint32_t)step * n) * 56924617u + j * 19349663u + 96925573u) % GRID_CELLS;
            idx += level * GRID_CELLS;
            if (grid[idx] > thresh) break;
        }
        uint32_t pos_idx = idx % GRID_CELLS;
        uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
        float u0 = pcg_next_float(&r), u1 = pcg_next_float(&r), u2 = pcg_next_float(&r);
        float s = scalbnf(1.0f, (int)level);
        float p[3] = {(((float)x + u0) / NERF_GRIDSIZE - 0.5f) * s + 0.5f, (((float)y + u1) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
                      (((float)z + u2) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
        for (int k = 0; k < 3; ++k) positions[3 * (size_t)i + k] = (p[k] - aabb0) / diag;
        indices[i] = (int32_t)idx;
    }
}

/* splat_grid_samples_nerf_max_nearest_neighbor_cuda, src/splat_...cu:6-27 (density activation hard-wired Exponential, :49) */
void oracle_splat(const float *mlp_out, const int32_t *indices, int padded_width, int n, float *grid_tmp) {
    for (int i = 0; i < n; ++i) {
        float mlp = expf(mlp_out[(size_t)i * padded_width]);
        float thickness = mlp * scalbnf(MIN_CONE_STEPSIZE(), 0);
        uint32_t u; memcpy(&u, &thickness, 4);
        uint32_t *cell = (uint32_t *)&grid_tmp[(uint32_t)indices[i]];
        if (*cell < u) *cell = u;
    }
}

/* ema_grid_samples_nerf_cuda, src/ema_grid_samples_nerf.cu:3-26 */
void oracle_ema(const float *grid_tmp, int n_elements, float decay, float *grid) {
#pragma omp parallel for schedule(static, 65536)
    for (int64_t i = 0; i < n_elements; ++i) {
        float imp = grid_tmp[i], prev = grid[i];
        grid[i] = (prev < 0.f) ? prev : fmaxf(prev * decay, imp);
    }
}

/* update_bitfield_api, src/update_bitfield.cu:24-116: mean of max(level0,0)/N (any summation order is the
 * reference's — its block sums are atomically accumulated), threshold, 8 cells -> 1 byte over ALL cascades, then
 * 7 OR-pools written at the centred (x+16,y+16,z+16) cell of the next level (|= onto that level's own bits). */
void oracle_update_bitfield(const float *grid, float *mean, uint8_t *bitfield) {
    /* blocked pairwise order that mirrors reduce_sum's 1024-thread x float4 blocks (raymarch_shared.h:657-745) */
    float total = 0.f;
    for (uint32_t b = 0; b < GRID_CELLS / 4096; ++b) {
        float v[1024];
        for (uint32_t t = 0; t < 1024; ++t) {
            const float *q = grid + 4 * ((size_t)b * 1024 + t);
            v[t] = ((fmaxf(q[0], 0.f) / GRID_CELLS + fmaxf(q[1], 0.f) / GRID_CELLS) + fmaxf(q[2], 0.f) / GRID_CELLS) + fmaxf(q[3], 0.f) / GRID_CELLS;
        }
        float sdata[32];
        for (uint32_t w = 0; w < 32; ++w) {
            float *x = v + 32 * w;
            for (int off = 16; off > 0; off /= 2) { float tmp[32]; for (int l = 0; l < 32; ++l) tmp[l] = x[l] + x[l ^ off]; memcpy(x, tmp, sizeof tmp); }
            sdata[w] = x[0];
        }
        for (int off = 16; off > 0; off /= 2) { float tmp[32]; for (int l = 0; l < 32; ++l) tmp[l] = sdata[l] + sdata[l ^ off]; memcpy(sdata, tmp, sizeof tmp); }
        total += sdata[0];
    }
    mean[0] = total;
    float thresh = NERF_MIN_OPTICAL_THICKNESS < total ? NERF_MIN_OPTICAL_THICKNESS : total;
    for (uint32_t i = 0; i < GRID_CELLS / 8 * NERF_CASCADES; ++i) {
        uint8_t bits = 0;
        for (int j = 0; j < 8; ++j) bits |= grid[(size_t)i * 8 + j] > thresh ? (uint8_t)(1 << j) : 0;
        bitfield[i] = bits;
    }
    for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
        const uint8_t *prev = bitfield + (GRID_CELLS * (level - 1)) / 8; uint8_t *next = bitfield + (GRID_CELLS * level) / 8;
        for (uint32_t i = 0; i < GRID_CELLS / 64; ++i) {
            uint8_t bits = 0;
            for (int j = 0; j < 8; ++j) bits |= prev[(size_t)i * 8 + j] > 0 ? (uint8_t)(1 << j) : 0;
            uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8, y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8, z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
            next[morton3D(x, y, z)] |= bits;
        }
    }
}
