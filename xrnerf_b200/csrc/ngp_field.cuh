// xrnerf_b200 — device building blocks of the Instant-NGP field (hash-grid encoding, SH, tiny MLP) shared by the
// stand-alone tcnn-shaped kernels (ngp_mlp.cu) and the fused render kernel (ngp_render.cu).
//
// Semantics restated from tiny-cuda-nn (see oracle/tcnn_oracle.c header: the reference delegates this arithmetic to an
// unpinned third-party package; call sites /root/reference/xrnerf/models/mlps/hashnerf_mlp.py:36-45,:55-79).
#pragma once
#include "common.cuh"

namespace xrb {

constexpr int MAX_LEVELS = 16;
struct HashGridDev {
    uint32_t offset[MAX_LEVELS + 1];  // in feature vectors
    float scale[MAX_LEVELS];
    uint32_t res[MAX_LEVELS];
    int n_levels;
    uint32_t hashed_mask;             // bit l: level l is hashed (its table is the 2^log2_hashmap_size cap), else dense
};

inline int64_t hashgrid_build(const xrb_ngp_config *cfg, HashGridDev *g) {
    g->n_levels = cfg->n_levels; g->hashed_mask = 0;
    const float log2_pls = log2f(cfg->per_level_scale);
    uint32_t off = 0;
    for (int l = 0; l < cfg->n_levels; ++l) {
        float scale = exp2f(l * log2_pls) * (float)cfg->base_resolution - 1.0f;
        uint32_t res = (uint32_t)ceilf(scale) + 1;
        uint64_t cube = (uint64_t)res * res * res; uint32_t maxp = 0xffffffffu / 2;
        uint32_t p = cube > maxp ? maxp : (uint32_t)cube;
        p = (p + 7u) / 8u * 8u;
        uint32_t cap = 1u << cfg->log2_hashmap_size; if (p > cap) p = cap;
        if (cube > p) g->hashed_mask |= 1u << l;   // then p == cap, a power of two: `% hashmap_size` is a mask
        g->offset[l] = off; g->scale[l] = scale; g->res[l] = res; off += p;
    }
    g->offset[cfg->n_levels] = off;
    return (int64_t)off * cfg->n_features;
}
__host__ __device__ inline int64_t mlp_num_params(int in_w, int width, int n_hidden, int out_pad) { return (int64_t)width * in_w + (int64_t)(n_hidden - 1) * width * width + (int64_t)out_pad * width; }

inline int check_cfg(const xrb_ngp_config *cfg) {
    if (!cfg) { set_error("ngp config is null"); return XRB_E_BADARG; }
    if (cfg->n_levels != 16 || cfg->n_features != 2 || cfg->width != 64 || cfg->density_hidden < 1 || cfg->density_hidden > 4 || cfg->color_hidden < 1 || cfg->color_hidden > 4 ||
        cfg->log2_hashmap_size < 8 || cfg->log2_hashmap_size > 24 || cfg->base_resolution < 1) {
        set_error("ngp config unsupported: need n_levels=16, n_features=2, n_neurons=64, 1..4 hidden layers");
        return XRB_E_UNSUPPORTED;
    }
    return XRB_OK;
}

// defined in ngp_mlp.cu: launches the field (impl 0 CUDA cores / 1 tcgen05) on n rows; if n_dev != NULL the effective row
// count is min(n, *n_dev) read on the device.
int launch_field(const xrb_ngp_config *cfg, const void *table, const void *dens, const void *color, const void *image, const float *pts, int pts_stride, const float *dirs,
                 int dirs_stride, int n, const int32_t *n_dev, float *out, int impl, bool density_only, cudaStream_t s);

#ifdef __CUDACC__
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t hashmap_size, uint32_t res) {
    // tcnn grid.h grid_index(): dense strides while they fit, else the coherent prime hash; then `% hashmap_size`.
    // The modulo is strength-reduced without changing its value (an integer division costs 3 XU-pipe ops on sm_100):
    //   hashed levels have hashmap_size == 2^log2_hashmap_size (the cap)          -> mask (never the generic `%`: it costs ~130 issue slots);
    //   dense levels have index <= res^3 + res^2 + res < 2 * hashmap_size          -> one conditional subtract.
    if ((uint64_t)res * res * res > hashmap_size) {   // uniform per level
        uint32_t h = x ^ (y * 2654435761u) ^ (z * 805459861u);
        return h & (hashmap_size - 1);   // a hashed level's table is always the 2^log2_hashmap_size cap (hashgrid_build)
    }
    uint32_t index = x + y * res + z * res * res;
    return index >= hashmap_size ? index - hashmap_size : index;
}

// one level of the multiresolution hash encoding: returns the two interpolated features (fp32, NOT yet rounded).
// Index arithmetic is tcnn's grid_index() strength-reduced without changing any value:
//   hashed level (res^3 > table size == 2^log2_hashmap_size): (x ^ y*P1 ^ z*P2) & (size-1); the two products are shared by the 8 corners
//   dense level: x + y*res + z*res^2 < 2*size -> one conditional subtract instead of `% size`
// (the generic `%` compiled to ~130 predicated-off but ISSUED instructions per level: ncu r01d source view)
__device__ __forceinline__ float2 hash_level(const __half2 *__restrict__ table, const HashGridDev &g, int l, float x, float y, float z) {
    const uint32_t off = g.offset[l], hs = g.offset[l + 1] - off, res = g.res[l];
    const __half2 *tl = table + off;
    asm volatile("" : "+l"(tl));   // keep the level base in a register pair: each gather address is then ONE imad.wide (base + idx*4), not a 64-bit add chain
    const float sc = g.scale[l];
    float px = __fmaf_rn(sc, x, 0.5f), py = __fmaf_rn(sc, y, 0.5f), pz = __fmaf_rn(sc, z, 0.5f);
    int ix_, iy_, iz_;
    float fx = floor_small(px, &ix_), fy = floor_small(py, &iy_), fz = floor_small(pz, &iz_);
    const uint32_t gx = (uint32_t)ix_, gy = (uint32_t)iy_, gz = (uint32_t)iz_;
    fx = px - fx; fy = py - fy; fz = pz - fz;
    __half2 v[8];
    if ((g.hashed_mask >> l) & 1u) {   // uniform per level
        const uint32_t mask = hs - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
        const uint32_t a00 = hy0 ^ hz0, a10 = hy1 ^ hz0, a01 = hy0 ^ hz1, a11 = hy1 ^ hz1, gx1 = gx + 1u;
        v[0] = __ldg(tl + ((gx ^ a00) & mask)); v[1] = __ldg(tl + ((gx1 ^ a00) & mask));
        v[2] = __ldg(tl + ((gx ^ a10) & mask)); v[3] = __ldg(tl + ((gx1 ^ a10) & mask));
        v[4] = __ldg(tl + ((gx ^ a01) & mask)); v[5] = __ldg(tl + ((gx1 ^ a01) & mask));
        v[6] = __ldg(tl + ((gx ^ a11) & mask)); v[7] = __ldg(tl + ((gx1 ^ a11) & mask));
    } else {
        const uint32_t sy = res, sz = res * res, b = gx + gy * sy + gz * sz;
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // all 8 gathers are issued before any is consumed (8 independent loads in flight per level)
            uint32_t idx = b + (c & 1) + ((c >> 1) & 1) * sy + ((c >> 2) & 1) * sz;
            idx = idx >= hs ? idx - hs : idx;
            v[c] = __ldg(tl + idx);
        }
    }
    const float ux = 1.f - fx, uy = 1.f - fy, uz = 1.f - fz;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = ((c & 1) ? fx : ux) * ((c & 2) ? fy : uy) * ((c & 4) ? fz : uz);
        float2 f = __half22float2(v[c]);
        acc.x += w * f.x; acc.y += w * f.y;
    }
    return acc;
}

// SH degree 4 on dir*2-1 (tcnn spherical_harmonics.h), fp32 values (caller rounds to fp16)
__device__ __forceinline__ void sh4(float dx01, float dy01, float dz01, float *out) {
    float x = dx01 * 2.f - 1.f, y = dy01 * 2.f - 1.f, z = dz01 * 2.f - 1.f;
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    out[0] = 0.28209479177387814f;
    out[1] = -0.48860251190291987f * y;
    out[2] = 0.48860251190291987f * z;
    out[3] = -0.48860251190291987f * x;
    out[4] = 1.0925484305920792f * xy;
    out[5] = -1.0925484305920792f * yz;
    out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    out[7] = -1.0925484305920792f * xz;
    out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    out[10] = 2.8906114426405538f * xy * z;
    out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    out[14] = 1.4453057213202769f * z * (x2 - y2);
    out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// ---------------------------------------------------------------------------- UMMA shared-memory image helpers
// K-major, 128-byte swizzle: row r (one sample / one output neuron) occupies 128 bytes = 64 halfs; 8-row atoms of 1024 B.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {  // byte offset of 16-byte chunk `chunk16` (0..7) of `row`
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk16 ^ (row & 7u)) << 4);
}
struct WeightImageLayout {  // byte offsets of each matrix inside the packed image
    uint32_t d_in, d_hid[3], d_out, c_in, c_hid[3], c_out, total;
};
__host__ __device__ inline WeightImageLayout weight_image_layout(int density_hidden, int color_hidden) {
    WeightImageLayout L{}; uint32_t o = 0;
    L.d_in = o; o += 64 * 128;
    for (int k = 0; k < density_hidden - 1; ++k) { L.d_hid[k] = o; o += 64 * 128; }
    L.d_out = o; o += 16 * 128;
    L.c_in = o; o += 64 * 128;
    for (int k = 0; k < color_hidden - 1; ++k) { L.c_hid[k] = o; o += 64 * 128; }
    L.c_out = o; o += 16 * 128;
    L.total = o;
    return L;
}
#endif  // __CUDACC__

}  // namespace xrb
