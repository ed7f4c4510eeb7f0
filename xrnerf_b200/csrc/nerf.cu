// xrnerf_b200 — NeRF / Mip-NeRF per-ray kernels that are pure PyTorch in the reference (≈25 tiny ATen kernels per
// render call, a full sort + expanded gathers per sample_pdf call): one launch each, one WARP per ray,
// lanes over samples with shuffle scans. References relative to /root/reference/xrnerf/models/.
#include "common.cuh"
#include <stdlib.h>

namespace xrb {

__device__ __forceinline__ float wincl_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    return v;
}
__device__ __forceinline__ float wincl_prod(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v *= n; }
    return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus(beta=1, threshold=20)

struct CompositeParams { int n_rays, n_samples, mip, white_bkgd, density_act; float rgb_padding, density_bias; };

// per-sample quantities of renders/nerf_render.py:61-83 (and mipnerf_render.py:27-33 when mip)
__device__ __forceinline__ void sample_terms(const CompositeParams &p, const float *__restrict__ raw, const float *__restrict__ z, float dnorm, int k, float4 &r, float &sd,
                                             float &alpha) {
    r = __ldg(reinterpret_cast<const float4 *>(raw) + k);
    float dist = p.mip ? (z[k + 1] - z[k]) : (k + 1 < p.n_samples ? z[k + 1] - z[k] : 1e10f);
    dist *= dnorm;
    float a = r.w + p.density_bias;
    float dens = p.density_act == XRB_ACT_SOFTPLUS ? softplusf_(a) : fmaxf(a, 0.f);
    sd = dens * dist;
    alpha = 1.f - expf(-sd);
}

// BACKWARD == false: rgb/disp/acc/weights.  BACKWARD == true: d_raw given grad_rgb (loss on rgb only, nerf.py:79-84 / mipnerf.py:49-57)
template <bool BACKWARD>
__global__ void __launch_bounds__(256) nerf_composite_kernel(CompositeParams p, const float *__restrict__ raw, const float *__restrict__ z_vals, const float *__restrict__ rays_d,
                                                             const float *__restrict__ grad_rgb, float *__restrict__ rgb_out, float *__restrict__ disp_out, float *__restrict__ acc_out,
                                                             float *__restrict__ weights_out, float *__restrict__ d_raw) {
    const int lane = threadIdx.x & 31;
    const int ray = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (ray >= p.n_rays) return;
    const int S = p.n_samples, zs = p.mip ? S + 1 : S;
    const float *rw = raw + (size_t)ray * S * 4, *z = z_vals + (size_t)ray * zs;
    const float dx = rays_d[3 * (size_t)ray], dy = rays_d[3 * (size_t)ray + 1], dz = rays_d[3 * (size_t)ray + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float pad_mul = 1.f + 2.f * p.rgb_padding;
    float gx = 0, gy = 0, gz = 0;
    if (BACKWARD) { gx = grad_rgb[3 * (size_t)ray]; gy = grad_rgb[3 * (size_t)ray + 1]; gz = grad_rgb[3 * (size_t)ray + 2]; }
    // ---- pass 1: forward accumulation
    float carry = p.mip ? 0.f : 1.f;  // mip: running sum of sigma*delta; nerf: running product of (1-alpha+1e-10)
    float ax = 0, ay = 0, az = 0, acc = 0, depth = 0, total_wg = 0;
    for (int k0 = 0; k0 < S; k0 += 32) {
        int k = k0 + lane; bool valid = k < S;
        float4 r = make_float4(0, 0, 0, 0); float sd = 0, alpha = 0;
        if (valid) sample_terms(p, rw, z, dnorm, k, r, sd, alpha);
        float T;
        if (p.mip) { float inc = wincl_sum(sd, lane); T = expf(-(carry + inc - sd)); carry += __shfl_sync(0xffffffffu, inc, 31); }
        else { float f = valid ? (1.f - alpha + 1e-10f) : 1.f; float inc = wincl_prod(f, lane); float ex = __shfl_up_sync(0xffffffffu, inc, 1); T = carry * (lane == 0 ? 1.f : ex); carry *= __shfl_sync(0xffffffffu, inc, 31); }
        float w = valid ? alpha * T : 0.f;
        float cx = sigmoidf_(r.x) * pad_mul - p.rgb_padding, cy = sigmoidf_(r.y) * pad_mul - p.rgb_padding, cz = sigmoidf_(r.z) * pad_mul - p.rgb_padding;
        if (!BACKWARD) {
            ax += w * cx; ay += w * cy; az += w * cz; acc += w;
            if (valid) { depth += w * (p.mip ? 0.5f * (z[k] + z[k + 1]) : z[k]); weights_out[(size_t)ray * S + k] = w; }
        } else {
            float wb = p.white_bkgd ? 1.f : 0.f;
            total_wg += w * (gx * (cx - wb) + gy * (cy - wb) + gz * (cz - wb));
        }
    }
    if (!BACKWARD) {
        ax = wsum(ax); ay = wsum(ay); az = wsum(az); acc = wsum(acc); depth = wsum(depth);
        if (lane == 0) {
            if (p.white_bkgd) { ax += 1.f - acc; ay += 1.f - acc; az += 1.f - acc; }
            rgb_out[3 * (size_t)ray] = ax; rgb_out[3 * (size_t)ray + 1] = ay; rgb_out[3 * (size_t)ray + 2] = az;
            acc_out[ray] = acc;
            float q = depth / acc, disp;
            if (p.mip) { if (isnan(q)) q = INFINITY; disp = fmaxf(fminf(q, z[S]), z[0]); }          // mipnerf_render.py:18-24
            else disp = 1.f / fmaxf(1e-10f, q);                                                       // nerf_render.py:32-36 (torch.max propagates NaN)
            if (!p.mip && isnan(q)) disp = q;
            disp_out[ray] = disp;
        }
        return;
    }
    // ---- pass 2 (backward): dL/d(sigma*delta_i) needs the suffix sum of w_k*G_k over k > i
    total_wg = wsum(total_wg);
    carry = p.mip ? 0.f : 1.f;
    float prefix_wg = 0.f;
    for (int k0 = 0; k0 < S; k0 += 32) {
        int k = k0 + lane; bool valid = k < S;
        float4 r = make_float4(0, 0, 0, 0); float sd = 0, alpha = 0;
        if (valid) sample_terms(p, rw, z, dnorm, k, r, sd, alpha);
        float T;
        if (p.mip) { float inc = wincl_sum(sd, lane); T = expf(-(carry + inc - sd)); carry += __shfl_sync(0xffffffffu, inc, 31); }
        else { float f = valid ? (1.f - alpha + 1e-10f) : 1.f; float inc = wincl_prod(f, lane); float ex = __shfl_up_sync(0xffffffffu, inc, 1); T = carry * (lane == 0 ? 1.f : ex); carry *= __shfl_sync(0xffffffffu, inc, 31); }
        float w = valid ? alpha * T : 0.f;
        float sx = sigmoidf_(r.x), sy = sigmoidf_(r.y), sz = sigmoidf_(r.z);
        float cx = sx * pad_mul - p.rgb_padding, cy = sy * pad_mul - p.rgb_padding, cz = sz * pad_mul - p.rgb_padding;
        float wb = p.white_bkgd ? 1.f : 0.f;
        float G = gx * (cx - wb) + gy * (cy - wb) + gz * (cz - wb);
        float inc_wg = wincl_sum(w * G, lane);
        float suffix = total_wg - (prefix_wg + inc_wg);
        prefix_wg += __shfl_sync(0xffffffffu, inc_wg, 31);
        if (valid) {
            float one_m = 1.f - alpha;
            float dL_dsd = p.mip ? (one_m * T * G - suffix) : (one_m * (T * G - suffix / (one_m + 1e-10f)));
            float dist = p.mip ? (z[k + 1] - z[k]) : (k + 1 < S ? z[k + 1] - z[k] : 1e10f);
            dist *= dnorm;
            float a = r.w + p.density_bias;
            float dact = p.density_act == XRB_ACT_SOFTPLUS ? (a > 20.f ? 1.f : sigmoidf_(a)) : (a > 0.f ? 1.f : 0.f);
            float4 o;
            o.x = gx * w * sx * (1.f - sx) * pad_mul; o.y = gy * w * sy * (1.f - sy) * pad_mul; o.z = gz * w * sz * (1.f - sz) * pad_mul;
            o.w = dL_dsd * dist * dact;
            reinterpret_cast<float4 *>(d_raw)[(size_t)ray * S + k] = o;
        }
    }
}

// sample_pdf (networks/utils/hierarchical_sample.py:6-53): one warp per ray, everything in shared memory, bitonic merge sort
constexpr int PDF_WARPS = 4;
constexpr int PDF_MAX_S = 256;   // coarse samples
constexpr int PDF_MAX_OUT = 512; // coarse + importance
__global__ void __launch_bounds__(PDF_WARPS * 32) sample_pdf_kernel(int n_rays, int S, int n_imp, const float *__restrict__ z_vals, const float *__restrict__ weights,
                                                                    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ u_in,
                                                                    float *__restrict__ z_out, float *__restrict__ pts_out) {
    __shared__ float s_cdf[PDF_WARPS][PDF_MAX_S];
    __shared__ float s_bins[PDF_WARPS][PDF_MAX_S];
    __shared__ float s_z[PDF_WARPS][PDF_MAX_OUT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ray = blockIdx.x * PDF_WARPS + wid;
    if (ray >= n_rays) return;
    const float *z = z_vals + (size_t)ray * S, *w = weights + (size_t)ray * S;
    float *cdf = s_cdf[wid], *bins = s_bins[wid], *zz = s_z[wid];
    const int nb = S - 1;   // bins (mid-points): S-1 ; pdf entries: S-2 ; cdf entries: S-1
    float part = 0.f;
    for (int k = lane; k < S - 2; k += 32) part += w[k + 1] + 1e-5f;
    const float wsum_all = wsum(part);
    for (int k = lane; k < nb; k += 32) bins[k] = 0.5f * (z[k + 1] + z[k]);
    __syncwarp();
    if (lane == 0) {  // sequential cumsum like torch.cumsum on the reference's CPU path
        float c = 0.f; cdf[0] = 0.f;
        for (int k = 0; k < S - 2; ++k) { c += (w[k + 1] + 1e-5f) / wsum_all; cdf[k + 1] = c; }
    }
    __syncwarp();
    const int ncdf = S - 1;
    for (int j = lane; j < n_imp; j += 32) {
        // torch.linspace(0,1,n): step = 1/(n-1); second half computed from the end (ATen's symmetric formulation)
        float u;
        if (u_in) u = u_in[(size_t)ray * n_imp + j];
        else { float step = 1.0f / (float)(n_imp - 1); u = j < n_imp / 2 ? step * (float)j : 1.0f - step * (float)(n_imp - 1 - j); }
        int lo = 0, hi = ncdf;   // searchsorted(right=True): first index with cdf[idx] > u
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
        int below = max(0, lo - 1), above = min(ncdf - 1, lo);
        float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
        float denom = c1 - c0; if (denom < 1e-5f) denom = 1.f;
        float t = (u - c0) / denom;
        zz[S + j] = b0 + t * (b1 - b0);
    }
    for (int k = lane; k < S; k += 32) zz[k] = z[k];
    const int total = S + n_imp;
    int np2 = 1; while (np2 < total) np2 <<= 1;
    for (int k = total + lane; k < np2; k += 32) zz[k] = INFINITY;
    __syncwarp();
    for (int size = 2; size <= np2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < np2 / 2; t += 32) {
                int i = 2 * t - (t & (stride - 1)), j = i + stride;
                bool up = ((i & size) == 0);
                float a = zz[i], b = zz[j];
                if ((a > b) == up) { zz[i] = b; zz[j] = a; }
            }
            __syncwarp();
        }
    const float ox = rays_o[3 * (size_t)ray], oy = rays_o[3 * (size_t)ray + 1], oz = rays_o[3 * (size_t)ray + 2];
    const float dx = rays_d[3 * (size_t)ray], dy = rays_d[3 * (size_t)ray + 1], dz = rays_d[3 * (size_t)ray + 2];
    for (int k = lane; k < total; k += 32) {
        float v = zz[k];
        z_out[(size_t)ray * total + k] = v;
        if (pts_out) { float *pp = pts_out + ((size_t)ray * total + k) * 3; pp[0] = ox + dx * v; pp[1] = oy + dy * v; pp[2] = oz + dz * v; }
    }
}

// BaseEmbedder.forward (embedders/base.py:57-74): one thread per output element (coalesced row writes)
__global__ void __launch_bounds__(256) posenc_kernel(int64_t n_pts, int samples_per_ray, int multires, int multires_dirs, const float *__restrict__ pts,
                                                     const float *__restrict__ viewdirs, float *__restrict__ out) {
    const int c_pts = 3 + 6 * multires, c_dir = 3 + 6 * multires_dirs, C = c_pts + c_dir;
    const int64_t total = n_pts * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = idx / C; int c = (int)(idx - i * C);
        const float *src; int cc;
        if (c < c_pts) { src = pts + 3 * i; cc = c; } else { src = viewdirs + 3 * (i / samples_per_ray); cc = c - c_pts; }
        float v;
        if (cc < 3) v = src[cc];
        else { int q = cc - 3, band = q / 6, r = q % 6; float x = src[r % 3] * exp2f((float)band); v = r < 3 ? sinf(x) : cosf(x); }
        out[idx] = v;
    }
}


// Mip-NeRF: cast_rays (networks/utils/mip.py:117-129, conical frustum -> Gaussian :91-114, lift_gaussian :66-88, diag) fused with
// MipNerfEmbedder.forward (embedders/mipnerf_embedder.py:43-99): one thread per output element of embedded[N*S, 6*n_deg + C_dir].
// The Gaussians (means/covs) are recomputed per element in registers and optionally also written out (API parity: data['samples']).
__device__ __forceinline__ void frustum_gaussian(float t0, float t1, float radius, const float d[3], const float o[3], float dmag_sq, float mean[3], float cov[3]) {
    float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
    float mu2 = mu * mu, hw2 = hw * hw, hw4 = hw2 * hw2, den = 3.f * mu2 + hw2;
    float t_mean = mu + (2.f * mu * hw2) / den;
    float t_var = hw2 / 3.f - (4.f / 15.f) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
    float r_var = radius * radius * (mu2 / 4.f + (5.f / 12.f) * hw2 - 4.f / 15.f * hw4 / den);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float dd = d[c] * d[c];
        mean[c] = d[c] * t_mean + o[c];
        cov[c] = t_var * dd + r_var * (1.f - dd / dmag_sq);
    }
}
__global__ void __launch_bounds__(256) mip_embed_kernel(int n_rays, int S, int min_deg, int max_deg, int min_deg_view, int max_deg_view, const float *__restrict__ z_vals,
                                                        const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ radii,
                                                        const float *__restrict__ viewdirs, float *__restrict__ embedded, float *__restrict__ means_out, float *__restrict__ covs_out) {
    const int n_deg = max_deg - min_deg, n_deg_v = max_deg_view - min_deg_view;
    const int c_ipe = 6 * n_deg, c_dir = 3 + 6 * n_deg_v, C = c_ipe + c_dir;
    const int64_t total = (int64_t)n_rays * S * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = idx / C; int c = (int)(idx - row * C);
        int ray = (int)(row / S), k = (int)(row - (int64_t)ray * S);
        float v;
        if (c < c_ipe) {
            const float d[3] = {rays_d[3 * (size_t)ray], rays_d[3 * (size_t)ray + 1], rays_d[3 * (size_t)ray + 2]};
            const float o[3] = {rays_o[3 * (size_t)ray], rays_o[3 * (size_t)ray + 1], rays_o[3 * (size_t)ray + 2]};
            float dmag = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            float mean[3], cov[3];
            frustum_gaussian(z_vals[(size_t)ray * (S + 1) + k], z_vals[(size_t)ray * (S + 1) + k + 1], radii[ray], d, o, dmag, mean, cov);
            int q = c < c_ipe / 2 ? c : c - c_ipe / 2;   // [y | y + pi/2]
            int deg = q / 3 + min_deg, ax = q % 3;
            float sc = exp2f((float)deg);
            float y = mean[ax] * sc, yv = cov[ax] * sc * sc;
            if (c >= c_ipe / 2) y += 1.5707963267948966f;
            v = expf(-0.5f * yv) * sinf(y);
            if (means_out && c < 3) { means_out[row * 3 + c] = mean[c]; covs_out[row * 3 + c] = cov[c]; }
        } else {
            int cc = c - c_ipe;
            const float *vd = viewdirs + 3 * (size_t)ray;
            if (cc < 3) v = vd[cc];
            else { int q = cc - 3; int halfn = 3 * n_deg_v; int qq = q < halfn ? q : q - halfn; float x = vd[qq % 3] * exp2f((float)(qq / 3 + min_deg_view)); if (q >= halfn) x += 1.5707963267948966f; v = sinf(x); }
        }
        embedded[idx] = v;
    }
}

// The same cast_rays + IPE + view-direction encoding written straight into the fp16 UMMA tile images the tcgen05 NerfMLP kernel TMA-loads
// (csrc/nerf_mlp_tc.cu: per 128-row tile, point blocks [128 x 64] x 2 (96 -> 128 columns, zero padded), then the direction block; K-major,
// 128-byte swizzle): no fp32 `embedded` [N*S,123] round trip (63 GB per 800x800 image in the reference). One thread per (row, 16-byte chunk);
// the frustum Gaussian is formed once per thread in registers. Values are the fp16 roundings of exactly what mip_embed_kernel produces.
__device__ __forceinline__ uint32_t ipe_sw128(uint32_t row, uint32_t chunk16) { return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk16 ^ (row & 7u)) << 4); }
__device__ __forceinline__ uint32_t ipe_pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t *>(&h); }
__global__ void __launch_bounds__(256) mip_ipe_tiles_kernel(int64_t n_rays, int S, int min_deg, int max_deg, int min_deg_view, int max_deg_view, const float *__restrict__ z_vals,
                                                            const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ radii,
                                                            const float *__restrict__ viewdirs, uint8_t *__restrict__ image) {
    const int n_deg = max_deg - min_deg, n_deg_v = max_deg_view - min_deg_view;
    const int c_ipe = 6 * n_deg, c_dir = 3 + 6 * n_deg_v, aux = (c_ipe + 63) / 64, chunks_per_row = (aux + 1) * 8;
    const int64_t n_rows = n_rays * S, n_tiles = (n_rows + 127) / 128, total = n_tiles * 128 * chunks_per_row;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / chunks_per_row; const int cc = (int)(idx - row * chunks_per_row);
        const int blk = cc >> 3, ch = cc & 7;
        const bool is_dir = blk == aux, live = row < n_rows;
        const int width = is_dir ? c_dir : c_ipe, k0 = (is_dir ? 0 : blk * 64) + ch * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 0.f;
        if (live && k0 < width) {
            const int64_t ray = row / S; const int k = (int)(row - ray * S);
            if (!is_dir) {
                const float d[3] = {rays_d[3 * ray], rays_d[3 * ray + 1], rays_d[3 * ray + 2]};
                const float o[3] = {rays_o[3 * ray], rays_o[3 * ray + 1], rays_o[3 * ray + 2]};
                const float dmag = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                float mean[3], cov[3];
                frustum_gaussian(z_vals[ray * (S + 1) + k], z_vals[ray * (S + 1) + k + 1], radii[ray], d, o, dmag, mean, cov);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = k0 + q;
                    if (c < c_ipe) {
                        const int qq = c < c_ipe / 2 ? c : c - c_ipe / 2;
                        const int deg = qq / 3 + min_deg, ax = qq % 3;
                        const float sc = exp2f((float)deg);
                        float y = (ax == 0 ? mean[0] : ax == 1 ? mean[1] : mean[2]) * sc;
                        const float yv = (ax == 0 ? cov[0] : ax == 1 ? cov[1] : cov[2]) * sc * sc;
                        if (c >= c_ipe / 2) y += 1.5707963267948966f;
                        v[q] = expf(-0.5f * yv) * sinf(y);
                    }
                }
            } else {
                const float vd[3] = {viewdirs[3 * ray], viewdirs[3 * ray + 1], viewdirs[3 * ray + 2]};
                const int halfn = 3 * n_deg_v;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = k0 + q;
                    if (c < 3) v[q] = c == 0 ? vd[0] : c == 1 ? vd[1] : vd[2];
                    else if (c < c_dir) {
                        const int qd = c - 3, qq = qd < halfn ? qd : qd - halfn, ax = qq % 3;
                        float x = (ax == 0 ? vd[0] : ax == 1 ? vd[1] : vd[2]) * exp2f((float)(qq / 3 + min_deg_view));
                        if (qd >= halfn) x += 1.5707963267948966f;
                        v[q] = sinf(x);
                    }
                }
            }
        }
        const int64_t tile = row >> 7; const uint32_t r = (uint32_t)(row & 127);
        *reinterpret_cast<uint4 *>(image + ((size_t)tile * (aux + 1) + blk) * 16384 + ipe_sw128(r, ch)) =
            make_uint4(ipe_pack_h2(v[0], v[1]), ipe_pack_h2(v[2], v[3]), ipe_pack_h2(v[4], v[5]), ipe_pack_h2(v[6], v[7]));
    }
}

// ---- fast tile-image encoders for the reference configurations (thread == (tile, encoding block, row): a warp works on 32 rows of ONE block
// kind, so there is no intra-warp divergence; every trig evaluation is shared by the columns that use it and all column indices are compile-time).
// The generic kernels above / in nerf_mlp_tc.cu stay as the fallback for other degrees. Launch list before this change (gpurun_out, 32768 rays):
// mip_ipe_tiles_kernel 4.29 ms per 4.2 M rows and posenc_tiles_kernel 1.92 ms per 6.3 M rows next to 8.46 ms of MLP.
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((127 + e) << 23); }   // 2^e, exact, -126 <= e <= 127
__device__ __forceinline__ void store_chunk(uint8_t *blk_row, uint32_t chunk, uint32_t r7, const float *v) {
    *reinterpret_cast<uint4 *>(blk_row + ((chunk ^ r7) << 4)) = make_uint4(ipe_pack_h2(v[0], v[1]), ipe_pack_h2(v[2], v[3]), ipe_pack_h2(v[4], v[5]), ipe_pack_h2(v[6], v[7]));
}
// Mip-NeRF, 6*NDEG <= 128 point columns (two blocks) + direction block. item 0 = both point blocks of the row, item 1 = direction block.
// IPE column q (< 3*NDEG): exp(-0.5 * cov_ax * 4^deg) * sin(mean_ax * 2^deg), column 3*NDEG + q: the same with sin(y + pi/2) — the reference's
// expression, NOT cos(y): fl(y + pi/2) differs from y + pi/2 by up to ulp(y)/2 and the reference's value includes that (mipnerf_embedder.py:35-41,:55-60).
// |value| <= exp(-0.5 y_var) < 2^-25 rounds to (+-)0 in fp16, so the trig evaluation is skipped when 0.5 * y_var > 20.
template <int NDEG, int NDEGV>
__global__ void __launch_bounds__(256) mip_ipe_tiles_fast_kernel(int64_t n_rays, int S, int min_deg, int min_deg_view, const float *__restrict__ z_vals, const float *__restrict__ rays_o,
                                                                 const float *__restrict__ rays_d, const float *__restrict__ radii, const float *__restrict__ viewdirs, uint8_t *__restrict__ image) {
    static_assert(3 * NDEG % 8 == 0 && 6 * NDEG > 64 && 6 * NDEG <= 128 && 3 + 6 * NDEGV <= 64, "tile-image geometry");
    constexpr int HALF = 3 * NDEG, NJ = HALF / 8;
    const int64_t n_rows = n_rays * S, n_tiles = (n_rows + 127) / 128, total = n_tiles * 256;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(idx & 127); const int item = (int)((idx >> 7) & 1); const int64_t tile = idx >> 8;
        const int64_t row = tile * 128 + r;
        const bool live = row < n_rows;
        const uint32_t r7 = r & 7u;
        uint8_t *trow = image + (size_t)tile * 3 * 16384 + (r >> 3) * 1024u + r7 * 128u;     // my row inside block 0 of the tile
        const int64_t ray = live ? row / S : 0; const int k = (int)(row - ray * S);
        if (item == 0) {
            float mean[3] = {0.f, 0.f, 0.f}, cov[3] = {0.f, 0.f, 0.f};
            if (live) {
                const float d[3] = {rays_d[3 * ray], rays_d[3 * ray + 1], rays_d[3 * ray + 2]};
                const float o[3] = {rays_o[3 * ray], rays_o[3 * ray + 1], rays_o[3 * ray + 2]};
                const float dmag = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                frustum_gaussian(z_vals[ray * (S + 1) + k], z_vals[ray * (S + 1) + k + 1], radii[ray], d, o, dmag, mean, cov);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float sv[8], cv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int q = 8 * j + e, deg = q / 3, ax = q % 3;
                    const float sc = pow2i(deg + min_deg);
                    const float y = mean[ax] * sc, hv = 0.5f * (cov[ax] * sc * sc);
                    float s = 0.f, c = 0.f;
                    if (live && hv <= 20.f) { const float ex = expf(-hv); s = ex * sinf(y); c = ex * sinf(y + 1.5707963267948966f); }
                    sv[e] = s; cv[e] = c;
                }
                store_chunk(trow, j, r7, sv);                                                   // columns 8j .. 8j+7 of block 0
                const int cc = HALF + 8 * j;                                                    // first "cos" column of this group
                store_chunk(trow + (cc >> 6) * 16384, (uint32_t)(cc & 63) >> 3, r7, cv);
            }
            const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ch = (6 * NDEG - 64) / 8; ch < 8; ++ch) store_chunk(trow + 16384, ch, r7, z8);   // zero padding of block 1
        } else {
            float v[64];
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = 0.f;
            if (live) {
                const float vd[3] = {viewdirs[3 * ray], viewdirs[3 * ray + 1], viewdirs[3 * ray + 2]};
#pragma unroll
                for (int a = 0; a < 3; ++a) v[a] = vd[a];
#pragma unroll
                for (int q = 0; q < 3 * NDEGV; ++q) {
                    const float x = vd[q % 3] * pow2i(q / 3 + min_deg_view);
                    v[3 + q] = sinf(x); v[3 + 3 * NDEGV + q] = sinf(x + 1.5707963267948966f);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) store_chunk(trow + 2 * 16384, ch, r7, v + 8 * ch);
        }
    }
}

// BaseEmbedder (embedders/base.py:26-52), 3 + 6*LP <= 64 point columns (one block) + direction block; item 0 = point block, item 1 = direction block.
// column 3 + 6b + a = sin(2^b x_a), column 3 + 6b + 3 + a = cos(2^b x_a). ray mode as posenc_tiles_kernel (rays_o != NULL: `pts` is z_vals).
template <int LP, int LD>
__global__ void __launch_bounds__(256) posenc_tiles_fast_kernel(const float *__restrict__ pts, const float *__restrict__ viewdirs, int64_t n_rows, int samples_per_ray,
                                                                uint8_t *__restrict__ image, const float *__restrict__ rays_o, const float *__restrict__ rays_d) {
    static_assert(3 + 6 * LP <= 64 && 3 + 6 * LD <= 64, "one block each");
    const int64_t n_tiles = (n_rows + 127) / 128, total = n_tiles * 256;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(idx & 127); const int item = (int)((idx >> 7) & 1); const int64_t tile = idx >> 8;
        const int64_t row = tile * 128 + r;
        const bool live = row < n_rows;
        const uint32_t r7 = r & 7u;
        uint8_t *trow = image + ((size_t)tile * 2 + item) * 16384 + (r >> 3) * 1024u + r7 * 128u;
        float v[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) v[e] = 0.f;
        if (live) {
            float x[3];
            const int64_t ray = row / samples_per_ray;
            if (item == 1) { x[0] = viewdirs[3 * ray]; x[1] = viewdirs[3 * ray + 1]; x[2] = viewdirs[3 * ray + 2]; }
            else if (rays_o) {
                const float z = pts[row];
                // GetPts (create.py:588-597): o + d * z as a rounded product then a rounded sum (torch), never an FMA — the 2^9-scaled sin arguments amplify the difference
                x[0] = __fadd_rn(rays_o[3 * ray], __fmul_rn(rays_d[3 * ray], z)); x[1] = __fadd_rn(rays_o[3 * ray + 1], __fmul_rn(rays_d[3 * ray + 1], z)); x[2] = __fadd_rn(rays_o[3 * ray + 2], __fmul_rn(rays_d[3 * ray + 2], z));
            } else { x[0] = pts[3 * row]; x[1] = pts[3 * row + 1]; x[2] = pts[3 * row + 2]; }
#pragma unroll
            for (int a = 0; a < 3; ++a) v[a] = x[a];
            if (item == 0) {
#pragma unroll
                for (int bnd = 0; bnd < LP; ++bnd)
#pragma unroll
                    for (int a = 0; a < 3; ++a) { float sn, cs; sincosf(x[a] * pow2i(bnd), &sn, &cs); v[3 + 6 * bnd + a] = sn; v[6 + 6 * bnd + a] = cs; }
            } else {
#pragma unroll
                for (int bnd = 0; bnd < LD; ++bnd)
#pragma unroll
                    for (int a = 0; a < 3; ++a) { float sn, cs; sincosf(x[a] * pow2i(bnd), &sn, &cs); v[3 + 6 * bnd + a] = sn; v[6 + 6 * bnd + a] = cs; }
            }
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) store_chunk(trow, ch, r7, v + 8 * ch);
    }
}

// Mip-NeRF resample_along_rays + sorted_piecewise_constant_pdf (networks/utils/mip.py:146-176, :7-63), randomized=False or
// caller-supplied jitter: one warp per ray, O(S log S) interval search instead of the reference's [N,S+2,S+1] mask.
constexpr int MIP_MAX_S = 256;
__global__ void __launch_bounds__(PDF_WARPS * 32) mip_resample_kernel(int n_rays, int S, float resample_padding, const float *__restrict__ z_vals, const float *__restrict__ weights,
                                                                      const float *__restrict__ u_in, float *__restrict__ z_out) {
    __shared__ float s_w[PDF_WARPS][MIP_MAX_S + 2];
    __shared__ float s_cdf[PDF_WARPS][MIP_MAX_S + 2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ray = blockIdx.x * PDF_WARPS + wid;
    if (ray >= n_rays) return;
    const float *z = z_vals + (size_t)ray * (S + 1), *w_in = weights + (size_t)ray * S;
    float *w = s_w[wid], *cdf = s_cdf[wid];
    // blur: pad, pairwise max, pairwise mean, + padding constant (mip.py:154-163)
    float part = 0.f;
    for (int k = lane; k < S; k += 32) {
        float wm1 = w_in[max(k - 1, 0)], w0 = w_in[k], wp1 = w_in[min(k + 1, S - 1)];
        float v = 0.5f * (fmaxf(wm1, w0) + fmaxf(w0, wp1)) + resample_padding;
        w[k] = v; part += v;
    }
    float wsum_all = wsum(part);
    const float eps = 1e-5f;
    float padding = fmaxf(0.f, eps - wsum_all);
    const float add = padding / (float)S;
    wsum_all += padding;
    __syncwarp();
    if (lane == 0) {
        float c = 0.f; cdf[0] = 0.f;
        for (int k = 0; k < S - 1; ++k) { c += (w[k] + add) / wsum_all; cdf[k + 1] = fminf(1.f, c); }
        cdf[S] = 1.f;
    }
    __syncwarp();
    const int num = S + 1;   // samples drawn == len(z_vals)
    for (int j = lane; j < num; j += 32) {
        float u;
        if (u_in) u = u_in[(size_t)ray * num + j];
        else { const float end = 1.f - 1.1920929e-07f; float step = end / (float)(num - 1); u = j < num / 2 ? step * (float)j : end - step * (float)(num - 1 - j); }
        int lo = 0, hi = num;   // first index with cdf[idx] > u  (cdf has S+1 entries; cdf[0]=0 <= u always)
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
        int i0 = lo - 1, i1 = lo < num ? lo : num - 1;
        float c0 = cdf[i0], c1 = cdf[i1], b0 = z[i0], b1 = z[i1];
        float t = (u - c0) / (c1 - c0);
        if (isnan(t)) t = 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        z_out[(size_t)ray * num + j] = b0 + t * (b1 - b0);
    }
}


// Ray generation: GetRays (+ Mip-NeRF radii) + GetViewdirs (datasets/pipelines/create.py:205-245, :437-448) and the
// NGP-convention get_rays_np_hash (datasets/load_data/get_rays.py:35-69), one thread per pixel, any pixel subset.
//   convention 0 (NeRF / Mip): dirs = ((i-cx)/fx, -(j-cy)/fy, -1), rays_d = c2w[:3,:3] dirs, viewdirs = rays_d/|rays_d|,
//                              radii = |rays_d(j,i) - rays_d(j+1,i)| * 2/sqrt(12) (last row reuses row H-3, create.py:237-243)
//   convention 1 (NGP):        pixel centres +0.5, dirs = ((i-cx)/fx, (j-cy)/fy, 1), rays_d normalised, pose given as [4,3] (ngp xforms)
struct RayGenParams { float c2w[12]; float fx, fy, cx, cy; int H, W, convention; };
__device__ __forceinline__ void raygen_dir(const RayGenParams &p, float i, float j, float d[3]) {
    float x, y, z;
    if (p.convention == 0) { x = (i - p.cx) / p.fx; y = -(j - p.cy) / p.fy; z = -1.f; }
    else { x = (i + 0.5f - p.cx) / p.fx; y = (j + 0.5f - p.cy) / p.fy; z = 1.f; }
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (x * p.c2w[4 * c] + y * p.c2w[4 * c + 1]) + z * p.c2w[4 * c + 2];
}
// The NGP training-ray source on the device: what HashNerfDataset + HashBatchSample + RandomBGColor produce on the host
//   load_rays_hash (datasets/load_data/get_rays.py:72-98): row r of the [I*H*W, 11] table = (rays_o3, rays_d3 by get_rays_np_hash, rgba4, image id) of pixel r % (H*W) of image r / (H*W)
//   the table shuffle (hashnerf_dataset.py:41-44) = the caller's row permutation; HashBatchSample (create.py:153-190) = a slice of it
//   RandomBGColor (augment.py:290-313): bg = U[0,1)^3, target = rgb * alpha + bg * (1 - alpha) (float64 on the host, rounded to fp32)
// without the 2.8 GB host table: one thread per ray regenerates its row from (pose, pixel) and blends the background.
__global__ void __launch_bounds__(256) ngp_batch_sample_kernel(const float *__restrict__ poses, const float *__restrict__ images, int n_img, int H, int W, float fx, float fy, float cx,
                                                               float cy, const int64_t *__restrict__ row_idx, int64_t n, const float *__restrict__ u_bg, Pcg32 rng,
                                                               float *__restrict__ rays_o, float *__restrict__ rays_d, float *__restrict__ target_s, float *__restrict__ alpha_out,
                                                               float *__restrict__ img_ids, float *__restrict__ bg_color) {
    const int64_t hw = (int64_t)H * W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = row_idx[t];
        const int img = (int)(row / hw), pix = (int)(row % hw), j = pix / W, i = pix % W;
        const float *P = poses + (size_t)img * 12;                 // [4,3]: rows 0..2 = rotation columns (c2w = P^T), row 3 = camera origin
        const float x = ((float)i + 0.5f - cx) / fx, y = ((float)j + 0.5f - cy) / fy;
        float d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = (x * P[c] + y * P[3 + c]) + P[6 + c];
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const float4 px = *reinterpret_cast<const float4 *>(images + ((size_t)img * hw + pix) * 4);
        float u[3];
        if (u_bg) { u[0] = u_bg[3 * t]; u[1] = u_bg[3 * t + 1]; u[2] = u_bg[3 * t + 2]; }
        else { Pcg32 r = rng; r.advance((uint64_t)t * 3u); u[0] = r.next_float(); u[1] = r.next_float(); u[2] = r.next_float(); }
        const float rgb[3] = {px.x, px.y, px.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rays_o[3 * t + c] = P[9 + c];
            rays_d[3 * t + c] = d[c] / nrm;
            target_s[3 * t + c] = (float)((double)(rgb[c] * px.w) + (double)u[c] * (double)(1.f - px.w));
            bg_color[3 * t + c] = u[c];
        }
        alpha_out[t] = px.w;
        img_ids[t] = (float)img;
    }
}

__global__ void __launch_bounds__(256) get_rays_kernel(RayGenParams p, const int32_t *__restrict__ pixel_idx, int64_t n, float *__restrict__ rays_o, float *__restrict__ rays_d,
                                                       float *__restrict__ viewdirs, float *__restrict__ radii) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int pix = pixel_idx ? pixel_idx[t] : (int)t;
        const int j = pix / p.W, i = pix % p.W;
        float d[3]; raygen_dir(p, (float)i, (float)j, d);
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (p.convention == 1) { d[0] /= nrm; d[1] /= nrm; d[2] /= nrm; }
#pragma unroll
        for (int c = 0; c < 3; ++c) { rays_o[3 * t + c] = p.c2w[4 * c + 3]; rays_d[3 * t + c] = d[c]; if (viewdirs) viewdirs[3 * t + c] = p.convention == 1 ? d[c] : d[c] / nrm; }
        if (radii) {
            const int ja = j < p.H - 1 ? j : p.H - 3;   // dx = cat([dx, dx[-2:-1]]): the last row takes dx of row H-3
            float a[3], b[3]; raygen_dir(p, (float)i, (float)ja, a); raygen_dir(p, (float)i, (float)(ja + 1), b);
            const float dx = sqrtf((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
            radii[t] = dx * 2.f / sqrtf(12.f);
        }
    }
}
// GetZvals (create.py:502-531, randomized=False) + PerturbZvals (augment.py:269-283) when u != NULL: one thread per (ray, sample)
__global__ void __launch_bounds__(256) zvals_kernel(int64_t n_rays, int S, float near_, float far_, int lindisp, const float *__restrict__ u, float *__restrict__ z_out) {
    const float step = 1.0f / (float)(S - 1);
    auto zat = [&](int k) { float t = k < S / 2 ? step * (float)k : 1.0f - step * (float)(S - 1 - k);   // torch.linspace
                            return lindisp ? 1.f / (1.f / near_ * (1.f - t) + 1.f / far_ * t) : near_ * (1.f - t) + far_ * t; };
    const int64_t total = n_rays * S;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % S);
        float z = zat(k);
        if (u) {
            float lower = k == 0 ? z : 0.5f * (z + zat(k - 1)), upper = k == S - 1 ? z : 0.5f * (zat(k + 1) + z);
            z = lower + (upper - lower) * u[idx];
        }
        z_out[idx] = z;
    }
}

}  // namespace xrb

using namespace xrb;

extern "C" {

static int composite_args_ok(const float *raw, const float *z, const float *d, int n_rays, int n_samples, int density_act) {
    if (n_rays < 0 || n_samples < 1) { set_error("nerf_composite: bad size"); return XRB_E_BADARG; }
    if (!(density_act == XRB_ACT_RELU || density_act == XRB_ACT_SOFTPLUS)) { set_error("nerf_composite: density activation must be relu(1) or softplus(4)"); return XRB_E_UNSUPPORTED; }
    if (n_rays && (!raw || !z || !d)) { set_error("nerf_composite: null pointer"); return XRB_E_BADARG; }
    if (((uintptr_t)raw & 15) != 0) { set_error("nerf_composite: raw must be 16-byte aligned"); return XRB_E_BADARG; }
    return XRB_OK;
}

int xrb_nerf_composite_forward(const float *raw, const float *z_vals, const float *rays_d, int n_rays, int n_samples, int mip, int white_bkgd, float rgb_padding, float density_bias,
                               int density_act, float *rgb, float *disp, float *acc, float *weights, void *stream) {
    int e = composite_args_ok(raw, z_vals, rays_d, n_rays, n_samples, density_act); if (e) return e;
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(rgb && disp && acc && weights, "nerf_composite_forward: null output");
    CompositeParams p{n_rays, n_samples, mip, white_bkgd, density_act, rgb_padding, density_bias};
    int blocks = (int)(((size_t)n_rays * 32 + 255) / 256);
    nerf_composite_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, raw, z_vals, rays_d, nullptr, rgb, disp, acc, weights, nullptr);
    return check_launch("nerf_composite_forward");
}

int xrb_nerf_composite_backward(const float *raw, const float *z_vals, const float *rays_d, const float *grad_rgb, int n_rays, int n_samples, int mip, int white_bkgd, float rgb_padding,
                                float density_bias, int density_act, float *d_raw, void *stream) {
    int e = composite_args_ok(raw, z_vals, rays_d, n_rays, n_samples, density_act); if (e) return e;
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(grad_rgb && d_raw && ((uintptr_t)d_raw & 15) == 0, "nerf_composite_backward: null/misaligned pointer");
    CompositeParams p{n_rays, n_samples, mip, white_bkgd, density_act, rgb_padding, density_bias};
    int blocks = (int)(((size_t)n_rays * 32 + 255) / 256);
    nerf_composite_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, raw, z_vals, rays_d, grad_rgb, nullptr, nullptr, nullptr, nullptr, d_raw);
    return check_launch("nerf_composite_backward");
}

int xrb_nerf_sample_pdf(const float *z_vals, const float *weights, const float *rays_o, const float *rays_d, const float *u, int n_rays, int n_samples, int n_importance, float *z_out,
                        float *pts_out, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && n_samples >= 3 && n_importance >= 2, "sample_pdf: bad size");
    if (n_samples > PDF_MAX_S || n_samples + n_importance > PDF_MAX_OUT) { set_error("sample_pdf: n_samples<=256 and n_samples+n_importance<=512 supported"); return XRB_E_UNSUPPORTED; }
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(z_vals && weights && rays_o && rays_d && z_out, "sample_pdf: null pointer");
    sample_pdf_kernel<<<(n_rays + PDF_WARPS - 1) / PDF_WARPS, PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(n_rays, n_samples, n_importance, z_vals, weights, rays_o, rays_d, u, z_out, pts_out);
    return check_launch("sample_pdf");
}

int xrb_nerf_posenc(const float *pts, const float *viewdirs, int64_t n_pts, int samples_per_ray, int multires, int multires_dirs, float *embedded, void *stream) {
    XRB_REQUIRE(n_pts >= 0 && samples_per_ray >= 1 && multires >= 0 && multires_dirs >= 0, "posenc: bad size");
    if (n_pts == 0) return XRB_OK;
    XRB_REQUIRE(pts && viewdirs && embedded, "posenc: null pointer");
    int64_t total = n_pts * (6 + 6 * (int64_t)(multires + multires_dirs));
    int64_t blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    posenc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_pts, samples_per_ray, multires, multires_dirs, pts, viewdirs, embedded);
    return check_launch("posenc");
}

int xrb_mip_embed(const float *z_vals, const float *rays_o, const float *rays_d, const float *radii, const float *viewdirs, int n_rays, int n_samples, int min_deg_point,
                  int max_deg_point, int min_deg_view, int max_deg_view, float *embedded, float *means_out, float *covs_out, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && n_samples >= 1 && max_deg_point > min_deg_point && max_deg_view >= min_deg_view, "mip_embed: bad size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(z_vals && rays_o && rays_d && radii && viewdirs && embedded && ((means_out == nullptr) == (covs_out == nullptr)), "mip_embed: null pointer");
    int64_t total = (int64_t)n_rays * n_samples * (6 * (max_deg_point - min_deg_point) + 3 + 6 * (max_deg_view - min_deg_view));
    int64_t blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    mip_embed_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, n_samples, min_deg_point, max_deg_point, min_deg_view, max_deg_view, z_vals, rays_o, rays_d, radii, viewdirs, embedded,
                                                                   means_out, covs_out);
    return check_launch("mip_embed");
}

int xrb_mip_ipe_tiles_rays(const float *z_vals, const float *rays_o, const float *rays_d, const float *radii, const float *viewdirs, int64_t n_rays, int n_samples, int min_deg_point,
                           int max_deg_point, int min_deg_view, int max_deg_view, void *enc_image, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && n_samples >= 1 && max_deg_point > min_deg_point && max_deg_view >= min_deg_view, "mip_ipe_tiles_rays: bad size");
    XRB_REQUIRE(6 * (max_deg_point - min_deg_point) <= 128 && 3 + 6 * (max_deg_view - min_deg_view) <= 64, "mip_ipe_tiles_rays: encoding wider than the tile image (128 + 64 columns)");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(z_vals && rays_o && rays_d && radii && viewdirs && enc_image && ((uintptr_t)enc_image & 15) == 0, "mip_ipe_tiles_rays: null/misaligned pointer");
    const int aux = (6 * (max_deg_point - min_deg_point) + 63) / 64;
    int64_t total = ((n_rays * n_samples + 127) / 128) * 128 * (aux + 1) * 8, blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    if (max_deg_point - min_deg_point == 16 && max_deg_view - min_deg_view == 4 && min_deg_point >= -100 && max_deg_point <= 100 && min_deg_view >= -100 && max_deg_view <= 100 &&
        !getenv("XRB_GENERIC_ENCODERS")) {                     // the reference configuration (configs/mipnerf/*.py): specialised kernel
        int64_t items = ((n_rays * n_samples + 127) / 128) * 256, fb = (items + 255) / 256; if (fb > NUM_SMS * 16) fb = NUM_SMS * 16;
        mip_ipe_tiles_fast_kernel<16, 4><<<(int)fb, 256, 0, (cudaStream_t)stream>>>(n_rays, n_samples, min_deg_point, min_deg_view, z_vals, rays_o, rays_d, radii, viewdirs, (uint8_t *)enc_image);
        return check_launch("mip_ipe_tiles_rays");
    }
    mip_ipe_tiles_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, n_samples, min_deg_point, max_deg_point, min_deg_view, max_deg_view, z_vals, rays_o, rays_d, radii, viewdirs,
                                                                       (uint8_t *)enc_image);
    return check_launch("mip_ipe_tiles_rays");
}

// called by xrb_nerf_posenc_tiles / xrb_nerf_posenc_tiles_rays (nerf_mlp_tc.cu) for the reference configuration multires=10, multires_dirs=4
int xrb_internal_posenc_tiles_fast(const float *pts, const float *viewdirs, int64_t n_rows, int samples_per_ray, void *enc_image, const float *rays_o, const float *rays_d, void *stream) {
    int64_t items = ((n_rows + 127) / 128) * 256, fb = (items + 255) / 256; if (fb > NUM_SMS * 16) fb = NUM_SMS * 16;
    posenc_tiles_fast_kernel<10, 4><<<(int)fb, 256, 0, (cudaStream_t)stream>>>(pts, viewdirs, n_rows, samples_per_ray, (uint8_t *)enc_image, rays_o, rays_d);
    return check_launch("posenc_tiles");
}

int xrb_mip_resample(const float *z_vals, const float *weights, const float *u, int n_rays, int n_samples, float resample_padding, float *z_out, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && n_samples >= 2, "mip_resample: bad size");
    if (n_samples > MIP_MAX_S) { set_error("mip_resample: n_samples <= 256 supported"); return XRB_E_UNSUPPORTED; }
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(z_vals && weights && z_out, "mip_resample: null pointer");
    mip_resample_kernel<<<(n_rays + PDF_WARPS - 1) / PDF_WARPS, PDF_WARPS * 32, 0, (cudaStream_t)stream>>>(n_rays, n_samples, resample_padding, z_vals, weights, u, z_out);
    return check_launch("mip_resample");
}

int xrb_ngp_batch_sample(const float *poses, const float *images_rgba, int n_images, int H, int W, float fx, float fy, float cx, float cy, const int64_t *row_idx, int64_t n,
                         const float *u_bg, uint64_t seed, int64_t n_prior_calls, float *rays_o, float *rays_d, float *target_s, float *alpha, float *img_ids, float *bg_color, void *stream) {
    XRB_REQUIRE(n_images >= 1 && H >= 1 && W >= 1 && n >= 0, "ngp_batch_sample: bad arguments");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(poses && images_rgba && row_idx && rays_o && rays_d && target_s && alpha && img_ids && bg_color, "ngp_batch_sample: null pointer");
    XRB_REQUIRE(((uintptr_t)images_rgba & 15) == 0, "ngp_batch_sample: images must be 16-byte aligned");
    int64_t blocks = (n + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    ngp_batch_sample_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(poses, images_rgba, n_images, H, W, fx, fy, cx, cy, row_idx, n, u_bg, host_rng(seed, n_prior_calls), rays_o, rays_d,
                                                                          target_s, alpha, img_ids, bg_color);
    return check_launch("ngp_batch_sample");
}

int xrb_nerf_get_rays(const float *c2w_host, int H, int W, float fx, float fy, float cx, float cy, int convention, const int32_t *pixel_idx, int64_t n, float *rays_o, float *rays_d,
                      float *viewdirs, float *radii, void *stream) {
    XRB_REQUIRE(H >= 1 && W >= 1 && n >= 0 && (convention == 0 || convention == 1), "get_rays: bad arguments");
    XRB_REQUIRE(!(radii && H < 3), "get_rays: radii need H >= 3");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(c2w_host && rays_o && rays_d, "get_rays: null pointer");
    RayGenParams p; for (int k = 0; k < 12; ++k) p.c2w[k] = c2w_host[k];
    p.fx = fx; p.fy = fy; p.cx = cx; p.cy = cy; p.H = H; p.W = W; p.convention = convention;
    int64_t blocks = (n + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    get_rays_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, pixel_idx, n, rays_o, rays_d, viewdirs, radii);
    return check_launch("get_rays");
}

int xrb_nerf_zvals(int64_t n_rays, int n_samples, float near_, float far_, int lindisp, const float *u, float *z_vals, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && n_samples >= 2, "zvals: bad size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(z_vals, "zvals: null pointer");
    int64_t blocks = (n_rays * n_samples + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    zvals_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, n_samples, near_, far_, lindisp, u, z_vals);
    return check_launch("zvals");
}

}  // extern "C"
