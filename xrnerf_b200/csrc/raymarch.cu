// xrnerf_b200 — sm_100a replacements for the 10 `raymarch_cuda` ops (C ABI: xrb_rm_*, include/xrnerf_b200.h).
//
// What differs from the reference kernels (all <<<n/128,128>>> thread-per-element, extensions/ngp_raymarch/src/*.cu):
//   * ray march: count pass + 2-level exclusive scan + emit pass => deterministic ray-order layout, no global
//     atomics, no device sync, start-t (the only RNG use) computed once and cached in the workspace;
//   * compositing: one WARP per ray, lanes over samples (coalesced float4 reads of raw[S,4]), transmittance by a
//     warp product-scan instead of a serial loop per thread;
//   * grid kernels: 128-bit vector loads/stores, grid sized to a multiple of 148 SMs with grid-stride loops,
//     bitfield + 7 max-pool levels produced by ONE kernel launch sequence without atomics, fixed-order mean.
// HBM roofline per op (algorithmic bytes) is listed in DESIGN.md §4.
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace xrb {

static thread_local char g_err[256] = "";
void set_error(const char *msg) { strncpy(g_err, msg, sizeof(g_err) - 1); g_err[sizeof(g_err) - 1] = 0; }
int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e)); return (int)e; }
    return XRB_OK;
}

// ============================================================================ ray march
// One thread per ray. emit==false: count steps (ray_sampler.cu:58-72). emit==true: write rows (:99-115).
constexpr uint32_t TCAP = 64;                      // per-ray sample-t slots cached inline by the count pass
constexpr uint32_t OVF_CAP = NERF_STEPS - TCAP;    // a ray with more samples gets ONE overflow chunk of this many slots (bump-allocated), so the emit
                                                   // pass can rebuild all of its rows in parallel; only when the chunk pool is exhausted is it re-marched
struct OvfPool { uint32_t *counter; float *area; uint32_t n_chunks; int32_t *chunk_of; };

template <bool EMIT>
__device__ __forceinline__ uint32_t march_ray(const float o[3], const float d[3], float lo, float hi, float startt, float cone,
                                              const uint8_t *__restrict__ bitfield, uint32_t limit, float *__restrict__ coords, float *__restrict__ tbuf = nullptr,
                                              const uint32_t *lut = nullptr, const OvfPool *pool = nullptr, uint32_t ray = 0) {
    const float idir[3] = {div_(1.0f, d[0]), div_(1.0f, d[1]), div_(1.0f, d[2])};
    float wdir[3], diag = sub_(hi, lo);
    if (EMIT) { wdir[0] = mul_(add_(d[0], 1.0f), 0.5f); wdir[1] = mul_(add_(d[1], 1.0f), 0.5f); wdir[2] = mul_(add_(d[2], 1.0f), 0.5f); }
    uint32_t j = 0; float t = startt;
    float *ovf = nullptr;
    while (true) {
        float p[3] = {add_(o[0], mul_(t, d[0])), add_(o[1], mul_(t, d[1])), add_(o[2], mul_(t, d[2]))};
        if (!(aabb_contains(lo, hi, p[0], p[1], p[2]) && j < limit)) break;
        float dt = calc_dt(t, cone);
        uint32_t mip = (uint32_t)mip_from_dt(dt, p[0], p[1], p[2]);
        if (occupied_at(p[0], p[1], p[2], bitfield, mip, lut)) {
            if (EMIT) {
                float *c = coords + 7 * (size_t)j;
                c[0] = div_(sub_(p[0], lo), diag); c[1] = div_(sub_(p[1], lo), diag); c[2] = div_(sub_(p[2], lo), diag);  // warp_position
                c[3] = warp_dt(dt);
                c[4] = wdir[0]; c[5] = wdir[1]; c[6] = wdir[2];
            } else if (tbuf) {
                if (j < TCAP) tbuf[j] = t;
                else if (pool) {
                    if (j == TCAP) {
                        const uint32_t c = atomicAdd(pool->counter, 1u);
                        const bool got = c < pool->n_chunks;
                        pool->chunk_of[ray] = got ? (int32_t)c : -1;
                        ovf = got ? pool->area + (size_t)c * OVF_CAP : nullptr;
                    }
                    if (ovf) ovf[j - TCAP] = t;
                }
            }
            ++j; t = add_(t, dt);
        } else {
            t = advance_to_next_voxel(t, cone, p, d, idir, NERF_GRIDSIZE >> mip);
        }
    }
    return j;
}

constexpr int MARCH_BLOCK = 128;
// grid for warp-per-ray kernels with persistent (ray-strided) warps
static inline int warp_grid(int n_rays) { size_t b = ((size_t)n_rays * 32 + 255) / 256, cap = (size_t)NUM_SMS * 8; return (int)(b < cap ? (b ? b : 1) : cap); }
// The march is a long, divergent, latency-bound chain per ray (ncu r01: kernel time == the slowest warp's serialised union of 32
// rays' paths). Only every LANE_STRIDE-th lane owns a ray: 4x more warps (the SMs have the slots: 14 -> 55 warps/SM at 65 536 rays),
// each serialising 8 rays instead of 32.
constexpr int MAX_LANE_STRIDE = 4;
constexpr int MARCH_RAYS_PER_BLOCK_MIN = MARCH_BLOCK / MAX_LANE_STRIDE;
static int lane_stride() { static int v = -1; if (v < 0) { const char *e = getenv("XRB_MARCH_LANE_STRIDE"); v = e ? atoi(e) : 1; if (v != 1 && v != 2 && v != 4) v = 1; } return v; }

struct MarchWs {  // device workspace layout for N rays
    uint32_t *local_excl;  // [N]
    float *startt;         // [N]
    uint32_t *block_sum;   // [n_blocks] -> exclusive block offsets after scan
    uint32_t *misc;        // [16]: base0, ray0, total, overflow-chunk counter
    float *tbuf;           // [N][TCAP] sample t values found by the count pass
    int32_t *chunk_of;     // [N] overflow chunk of a ray with more than TCAP samples (-1: pool exhausted); valid only for such rays
    float *ovf;            // [n_chunks][OVF_CAP]
    uint32_t n_chunks;
};
__host__ __device__ inline uint32_t march_n_chunks(int n) { uint32_t c = (uint32_t)n / 8u; return c < 64u ? 64u : c; }   // every 8th ray may be longer than TCAP samples
__host__ __device__ inline size_t march_ws_bytes(int n) {
    size_t nb = (size_t)(n + MARCH_RAYS_PER_BLOCK_MIN - 1) / MARCH_RAYS_PER_BLOCK_MIN;
    return sizeof(uint32_t) * ((size_t)n * 3 + nb + 16 + (size_t)n * TCAP + (size_t)march_n_chunks(n) * OVF_CAP);
}
inline MarchWs march_ws(void *ws, int n) {
    size_t nb = (size_t)(n + MARCH_RAYS_PER_BLOCK_MIN - 1) / MARCH_RAYS_PER_BLOCK_MIN;
    MarchWs w; w.local_excl = (uint32_t *)ws; w.startt = (float *)(w.local_excl + n); w.block_sum = (uint32_t *)(w.startt + n); w.misc = w.block_sum + nb; w.tbuf = (float *)(w.misc + 16);
    w.chunk_of = (int32_t *)(w.tbuf + (size_t)n * TCAP); w.ovf = (float *)(w.chunk_of + n); w.n_chunks = march_n_chunks(n);
    return w;
}

// block-wide exclusive scan of one uint32 per thread (MARCH_BLOCK threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_tot[MARCH_BLOCK / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan_u32(v, lane);
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < MARCH_BLOCK / 32; ++k) { uint32_t s = warp_tot[k]; if (k < w) off += s; tot += s; }
    *total = tot;
    return off + incl - v;
}

template <int LANE_STRIDE>
__global__ void __launch_bounds__(MARCH_BLOCK) march_count_kernel(int n_rays, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                                  const uint8_t *__restrict__ bitfield, float lo, float hi, float near_distance, float cone,
                                                                  Pcg32 rng, uint32_t *__restrict__ local_excl, float *__restrict__ startt_out,
                                                                  uint32_t *__restrict__ block_sum, int32_t *__restrict__ numsteps, float *__restrict__ tbuf, OvfPool pool) {
    constexpr int MARCH_RAYS_PER_BLOCK = MARCH_BLOCK / LANE_STRIDE;
    __shared__ uint32_t lut[128];
    morton_lut_init(lut, threadIdx.x, MARCH_BLOCK);
    __syncthreads();
    const bool owner = (threadIdx.x % LANE_STRIDE) == 0;
    uint32_t i = blockIdx.x * MARCH_RAYS_PER_BLOCK + threadIdx.x / LANE_STRIDE;
    uint32_t n = 0;
    if (owner && i < (uint32_t)n_rays) {
        float o[3] = {rays_o[3 * (size_t)i], rays_o[3 * (size_t)i + 1], rays_o[3 * (size_t)i + 2]};
        float d[3] = {rays_d[3 * (size_t)i], rays_d[3 * (size_t)i + 1], rays_d[3 * (size_t)i + 2]};
        float st = ray_start_t(rng, i, lo, hi, o, d, near_distance, cone);
        startt_out[i] = st;
        n = march_ray<false>(o, d, lo, hi, st, cone, bitfield, NERF_STEPS, nullptr, tbuf ? tbuf + (size_t)i * TCAP : nullptr, lut, tbuf ? &pool : nullptr, i);
        numsteps[2 * (size_t)i] = (int32_t)n;
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(n, &tot);
    if (owner && i < (uint32_t)n_rays) local_excl[i] = ex;
    if (threadIdx.x == 0) block_sum[blockIdx.x] = tot;
}

// single block: exclusive scan of block sums, base/ray counters bookkeeping (counters[] accumulate like the reference's atomics)
__global__ void __launch_bounds__(1024) march_scan_kernel(int n_blocks, uint32_t *__restrict__ block_sum, uint32_t *__restrict__ misc, int32_t *__restrict__ counters) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        int b = b0 + threadIdx.x;
        uint32_t v = b < n_blocks ? block_sum[b] : 0u;
        uint32_t incl = warp_incl_scan_u32(v, lane);
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        uint32_t off = 0, tot = 0;
        for (int k = 0; k < 32; ++k) { uint32_t s = warp_tot[k]; if (k < w) off += s; tot += s; }
        uint32_t carry = carry_s;
        if (b < n_blocks) block_sum[b] = carry + off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t base0 = counters ? (uint32_t)counters[1] : 0u, ray0 = counters ? (uint32_t)counters[0] : 0u;
        misc[0] = base0; misc[1] = ray0; misc[2] = carry_s;
        if (counters) counters[1] = (int32_t)(base0 + carry_s);
    }
}

// Emit pass, one WARP per ray: lanes rebuild rows from the cached t values (pos = o + t d and dt = calc_dt(t) are the very
// expressions of the march, so the rows are bit-identical to a second march) and write 32 consecutive 28-byte rows per step.
// Rays longer than TCAP are re-marched by lane 0 (ray_sampler.cu:99-115).
__global__ void __launch_bounds__(256) march_emit_kernel(int n_rays, const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ bitfield, float lo,
                                                         float hi, float cone, uint32_t max_samples, const uint32_t *__restrict__ local_excl, const float *__restrict__ startt,
                                                         const uint32_t *__restrict__ block_off, const uint32_t *__restrict__ misc, const float *__restrict__ tbuf,
                                                         const int32_t *__restrict__ chunk_of, const float *__restrict__ ovf, float *__restrict__ coords_out, int32_t *__restrict__ rays_index, int32_t *__restrict__ numsteps,
                                                         int32_t *__restrict__ counters, int rays_per_scan_block) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * 256) >> 5;
    uint32_t ok_count = 0;
    for (uint32_t i = (blockIdx.x * 256 + threadIdx.x) >> 5; i < (uint32_t)n_rays; i += n_warps) {   // persistent warps: 8192 tiny blocks cost more to schedule than to run
        uint32_t n = (uint32_t)numsteps[2 * (size_t)i];
        uint32_t base = misc[0] + block_off[i / rays_per_scan_block] + local_excl[i];
        __syncwarp();
        if (base + n > max_samples) {  // ray_sampler.cu:76-82
            if (lane == 0) { numsteps[2 * (size_t)i] = 0; numsteps[2 * (size_t)i + 1] = (int32_t)base; }
        } else {
            ok_count += 1;
            if (lane == 0) {
                numsteps[2 * (size_t)i + 1] = (int32_t)base;
                // serial-order equivalent of `ray_idx = atomicAdd(ray_counter,1)`: slots are a prefix of the rays
                rays_index[i] = n == 0 ? -1 : (int32_t)(misc[1] + i);
            }
            if (n > 0) {
                const float o[3] = {rays_o[3 * (size_t)i], rays_o[3 * (size_t)i + 1], rays_o[3 * (size_t)i + 2]};
                const float d[3] = {rays_d[3 * (size_t)i], rays_d[3 * (size_t)i + 1], rays_d[3 * (size_t)i + 2]};
                const float diag = sub_(hi, lo);
                const float w0 = mul_(add_(d[0], 1.0f), 0.5f), w1 = mul_(add_(d[1], 1.0f), 0.5f), w2 = mul_(add_(d[2], 1.0f), 0.5f);
                // rows whose t is cached (inline slots, then the ray's overflow chunk) are rebuilt by all lanes
                const int32_t chunk = n > TCAP ? chunk_of[i] : -1;
                const float *tl = tbuf + (size_t)i * TCAP, *to = chunk >= 0 ? ovf + (size_t)chunk * OVF_CAP : nullptr;
                const uint32_t n_par = n <= TCAP ? n : (to ? n : TCAP - 1);
                for (uint32_t j = lane; j < n_par; j += 32) {
                    float t = j < TCAP ? tl[j] : to[j - TCAP];
                    float *c = coords_out + 7 * (size_t)(base + j);
                    c[0] = div_(sub_(add_(o[0], mul_(t, d[0])), lo), diag); c[1] = div_(sub_(add_(o[1], mul_(t, d[1])), lo), diag); c[2] = div_(sub_(add_(o[2], mul_(t, d[2])), lo), diag);
                    c[3] = warp_dt(calc_dt(t, cone));
                    c[4] = w0; c[5] = w1; c[6] = w2;
                }
                // chunk pool exhausted: lane 0 resumes the march AT cached sample TCAP-1 (the march state is just t) and emits the remaining rows (ray_sampler.cu:99-115)
                if (n_par < n && lane == 0) march_ray<true>(o, d, lo, hi, tl[TCAP - 1], cone, bitfield, n - (TCAP - 1), coords_out + 7 * (size_t)(base + TCAP - 1));
            }
        }
    }
    if (lane == 0 && ok_count && counters) atomicAdd((unsigned int *)counters, ok_count);
}

// ============================================================================ compaction (compacted_coord.cu:5-77)
__global__ void __launch_bounds__(MARCH_BLOCK) compact_count_kernel(int n_rays, const int32_t *__restrict__ numsteps, uint32_t *__restrict__ local_excl,
                                                                    uint32_t *__restrict__ block_sum) {
    uint32_t i = blockIdx.x * MARCH_BLOCK + threadIdx.x;
    uint32_t n = i < (uint32_t)n_rays ? (uint32_t)numsteps[2 * (size_t)i] : 0u;
    uint32_t tot, ex = block_excl_scan(n, &tot);
    if (i < (uint32_t)n_rays) local_excl[i] = ex;
    if (threadIdx.x == 0) block_sum[blockIdx.x] = tot;
}
// warp per ray: copies its rows as a flat stream of floats (coalesced)
__global__ void __launch_bounds__(256) compact_copy_kernel(int n_rays, uint32_t max_compacted, const float *__restrict__ coords_in, const int32_t *__restrict__ numsteps,
                                                           const uint32_t *__restrict__ local_excl, const uint32_t *__restrict__ block_off, const uint32_t *__restrict__ misc,
                                                           float *__restrict__ coords_out, int32_t *__restrict__ numsteps_c, int32_t *__restrict__ ray_counter) {
    int lane = threadIdx.x & 31;
    uint32_t ray = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (ray >= (uint32_t)n_rays) return;
    uint32_t n = (uint32_t)numsteps[2 * (size_t)ray], base = (uint32_t)numsteps[2 * (size_t)ray + 1];
    uint32_t cbase = misc[0] + block_off[ray / MARCH_BLOCK] + local_excl[ray];
    uint32_t nc = min(max_compacted - min(max_compacted, cbase), n);  // :64
    if (lane == 0) { numsteps_c[2 * (size_t)ray] = (int32_t)nc; numsteps_c[2 * (size_t)ray + 1] = (int32_t)cbase; if (nc) atomicAdd((unsigned int *)ray_counter, 1u); }
    const float *src = coords_in + 7 * (size_t)base; float *dst = coords_out + 7 * (size_t)cbase;
    for (uint32_t k = lane; k < nc * 7; k += 32) dst[k] = src[k];
}

// rays of a compacted batch whose every marched sample survived the compaction (numsteps_c.count == numsteps.count): the rays a training step really trains
__global__ void __launch_bounds__(256) count_trained_kernel(int n_rays, const int32_t *__restrict__ numsteps, const int32_t *__restrict__ numsteps_c, unsigned long long *__restrict__ accum) {
    unsigned int local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += gridDim.x * blockDim.x) local += numsteps_c[2 * (size_t)i] == numsteps[2 * (size_t)i] ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(accum, (unsigned long long)local);
}

// ============================================================================ compositing (calc_rgb.cu)
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float *incl_out) {
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= n; }
    float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    *incl_out = incl;
    return lane == 0 ? 1.f : ex;
}
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MODE 0: training forward (per-ray bg, background only if not truncated; calc_rgb.cu:5-67)
// MODE 1: inference (global bg, alpha out; calc_rgb.cu:143-206)
template <int MODE>
__global__ void __launch_bounds__(256) composite_fwd_kernel(int n_rays, const float4 *__restrict__ raw, const float *__restrict__ coords, const int32_t *__restrict__ numsteps,
                                                            const int32_t *__restrict__ numsteps_c, const float *__restrict__ bg, float3 bg_global, int rgb_act, int dens_act,
                                                            float *__restrict__ rgb_out, float *__restrict__ alpha_out) {
    int lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * 256) >> 5;
    for (uint32_t ray = (blockIdx.x * 256 + threadIdx.x) >> 5; ray < (uint32_t)n_rays; ray += n_warps) {
    const int32_t *ns = MODE == 0 ? numsteps_c : numsteps;
    uint32_t n = (uint32_t)ns[2 * (size_t)ray], base = (uint32_t)ns[2 * (size_t)ray + 1];
    float3 b = MODE == 0 ? make_float3(bg[3 * (size_t)ray], bg[3 * (size_t)ray + 1], bg[3 * (size_t)ray + 2]) : bg_global;
    if (n == 0) {
        if (lane == 0) { rgb_out[3 * (size_t)ray] = b.x; rgb_out[3 * (size_t)ray + 1] = b.y; rgb_out[3 * (size_t)ray + 2] = b.z; if (MODE == 1) alpha_out[ray] = 0.f; }
        continue;
    }
    float T = 1.f, ax = 0.f, ay = 0.f, az = 0.f;
    for (uint32_t k0 = 0; k0 < n; k0 += 32) {
        uint32_t k = k0 + lane; bool valid = k < n;
        float alpha = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
        if (valid) {
            float4 r = __ldg(raw + base + k);
            float dt = unwarp_dt(__ldg(coords + 7 * (size_t)(base + k) + 3));
            float density = net_to_density(r.w, dens_act);
            alpha = 1.f - __expf(-density * dt);
            cx = net_to_rgb(r.x, rgb_act); cy = net_to_rgb(r.y, rgb_act); cz = net_to_rgb(r.z, rgb_act);
        }
        float incl, ex = warp_excl_prod(1.f - alpha, lane, &incl);
        float w = alpha * (T * ex);
        ax += w * cx; ay += w * cy; az += w * cz;
        T *= __shfl_sync(0xffffffffu, incl, 31);
    }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane == 0) {
        bool add_bg = MODE == 1 ? true : (n == (uint32_t)numsteps[2 * (size_t)ray]);  // :61-64 / :199-202
        if (add_bg) { ax += T * b.x; ay += T * b.y; az += T * b.z; }
        rgb_out[3 * (size_t)ray] = ax; rgb_out[3 * (size_t)ray + 1] = ay; rgb_out[3 * (size_t)ray + 2] = az;
        if (MODE == 1) alpha_out[ray] = 1.f - T;
    }
    }
}

// compute_rgbs_grad (calc_rgb.cu:70-140)
__global__ void __launch_bounds__(256) composite_bwd_kernel(int n_rays, const float4 *__restrict__ raw, const float *__restrict__ coords, const int32_t *__restrict__ numsteps_c,
                                                            const float *__restrict__ grad_rgb, const float *__restrict__ rgb_final, const float *__restrict__ grid_mean,
                                                            int rgb_act, int dens_act, float4 *__restrict__ dl_draw) {
    int lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * 256) >> 5;
    for (uint32_t ray = (blockIdx.x * 256 + threadIdx.x) >> 5; ray < (uint32_t)n_rays; ray += n_warps) {
    uint32_t n = (uint32_t)numsteps_c[2 * (size_t)ray], base = (uint32_t)numsteps_c[2 * (size_t)ray + 1];
    if (n == 0) continue;
    float loss_scale = 128.f / (float)n_rays;                                         // :92-93
    const float l2 = rgb_act == XRB_ACT_EXPONENTIAL ? 1e-4f : 0.0f;                   // :103
    const float l1 = __ldg(grid_mean) < NERF_MIN_OPTICAL_THICKNESS ? 1e-4f : 0.0f;    // :104
    float gx = grad_rgb[3 * (size_t)ray], gy = grad_rgb[3 * (size_t)ray + 1], gz = grad_rgb[3 * (size_t)ray + 2];
    float fx = rgb_final[3 * (size_t)ray], fy = rgb_final[3 * (size_t)ray + 1], fz = rgb_final[3 * (size_t)ray + 2];
    float T = 1.f, px = 0.f, py = 0.f, pz = 0.f;  // running transmittance / prefix colour carried across 32-sample chunks
    for (uint32_t k0 = 0; k0 < n; k0 += 32) {
        uint32_t k = k0 + lane; bool valid = k < n;
        float4 r = make_float4(0, 0, 0, 0); float dt = 0.f, alpha = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
        if (valid) {
            r = __ldg(raw + base + k);
            dt = unwarp_dt(__ldg(coords + 7 * (size_t)(base + k) + 3));
            alpha = 1.f - __expf(-net_to_density(r.w, dens_act) * dt);
            cx = net_to_rgb(r.x, rgb_act); cy = net_to_rgb(r.y, rgb_act); cz = net_to_rgb(r.z, rgb_act);
        }
        float incl, ex = warp_excl_prod(1.f - alpha, lane, &incl);
        float Tb = T * ex, w = alpha * Tb, Ta = Tb * (1.f - alpha);
        float sx = px + warp_incl_sum(w * cx, lane), sy = py + warp_incl_sum(w * cy, lane), sz = pz + warp_incl_sum(w * cz, lane);  // rgb_ray2 after this sample
        if (valid) {
            float4 o;
            o.x = loss_scale * (w * gx * net_to_rgb_deriv(r.x, rgb_act) + fmaxf(0.0f, l2 * r.x));
            o.y = loss_scale * (w * gy * net_to_rgb_deriv(r.y, rgb_act) + fmaxf(0.0f, l2 * r.y));
            o.z = loss_scale * (w * gz * net_to_rgb_deriv(r.z, rgb_act) + fmaxf(0.0f, l2 * r.z));
            float dot = gx * (Ta * cx - (fx - sx)) + (gy * (Ta * cy - (fy - sy)) + gz * (Ta * cz - (fz - sz)));
            o.w = loss_scale * (net_to_density_deriv(r.w, dens_act) * (dt * dot)) + (r.w < 0.f ? -l1 : 0.0f);
            dl_draw[base + k] = o;
        }
        T *= __shfl_sync(0xffffffffu, incl, 31);
        px = __shfl_sync(0xffffffffu, sx, 31); py = __shfl_sync(0xffffffffu, sy, 31); pz = __shfl_sync(0xffffffffu, sz, 31);
    }
    }
}

// ============================================================================ occupancy grid kernels
constexpr int GRID_THREADS = 256;
static inline int stream_grid(size_t work_items) {
    size_t blocks = (work_items + GRID_THREADS - 1) / GRID_THREADS;
    size_t cap = (size_t)NUM_SMS * 8;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// mark_untrained_density_grid.cu:5-51; always writes (Q1)
__global__ void __launch_bounds__(GRID_THREADS) mark_untrained_kernel(uint32_t n_elements, float *__restrict__ grid, uint32_t n_images, const float2 *__restrict__ focal,
                                                                      const float *__restrict__ xforms, int res0, int res1) {
    extern __shared__ float s_cam[];  // [n_images][14]: 12 xform + 2 focal (cameras are read by every thread)
    for (uint32_t k = threadIdx.x; k < n_images * 14; k += blockDim.x) {
        uint32_t j = k / 14, c = k % 14;
        s_cam[k] = c < 12 ? xforms[12 * (size_t)j + c] : (c == 12 ? focal[j].x : focal[j].y);
    }
    __syncthreads();
    const float half_resx = mul_((float)res0, 0.5f), half_resy = mul_((float)res1, 0.5f);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elements; i += gridDim.x * blockDim.x) {
        uint32_t level = i / GRID_CELLS, pos_idx = i % GRID_CELLS;
        float x = (float)morton3D_invert(pos_idx >> 0), y = (float)morton3D_invert(pos_idx >> 1), z = (float)morton3D_invert(pos_idx >> 2);
        float s = __uint_as_float((127u + level) << 23);
        float px = add_(mul_(sub_(div_(add_(x, 0.5f), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        float py = add_(mul_(sub_(div_(add_(y, 0.5f), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        float pz = add_(mul_(sub_(div_(add_(z, 0.5f), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        float voxel_radius = div_(mul_(mul_(0.5f, SQRT3()), s), (float)NERF_GRIDSIZE);
        int count = 0;
        for (uint32_t j = 0; j < n_images; ++j) {
            const float *m = s_cam + 14 * j;
            float lx = sub_(px, m[9]), ly = sub_(py, m[10]), lz = sub_(pz, m[11]);
            // Eigen fixed-size-3 dot = p0 + (p1 + p2)
            float cx = add_(mul_(lx, m[0]), add_(mul_(ly, m[1]), mul_(lz, m[2])));
            float cy = add_(mul_(lx, m[3]), add_(mul_(ly, m[4]), mul_(lz, m[5])));
            float cz = add_(mul_(lx, m[6]), add_(mul_(ly, m[7]), mul_(lz, m[8])));
            if (cz > 0.f) {
                if (sub_(fabsf(cx), voxel_radius) < mul_(div_(cz, m[12]), half_resx) && sub_(fabsf(cy), voxel_radius) < mul_(div_(cz, m[13]), half_resy)) { count++; break; }
            }
        }
        grid[i] = count > 0 ? 0.f : -1.f;
    }
}

// generate_grid_samples_nerf_nonuniform.cu:6-42
__global__ void __launch_bounds__(GRID_THREADS) generate_grid_samples_kernel(uint32_t n_elements, Pcg32 rng0, uint32_t step, float lo, float hi, const float *__restrict__ grid,
                                                                             float *__restrict__ positions, int32_t *__restrict__ indices, uint32_t n_cascades, float thresh) {
    const float diag = sub_(hi, lo);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elements; i += gridDim.x * blockDim.x) {
        Pcg32 rng = rng0; rng.advance((uint64_t)(i * 4u));
        uint32_t level = (uint32_t)mul_(rng.next_float(), (float)n_cascades) % n_cascades;
        uint32_t idx = 0;
        for (uint32_t j = 0; j < 10; ++j) {
            idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % GRID_CELLS;
            idx += level * GRID_CELLS;
            if (__ldg(grid + idx) > thresh) break;
        }
        uint32_t pos_idx = idx % GRID_CELLS;
        float x = (float)morton3D_invert(pos_idx >> 0), y = (float)morton3D_invert(pos_idx >> 1), z = (float)morton3D_invert(pos_idx >> 2);
        float u0 = rng.next_float(), u1 = rng.next_float(), u2 = rng.next_float();
        float s = __uint_as_float((127u + level) << 23);
        float px = add_(mul_(sub_(div_(add_(x, u0), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        float py = add_(mul_(sub_(div_(add_(y, u1), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        float pz = add_(mul_(sub_(div_(add_(z, u2), (float)NERF_GRIDSIZE), 0.5f), s), 0.5f);
        positions[3 * (size_t)i] = div_(sub_(px, lo), diag); positions[3 * (size_t)i + 1] = div_(sub_(py, lo), diag); positions[3 * (size_t)i + 2] = div_(sub_(pz, lo), diag);
        indices[i] = (int32_t)idx;
    }
}

// splat_grid_samples_nerf_max_nearest_neighbor.cu:6-27
__global__ void __launch_bounds__(GRID_THREADS) splat_kernel(uint32_t n, const int32_t *__restrict__ indices, int padded_width, const float *__restrict__ mlp_out,
                                                             float *__restrict__ grid_tmp) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float thickness = mul_(__expf(__ldg(mlp_out + (size_t)i * padded_width)), MIN_CONE_STEPSIZE());
        atomicMax((unsigned int *)&grid_tmp[(uint32_t)indices[i]], __float_as_uint(thickness));
    }
}

// ema_grid_samples_nerf.cu:3-26, float4-vectorised stream
__global__ void __launch_bounds__(GRID_THREADS) ema_kernel(uint32_t n4, float decay, float4 *__restrict__ grid, const float4 *__restrict__ tmp) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 g = grid[i], t = __ldg(tmp + i);
        g.x = g.x < 0.f ? g.x : fmaxf(mul_(g.x, decay), t.x); g.y = g.y < 0.f ? g.y : fmaxf(mul_(g.y, decay), t.y);
        g.z = g.z < 0.f ? g.z : fmaxf(mul_(g.z, decay), t.z); g.w = g.w < 0.f ? g.w : fmaxf(mul_(g.w, decay), t.w);
        grid[i] = g;
    }
}
__global__ void ema_tail_kernel(uint32_t start, uint32_t n, float decay, float *__restrict__ grid, const float *__restrict__ tmp) {
    uint32_t i = start + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float p = grid[i]; grid[i] = p < 0.f ? p : fmaxf(mul_(p, decay), tmp[i]); }
}

// update_bitfield.cu:74-116. Mean: the reference sums fmaxf(v,0)/N per element in 1024-thread float4 blocks combined by an
// xor-butterfly, then atomically across its 512 blocks (order undefined). Here: same per-block association, blocks combined in
// index order by a second single-warp kernel => deterministic and equal to the reference run serially.
constexpr int MEAN_BLOCKS = GRID_CELLS / 4096;  // 512
__global__ void __launch_bounds__(1024) grid_mean_partial_kernel(const float4 *__restrict__ grid, float *__restrict__ partial) {
    __shared__ float sdata[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 v = __ldg(grid + (size_t)blockIdx.x * 1024 + threadIdx.x);
    const float N = (float)GRID_CELLS;
    float val = add_(add_(add_(div_(fmaxf(v.x, 0.f), N), div_(fmaxf(v.y, 0.f), N)), div_(fmaxf(v.z, 0.f), N)), div_(fmaxf(v.w, 0.f), N));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val = add_(val, __shfl_xor_sync(0xffffffffu, val, o));
    if (lane == 0) sdata[w] = val;
    __syncthreads();
    if (w == 0) {
        val = sdata[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val = add_(val, __shfl_xor_sync(0xffffffffu, val, o));
        if (lane == 0) partial[blockIdx.x] = val;
    }
}
__global__ void grid_mean_final_kernel(const float *__restrict__ partial, float *__restrict__ mean) {
    if (threadIdx.x == 0) { float t = 0.f; for (int b = 0; b < MEAN_BLOCKS; ++b) t = add_(t, partial[b]); mean[0] = t; }
}
// 8 cells -> 1 byte over all cascades; each thread builds 4 bytes (one 32-bit store) from 8 float4 loads
__global__ void __launch_bounds__(GRID_THREADS) grid_to_bitfield_kernel(uint32_t n_words, const float4 *__restrict__ grid, uint32_t *__restrict__ bitfield, const float *__restrict__ mean) {
    float m = __ldg(mean);
    float thresh = NERF_MIN_OPTICAL_THICKNESS < m ? NERF_MIN_OPTICAL_THICKNESS : m;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) {
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 v = __ldg(grid + (size_t)i * 8 + q);
            uint32_t nib = (v.x > thresh ? 1u : 0u) | (v.y > thresh ? 2u : 0u) | (v.z > thresh ? 4u : 0u) | (v.w > thresh ? 8u : 0u);
            word |= nib << (4 * q);
        }
        bitfield[i] = word;
    }
}
// update_bitfield.cu:47-71: one thread per output byte of the next level
__global__ void __launch_bounds__(GRID_THREADS) bitfield_max_pool_kernel(uint32_t n_elements, const uint2 *__restrict__ prev_level, uint8_t *__restrict__ next_level) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elements) return;
    uint2 p = __ldg(prev_level + i);
    uint8_t bits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { bits |= ((p.x >> (8 * j)) & 0xffu) ? (uint8_t)(1u << j) : 0; bits |= ((p.y >> (8 * j)) & 0xffu) ? (uint8_t)(1u << (j + 4)) : 0; }
    uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8, y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8, z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
    next_level[morton3D(x, y, z)] |= bits;
}

}  // namespace xrb

using namespace xrb;

// ============================================================================ C ABI
extern "C" {

int xrb_abi_version(void) { return 2; }   // 2: xrb_ngp_table (cell image), explicit profile events in xrb_ngp_render
int xrb_built_for_sm(void) { return 100; }
const char *xrb_last_error(void) { return g_err; }

size_t xrb_rm_rays_sampler_workspace(int n_rays) { return march_ws_bytes(n_rays) + 2048 * sizeof(float); }

int xrb_rm_rays_sampler(const float *rays_o, const float *rays_d, const uint8_t *bitfield, const float *metadata, const int32_t *img_ids, const float *xforms,
                        int n_rays, int max_samples, float aabb0, float aabb1, float near_distance, float cone_angle, uint64_t seed, int64_t n_prior_calls,
                        float *coords_out, int32_t *rays_index, int32_t *numsteps, int32_t *counters, void *workspace, void *stream) {
    (void)metadata; (void)img_ids; (void)xforms;
    XRB_REQUIRE(n_rays >= 0 && max_samples >= 0, "rays_sampler: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(rays_o && rays_d && bitfield && coords_out && rays_index && numsteps && counters && workspace, "rays_sampler: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    MarchWs w = march_ws(workspace, n_rays);
    const int ls = lane_stride(), rpb = MARCH_BLOCK / ls;
    int nb = (n_rays + rpb - 1) / rpb;
    Pcg32 rng = host_rng(seed, n_prior_calls);
    cudaMemsetAsync(w.misc + 3, 0, sizeof(uint32_t), s);   // overflow-chunk counter
    OvfPool pool{w.misc + 3, w.ovf, w.n_chunks, w.chunk_of};
    if (ls == 1) march_count_kernel<1><<<nb, MARCH_BLOCK, 0, s>>>(n_rays, rays_o, rays_d, bitfield, aabb0, aabb1, near_distance, cone_angle, rng, w.local_excl, w.startt, w.block_sum, numsteps, w.tbuf, pool);
    else if (ls == 2) march_count_kernel<2><<<nb, MARCH_BLOCK, 0, s>>>(n_rays, rays_o, rays_d, bitfield, aabb0, aabb1, near_distance, cone_angle, rng, w.local_excl, w.startt, w.block_sum, numsteps, w.tbuf, pool);
    else march_count_kernel<4><<<nb, MARCH_BLOCK, 0, s>>>(n_rays, rays_o, rays_d, bitfield, aabb0, aabb1, near_distance, cone_angle, rng, w.local_excl, w.startt, w.block_sum, numsteps, w.tbuf, pool);
    march_scan_kernel<<<1, 1024, 0, s>>>(nb, w.block_sum, w.misc, counters);
    march_emit_kernel<<<warp_grid(n_rays), 256, 0, s>>>(n_rays, rays_o, rays_d, bitfield, aabb0, aabb1, cone_angle, (uint32_t)max_samples, w.local_excl, w.startt,
                                                                             w.block_sum, w.misc, w.tbuf, w.chunk_of, w.ovf, coords_out, rays_index, numsteps, counters, rpb);
    return check_launch("rays_sampler");
}

size_t xrb_rm_compacted_coord_workspace(int n_rays) {   // the compaction uses only the scan part of the march workspace layout (local_excl, startt, block_sum, misc)
    size_t nb = (size_t)(n_rays + MARCH_RAYS_PER_BLOCK_MIN - 1) / MARCH_RAYS_PER_BLOCK_MIN;
    return sizeof(uint32_t) * ((size_t)n_rays * 2 + nb + 16);
}

int xrb_rm_compacted_coord(const float *network_output, const float *coords_in, const int32_t *numsteps, int n_rays, int max_compacted, float *coords_out,
                           int32_t *numsteps_compacted, int32_t *ray_counter, int32_t *step_counter, void *workspace, void *stream) {
    (void)network_output;
    XRB_REQUIRE(n_rays >= 0 && max_compacted >= 0, "compacted_coord: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(coords_in && numsteps && coords_out && numsteps_compacted && ray_counter && step_counter && workspace, "compacted_coord: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    MarchWs w = march_ws(workspace, n_rays);
    int nb = (n_rays + MARCH_BLOCK - 1) / MARCH_BLOCK;
    compact_count_kernel<<<nb, MARCH_BLOCK, 0, s>>>(n_rays, numsteps, w.local_excl, w.block_sum);
    // reuse the scan kernel: "counters" = {ray_counter, step_counter} must be adjacent for it; they are separate tensors in the
    // reference API, so run it on a 2-int staging area in the workspace instead.
    int32_t *stage = (int32_t *)(w.misc + 4);
    cudaMemcpyAsync(stage, ray_counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(stage + 1, step_counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    march_scan_kernel<<<1, 1024, 0, s>>>(nb, w.block_sum, w.misc, stage);
    cudaMemcpyAsync(step_counter, stage + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    int blocks = (int)(((size_t)n_rays * 32 + 255) / 256);
    compact_copy_kernel<<<blocks, 256, 0, s>>>(n_rays, (uint32_t)max_compacted, coords_in, numsteps, w.local_excl, w.block_sum, w.misc, coords_out, numsteps_compacted, ray_counter);
    return check_launch("compacted_coord");
}

int xrb_ngp_count_trained_rays(const int32_t *numsteps, const int32_t *numsteps_compacted, int n_rays, int64_t *accum, void *stream) {
    XRB_REQUIRE(n_rays >= 0, "count_trained_rays: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(numsteps && numsteps_compacted && accum && ((uintptr_t)accum & 7) == 0, "count_trained_rays: null / misaligned pointer");
    int blocks = (n_rays + 255) / 256; if (blocks > NUM_SMS * 4) blocks = NUM_SMS * 4;
    count_trained_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, numsteps, numsteps_compacted, (unsigned long long *)accum);
    return check_launch("count_trained_rays");
}

int xrb_rm_calc_rgb_forward(const float *raw, const float *coords, const int32_t *numsteps, const int32_t *numsteps_compacted, const float *bg, int n_rays, int rgb_act,
                            int dens_act, float *rgb_out, void *stream) {
    XRB_REQUIRE(n_rays >= 0, "calc_rgb_forward: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(raw && coords && numsteps && numsteps_compacted && bg && rgb_out, "calc_rgb_forward: null pointer");
    XRB_REQUIRE(((uintptr_t)raw & 15) == 0, "calc_rgb_forward: raw must be 16-byte aligned");
    int blocks = warp_grid(n_rays);
    composite_fwd_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, (const float4 *)raw, coords, numsteps, numsteps_compacted, bg, make_float3(0, 0, 0), rgb_act, dens_act,
                                                                      rgb_out, nullptr);
    return check_launch("calc_rgb_forward");
}

int xrb_rm_calc_rgb_backward(const float *raw, const int32_t *numsteps_compacted, const float *coords, const float *grad_rgb, const float *rgb, const float *grid_mean,
                             int n_rays, int rgb_act, int dens_act, float *dl_draw, void *stream) {
    XRB_REQUIRE(n_rays >= 0, "calc_rgb_backward: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(raw && coords && numsteps_compacted && grad_rgb && rgb && grid_mean && dl_draw, "calc_rgb_backward: null pointer");
    XRB_REQUIRE(((uintptr_t)raw & 15) == 0 && ((uintptr_t)dl_draw & 15) == 0, "calc_rgb_backward: raw/dl_draw must be 16-byte aligned");
    int blocks = warp_grid(n_rays);
    composite_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, (const float4 *)raw, coords, numsteps_compacted, grad_rgb, rgb, grid_mean, rgb_act, dens_act,
                                                                   (float4 *)dl_draw);
    return check_launch("calc_rgb_backward");
}

int xrb_rm_calc_rgb_inference(const float *raw, const float *coords, const int32_t *numsteps, const float *bg3_host, int n_rays, int rgb_act, int dens_act, float *rgb_out,
                              float *alpha_out, void *stream) {
    XRB_REQUIRE(n_rays >= 0, "calc_rgb_inference: negative size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(raw && coords && numsteps && bg3_host && rgb_out && alpha_out, "calc_rgb_inference: null pointer");
    XRB_REQUIRE(((uintptr_t)raw & 15) == 0, "calc_rgb_inference: raw must be 16-byte aligned");
    int blocks = warp_grid(n_rays);
    composite_fwd_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(n_rays, (const float4 *)raw, coords, numsteps, nullptr, nullptr,
                                                                      make_float3(bg3_host[0], bg3_host[1], bg3_host[2]), rgb_act, dens_act, rgb_out, alpha_out);
    return check_launch("calc_rgb_inference");
}

int xrb_rm_mark_untrained_density_grid(const float *focal, const float *xforms, int n_elements, int n_images, int res0, int res1, float *grid, void *stream) {
    XRB_REQUIRE(n_elements >= 0 && n_images >= 0, "mark_untrained: negative size");
    if (n_elements == 0) return XRB_OK;
    XRB_REQUIRE(focal && xforms && grid, "mark_untrained: null pointer");
    size_t smem = (size_t)n_images * 14 * sizeof(float);
    XRB_REQUIRE(smem <= 200 * 1024, "mark_untrained: too many images for one shared-memory camera table (max 3657)");
    if (smem > 48 * 1024) cudaFuncSetAttribute(mark_untrained_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mark_untrained_kernel<<<stream_grid(n_elements), GRID_THREADS, smem, (cudaStream_t)stream>>>((uint32_t)n_elements, grid, (uint32_t)n_images, (const float2 *)focal, xforms, res0, res1);
    return check_launch("mark_untrained");
}

int xrb_rm_generate_grid_samples(const float *grid, int ema_step, int n_elements, int max_cascade, float thresh, float aabb0, float aabb1, uint64_t seed, int64_t n_prior_calls,
                                 float *positions, int32_t *indices, void *stream) {
    XRB_REQUIRE(n_elements >= 0 && max_cascade >= 0 && max_cascade < (int)NERF_CASCADES, "generate_grid_samples: bad size/cascade");
    if (n_elements == 0) return XRB_OK;
    XRB_REQUIRE(grid && positions && indices, "generate_grid_samples: null pointer");
    Pcg32 rng = host_rng(seed, n_prior_calls);
    generate_grid_samples_kernel<<<stream_grid(n_elements), GRID_THREADS, 0, (cudaStream_t)stream>>>((uint32_t)n_elements, rng, (uint32_t)ema_step, aabb0, aabb1, grid, positions, indices,
                                                                                                     (uint32_t)max_cascade + 1, thresh);
    return check_launch("generate_grid_samples");
}

int xrb_rm_splat_grid_samples(const float *mlp_out, const int32_t *indices, int padded_width, int n, float *grid_tmp, void *stream) {
    XRB_REQUIRE(n >= 0 && padded_width >= 1, "splat: bad size");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(mlp_out && indices && grid_tmp, "splat: null pointer");
    splat_kernel<<<stream_grid(n), GRID_THREADS, 0, (cudaStream_t)stream>>>((uint32_t)n, indices, padded_width, mlp_out, grid_tmp);
    return check_launch("splat");
}

int xrb_rm_ema_grid_samples(const float *grid_tmp, int n_elements, float decay, float *grid, void *stream) {
    XRB_REQUIRE(n_elements >= 0, "ema: negative size");
    if (n_elements == 0) return XRB_OK;
    XRB_REQUIRE(grid_tmp && grid, "ema: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    bool aligned = (((uintptr_t)grid | (uintptr_t)grid_tmp) & 15) == 0;
    uint32_t n4 = aligned ? (uint32_t)n_elements / 4 : 0;
    if (n4) ema_kernel<<<stream_grid(n4), GRID_THREADS, 0, s>>>(n4, decay, (float4 *)grid, (const float4 *)grid_tmp);
    uint32_t rest = (uint32_t)n_elements - n4 * 4;
    if (rest) ema_tail_kernel<<<(rest + 255) / 256, 256, 0, s>>>(n4 * 4, (uint32_t)n_elements, decay, grid, grid_tmp);
    return check_launch("ema");
}

size_t xrb_rm_update_bitfield_workspace(void) { return MEAN_BLOCKS * sizeof(float); }

int xrb_rm_update_bitfield(const float *grid, float *mean, uint8_t *bitfield, void *workspace, void *stream) {
    XRB_REQUIRE(grid && mean && bitfield && workspace, "update_bitfield: null pointer");
    XRB_REQUIRE(((uintptr_t)grid & 15) == 0 && ((uintptr_t)bitfield & 7) == 0 && ((uintptr_t)workspace & 3) == 0, "update_bitfield: grid must be 16-byte, bitfield 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    // the per-block partial sums of the fixed-order mean live in caller-owned scratch: no allocation inside the call (CUDA-graph capturable),
    // nothing shared between streams or host threads
    float *partial = (float *)workspace;
    grid_mean_partial_kernel<<<MEAN_BLOCKS, 1024, 0, s>>>((const float4 *)grid, partial);
    grid_mean_final_kernel<<<1, 32, 0, s>>>(partial, mean);
    uint32_t n_words = GRID_CELLS / 8 * NERF_CASCADES / 4;
    grid_to_bitfield_kernel<<<stream_grid(n_words), GRID_THREADS, 0, s>>>(n_words, (const float4 *)grid, (uint32_t *)bitfield, mean);
    for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
        uint32_t n = GRID_CELLS / 64;
        bitfield_max_pool_kernel<<<(n + GRID_THREADS - 1) / GRID_THREADS, GRID_THREADS, 0, s>>>(n, (const uint2 *)(bitfield + (size_t)GRID_CELLS * (level - 1) / 8), bitfield + (size_t)GRID_CELLS * level / 8);
    }
    return check_launch("update_bitfield");
}

}  // extern "C"
