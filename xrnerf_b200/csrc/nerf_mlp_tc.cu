// xrnerf_b200 — NerfMLP.run_mlp (/root/reference/xrnerf/models/mlps/nerf_mlp.py:70-94: 8x256 ReLU MLP, skip-concat after layer 4,
// alpha head, 256-wide feature layer, 128-wide view branch, rgb head) as ONE persistent tcgen05 kernel.
// The reference runs it as 11 cuBLAS fp32 GEMMs + 2 cat per 32 768-row chunk with every [chunk,256] activation round-tripping HBM.
//
// Per CTA (1 per SM, 192 threads), per 128-row tile:
//   warps 0-3  compute warpgroup: thread r owns row r — converts its `embedded` row to fp16 into the A operand blocks, and after each
//              layer reads its TMEM accumulator row (tcgen05.ld), adds the bias, applies ReLU, rounds to fp16 and rewrites the A blocks;
//   warp 4     TMA producer: streams the pre-swizzled weight slabs ([N x 64] fp16, K-major SWIZZLE_128B) of all 12 layers from the
//              L2-resident weight image through a 3-slot shared-memory ring (cp.async.bulk + mbarrier expect_tx);
//   warp 5     MMA issuer: per slab 4 x tcgen05.mma (M=128, N=256|128|16, K=16), tcgen05.commit frees the slab slot; after the last
//              slab of a layer a commit signals the epilogue.
// A operand: [128 x 64]-blocks, K-major, 128-byte swizzle: blocks 0-3 hidden state h (256 wide), 4..4+AUX-1 the point encoding
// (63 -> 64 for NeRF, 96 -> 128 for Mip-NeRF), last block the view-direction encoding (27 -> 64). Accumulators: TMEM columns 0-255
// (main) and 256-271 (alpha head, issued next to the feature layer on the same A tile).
// Numeric contract: fp16 operands (weights and activations), fp32 accumulate + fp32 bias; parity tolerance vs the fp32 reference
// is stated in tests/test_gpu_nerf_mlp.py.
#include "tc.cuh"
#include "common.cuh"
#include <cuda_fp16.h>

namespace xrb {

constexpr int NM_W = 256;
constexpr int NM_MAX_LAYERS = 12;
constexpr int NM_SLOTS = 3;
constexpr uint32_t NM_SLAB_BYTES = 272 * 128;   // largest slab: feature (256) + alpha (16) rows
constexpr int NM_THREADS = 192;

struct NmLayer {
    int n_kblocks;        // K-blocks of 64
    int kblock_src[6];    // which A block feeds each K-block
    int N;                // main output width (256 / 128 / 16)
    int n_alpha;          // 16 when the alpha head rides along (extra rows in every slab of this layer), else 0
    int bias_off;         // offset into the bias vector (main), alpha bias follows at bias_off + N when n_alpha
    int relu;
    int out_blocks;       // how many 64-wide A blocks the epilogue writes (N/64; 0 for the last layer)
};
struct NmPlan {
    int n_layers, aux_blocks, input_ch, input_ch_dirs;
    NmLayer layer[NM_MAX_LAYERS];
};

__device__ __forceinline__ uint32_t nm_pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t nm_sw128(uint32_t row, uint32_t chunk16) { return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk16 ^ (row & 7u)) << 4); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }

__global__ void __launch_bounds__(NM_THREADS, 1) nerf_mlp_tc_kernel(NmPlan plan, const uint8_t *__restrict__ weight_image, const float *__restrict__ bias, const float *__restrict__ embedded,
                                                                     int64_t n_rows, float *__restrict__ raw) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    const int n_ablocks = 5 + plan.aux_blocks;
    uint8_t *A = base;                                        // n_ablocks x 16 KB
    uint8_t *slab = A + (size_t)n_ablocks * 16384;            // NM_SLOTS x NM_SLAB_BYTES (1024-aligned: 34816 = 34 x 1024)
    uint64_t *bars = (uint64_t *)(slab + (size_t)NM_SLOTS * NM_SLAB_BYTES);
    uint64_t *full = bars, *empty = bars + NM_SLOTS, *a_ready = bars + 2 * NM_SLOTS, *acc_full = bars + 2 * NM_SLOTS + 1;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NM_SLOTS + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NM_SLOTS; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        tc::mbar_init(a_ready, 128);
        tc::mbar_init(acc_full, 1);
        tc::fence_mbar_init();
    }
    if (warp == 4) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int64_t n_tiles = (n_rows + 127) / 128;
    const int C = plan.input_ch + plan.input_ch_dirs;

    if (warp == 4) {
        // ===================================================== TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                size_t off = 0;
                for (int l = 0; l < plan.n_layers; ++l) {
                    const uint32_t bytes = (uint32_t)(plan.layer[l].N + plan.layer[l].n_alpha) * 128u;
                    for (int kb = 0; kb < plan.layer[l].n_kblocks; ++kb, ++it) {
                        const uint32_t slot = it % NM_SLOTS, round = it / NM_SLOTS;
                        if (round > 0) tc::mbar_wait(empty + slot, (round - 1) & 1);
                        tc::mbar_expect_tx(full + slot, bytes);
                        tc::tma_bulk_g2s(slab + (size_t)slot * NM_SLAB_BYTES, weight_image + off, bytes, full + slot);
                        off += bytes;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            uint32_t it = 0, a_phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < plan.n_layers; ++l) {
                    const NmLayer &L = plan.layer[l];
                    tc::mbar_wait(a_ready, a_phase); a_phase ^= 1;   // all 128 rows of this layer's input are in smem, previous accumulator drained
                    tc::tc_fence_after_sync();
                    const uint32_t idesc = tc::idesc_f16_m128((uint32_t)L.N), idesc_a = tc::idesc_f16_m128(16);
                    for (int kb = 0; kb < L.n_kblocks; ++kb, ++it) {
                        const uint32_t slot = it % NM_SLOTS, round = it / NM_SLOTS;
                        tc::mbar_wait(full + slot, round & 1);
                        tc::tc_fence_after_sync();
                        const uint32_t a0 = tc::smem_u32(A + (size_t)L.kblock_src[kb] * 16384), b0 = tc::smem_u32(slab + (size_t)slot * NM_SLAB_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (kb | k) ? 1u : 0u;
                            tc::mma_f16_ss(tmem, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), idesc, acc);
                            if (L.n_alpha) tc::mma_f16_ss(tmem + 256, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + (uint32_t)L.N * 128u + k * 32), idesc_a, acc);
                        }
                        tc::mma_commit(empty + slot);              // slab slot reusable once these MMAs have read it
                    }
                    tc::mma_commit(acc_full);                      // accumulator of layer l complete
                }
            }
        }
    } else {
        // ===================================================== compute warpgroup (thread == row)
        const uint32_t row = threadIdx.x;
        const uint32_t taddr = tmem + ((warp * 32u) << 16);
        uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * 128 + row;
            const bool valid = i < n_rows;
            // ---- stage the encodings: point encoding -> blocks 4.., direction encoding -> last block (zero padded)
            {
                const float *e = embedded + (size_t)(valid ? i : 0) * C;
                for (int blk = 0; blk < plan.aux_blocks + 1; ++blk) {
                    const bool is_dir = blk == plan.aux_blocks;
                    const int width = is_dir ? plan.input_ch_dirs : plan.input_ch, c0 = is_dir ? plan.input_ch : blk * 64;
                    const int local0 = is_dir ? 0 : blk * 64;
                    uint8_t *dst = A + (size_t)(4 + blk) * 16384;
                    for (int ch = 0; ch < 8; ++ch) {
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) { int k = local0 + ch * 8 + q; v[q] = (valid && k < width) ? __ldg(e + c0 + (is_dir ? ch * 8 + q : ch * 8 + q)) : 0.f; }
                        *reinterpret_cast<uint4 *>(dst + nm_sw128(row, ch)) = make_uint4(nm_pack_h2(v[0], v[1]), nm_pack_h2(v[2], v[3]), nm_pack_h2(v[4], v[5]), nm_pack_h2(v[6], v[7]));
                    }
                }
            }
            tc::fence_proxy_async_smem();
            mbar_arrive(a_ready);
            float alpha_out = 0.f;
            for (int l = 0; l < plan.n_layers; ++l) {
                const NmLayer &L = plan.layer[l];
                tc::mbar_wait(acc_full, acc_phase); acc_phase ^= 1;
                tc::tc_fence_after_sync();
                const float *b = bias + L.bias_off;
                if (L.n_alpha) { float a16[16]; tc::tmem_ld16(taddr + 256, a16); alpha_out = a16[0] + __ldg(b + L.N); }
                if (L.out_blocks > 0) {
                    for (int blk = 0; blk < L.out_blocks; ++blk) {
                        uint8_t *dst = A + (size_t)blk * 16384;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t r[32];
                            tc::tmem_ld32(taddr + blk * 64 + half * 32, r);
                            const float *bb = b + blk * 64 + half * 32;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float v[8];
#pragma unroll
                                for (int e2 = 0; e2 < 8; ++e2) { float x = __uint_as_float(r[8 * q + e2]) + __ldg(bb + 8 * q + e2); v[e2] = L.relu ? fmaxf(x, 0.f) : x; }
                                *reinterpret_cast<uint4 *>(dst + nm_sw128(row, half * 4 + q)) = make_uint4(nm_pack_h2(v[0], v[1]), nm_pack_h2(v[2], v[3]), nm_pack_h2(v[4], v[5]), nm_pack_h2(v[6], v[7]));
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    tc::tc_fence_before_sync();
                    mbar_arrive(a_ready);
                } else {
                    float o16[16];
                    tc::tmem_ld16(taddr, o16);
                    if (valid) reinterpret_cast<float4 *>(raw)[i] = make_float4(o16[0] + __ldg(b), o16[1] + __ldg(b + 1), o16[2] + __ldg(b + 2), alpha_out);
                    tc::tc_fence_before_sync();
                    // the next tile's staging writes only the encoding blocks, which the (completed) last layers no longer read
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc<512>(tmem);
}

}  // namespace xrb

using namespace xrb;

extern "C" {

// weight image / bias layout are produced by the host (xrnerf_b200/nerf_mlp.py: pack_nerf_mlp) following the same NmPlan.
int xrb_nerf_mlp_forward(const void *weight_image, const float *bias, const float *embedded, int64_t n_rows, int input_ch, int input_ch_dirs, float *raw, void *stream) {
    XRB_REQUIRE(n_rows >= 0, "nerf_mlp_forward: negative size");
    if (!((input_ch == 63 && input_ch_dirs == 27) || (input_ch == 96 && input_ch_dirs == 27))) {
        set_error("nerf_mlp_forward: implemented for NerfMLP(netdepth=8, netwidth=256, skips=[4], use_viewdirs) with (63,27) or (96,27) input channels");
        return XRB_E_UNSUPPORTED;
    }
    if (n_rows == 0) return XRB_OK;
    XRB_REQUIRE(weight_image && bias && embedded && raw, "nerf_mlp_forward: null pointer");
    XRB_REQUIRE(((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)raw & 15) == 0, "nerf_mlp_forward: weight image / raw must be 16-byte aligned");
    NmPlan p{};
    p.input_ch = input_ch; p.input_ch_dirs = input_ch_dirs; p.aux_blocks = (input_ch + 63) / 64;
    const int aux = p.aux_blocks, dir_blk = 4 + aux;
    int nl = 0, boff = 0;
    auto add = [&](int nk, const int *src, int N, int n_alpha, int relu, int out_blocks) {
        NmLayer &L = p.layer[nl++]; L.n_kblocks = nk; for (int k = 0; k < nk; ++k) L.kblock_src[k] = src[k];
        L.N = N; L.n_alpha = n_alpha; L.bias_off = boff; L.relu = relu; L.out_blocks = out_blocks; boff += N + n_alpha;
    };
    const int h4[4] = {0, 1, 2, 3};
    int pts_src[2] = {4, 5};
    add(aux, pts_src, 256, 0, 1, 4);                                        // pts_linears.0
    for (int l = 1; l <= 4; ++l) add(4, h4, 256, 0, 1, 4);                  // pts_linears.1-4
    { int s[6] = {0, 1, 2, 3, 4, 5}; add(4 + aux, s, 256, 0, 1, 4); }       // pts_linears.5 on cat([pts, h]) (image columns permuted to [h | pts])
    add(4, h4, 256, 0, 1, 4); add(4, h4, 256, 0, 1, 4);                     // pts_linears.6-7
    add(4, h4, 256, 16, 0, 4);                                              // feature_linear (+ alpha_linear rows)
    { int s[6] = {0, 1, 2, 3, dir_blk, 0}; add(5, s, 128, 0, 1, 2); }       // views_linears.0 on cat([feature, dirs])
    { int s[6] = {0, 1, 0, 0, 0, 0}; add(2, s, 16, 0, 0, 0); }              // rgb_linear
    p.n_layers = nl;
    size_t smem = 1024 + (size_t)(5 + aux) * 16384 + (size_t)NM_SLOTS * NM_SLAB_BYTES + 8 * (2 * NM_SLOTS + 2) + 16;
    cudaFuncSetAttribute(nerf_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t n_tiles = (n_rows + 127) / 128;
    int grid = (int)(n_tiles < sms ? n_tiles : sms);
    nerf_mlp_tc_kernel<<<grid, NM_THREADS, smem, (cudaStream_t)stream>>>(p, (const uint8_t *)weight_image, bias, embedded, n_rows, raw);
    return check_launch("nerf_mlp_forward");
}

}  // extern "C"
