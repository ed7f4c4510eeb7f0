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
#include <stdlib.h>

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


// =====================================================================================================================
// v2: the epilogue is the bottleneck of v1 (one warpgroup drains 256 accumulator columns per layer while the tensor core idles).
//   * TWO compute warpgroups per tile: WG0 owns output columns [0,N/2), WG1 owns [N/2,N) of the same 128 rows;
//   * every layer is issued as four MMA phases  S1=(half0; aux+lo K-blocks) S2=(half1; aux+lo) S3=(half0; hi) S4=(half1; hi)
//     where "lo"/"hi" are the A blocks written by WG0/WG1 in the previous layer: half0's accumulator completes after S3, so WG0
//     drains it while S4 runs, and the next layer's S1/S2 start as soon as WG0 is done (they only read lo blocks);
//   * accumulators double-buffered in TMEM by layer parity (2 x 256 columns);
//   * weights stream as HALF slabs ([N/2 x 64] fp16, <=16 KB) in exactly that issue order through a 6-slot ring;
//   * biases live in shared memory (LDS.128 instead of 256 LDG per thread per layer); the 1-wide alpha head is a register dot
//     product inside the epilogue of pts_linears.7 instead of an N=16 MMA.
constexpr int N2_SLOTS = 6;        // two rings of 3 half-slab slots: ring h feeds the issuer of output half h
constexpr uint32_t N2_SLAB_BYTES = 128 * 128;
constexpr int N2_THREADS = 384;   // 8 compute warps + slab producer (8) + MMA issuer half 0 (9) + encoding producer (10) + MMA issuer half 1 (11)
constexpr int N2_MAX_LAYERS = 11;

struct N2Layer {
    int n_first, n_second;      // K-blocks issued in S1/S2 (aux + lo) and in S3/S4 (hi)
    int src[7];                 // A block per K-block, first-phase blocks then second-phase blocks
    int N, n_halves, relu, writes_a, alpha_dot, bias_off;
};
struct N2Plan { int n_layers, aux_blocks, input_ch, input_ch_dirs, bias_total, dbg; N2Layer layer[N2_MAX_LAYERS]; };  // dbg bit0: skip TMA copies, bit1: skip MMAs, bit2: skip epilogue math (attribution experiments)

__global__ void __launch_bounds__(N2_THREADS, 1) nerf_mlp_tc2_kernel(N2Plan plan, const uint8_t *__restrict__ weight_image, const float *__restrict__ bias_g, const uint8_t *__restrict__ enc_image,
                                                                      int64_t n_rows, float *__restrict__ raw) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    const int n_ablocks = 5 + plan.aux_blocks;
    uint8_t *A = base;
    uint8_t *slab = A + (size_t)n_ablocks * 16384;
    float *bias = (float *)(slab + (size_t)N2_SLOTS * N2_SLAB_BYTES);          // all biases, then Wa[256], ba
    float *alpha_part = bias + ((plan.bias_total + 3) & ~3);                    // [128] partial alpha of WG1
    __half *bias_h = (__half *)(alpha_part + 128);                              // fp16 copy of the biases for the packed-half2 epilogue
    uint64_t *bars = (uint64_t *)(bias_h + ((plan.bias_total + 7) & ~7));
    uint64_t *full = bars, *empty = bars + N2_SLOTS, *a_ready = bars + 2 * N2_SLOTS /*[2]*/, *acc_ready = bars + 2 * N2_SLOTS + 2 /*[2]*/;
    uint64_t *enc_full = bars + 2 * N2_SLOTS + 4 /*[2]: pts, dir*/, *enc_free = bars + 2 * N2_SLOTS + 6 /*[2]*/;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * N2_SLOTS + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < N2_SLOTS; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        tc::mbar_init(a_ready, 4); tc::mbar_init(a_ready + 1, 4);   // one elected arrival per warp (128 same-address arrives cost ~0.3 us per hop)
        tc::mbar_init(acc_ready, 1); tc::mbar_init(acc_ready + 1, 1);
        for (int q = 0; q < 2; ++q) { tc::mbar_init(enc_full + q, 1); tc::mbar_init(enc_free + q, 2); }   // both issuers release the encoding blocks
        tc::fence_mbar_init();
    }
    for (int k = threadIdx.x; k < plan.bias_total; k += N2_THREADS) { float bv = bias_g[k]; bias[k] = bv; bias_h[k] = __float2half_rn(bv); }
    if (warp == 8) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int64_t n_tiles = (n_rows + 127) / 128;
    const uint32_t enc_tile_bytes = (uint32_t)(plan.aux_blocks + 1) * 16384u;

    if (warp == 10) {
        // ===================================================== encoding producer: the tile's fp16, pre-swizzled encoding blocks land straight in
        // the A operand blocks by TMA (no staging instructions, HBM latency off the critical path): the point blocks of tile t+1 are fetched as
        // soon as pts_linears.5 of tile t has consumed them, the direction block as soon as views_linears.0 has.
        if (lane == 0) {
            uint32_t n = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
                const uint8_t *src = enc_image + (size_t)tile * enc_tile_bytes;
                if (n > 0) tc::mbar_wait(enc_free, (n - 1) & 1);
                tc::mbar_expect_tx(enc_full, (uint32_t)plan.aux_blocks * 16384u);
                tc::tma_bulk_g2s(A + (size_t)4 * 16384, src, (uint32_t)plan.aux_blocks * 16384u, enc_full);
                if (n > 0) tc::mbar_wait(enc_free + 1, (n - 1) & 1);
                tc::mbar_expect_tx(enc_full + 1, 16384u);
                tc::tma_bulk_g2s(A + (size_t)(4 + plan.aux_blocks) * 16384, src + (size_t)plan.aux_blocks * 16384, 16384u, enc_full + 1);
            }
        }
    } else if (warp == 8) {
        // ===================================================== TMA producer: per K-block, half-slab 0 -> ring 0 (issuer 0), half-slab 1 -> ring 1
        if (lane == 0) {
            uint32_t it[2] = {0, 0};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                size_t off = 0;
                for (int l = 0; l < plan.n_layers; ++l) {
                    const N2Layer &L = plan.layer[l];
                    const uint32_t bytes = (uint32_t)(L.N / L.n_halves) * 128u;
                    for (int kb = 0; kb < L.n_first + L.n_second; ++kb)
                        for (int nh = 0; nh < L.n_halves; ++nh) {
                            const uint32_t i = it[nh]++, slot = nh * 3 + i % 3, round = i / 3;
                            if (round > 0) tc::mbar_wait(empty + slot, (round - 1) & 1);
                            if (plan.dbg & 1) { mbar_arrive(full + slot); }
                            else { tc::mbar_expect_tx(full + slot, bytes); tc::tma_bulk_g2s(slab + (size_t)slot * N2_SLAB_BYTES, weight_image + off, bytes, full + slot); }
                            off += bytes;
                        }
                }
            }
        }
    } else if (warp == 9 || warp == 11) {
        // ===================================================== two MMA issuers, one per output half (a single issuing thread is the bottleneck:
        // ~60-90 cycles per tcgen05.mma issue + ~200 per slab wait/commit, scripts/micro/lat.cu). Issuer h: S(h; aux+lo) after WG0's rows are in,
        // S(h; hi) after WG1's; then commits accumulator half h.
        if (lane == 0) {
            const int nh = warp == 9 ? 0 : 1;
            uint32_t it = 0, lo_phase = 0, hi_phase = 0, lcount = 0, tcount = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                for (int l = 0; l < plan.n_layers; ++l, ++lcount) {
                    const N2Layer &L = plan.layer[l];
                    if (l == 0) { tc::mbar_wait(enc_full, tcount & 1); tc::tc_fence_after_sync(); }
                    if (l == plan.n_layers - 2) { tc::mbar_wait(enc_full + 1, tcount & 1); tc::tc_fence_after_sync(); }
                    const uint32_t nh_w = (uint32_t)(L.N / L.n_halves);
                    const uint32_t idesc = tc::idesc_f16_m128(nh_w);
                    const uint32_t dbase = tmem + (lcount & 1u) * 256u + nh * nh_w;
                    const bool active = nh < L.n_halves;
                    for (int second = 0; second < 2; ++second) {
                        if (second == 0) { tc::mbar_wait(a_ready, lo_phase); lo_phase ^= 1; } else { tc::mbar_wait(a_ready + 1, hi_phase); hi_phase ^= 1; }
                        tc::tc_fence_after_sync();
                        if (!active) continue;
                        const int kb0 = second ? L.n_first : 0, kb1 = second ? L.n_first + L.n_second : L.n_first;
                        for (int kb = kb0; kb < kb1; ++kb, ++it) {
                            const uint32_t slot = nh * 3 + it % 3, round = it / 3;
                            tc::mbar_wait(full + slot, round & 1);
                            tc::tc_fence_after_sync();
                            const uint32_t a0 = tc::smem_u32(A + (size_t)L.src[kb] * 16384), b0 = tc::smem_u32(slab + (size_t)slot * N2_SLAB_BYTES);
                            if (!(plan.dbg & 2)) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) tc::mma_f16_ss(dbase, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), idesc, (kb | k) ? 1u : 0u);
                            }
                            tc::mma_commit(empty + slot);
                        }
                    }
                    if (active) tc::mma_commit(acc_ready + nh);
                    if (l == 5) tc::mma_commit(enc_free);                          // my MMAs of pts_linears.5 were the last readers of the point-encoding blocks
                    if (l == plan.n_layers - 2) tc::mma_commit(enc_free + 1);      // ... of the direction block
                }
            }
        }
    } else {
        // ===================================================== two compute warpgroups (thread == row, WG == column half)
        const uint32_t wg = warp >> 2, row = threadIdx.x & 127;
        const uint32_t lane_base = ((warp & 3) * 32u) << 16;
        const uint32_t r7 = row & 7u, row_off = (row >> 3) * 1024u + r7 * 128u;   // my row inside a [128x64] swizzled block; chunk c lives at ((c ^ r7) << 4)
        uint32_t acc_phase = 0, lcount = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * 128 + row;
            const bool valid = i < n_rows;
            // (encodings arrive by TMA; this arrival only tells the MMA issuer that my warp's reads of the previous tile's accumulators are done)
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready + wg);
            float alpha_acc = 0.f;
            for (int l = 0; l < plan.n_layers; ++l, ++lcount) {
                const N2Layer &L = plan.layer[l];
                const int nh_w = L.N / L.n_halves;
                const bool last = l == plan.n_layers - 1;
                if ((int)wg < L.n_halves) {
                    if (lane == 0) tc::mbar_wait(acc_ready + wg, acc_phase);   // one poller per warp
                    acc_phase ^= 1;
                    __syncwarp();
                    tc::tc_fence_after_sync();
                    const uint32_t taddr = tmem + lane_base + (lcount & 1u) * 256u + wg * nh_w;
                    const float *b = bias + L.bias_off + wg * nh_w;
                    if (!last) {
                        // lean epilogue: acc -> fp16 (F2FP), + bias and ReLU as packed half2 ops (HADD2/HMNMX2), one LDS.128 of fp16 biases and one
                        // STS.128 per 8 columns, swizzled chunk offsets precomputed per thread (ncu/attribution: the epilogue was 1.6 of 4.3 ms)
                        const __half2 zero2 = __float2half2_rn(0.f);
                        const __half *bh = bias_h + L.bias_off + wg * nh_w;
                        for (int c2 = 0; c2 < ((plan.dbg & 4) ? 0 : nh_w / 64); ++c2) {
                            uint32_t r64[64];
                            tc::tmem_ld64(taddr + c2 * 64, r64);          // one TMEM round trip per 64 columns
                            const int colb = wg * nh_w + c2 * 64;         // first column of this 64-block inside the layer output (64-aligned)
                            uint8_t *dst = A + (size_t)(colb >> 6) * 16384 + row_off;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const uint4 bq = *reinterpret_cast<const uint4 *>(bh + c2 * 64 + 8 * q);
                                const __half2 *b2 = reinterpret_cast<const __half2 *>(&bq);
                                __half2 h[4];
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2) {
                                    h[e2] = __hadd2(__floats2half2_rn(__uint_as_float(r64[8 * q + 2 * e2]), __uint_as_float(r64[8 * q + 2 * e2 + 1])), b2[e2]);
                                    if (L.relu) h[e2] = __hmax2(h[e2], zero2);
                                }
                                if (L.alpha_dot) {   // alpha_linear on the fp16 output of pts_linears.7: partial dot over my columns
                                    const float *wa = bias + plan.bias_total - 257 + colb + 8 * q;
#pragma unroll
                                    for (int e2 = 0; e2 < 4; ++e2) { float2 f = __half22float2(h[e2]); alpha_acc = fmaf(f.x, wa[2 * e2], alpha_acc); alpha_acc = fmaf(f.y, wa[2 * e2 + 1], alpha_acc); }
                                }
                                *reinterpret_cast<uint4 *>(dst + (((uint32_t)q ^ r7) << 4)) = *reinterpret_cast<uint4 *>(h);
                            }
                        }
                        if (L.alpha_dot && wg == 1) alpha_part[row] = alpha_acc;
                    } else {
                        float o16[16];
                        tc::tmem_ld16(taddr, o16);
                        tc::named_bar_sync(1, 256);   // WG1's partial alpha is in shared memory
                        const float alpha = alpha_acc + alpha_part[row] + bias[plan.bias_total - 1];
                        if (valid) reinterpret_cast<float4 *>(raw)[i] = make_float4(o16[0] + b[0], o16[1] + b[1], o16[2] + b[2], alpha);
                    }
                } else if (last) {
                    tc::named_bar_sync(1, 256);
                }
                if (!last) {
                    tc::fence_proxy_async_smem();
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_ready + wg);
                } else {
                    tc::tc_fence_before_sync();
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc<512>(tmem);
}


// ---- encoding tile images for v2: per 128-row tile, (aux+1) blocks of [128 x 64] fp16 in the UMMA layout (K-major, 128-byte swizzle),
// point-encoding block(s) first, direction block last, zero padded; written by fully coalesced 16-byte stores.
__global__ void __launch_bounds__(256) pack_embedded_tiles_kernel(const float *__restrict__ embedded, int64_t n_rows, int ic, int icd, int aux, uint8_t *__restrict__ image) {
    const int chunks_per_row = (aux + 1) * 8;
    const int64_t n_tiles = (n_rows + 127) / 128, total = n_tiles * 128 * chunks_per_row;
    const int C = ic + icd;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / chunks_per_row; const int cc = (int)(idx - row * chunks_per_row);
        const int blk = cc >> 3, ch = cc & 7;
        const bool is_dir = blk == aux;
        const int width = is_dir ? icd : ic, c0 = is_dir ? ic : 0, k0 = (is_dir ? 0 : blk * 64) + ch * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (row < n_rows && k0 + q < width) ? __ldg(embedded + (size_t)row * C + c0 + k0 + q) : 0.f;
        const int64_t tile = row >> 7; const uint32_t r = (uint32_t)(row & 127);
        *reinterpret_cast<uint4 *>(image + ((size_t)tile * (aux + 1) + blk) * 16384 + nm_sw128(r, ch)) =
            make_uint4(nm_pack_h2(v[0], v[1]), nm_pack_h2(v[2], v[3]), nm_pack_h2(v[4], v[5]), nm_pack_h2(v[6], v[7]));
    }
}
// BaseEmbedder positional encoding (embedders/base.py:26-52) computed straight into the tile image: pts f32[n,3], viewdirs f32[n/S,3]
// ray mode (rays_o != NULL): `pts` is z_vals f32[n_rays, S] and the sample position o + d*z is formed in registers (GetPts, create.py:588-597,
// never materialised: 491 MB per 800x800 image in the reference)
__global__ void __launch_bounds__(256) posenc_tiles_kernel(const float *__restrict__ pts, const float *__restrict__ viewdirs, int64_t n_rows, int samples_per_ray, int multires,
                                                           int multires_dirs, uint8_t *__restrict__ image, const float *__restrict__ rays_o, const float *__restrict__ rays_d) {
    const int ic = 3 + 6 * multires, icd = 3 + 6 * multires_dirs, aux = (ic + 63) / 64, chunks_per_row = (aux + 1) * 8;
    const int64_t n_tiles = (n_rows + 127) / 128, total = n_tiles * 128 * chunks_per_row;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / chunks_per_row; const int cc = (int)(idx - row * chunks_per_row);
        const int blk = cc >> 3, ch = cc & 7;
        const bool is_dir = blk == aux;
        const int width = is_dir ? icd : ic, k0 = (is_dir ? 0 : blk * 64) + ch * 8;
        float v[8];
        const float *src = row < n_rows ? (is_dir ? viewdirs + 3 * (row / samples_per_ray) : pts + 3 * row) : nullptr;
        float p3[3];
        if (src && !is_dir && rays_o) {
            const int64_t ray = row / samples_per_ray; const float z = pts[row];
            p3[0] = __fadd_rn(rays_o[3 * ray], __fmul_rn(rays_d[3 * ray], z)); p3[1] = __fadd_rn(rays_o[3 * ray + 1], __fmul_rn(rays_d[3 * ray + 1], z)); p3[2] = __fadd_rn(rays_o[3 * ray + 2], __fmul_rn(rays_d[3 * ray + 2], z));   // torch: mul then add, no FMA
            src = p3;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + q;
            float out = 0.f;
            if (src && k < width) {
                if (k < 3) out = src[k];
                else { int qq = k - 3, band = qq / 6, r6 = qq % 6; float x = src[r6 % 3] * exp2f((float)band); out = r6 < 3 ? sinf(x) : cosf(x); }
            }
            v[q] = out;
        }
        const int64_t tile = row >> 7; const uint32_t r = (uint32_t)(row & 127);
        *reinterpret_cast<uint4 *>(image + ((size_t)tile * (aux + 1) + blk) * 16384 + nm_sw128(r, ch)) =
            make_uint4(nm_pack_h2(v[0], v[1]), nm_pack_h2(v[2], v[3]), nm_pack_h2(v[4], v[5]), nm_pack_h2(v[6], v[7]));
    }
}

}  // namespace xrb

using namespace xrb;

extern "C" int xrb_internal_posenc_tiles_fast(const float *pts, const float *viewdirs, int64_t n_rows, int samples_per_ray, void *enc_image, const float *rays_o, const float *rays_d, void *stream);   // nerf.cu

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

// v2 entry (see the kernel comment); image/bias from xrnerf_b200.nerf_mlp.pack_nerf_mlp_v2
int xrb_nerf_mlp_forward_v2(const void *weight_image, const float *bias, const void *enc_image, int64_t n_rows, int input_ch, int input_ch_dirs, float *raw, void *stream) {
    XRB_REQUIRE(n_rows >= 0, "nerf_mlp_forward_v2: negative size");
    if (!((input_ch == 63 && input_ch_dirs == 27) || (input_ch == 96 && input_ch_dirs == 27))) {
        set_error("nerf_mlp_forward_v2: implemented for NerfMLP(netdepth=8, netwidth=256, skips=[4], use_viewdirs) with (63,27) or (96,27) input channels");
        return XRB_E_UNSUPPORTED;
    }
    if (n_rows == 0) return XRB_OK;
    XRB_REQUIRE(weight_image && bias && enc_image && raw, "nerf_mlp_forward_v2: null pointer");
    XRB_REQUIRE(((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)raw & 15) == 0 && ((uintptr_t)enc_image & 15) == 0, "nerf_mlp_forward_v2: images / raw must be 16-byte aligned");
    N2Plan p{};
    p.input_ch = input_ch; p.input_ch_dirs = input_ch_dirs; p.aux_blocks = (input_ch + 63) / 64;
    const int aux = p.aux_blocks, dir_blk = 4 + aux;
    int nl = 0, boff = 0;
    auto add = [&](int n_first, int n_second, const int *src, int N, int n_halves, int relu, int writes_a, int alpha_dot) {
        N2Layer &L = p.layer[nl++]; L.n_first = n_first; L.n_second = n_second; for (int k = 0; k < n_first + n_second; ++k) L.src[k] = src[k];
        L.N = N; L.n_halves = n_halves; L.relu = relu; L.writes_a = writes_a; L.alpha_dot = alpha_dot; L.bias_off = boff; boff += N;
    };
    { int s[7] = {4, 5, 0, 0, 0, 0, 0}; add(aux, 0, s, 256, 2, 1, 1, 0); }                               // pts_linears.0 (point-encoding blocks only)
    const int hh[7] = {0, 1, 2, 3, 0, 0, 0};
    for (int l = 1; l <= 4; ++l) add(2, 2, hh, 256, 2, 1, 1, 0);                                         // pts_linears.1-4
    { int s[7]; int n = 0; for (int a = 0; a < aux; ++a) s[n++] = 4 + a; s[n++] = 0; s[n++] = 1; s[n++] = 2; s[n++] = 3; add(aux + 2, 2, s, 256, 2, 1, 1, 0); }   // pts_linears.5 on cat([pts, h])
    add(2, 2, hh, 256, 2, 1, 1, 0);                                                                       // pts_linears.6
    add(2, 2, hh, 256, 2, 1, 1, 1);                                                                       // pts_linears.7 (+ alpha dot in its epilogue)
    add(2, 2, hh, 256, 2, 0, 1, 0);                                                                       // feature_linear (no ReLU)
    { int s[7] = {dir_blk, 0, 1, 2, 3, 0, 0}; add(3, 2, s, 128, 2, 1, 1, 0); }                            // views_linears.0 on cat([feature, dirs]) -> blocks 0 (WG0), 1 (WG1)
    { int s[7] = {0, 1, 0, 0, 0, 0, 0}; add(1, 1, s, 16, 1, 0, 0, 0); }                                   // rgb_linear: K-block 0 (lo), 1 (hi)
    p.n_layers = nl;
    p.bias_total = boff + 257;                                                                            // + Wa[256] + ba
    p.dbg = getenv("XRB_NM_DBG") ? atoi(getenv("XRB_NM_DBG")) : 0;
    size_t smem = 1024 + (size_t)(5 + aux) * 16384 + (size_t)N2_SLOTS * N2_SLAB_BYTES + sizeof(float) * (((p.bias_total + 3) & ~3) + 128) + sizeof(__half) * ((p.bias_total + 7) & ~7) + 8 * (2 * N2_SLOTS + 8) + 16;
    cudaFuncSetAttribute(nerf_mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t n_tiles = (n_rows + 127) / 128;
    int grid = (int)(n_tiles < sms ? n_tiles : sms);
    nerf_mlp_tc2_kernel<<<grid, N2_THREADS, smem, (cudaStream_t)stream>>>(p, (const uint8_t *)weight_image, bias, (const uint8_t *)enc_image, n_rows, raw);
    return check_launch("nerf_mlp_forward_v2");
}

size_t xrb_nerf_enc_image_bytes(int64_t n_rows, int input_ch) { return (size_t)((n_rows + 127) / 128) * ((input_ch + 63) / 64 + 1) * 16384; }

int xrb_nerf_pack_embedded(const float *embedded, int64_t n_rows, int input_ch, int input_ch_dirs, void *enc_image, void *stream) {
    XRB_REQUIRE(n_rows >= 0 && input_ch >= 1 && input_ch <= 128 && input_ch_dirs >= 0 && input_ch_dirs <= 64, "nerf_pack_embedded: bad size");
    if (n_rows == 0) return XRB_OK;
    XRB_REQUIRE(embedded && enc_image && ((uintptr_t)enc_image & 15) == 0, "nerf_pack_embedded: null/misaligned pointer");
    const int aux = (input_ch + 63) / 64;
    int64_t total = ((n_rows + 127) / 128) * 128 * (aux + 1) * 8, blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    pack_embedded_tiles_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(embedded, n_rows, input_ch, input_ch_dirs, aux, (uint8_t *)enc_image);
    return check_launch("nerf_pack_embedded");
}

int xrb_nerf_posenc_tiles(const float *pts, const float *viewdirs, int64_t n_pts, int samples_per_ray, int multires, int multires_dirs, void *enc_image, void *stream) {
    XRB_REQUIRE(n_pts >= 0 && samples_per_ray >= 1 && multires >= 0 && multires <= 20 && multires_dirs >= 0 && multires_dirs <= 10, "nerf_posenc_tiles: bad size");
    if (n_pts == 0) return XRB_OK;
    XRB_REQUIRE(pts && viewdirs && enc_image && ((uintptr_t)enc_image & 15) == 0, "nerf_posenc_tiles: null/misaligned pointer");
    if (multires == 10 && multires_dirs == 4 && !getenv("XRB_GENERIC_ENCODERS")) return xrb_internal_posenc_tiles_fast(pts, viewdirs, n_pts, samples_per_ray, enc_image, nullptr, nullptr, stream);
    const int aux = (3 + 6 * multires + 63) / 64;
    int64_t total = ((n_pts + 127) / 128) * 128 * (aux + 1) * 8, blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    posenc_tiles_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pts, viewdirs, n_pts, samples_per_ray, multires, multires_dirs, (uint8_t *)enc_image, nullptr, nullptr);
    return check_launch("nerf_posenc_tiles");
}

int xrb_nerf_posenc_tiles_rays(const float *rays_o, const float *rays_d, const float *z_vals, const float *viewdirs, int64_t n_rays, int samples_per_ray, int multires, int multires_dirs,
                               void *enc_image, void *stream) {
    XRB_REQUIRE(n_rays >= 0 && samples_per_ray >= 1 && multires >= 0 && multires <= 20 && multires_dirs >= 0 && multires_dirs <= 10, "nerf_posenc_tiles_rays: bad size");
    if (n_rays == 0) return XRB_OK;
    XRB_REQUIRE(rays_o && rays_d && z_vals && viewdirs && enc_image && ((uintptr_t)enc_image & 15) == 0, "nerf_posenc_tiles_rays: null/misaligned pointer");
    if (multires == 10 && multires_dirs == 4 && !getenv("XRB_GENERIC_ENCODERS")) return xrb_internal_posenc_tiles_fast(z_vals, viewdirs, n_rays * samples_per_ray, samples_per_ray, enc_image, rays_o, rays_d, stream);
    const int aux = (3 + 6 * multires + 63) / 64;
    const int64_t n_pts = n_rays * samples_per_ray;
    int64_t total = ((n_pts + 127) / 128) * 128 * (aux + 1) * 8, blocks = (total + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    posenc_tiles_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(z_vals, viewdirs, n_pts, samples_per_ray, multires, multires_dirs, (uint8_t *)enc_image, rays_o, rays_d);
    return check_launch("nerf_posenc_tiles_rays");
}

}  // extern "C"
