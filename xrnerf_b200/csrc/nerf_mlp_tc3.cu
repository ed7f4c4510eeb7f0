// xrnerf_b200 — NerfMLP.run_mlp (/root/reference/xrnerf/models/mlps/nerf_mlp.py:70-94) on tcgen05, v3: TWO 128-row tiles in flight per SM.
//
// v2 (nerf_mlp_tc.cu) keeps one tile per CTA: every layer is MMA -> TMEM drain -> next MMA on the same rows, so the tensor pipe idles while
// the epilogue converts the accumulators (ncu: tensor pipe 22 % active, profiles/r01b_nerf_mlp_tc2_ncu.md). v3 runs two independent tile
// pipelines p = 0,1 on one SM; while pipeline 0 drains layer l, the tensor core works on pipeline 1's layer:
//
//   warps 0-15   four compute warpgroups WG(p, c): tile p, output-column half c. thread == row == TMEM lane. Per layer: tcgen05.ld 32 columns
//                at a time, + bias (packed half2), ReLU, fp16, STS.128 into the tile's swizzled H blocks (the next layer's A operand);
//   warps 16,17  producer of pipeline 0 / 1 (one lane): streams the layer weights as [<=128 x 64] fp16 half-slabs through a 2-slot ring and
//                TMA-loads the tile's encoding blocks;
//   warps 18,19  MMA issuer of pipeline 0 / 1 (one lane): per K-block and output half 4 x tcgen05.mma (M=128, N=128|16, K=16) into the
//                pipeline's 256 TMEM columns; everything a layer reads is complete before its epilogue rewrites H (no in-place race).
// Synchronisation: compute -> issuer is a hardware named barrier (bar.arrive x256 / bar.sync x32); issuer -> compute is one mbarrier
// (tcgen05.commit) polled by ONE thread per pipeline, fanned out by a named barrier; 6 polling threads per SM in total.
//
// Shared memory (227 KB): per pipeline H0-H3 (the 256-wide hidden state, 64 KB) + ONE 16 KB AUX block that holds, in turn, the point
// encoding (layers 0 and 5), the view-direction encoding (views_linears.0) and then the next tile's point encoding; Mip-NeRF's second point
// block (columns 64-95) is loaded into H3 for layer 0 (H is dead there) and re-loaded into AUX behind the first one for layer 5.
// Biases are read from global memory (L1-resident, an fp16 copy follows the fp32 vector) — no room for them in shared memory.
// Numeric contract identical to v2: fp16 operands, fp32 accumulate, bias added in fp16 after rounding the accumulator (tests/test_gpu_nerf_mlp.py).
#include "tc.cuh"
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace xrb {

constexpr int N3_THREADS = 640;
constexpr int N3_MAX_LAYERS = 11;
constexpr int N3_MAX_KB = 6;
constexpr uint32_t N3_BLOCK = 16384;                 // one [128 x 64] fp16 operand block / ring slot
constexpr uint32_t N3_PIPE_A = 5 * N3_BLOCK;         // H0..H3 + AUX
constexpr int N3_RING = 2;
constexpr int N3_AUX = 4;                            // A block index of AUX

struct N3Layer {
    int n_kb, N, n_halves, relu, alpha_dot, bias_off, commit_h3;
    int8_t src[N3_MAX_KB];      // A block per K-block: 0-3 = H0..H3, 4 = AUX
    int8_t wait_enc[N3_MAX_KB]; // issuer waits before K-block kb: 0 none, 1 = tile's first point block in AUX (+ second in H3, Mip), 2 = direction block, 3 = second point block in AUX
    int8_t reload[N3_MAX_KB];   // AUX is dead after K-block kb; the producer refills it with: 0 nothing, 1 = direction block, 2 = second point block, 3 = next tile's first point block
};
struct N3Plan { int n_layers, aux_blocks, bias_total, bias_h_off, dbg; N3Layer layer[N3_MAX_LAYERS]; };   // dbg (XRB_NM_DBG, attribution experiments): bit0 skip the weight TMA copies, bit1 skip the MMAs, bit2 skip the epilogue math

__device__ __forceinline__ void n3_bar_arrive(uint32_t id, uint32_t n_threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }
__device__ __forceinline__ void n3_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }

// barrier indices inside a pipeline's block of 16
enum { B_FULL = 0, B_EMPTY = 2, B_AREADY = 4, B_ACC = 5, B_E0 = 6, B_E1 = 7, B_E2 = 8, B_E3 = 9, B_AUXFREE = 10, B_H3FREE = 11, B_PER_PIPE = 16 };

__global__ void __launch_bounds__(N3_THREADS, 1) nerf_mlp_tc3_kernel(N3Plan plan, const uint8_t *__restrict__ weight_image, const float *__restrict__ bias_g, const uint8_t *__restrict__ enc_image,
                                                                      int64_t n_rows, float *__restrict__ raw) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    uint8_t *ring_base = base + 2 * N3_PIPE_A;
    float *alpha_part = (float *)(ring_base + 2 * N3_RING * N3_BLOCK);          // [2][128]
    uint64_t *bars = (uint64_t *)(alpha_part + 256);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * B_PER_PIPE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int p = 0; p < 2; ++p) {
            uint64_t *b = bars + p * B_PER_PIPE;
            for (int s = 0; s < N3_RING; ++s) { tc::mbar_init(b + B_FULL + s, 1); tc::mbar_init(b + B_EMPTY + s, 1); }
            tc::mbar_init(b + B_ACC, 1);
            tc::mbar_init(b + B_E0, 1); tc::mbar_init(b + B_E1, 1); tc::mbar_init(b + B_E2, 1); tc::mbar_init(b + B_E3, 1);
            tc::mbar_init(b + B_AUXFREE, 1); tc::mbar_init(b + B_H3FREE, 1);
        }
        tc::fence_mbar_init();
    }
    if (warp == 16) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int64_t n_tiles = (n_rows + 127) / 128;
    const uint32_t enc_tile_bytes = (uint32_t)(plan.aux_blocks + 1) * N3_BLOCK;
    const bool mip = plan.aux_blocks == 2;

    if (warp >= 16) {
        const int p = warp & 1;                                  // warps 16,18 -> pipeline 0; 17,19 -> pipeline 1
        uint64_t *b = bars + p * B_PER_PIPE;
        uint8_t *A = base + (size_t)p * N3_PIPE_A;
        uint8_t *ring = ring_base + (size_t)p * N3_RING * N3_BLOCK;
        const int64_t vcta = (int64_t)blockIdx.x * 2 + p, vstride = (int64_t)gridDim.x * 2;
        if (warp < 18) {
            // ===================================================== producer of pipeline p
            if (lane == 0) {
                uint32_t it = 0, af = 0, n = 0;
                for (int64_t tile = vcta; tile < n_tiles; tile += vstride, ++n) {
                    const uint8_t *enc = enc_image + (size_t)tile * enc_tile_bytes;
                    if (n == 0) { tc::mbar_expect_tx(b + B_E0, N3_BLOCK); tc::tma_bulk_g2s(A + N3_AUX * N3_BLOCK, enc, N3_BLOCK, b + B_E0); }   // later tiles: loaded behind the previous tile's direction block
                    if (mip) {
                        if (n > 0) tc::mbar_wait(b + B_H3FREE, (n - 1) & 1);          // views_linears.0 of the previous tile has read the feature block in H3
                        tc::mbar_expect_tx(b + B_E3, N3_BLOCK);
                        tc::tma_bulk_g2s(A + 3 * N3_BLOCK, enc + N3_BLOCK, N3_BLOCK, b + B_E3);
                    }
                    size_t off = 0;
                    int pending = 0;
                    for (int l = 0; l < plan.n_layers; ++l) {
                        const N3Layer &L = plan.layer[l];
                        const uint32_t bytes = (uint32_t)(L.N / L.n_halves) * 128u;
                        for (int kb = 0; kb < L.n_kb; ++kb) {
                            for (int h = 0; h < L.n_halves; ++h, ++it) {
                                const uint32_t slot = it % N3_RING, round = it / N3_RING;
                                if (round > 0) tc::mbar_wait(b + B_EMPTY + slot, (round - 1) & 1);
                                if (plan.dbg & 1) n3_arrive(b + B_FULL + slot);
                                else { tc::mbar_expect_tx(b + B_FULL + slot, bytes); tc::tma_bulk_g2s(ring + (size_t)slot * N3_BLOCK, weight_image + off, bytes, b + B_FULL + slot); }
                                off += bytes;
                            }
                            if (pending) {
                                // the K-block that last read AUX was issued one K-block ago: its ring slots have been released since, so its commit
                                // on AUXFREE (issued right behind the slot release) has landed or is about to — this wait does not stall the weight stream
                                const bool has_next = tile + vstride < n_tiles;
                                if (pending != 3 || has_next) {
                                    tc::mbar_wait(b + B_AUXFREE, af & 1);
                                    const uint8_t *src = pending == 1 ? enc + (size_t)plan.aux_blocks * N3_BLOCK : pending == 2 ? enc + N3_BLOCK : enc_image + (size_t)(tile + vstride) * enc_tile_bytes;
                                    uint64_t *eb = b + (pending == 1 ? B_E1 : pending == 2 ? B_E2 : B_E0);
                                    tc::mbar_expect_tx(eb, N3_BLOCK);
                                    tc::tma_bulk_g2s(A + N3_AUX * N3_BLOCK, src, N3_BLOCK, eb);
                                }
                                ++af;
                                pending = 0;
                            }
                            pending = L.reload[kb];
                        }
                    }
                }
            }
        } else {
            // ===================================================== MMA issuer of pipeline p (the whole warp walks the loop: the layer hand-off from
            // the 256 compute threads is a HARDWARE named barrier — bar.arrive x256 + bar.sync x32 — not a polled mbarrier; lane 0 issues)
            uint32_t it = 0, tcount = 0;
            const uint32_t tmem_p = tmem + (uint32_t)p * 256u;
            for (int64_t tile = vcta; tile < n_tiles; tile += vstride, ++tcount) {
                for (int l = 0; l < plan.n_layers; ++l) {
                    const N3Layer &L = plan.layer[l];
                    __syncwarp();
                    tc::named_bar_sync(5 + p, 288);                             // the layer's input rows are in H, the previous accumulator is drained
                    if (lane == 0) {
                        tc::tc_fence_after_sync();
                        const uint32_t hw = (uint32_t)(L.N / L.n_halves);
                        const uint32_t idesc = tc::idesc_f16_m128(hw);
                        for (int kb = 0; kb < L.n_kb; ++kb) {
                            const int we = L.wait_enc[kb];
                            if (we == 1) { tc::mbar_wait(b + B_E0, tcount & 1); if (mip) tc::mbar_wait(b + B_E3, tcount & 1); tc::tc_fence_after_sync(); }
                            else if (we == 2) { tc::mbar_wait(b + B_E1, tcount & 1); tc::tc_fence_after_sync(); }
                            else if (we == 3) { tc::mbar_wait(b + B_E2, tcount & 1); tc::tc_fence_after_sync(); }
                            const uint32_t a0 = tc::smem_u32(A + (size_t)L.src[kb] * N3_BLOCK);
                            for (int h = 0; h < L.n_halves; ++h, ++it) {
                                const uint32_t slot = it % N3_RING, round = it / N3_RING;
                                tc::mbar_wait(b + B_FULL + slot, round & 1);
                                tc::tc_fence_after_sync();
                                const uint32_t b0 = tc::smem_u32(ring + (size_t)slot * N3_BLOCK);
                                if (!(plan.dbg & 2))
#pragma unroll
                                for (int k = 0; k < 4; ++k) tc::mma_f16_ss(tmem_p + h * hw, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), idesc, (kb | k) ? 1u : 0u);
                                tc::mma_commit(b + B_EMPTY + slot);
                            }
                            if (L.reload[kb]) tc::mma_commit(b + B_AUXFREE);       // every MMA that reads AUX's current content has been issued
                        }
                        if (L.commit_h3) tc::mma_commit(b + B_H3FREE);
                        tc::mma_commit(b + B_ACC);                                  // accumulator of layer l complete, H/AUX reads of layer l done
                    }
                }
            }
        }
    } else {
        // ===================================================== compute warpgroups WG(p, c): thread == row, c == column half
        const int g = warp >> 2, p = g >> 1, c = g & 1;
        uint64_t *b = bars + p * B_PER_PIPE;
        uint8_t *A = base + (size_t)p * N3_PIPE_A;
        float *apart = alpha_part + p * 128;
        const int64_t vcta = (int64_t)blockIdx.x * 2 + p, vstride = (int64_t)gridDim.x * 2;
        const uint32_t row = threadIdx.x & 127;
        const uint32_t taddr = tmem + (uint32_t)p * 256u + (((uint32_t)(warp & 3) * 32u) << 16);
        const uint32_t r7 = row & 7u, row_off = (row >> 3) * 1024u + r7 * 128u;
        const __half *bias_h = reinterpret_cast<const __half *>(bias_g + plan.bias_h_off);
        const float *wa = bias_g + plan.bias_total - 257;
        uint32_t acc_phase = 0;
        for (int64_t tile = vcta; tile < n_tiles; tile += vstride) {
            const int64_t i = tile * 128 + row;
            const bool valid = i < n_rows;
            tc::tc_fence_before_sync();
            __syncwarp();
            n3_bar_arrive(5 + p, 288);                             // my reads of the previous tile's accumulators are done (hardware barrier towards the issuer warp)
            float alpha_acc = 0.f;
            for (int l = 0; l < plan.n_layers; ++l) {
                const N3Layer &L = plan.layer[l];
                const bool last = l == plan.n_layers - 1;
                // ONE thread per pipeline polls the accumulator mbarrier (tcgen05.commit can only signal an mbarrier); the other 255 block in a hardware
                // named barrier. 16 polling lanes executed 4.5 M try_waits per SM (ncu r01c: 47 % of all stall samples in the poll loop and its
                // __syncwarp) and the polled mbarrier traffic delayed every other barrier operation: the bare synchronisation skeleton cost 2.07 of 3.98 ms.
                if ((warp & 7) == 0 && lane == 0) tc::mbar_wait(b + B_ACC, acc_phase);
                __syncwarp();
                tc::named_bar_sync(3 + p, 256);
                tc::tc_fence_after_sync();
                acc_phase ^= 1;
                if (!last) {
                    const int cols = L.N >> 1, col0 = c * cols;
                    const __half2 zero2 = __float2half2_rn(0.f);
                    for (int ck = 0; ck < ((plan.dbg & 4) ? 0 : cols / 32); ++ck) {
                        const int colb = col0 + ck * 32;
                        uint32_t r[32];
                        tc::tmem_ld32(taddr + colb, r);
                        const uint4 *bh = reinterpret_cast<const uint4 *>(bias_h + L.bias_off + colb);
                        uint8_t *dst = A + (size_t)(colb >> 6) * N3_BLOCK + row_off;
                        const uint32_t cb = (uint32_t)(colb & 63) >> 3;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 bq = __ldg(bh + q);
                            const __half2 *b2 = reinterpret_cast<const __half2 *>(&bq);
                            __half2 h[4];
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                h[e2] = __hadd2(__floats2half2_rn(__uint_as_float(r[8 * q + 2 * e2]), __uint_as_float(r[8 * q + 2 * e2 + 1])), b2[e2]);
                                if (L.relu) h[e2] = __hmax2(h[e2], zero2);
                            }
                            *reinterpret_cast<uint4 *>(dst + (((cb + (uint32_t)q) ^ r7) << 4)) = *reinterpret_cast<uint4 *>(h);
                        }
                    }
                    if (L.alpha_dot) {   // alpha_linear on the fp16 output of pts_linears.7: partial dot over my columns, read back from my own row of H
                        // (a warp-uniform branch taken once per tile; inside the chunk loop it was predicated code issued for every layer: 160 M FFMA + 60 M LDG slots)
                        for (int ck = 0; ck < cols / 32; ++ck) {
                            const int colb = col0 + ck * 32;
                            const uint8_t *src = A + (size_t)(colb >> 6) * N3_BLOCK + row_off;
                            const uint32_t cb = (uint32_t)(colb & 63) >> 3;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 hv = *reinterpret_cast<const uint4 *>(src + (((cb + (uint32_t)q) ^ r7) << 4));
                                const __half2 *h = reinterpret_cast<const __half2 *>(&hv);
                                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wa + colb + 8 * q)), w1 = __ldg(reinterpret_cast<const float4 *>(wa + colb + 8 * q + 4));
                                float2 f;
                                f = __half22float2(h[0]); alpha_acc = fmaf(f.x, w0.x, alpha_acc); alpha_acc = fmaf(f.y, w0.y, alpha_acc);
                                f = __half22float2(h[1]); alpha_acc = fmaf(f.x, w0.z, alpha_acc); alpha_acc = fmaf(f.y, w0.w, alpha_acc);
                                f = __half22float2(h[2]); alpha_acc = fmaf(f.x, w1.x, alpha_acc); alpha_acc = fmaf(f.y, w1.y, alpha_acc);
                                f = __half22float2(h[3]); alpha_acc = fmaf(f.x, w1.z, alpha_acc); alpha_acc = fmaf(f.y, w1.w, alpha_acc);
                            }
                        }
                        if (c == 1) apart[row] = alpha_acc;
                    }
                    tc::fence_proxy_async_smem();
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    n3_bar_arrive(5 + p, 288);
                } else {
                    if (c == 0) {
                        float o16[16];
                        tc::tmem_ld16(taddr, o16);
                        tc::named_bar_sync(1 + p, 256);    // WG(p,1)'s partial alpha is in shared memory
                        const float *bl = bias_g + L.bias_off;
                        const float alpha = alpha_acc + apart[row] + __ldg(bias_g + plan.bias_total - 1);
                        if (valid) reinterpret_cast<float4 *>(raw)[i] = make_float4(o16[0] + __ldg(bl), o16[1] + __ldg(bl + 1), o16[2] + __ldg(bl + 2), alpha);
                    } else {
                        tc::named_bar_sync(1 + p, 256);
                    }
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) tc::tmem_dealloc<512>(tmem);
}

}  // namespace xrb

using namespace xrb;

extern "C" {

// weight image / bias vector from xrnerf_b200.nerf_mlp.pack_nerf_mlp_v3 (same plan, mirrored on the host)
int xrb_nerf_mlp_forward_v3(const void *weight_image, const float *bias, const void *enc_image, int64_t n_rows, int input_ch, int input_ch_dirs, float *raw, void *stream) {
    XRB_REQUIRE(n_rows >= 0, "nerf_mlp_forward_v3: negative size");
    if (!((input_ch == 63 && input_ch_dirs == 27) || (input_ch == 96 && input_ch_dirs == 27))) {
        set_error("nerf_mlp_forward_v3: implemented for NerfMLP(netdepth=8, netwidth=256, skips=[4], use_viewdirs) with (63,27) or (96,27) input channels");
        return XRB_E_UNSUPPORTED;
    }
    if (n_rows == 0) return XRB_OK;
    XRB_REQUIRE(weight_image && bias && enc_image && raw, "nerf_mlp_forward_v3: null pointer");
    XRB_REQUIRE(((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)raw & 15) == 0 && ((uintptr_t)enc_image & 15) == 0 && ((uintptr_t)bias & 15) == 0,
                "nerf_mlp_forward_v3: images / bias / raw must be 16-byte aligned");
    N3Plan p{};
    p.aux_blocks = (input_ch + 63) / 64;
    const bool mip = p.aux_blocks == 2;
    int nl = 0, boff = 0;
    auto add = [&](int n_kb, const int *src, const int *wait_enc, const int *reload, int N, int n_halves, int relu, int alpha_dot, int commit_h3) {
        N3Layer &L = p.layer[nl++];
        L.n_kb = n_kb; L.N = N; L.n_halves = n_halves; L.relu = relu; L.alpha_dot = alpha_dot; L.bias_off = boff; L.commit_h3 = commit_h3; boff += N;
        for (int k = 0; k < n_kb; ++k) { L.src[k] = (int8_t)src[k]; L.wait_enc[k] = (int8_t)(wait_enc ? wait_enc[k] : 0); L.reload[k] = (int8_t)(reload ? reload[k] : 0); }
    };
    const int hh[4] = {0, 1, 2, 3};
    if (!mip) { int s[1] = {N3_AUX}, w[1] = {1}; add(1, s, w, nullptr, 256, 2, 1, 0, 0); }                                     // pts_linears.0
    else { int s[2] = {N3_AUX, 3}, w[2] = {1, 0}; add(2, s, w, nullptr, 256, 2, 1, 0, 0); }
    for (int l = 1; l <= 4; ++l) add(4, hh, nullptr, nullptr, 256, 2, 1, 0, 0);                                                // pts_linears.1-4
    if (!mip) { int s[5] = {N3_AUX, 0, 1, 2, 3}, r[5] = {1, 0, 0, 0, 0}; add(5, s, nullptr, r, 256, 2, 1, 0, 0); }            // pts_linears.5 on cat([pts, h])
    else { int s[6] = {N3_AUX, 0, 1, 2, 3, N3_AUX}, w[6] = {0, 0, 0, 0, 0, 3}, r[6] = {2, 0, 0, 0, 0, 1}; add(6, s, w, r, 256, 2, 1, 0, 0); }
    add(4, hh, nullptr, nullptr, 256, 2, 1, 0, 0);                                                                              // pts_linears.6
    add(4, hh, nullptr, nullptr, 256, 2, 1, 1, 0);                                                                              // pts_linears.7 (+ alpha dot in its epilogue)
    add(4, hh, nullptr, nullptr, 256, 2, 0, 0, 0);                                                                              // feature_linear (no ReLU)
    { int s[5] = {N3_AUX, 0, 1, 2, 3}, w[5] = {2, 0, 0, 0, 0}, r[5] = {3, 0, 0, 0, 0}; add(5, s, w, r, 128, 1, 1, 0, 1); }    // views_linears.0 on cat([feature, dirs])
    { int s[2] = {0, 1}; add(2, s, nullptr, nullptr, 16, 1, 0, 0, 0); }                                                        // rgb_linear
    p.n_layers = nl;
    p.bias_total = boff + 257;                    // + Wa[256] + ba
    p.bias_h_off = (p.bias_total + 7) & ~7;       // fp16 copy of the per-layer biases follows the fp32 vector
    p.dbg = getenv("XRB_NM_DBG") ? atoi(getenv("XRB_NM_DBG")) : 0;
    const size_t smem = 1024 + 2 * (size_t)N3_PIPE_A + 2 * (size_t)N3_RING * N3_BLOCK + 256 * sizeof(float) + 8 * (2 * B_PER_PIPE) + 16;
    static_assert(1024 + 2 * (size_t)N3_PIPE_A + 2 * (size_t)N3_RING * N3_BLOCK + 256 * sizeof(float) + 8 * (2 * B_PER_PIPE) + 16 <= 232448, "v3 shared memory budget");
    cudaFuncSetAttribute(nerf_mlp_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (n_rows + 127) / 128, pairs = (n_tiles + 1) / 2;
    const int grid = (int)(pairs < sms ? pairs : sms);
    nerf_mlp_tc3_kernel<<<grid, N3_THREADS, smem, (cudaStream_t)stream>>>(p, (const uint8_t *)weight_image, bias, (const uint8_t *)enc_image, n_rows, raw);
    return check_launch("nerf_mlp_forward_v3");
}

}  // extern "C"
