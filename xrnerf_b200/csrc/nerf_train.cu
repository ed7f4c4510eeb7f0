// xrnerf_b200 — NerfMLP training (forward with saved activations + backward) on tcgen05 tensor cores.
//
// The reference trains NerfMLP (/root/reference/xrnerf/models/mlps/nerf_mlp.py:70-94: 8 x 256 ReLU trunk with the skip at layer 4, alpha / feature heads, a 128-wide view
// branch) with torch.autograd over 11 nn.Linear modules: cuBLAS fp32 GEMMs forward, two more per layer backward, every [rows, 256] activation through HBM in fp32
// (networks/nerf.py:71-92, networks/mipnerf.py:45-74). Here every dense contraction of the training step is a hand-written UMMA kernel over TILE IMAGES:
//
//   tile image   an activation matrix [rows x C] stored per 128-row tile as C/64 blocks of [128 x 64] fp16 in the UMMA shared-memory layout (K-major, 128-byte swizzle,
//                16 KB per block) — the layout the encoders already write (xrb_nerf_posenc_tiles*, xrb_mip_ipe_tiles_rays). A block moves between HBM and shared memory
//                with ONE 1-D bulk TMA copy and is an MMA operand as it lies; read K-major it contracts over its 64 columns, read MN-major over its 128 rows.
//   tg_kernel    Y = epilogue( sum_p A_p . op(W_p) ) per 128-row tile, accumulator in TMEM (double-buffered: tile t+1's MMAs run while tile t's epilogue drains):
//                  forward   A = layer input blocks, W slabs [N x 64] K-major, epilogue +bias, ReLU, fp16 -> output blocks (saved: they are the backward's operands)
//                  input grad dX = dZ . W : A = dZ blocks, the SAME weight slabs read MN-major (no transposed copy), epilogue * [activation > 0] -> dZ of the layer below
//   tg_dw_kernel dW += dZ^T . X over all rows: both operands read MN-major (contraction over the 128 rows of a tile), M = 128 output features per CTA (two adjacent
//                blocks through the leading-dimension offset), accumulators stay in TMEM across the CTA's whole tile range; the bias gradient is one more MMA against a
//                block of ones. One flush per CTA (fp32 atomics).
// Warp roles (320 threads): warps 0-7 two epilogue warpgroups (thread == row == TMEM lane), warp 8 TMA producer (one lane), warp 9 MMA issuer (one lane).
// dZ is staged in fp16 with a fixed scale (2^14, applied where dL/draw enters and removed at the dW flush): tcnn-style loss scaling for 1e-6-sized NeRF gradients.
#include "tc.cuh"
#include "common.cuh"
#include "ngp_field.cuh"   // sw128_offset
#include <cuda_fp16.h>
#include <stdio.h>

namespace xrb {

constexpr uint32_t TG_BLOCK = 16384;
constexpr int TG_MAX_A = 6, TG_MAX_SLABS = 10, TG_THREADS = 320;   // 6 A blocks: Mip-NeRF's skip layer reads 2 IPE blocks + 4 hidden blocks
constexpr uint32_t TG_RING_SLOT = 32768;   // one [256 x 64] fp16 slab
constexpr float TG_SCALE = 16384.f;

struct TgOp { const uint8_t *base; uint32_t tile_stride, blk_off; int n_blk; };   // blocks [blk_off/16K, +n_blk) of every tile of an image
struct TgSlab { uint32_t w_off, bytes; int n_sub, a_blk0, d_col, first, n_k; };    // one weight slab: loaded once per tile, used by n_sub MMA groups (A block a_blk0 + j) of n_k K-steps of 16
struct TgP {
    TgOp a[2]; int n_a;                    // A operand sources, concatenated block-wise into the shared-memory A buffer
    const uint8_t *w;                      // weight image base
    TgSlab slab[TG_MAX_SLABS]; int n_slabs;
    int b_mn;                              // 0: slabs are K-major B operands (forward), 1: MN-major (input gradient)
    int mma_n;                             // N of each MMA (forward: layer width 16/128/256; input gradient: 64)
    int out_cols;                          // accumulator columns the epilogue reads (multiple of 16, <= 256)
    int epi;                               // 0: +bias, relu?, fp16 -> y image; 1: +bias -> fp32 rows (yf); 2: * [x > 0] -> fp16 y image; 3: plain fp16 -> y image
    int relu; const float *bias;
    TgOp x;                                // epi 2: the activation whose sign masks the gradient (out_cols/64 blocks per tile)
    uint8_t *y; uint32_t y_tile_stride, y_blk_off;
    float *yf; int yf_stride, yf_col, yf_n; float yf_scale;
    long long n_rows;
};

__device__ __forceinline__ void tg_bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tg_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__host__ __device__ constexpr uint32_t tg_idesc(uint32_t m, uint32_t n, uint32_t a_mn, uint32_t b_mn) { return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24); }
// MN-major SWIZZLE_128B operand: 64 MN elements per 128-byte row, 8-row K groups 1024 B apart (SBO), next 64-element MN atom `lbo` bytes away (LBO)
__device__ __forceinline__ uint64_t tg_desc_mn(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

enum { TB_AFULL = 0, TB_AEMPTY = 1, TB_SFULL = 2, TB_SEMPTY = 4, TB_ACCFULL = 6, TB_ACCEMPTY = 8, TB_XFULL = 10, TB_YEMPTY = 11, TB_N = 12 };

__global__ void __launch_bounds__(TG_THREADS, 1) tg_kernel(const __grid_constant__ TgP P) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    uint8_t *A = base, *RING = A + TG_MAX_A * TG_BLOCK, *Y = RING + 2 * TG_RING_SLOT;
    uint64_t *bars = (uint64_t *)(Y + 4 * TG_BLOCK);
    uint32_t *tmem_slot = (uint32_t *)(bars + TB_N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int k = 0; k < TB_N; ++k) tc::mbar_init(bars + k, 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const long long n_tiles = (P.n_rows + 127) / 128;
    int n_ablk = 0;
    for (int s = 0; s < P.n_a; ++s) n_ablk += P.a[s].n_blk;
    const int n_xblk = P.epi == 2 ? P.out_cols / 64 : 0;

    if (warp == 8) {
        // ===================================================== producer (one lane): A blocks of the tile, then its weight slabs through the 2-slot ring, then the mask blocks
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                if (it > 0) tc::mbar_wait(bars + TB_AEMPTY, (it - 1) & 1);
                tc::mbar_expect_tx(bars + TB_AFULL, (uint32_t)n_ablk * TG_BLOCK);
                int ab = 0;
                for (int s = 0; s < P.n_a; ++s)
                    for (int k = 0; k < P.a[s].n_blk; ++k, ++ab)
                        tc::tma_bulk_g2s(A + (size_t)ab * TG_BLOCK, P.a[s].base + (size_t)tile * P.a[s].tile_stride + P.a[s].blk_off + (size_t)k * TG_BLOCK, TG_BLOCK, bars + TB_AFULL);
                for (int s = 0; s < P.n_slabs; ++s, ++g) {
                    const uint32_t slot = g & 1u, round = g >> 1;
                    if (round > 0) tc::mbar_wait(bars + TB_SEMPTY + slot, (round - 1) & 1);
                    tc::mbar_expect_tx(bars + TB_SFULL + slot, P.slab[s].bytes);
                    tc::tma_bulk_g2s(RING + (size_t)slot * TG_RING_SLOT, P.w + P.slab[s].w_off, P.slab[s].bytes, bars + TB_SFULL + slot);
                }
                if (n_xblk) {
                    if (it > 0) tc::mbar_wait(bars + TB_YEMPTY, (it - 1) & 1);     // the previous tile's output has left the Y region
                    tc::mbar_expect_tx(bars + TB_XFULL, (uint32_t)n_xblk * TG_BLOCK);
                    for (int k = 0; k < n_xblk; ++k)
                        tc::tma_bulk_g2s(Y + (size_t)k * TG_BLOCK, P.x.base + (size_t)tile * P.x.tile_stride + P.x.blk_off + (size_t)k * TG_BLOCK, TG_BLOCK, bars + TB_XFULL);
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== MMA issuer (one lane)
        if (lane == 0) {
            uint32_t it = 0, g = 0;
            const uint32_t idesc = tg_idesc(128, (uint32_t)P.mma_n, 0, (uint32_t)P.b_mn);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1u;
                if (it >= 2) tc::mbar_wait(bars + TB_ACCEMPTY + buf, ((it >> 1) - 1) & 1);
                tc::mbar_wait(bars + TB_AFULL, it & 1);
                tc::tc_fence_after_sync();
                const uint32_t d0 = tmem + buf * 256u;
                for (int s = 0; s < P.n_slabs; ++s, ++g) {
                    const uint32_t slot = g & 1u, round = g >> 1;
                    tc::mbar_wait(bars + TB_SFULL + slot, round & 1);
                    tc::tc_fence_after_sync();
                    const TgSlab &S = P.slab[s];
                    const uint32_t b_base = tc::smem_u32(RING + (size_t)slot * TG_RING_SLOT);
                    for (int j = 0; j < S.n_sub; ++j) {
                        const uint32_t a0 = tc::smem_u32(A + (size_t)(S.a_blk0 + j) * TG_BLOCK);
                        for (int k = 0; k < S.n_k; ++k) {
                            const uint64_t ad = tc::smem_desc_sw128(a0 + k * 32);
                            const uint64_t bd = P.b_mn ? tg_desc_mn(b_base + j * 8192 + k * 2048, 1024) : tc::smem_desc_sw128(b_base + k * 32);
                            tc::mma_f16_ss(d0 + (uint32_t)S.d_col, ad, bd, idesc, (S.first && j == 0 && k == 0) ? 0u : 1u);
                        }
                    }
                    tc::mma_commit(bars + TB_SEMPTY + slot);
                }
                tc::mma_commit(bars + TB_AEMPTY);
                tc::mma_commit(bars + TB_ACCFULL + buf);
            }
        }
    } else {
        // ===================================================== epilogue warpgroups: thread == row == TMEM lane
        const int wg = warp >> 2, row = threadIdx.x & 127;
        const int cols_wg = P.out_cols >= 128 ? P.out_cols / 2 : (wg == 0 ? P.out_cols : 0), col0 = P.out_cols >= 128 ? wg * cols_wg : 0;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u;
            const long long i = tile * 128 + row;
            tc::mbar_wait(bars + TB_ACCFULL + buf, (it >> 1) & 1);
            tc::tc_fence_after_sync();
            if (P.epi == 2) tc::mbar_wait(bars + TB_XFULL, it & 1);
            else tc::named_bar_sync(1, 256);                        // the previous tile's bulk store has finished reading Y (thread 0 waited before arriving here)
            const uint32_t taddr = tmem + buf * 256u + (((uint32_t)(row >> 5) * 32u) << 16);
            if (P.epi == 1) {
                if (wg == 0) {
                    float v[16];
                    tc::tmem_ld16(taddr, v);
                    if (i < P.n_rows)
                        for (int k = 0; k < P.yf_n; ++k) P.yf[(size_t)i * P.yf_stride + P.yf_col + k] = (v[k] + (P.bias ? __ldg(P.bias + k) : 0.f)) * P.yf_scale;
                }
            } else {
                for (int c = 0; c < cols_wg; c += 16) {
                    const int col = col0 + c;
                    float v[16];
                    tc::tmem_ld16(taddr + col, v);
                    uint8_t *yrow = Y + (size_t)(col >> 6) * TG_BLOCK;
                    const uint32_t ch = (uint32_t)(col & 63) >> 3;
                    uint4 o[2];
                    if (P.epi == 0) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) { v[k] += P.bias ? __ldg(P.bias + col + k) : 0.f; if (P.relu) v[k] = fmaxf(v[k], 0.f); }
                    }
                    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]); ow[k] = *reinterpret_cast<uint32_t *>(&h); }
                    if (P.epi == 2) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint4 xa = *reinterpret_cast<const uint4 *>(yrow + sw128_offset(row, ch + q));
                            const uint32_t *xw = reinterpret_cast<const uint32_t *>(&xa);
                            uint32_t *oq = reinterpret_cast<uint32_t *>(&o[q]);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __half2 m = __hgt2(*reinterpret_cast<const __half2 *>(&xw[k]), __float2half2_rn(0.f));
                                const __half2 r = __hmul2(*reinterpret_cast<const __half2 *>(&oq[k]), m);
                                oq[k] = *reinterpret_cast<const uint32_t *>(&r);
                            }
                        }
                    }
                    *reinterpret_cast<uint4 *>(yrow + sw128_offset(row, ch)) = o[0];
                    *reinterpret_cast<uint4 *>(yrow + sw128_offset(row, ch + 1)) = o[1];
                }
            }
            // accumulator drained; output rows are in Y
            tc::tc_fence_before_sync();
            tc::fence_proxy_async_smem();
            tc::named_bar_sync(2, 256);
            if (threadIdx.x == 0) {
                tc::mbar_arrive(bars + TB_ACCEMPTY + buf);
                if (P.epi != 1) {
                    const int n_out_blk = (P.out_cols + 63) / 64;
                    for (int k = 0; k < n_out_blk; ++k) tg_bulk_s2g(P.y + (size_t)tile * P.y_tile_stride + P.y_blk_off + (size_t)k * TG_BLOCK, Y + (size_t)k * TG_BLOCK, TG_BLOCK);
                    tg_bulk_commit_wait_read();
                    if (P.epi == 2) tc::mbar_arrive(bars + TB_YEMPTY);
                }
            }
        }
        if (threadIdx.x == 0) tg_bulk_wait_all();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------- weight gradient
struct TgDwP {
    TgOp z;                 // dZ image of the layer: N/64 blocks per tile (one block for the 16-output layers)
    TgOp x;                 // one source of the layer input: 1..4 blocks per tile
    int n_chunks;           // normal: N / 128 chunks of output rows; transposed: chunks of 128 input columns
    int transposed;         // 0: D[m = out (128 per chunk), n = in]; 1 (N <= 16 layers): D[m = in (128 per chunk), n = out (16)]
    int N, K;               // true layer shape (dW[N][K], fp32 parameter)
    int k_off, k_cols;      // this source's columns [k_off, k_off + k_cols) of the layer input
    float *dW; float *db;   // fp32 gradients (accumulated); db == NULL: no bias gradient from this launch
    long long n_rows;
};

__global__ void __launch_bounds__(192, 1) tg_dw_kernel(const __grid_constant__ TgDwP P) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    // two stages of (dZ chunk: 2 blocks | X: 4 blocks), then one block of ones
    constexpr uint32_t STAGE = 6 * TG_BLOCK;
    uint8_t *ONES = base + 2 * STAGE;
    uint64_t *bars = (uint64_t *)(ONES + TG_BLOCK);      // full[2], empty[2], done
    uint32_t *tmem_slot = (uint32_t *)(bars + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int k = 0; k < 5; ++k) tc::mbar_init(bars + k, 1); tc::fence_mbar_init(); }
    for (uint32_t k = threadIdx.x; k < TG_BLOCK / 4; k += blockDim.x) reinterpret_cast<uint32_t *>(ONES)[k] = 0x3C003C00u;   // fp16 1.0 everywhere
    if (warp == 4) tc::tmem_alloc<512>(tmem_slot);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const long long n_tiles = (P.n_rows + 127) / 128;
    const int chunk = blockIdx.x % P.n_chunks, cta = blockIdx.x / P.n_chunks, n_cta = gridDim.x / P.n_chunks;
    const int n_xblk = P.x.n_blk;
    const int z_blk = P.transposed ? 1 : 2;                                   // blocks of dZ per stage
    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = cta; tile < n_tiles; tile += n_cta, ++it) {
                const uint32_t st = it & 1u;
                if (it >= 2) tc::mbar_wait(bars + 2 + st, ((it >> 1) - 1) & 1);
                uint8_t *S = base + (size_t)st * STAGE;
                tc::mbar_expect_tx(bars + st, (uint32_t)(z_blk + n_xblk) * TG_BLOCK);
                const uint32_t z_first = P.transposed ? 0u : (uint32_t)chunk * 2u;
                for (int k = 0; k < z_blk; ++k)
                    tc::tma_bulk_g2s(S + (size_t)k * TG_BLOCK, P.z.base + (size_t)tile * P.z.tile_stride + P.z.blk_off + (size_t)(z_first + k) * TG_BLOCK, TG_BLOCK, bars + st);
                for (int k = 0; k < n_xblk; ++k)
                    tc::tma_bulk_g2s(S + (size_t)(2 + k) * TG_BLOCK, P.x.base + (size_t)tile * P.x.tile_stride + P.x.blk_off + (size_t)k * TG_BLOCK, TG_BLOCK, bars + st);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            uint32_t it = 0;
            for (long long tile = cta; tile < n_tiles; tile += n_cta, ++it) {
                const uint32_t st = it & 1u;
                tc::mbar_wait(bars + st, (it >> 1) & 1);
                tc::tc_fence_after_sync();
                const uint32_t S = tc::smem_u32(base + (size_t)st * STAGE), ones = tc::smem_u32(ONES);
                if (!P.transposed) {
                    // D[128 out x (64 * n_xblk) in] += dZchunk^T . X ; bias columns at column 64 * n_xblk
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t acc = (it | (uint32_t)k) ? 1u : 0u;
                        const uint64_t ad = tg_desc_mn(S + k * 2048, TG_BLOCK);                     // M = 128: two dZ blocks 16 KB apart
                        for (int xb = 0; xb < n_xblk; ++xb)
                            tc::mma_f16_ss(tmem + (uint32_t)xb * 64u, ad, tg_desc_mn(S + (2 + xb) * TG_BLOCK + k * 2048, 1024), tg_idesc(128, 64, 1, 1), acc);
                        if (P.db) tc::mma_f16_ss(tmem + (uint32_t)n_xblk * 64u, ad, tg_desc_mn(ones + k * 2048, 1024), tg_idesc(128, 16, 1, 1), acc);
                    }
                } else {
                    // D[128 in (blocks 2*chunk, 2*chunk+1 of X) x 16 out] += X^T . dZ
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t acc = (it | (uint32_t)k) ? 1u : 0u;
                        tc::mma_f16_ss(tmem, tg_desc_mn(S + (2 + 2 * chunk) * TG_BLOCK + k * 2048, TG_BLOCK), tg_desc_mn(S + k * 2048, 1024), tg_idesc(128, 16, 1, 1), acc);
                        if (chunk == 0 && P.db) tc::mma_f16_ss(tmem + 16u, tg_desc_mn(ones + k * 2048, 1024), tg_desc_mn(S + k * 2048, 1024), tg_idesc(64, 16, 1, 1), acc);   // bias: ones^T . dZ
                    }
                }
                tc::mma_commit(bars + 2 + st);
            }
            tc::mma_commit(bars + 4);
        }
    }
    // ---- flush (warps 0-3: thread == accumulator row)
    if (warp < 4 && cta < n_tiles) {
        tc::mbar_wait(bars + 4, 0);
        tc::tc_fence_after_sync();
        const int row = threadIdx.x;
        const uint32_t taddr = tmem + (((uint32_t)(row >> 5) * 32u) << 16);
        const float inv = 1.f / TG_SCALE;
        if (!P.transposed) {
            const int out = chunk * 128 + row;
            for (int b = 0; b < n_xblk; ++b) {
                for (int c = 0; c < 64; c += 16) {
                    float v[16];
                    tc::tmem_ld16(taddr + (uint32_t)b * 64u + c, v);
                    if (out < P.N)
                        for (int k = 0; k < 16; ++k) { const int kc = b * 64 + c + k; if (kc < P.k_cols) atomicAdd(P.dW + (size_t)out * P.K + P.k_off + kc, v[k] * inv); }
                }
            }
            if (P.db) {
                float v[16];
                tc::tmem_ld16(taddr + (uint32_t)n_xblk * 64u, v);     // every column of the ones block gives the same sum over rows of dZ
                if (out < P.N) atomicAdd(P.db + out, v[0] * inv);
            }
        } else {
            float v[16];
            tc::tmem_ld16(taddr, v);
            const int kin = chunk * 128 + row;
            if (kin < P.k_cols)
                for (int k = 0; k < P.N; ++k) atomicAdd(P.dW + (size_t)k * P.K + P.k_off + kin, v[k] * inv);
            if (chunk == 0 && P.db) {
                float b[16];
                tc::tmem_ld16(taddr + 16u, b);                     // M = 64 atom: row 0 sits in lane 0 of warp 0
                if (row == 0) for (int k = 0; k < P.N; ++k) atomicAdd(P.db + k, b[k] * inv);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------- small helpers
// nn.Linear weight W[N][K_in] (fp32, row-major) -> slabs: for every 64-column input block kb (columns map[kb] .. +64 of the layer input, zero beyond `valid`): [n_pad x 64] fp16 SW128 K-major
struct TgPackP { const float *W; int N, K_in, n_pad; int n_kb; int col0[TG_MAX_A]; int valid[TG_MAX_A]; uint8_t *out; };
__global__ void tg_pack_weights_kernel(const __grid_constant__ TgPackP P) {
    const int kb = blockIdx.x;
    uint8_t *slab = P.out + (size_t)kb * P.n_pad * 128;
    for (int t = threadIdx.x; t < P.n_pad * 8; t += blockDim.x) {
        const int row = t >> 3, chunk = t & 7;
        __half h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = chunk * 8 + e;
            h[e] = (row < P.N && c < P.valid[kb]) ? __float2half_rn(P.W[(size_t)row * P.K_in + P.col0[kb] + c]) : __float2half_rn(0.f);
        }
        *reinterpret_cast<uint4 *>(slab + sw128_offset(row, chunk)) = *reinterpret_cast<uint4 *>(h);
    }
}
// dL/draw fp32 [rows,4] -> two one-block tile images (fp16, scaled): d_rgb in columns 0..2, d_alpha in column 0; rows beyond n_rows are zero
__global__ void __launch_bounds__(128) tg_pack_draw_kernel(const float4 *__restrict__ d_raw, long long n_rows, uint8_t *__restrict__ z_rgb, uint8_t *__restrict__ z_alpha) {
    const long long tile = blockIdx.x, i = tile * 128 + threadIdx.x;
    const float4 g = i < n_rows ? d_raw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = threadIdx.x;
    uint8_t *br = z_rgb + (size_t)tile * TG_BLOCK, *ba = z_alpha + (size_t)tile * TG_BLOCK;
    __half2 a = __floats2half2_rn(g.x * TG_SCALE, g.y * TG_SCALE), b = __floats2half2_rn(g.z * TG_SCALE, 0.f), c = __floats2half2_rn(g.w * TG_SCALE, 0.f);
    uint4 r = make_uint4(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b), 0u, 0u), al = make_uint4(*reinterpret_cast<uint32_t *>(&c), 0u, 0u, 0u), z = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(br + sw128_offset(row, 0)) = r; *reinterpret_cast<uint4 *>(br + sw128_offset(row, 1)) = z;
    *reinterpret_cast<uint4 *>(ba + sw128_offset(row, 0)) = al; *reinterpret_cast<uint4 *>(ba + sw128_offset(row, 1)) = z;
}

static size_t tg_smem() { return 1024 + (size_t)TG_MAX_A * TG_BLOCK + 2 * TG_RING_SLOT + 4 * TG_BLOCK + 8 * TB_N + 64; }
static size_t tg_dw_smem() { return 1024 + 2 * 6 * (size_t)TG_BLOCK + TG_BLOCK + 8 * 5 + 64; }

}  // namespace xrb

using namespace xrb;

extern "C" {

// One dense layer of the NerfMLP training forward / input-gradient pass over tile images. See include/xrnerf_b200.h.
int xrb_nerf_tg_layer(const xrb_tg_layer *L, void *stream) {
    XRB_REQUIRE(L, "tg_layer: null descriptor");
    XRB_REQUIRE(L->n_rows >= 0 && L->n_a >= 1 && L->n_a <= 2 && L->n_slabs >= 1 && L->n_slabs <= TG_MAX_SLABS, "tg_layer: bad sizes");
    if (L->n_rows == 0) return XRB_OK;
    XRB_REQUIRE(L->w && L->a_base[0], "tg_layer: null pointer");
    XRB_REQUIRE(L->epi >= 0 && L->epi <= 3 && L->out_cols >= 16 && L->out_cols <= 256 && (L->out_cols % 16) == 0, "tg_layer: bad epilogue");
    XRB_REQUIRE(L->mma_n >= 16 && L->mma_n <= 256 && (L->mma_n % 16) == 0, "tg_layer: bad MMA N");
    TgP P{};
    int n_ablk = 0;
    for (int s = 0; s < L->n_a; ++s) {
        P.a[s] = TgOp{(const uint8_t *)L->a_base[s], L->a_tile_stride[s], L->a_blk_off[s], L->a_n_blk[s]};
        XRB_REQUIRE(((uintptr_t)L->a_base[s] & 127) == 0 && L->a_n_blk[s] >= 1, "tg_layer: A image must be 128-byte aligned");
        n_ablk += L->a_n_blk[s];
    }
    XRB_REQUIRE(n_ablk <= TG_MAX_A, "tg_layer: at most 6 input blocks per tile");
    P.n_a = L->n_a; P.w = (const uint8_t *)L->w; P.n_slabs = L->n_slabs;
    for (int s = 0; s < L->n_slabs; ++s) {
        P.slab[s] = TgSlab{L->slab_w_off[s], L->slab_bytes[s], L->slab_n_sub[s], L->slab_a_blk0[s], L->slab_d_col[s], L->slab_first[s], L->slab_n_k[s]};
        XRB_REQUIRE(L->slab_n_k[s] >= 1 && L->slab_n_k[s] <= 4 && L->slab_n_sub[s] >= 1, "tg_layer: slab K-steps");
        XRB_REQUIRE(L->slab_bytes[s] <= TG_RING_SLOT && (L->slab_bytes[s] % 1024) == 0 && (L->slab_w_off[s] % 128) == 0, "tg_layer: slab size/offset");
        XRB_REQUIRE(L->slab_a_blk0[s] + L->slab_n_sub[s] <= n_ablk && L->slab_d_col[s] + L->mma_n <= 256, "tg_layer: slab refers outside the tile");
    }
    P.b_mn = L->b_mn; P.mma_n = L->mma_n; P.out_cols = L->out_cols; P.epi = L->epi; P.relu = L->relu; P.bias = L->bias;
    P.x = TgOp{(const uint8_t *)L->x_base, L->x_tile_stride, L->x_blk_off, L->out_cols / 64};
    P.y = (uint8_t *)L->y; P.y_tile_stride = L->y_tile_stride; P.y_blk_off = L->y_blk_off;
    P.yf = L->yf; P.yf_stride = L->yf_stride; P.yf_col = L->yf_col; P.yf_n = L->yf_n; P.yf_scale = L->yf_scale; P.n_rows = L->n_rows;
    XRB_REQUIRE(L->epi == 1 ? (L->yf != nullptr && L->yf_n >= 1 && L->yf_n <= 16) : (L->y != nullptr && (L->out_cols % 64) == 0), "tg_layer: output");
    XRB_REQUIRE(L->epi != 2 || L->x_base != nullptr, "tg_layer: mask image missing");
    const size_t smem = tg_smem();
    cudaFuncSetAttribute(tg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long n_tiles = (L->n_rows + 127) / 128; int grid = n_tiles < sms ? (int)n_tiles : sms;
    tg_kernel<<<grid, TG_THREADS, smem, (cudaStream_t)stream>>>(P);
    return check_launch("nerf_tg_layer");
}

int xrb_nerf_tg_dw(const xrb_tg_dw *L, void *stream) {
    XRB_REQUIRE(L, "tg_dw: null descriptor");
    XRB_REQUIRE(L->n_rows >= 0 && L->N >= 1 && L->K >= 1 && L->x_n_blk >= 1 && L->x_n_blk <= 4 && L->k_off >= 0 && L->k_cols >= 1 && L->k_off + L->k_cols <= L->K, "tg_dw: bad sizes");
    if (L->n_rows == 0) return XRB_OK;
    XRB_REQUIRE(L->z_base && L->x_base && L->dW, "tg_dw: null pointer");
    TgDwP P{};
    P.z = TgOp{(const uint8_t *)L->z_base, L->z_tile_stride, L->z_blk_off, 0};
    P.x = TgOp{(const uint8_t *)L->x_base, L->x_tile_stride, L->x_blk_off, L->x_n_blk};
    P.N = L->N; P.K = L->K; P.k_off = L->k_off; P.k_cols = L->k_cols; P.dW = L->dW; P.db = L->db; P.n_rows = L->n_rows;
    P.transposed = L->N <= 16;
    if (P.transposed) { XRB_REQUIRE((L->x_n_blk % 2) == 0, "tg_dw: 16-output layers take an input image of 128 or 256 columns"); P.n_chunks = L->x_n_blk / 2; }
    else { XRB_REQUIRE(L->N == 128 || L->N == 256, "tg_dw: N must be 128 or 256 (or <= 16)"); P.n_chunks = L->N / 128; }
    const size_t smem = tg_dw_smem();
    cudaFuncSetAttribute(tg_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long n_tiles = (L->n_rows + 127) / 128;
    int per = sms / P.n_chunks; if (per > n_tiles) per = (int)n_tiles; if (per < 1) per = 1;
    tg_dw_kernel<<<per * P.n_chunks, 192, smem, (cudaStream_t)stream>>>(P);
    return check_launch("nerf_tg_dw");
}

int xrb_nerf_tg_pack_weights(const float *W, int N, int K_in, int n_pad, int n_kb, const int *col0_host, const int *valid_host, void *slabs, void *stream) {
    XRB_REQUIRE(W && slabs && col0_host && valid_host && N >= 1 && n_pad >= N && (n_pad % 8) == 0 && n_kb >= 1 && n_kb <= TG_MAX_A, "tg_pack_weights: bad arguments");
    XRB_REQUIRE(((uintptr_t)slabs & 127) == 0, "tg_pack_weights: slabs must be 128-byte aligned");
    TgPackP P{}; P.W = W; P.N = N; P.K_in = K_in; P.n_pad = n_pad; P.n_kb = n_kb; P.out = (uint8_t *)slabs;
    for (int k = 0; k < n_kb; ++k) { P.col0[k] = col0_host[k]; P.valid[k] = valid_host[k]; XRB_REQUIRE(col0_host[k] >= 0 && valid_host[k] >= 0 && valid_host[k] <= 64 && col0_host[k] + valid_host[k] <= K_in, "tg_pack_weights: block outside the matrix"); }
    tg_pack_weights_kernel<<<n_kb, 256, 0, (cudaStream_t)stream>>>(P);
    return check_launch("nerf_tg_pack_weights");
}

int xrb_nerf_tg_pack_draw(const float *d_raw, int64_t n_rows, void *z_rgb_image, void *z_alpha_image, void *stream) {
    XRB_REQUIRE(n_rows >= 0, "tg_pack_draw: negative size");
    if (n_rows == 0) return XRB_OK;
    XRB_REQUIRE(d_raw && z_rgb_image && z_alpha_image && ((uintptr_t)d_raw & 15) == 0, "tg_pack_draw: null / misaligned pointer");
    tg_pack_draw_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const float4 *)d_raw, n_rows, (uint8_t *)z_rgb_image, (uint8_t *)z_alpha_image);
    return check_launch("nerf_tg_pack_draw");
}

float xrb_nerf_tg_grad_scale(void) { return TG_SCALE; }

}  // extern "C"
