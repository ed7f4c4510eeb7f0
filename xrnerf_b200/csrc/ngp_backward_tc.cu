// xrnerf_b200 — backward of the Instant-NGP field (HashNerfMLP.run_mlp, /root/reference/xrnerf/models/mlps/hashnerf_mlp.py:55-79: what tcnn's
// backward computes behind it) on tcgen05 tensor-core tiles. Replaces the CUDA-core kernel of ngp_backward.cu on the training path (that kernel
// stays as impl 0: reference-grade arithmetic for every (density_hidden, color_hidden) and the comparator of the tests).
//
// One persistent kernel, two warpgroups per CTA, each working on its own 128-sample tile (thread r == sample r == TMEM lane r == row r of every
// shared-memory tile). Per tile:
//   forward recompute   hash gather -> ENC tile -> density net -> CIN tile (density features ++ SH ++ 1) -> colour hidden layers; every layer
//                       INPUT stays in shared memory as a [128 x 64] fp16 tile (K-major, 128-byte swizzle: the UMMA operand layout);
//   backward chain      per layer ONE round of tensor-core work issued by one thread:
//                         dX  = dZ . W            M=128 (samples), N=in, K=out : A = dZ tile (K-major), B = the FORWARD weight image read MN-major
//                                                 (W[out][in] rows are `in`-contiguous: no transposed weight copy exists)
//                         dW += dZ^T . X          M=64 (out features), N=in, K=128 samples : both operands are the sample tiles read MN-major
//                                                 (contraction over ROWS); accumulators live in TMEM for the whole persistent loop
//                       then the epilogue masks dX with ReLU'(activation) and writes the next dZ tile (fp16, fixed 2^12 scale like the CUDA-core
//                       kernel; tcnn uses fp16 gradients with a loss scale);
//   scatter             dL/d(encoding) -> fp32 hash-table gradient with vector atomics (red.global.add.v2.f32), as before.
// At the end each warpgroup adds its dW accumulators (TMEM, M=64 atoms: row m sits in lane 32*(m/16) + m%16) to the fp32 gradient with atomics.
// TMEM: 512 columns = 2 x (64 chain accumulator + <=160 dW).  Shared memory: weight image + 2 x (ENC, H.., CIN, C.., dZ) tiles.
#include "tc_field.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace xrb {

constexpr float BT_SCALE = 4096.f;        // dZ staging scale (fp16): |dL/draw| < 16 stays finite, 1e-8 stays normal
constexpr uint32_t BT_TILE = 16384;

// instruction descriptor: D=F32, A=B=F16; majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t bt_idesc(uint32_t m, uint32_t n, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// MN-major, SWIZZLE_128B: 64 MN elements (128 bytes) contiguous per K row, 8-row groups 1024 bytes apart (SBO); LBO (next 64-element MN atom) unused: extent <= 64
__device__ __forceinline__ uint64_t bt_desc_mn(uint32_t smem_addr_bytes) {
    return (uint64_t)((smem_addr_bytes >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct BtWg {
    uint8_t *tiles;      // this warpgroup's tiles, BT_TILE bytes each
    const uint8_t *W;    // forward weight image (shared by both warpgroups)
    uint64_t *mbar; uint32_t phase;
    uint32_t tmem_chain; // 64 columns, M=128
    uint32_t tmem_dw;    // dW accumulators, M=64 atoms
    uint32_t wg, row;
    bool first;          // first tile of this warpgroup: dW accumulators are overwritten, not accumulated
};

// all 128 threads: make my smem writes visible to the tensor core, meet, ONE thread issues, everybody waits for completion
template <class F>
__device__ __forceinline__ void bt_round(BtWg &c, F issue) {
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    tc::named_bar_sync(1 + c.wg, 128);
    if (c.row == 0) {
        tc::tc_fence_after_sync();
        issue();
        tc::mma_commit(c.mbar);
    }
    tc::mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
    tc::tc_fence_after_sync();
}
__device__ __forceinline__ uint8_t *bt_tile(const BtWg &c, int t) { return c.tiles + (size_t)t * BT_TILE; }
__device__ __forceinline__ void bt_store_chunk(const BtWg &c, int t, uint32_t chunk, uint4 v) { *reinterpret_cast<uint4 *>(bt_tile(c, t) + sw128_offset(c.row, chunk)) = v; }
__device__ __forceinline__ uint4 bt_load_chunk(const BtWg &c, int t, uint32_t chunk) { return *reinterpret_cast<const uint4 *>(bt_tile(c, t) + sw128_offset(c.row, chunk)); }

// forward layer: D[128 x N] = X[128 x K] . W[N x K]^T (both K-major), as tc_layer but with an explicit A tile
template <int K, int N>
__device__ __forceinline__ void bt_fwd(BtWg &c, int x_tile, uint32_t w_off) {
    bt_round(c, [&] {
        const uint32_t a0 = tc::smem_u32(bt_tile(c, x_tile)), b0 = tc::smem_u32(c.W + w_off);
#pragma unroll
        for (int k = 0; k < K / 16; ++k) tc::mma_f16_ss(c.tmem_chain, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), bt_idesc(128, N, 0, 0), k > 0 ? 1u : 0u);
    });
}
// backward round of one layer with OUT outputs and IN inputs:
//   dX[128 x IN] = dZ[128 x OUT] . W[OUT x IN]          (chain accumulator)
//   dW accumulators (M=64 atoms at tmem_dw + dw_col):
//     OUT == 64:  D[m = out][n = in]  += dZ^T X   (A = dZ tile, B = X tile, N = IN)
//     OUT == 16:  D[m = in][n = out]  += X^T dZ   (A = X tile (IN == 64), B = dZ tile, N = 16)   -> stored transposed at the flush
template <int IN, int OUT>
__device__ __forceinline__ void bt_bwd(BtWg &c, int dz_tile, int x_tile, uint32_t w_off, uint32_t dw_col) {
    bt_round(c, [&] {
        const uint32_t z0 = tc::smem_u32(bt_tile(c, dz_tile)), x0 = tc::smem_u32(bt_tile(c, x_tile)), w0 = tc::smem_u32(c.W + w_off);
        const uint32_t acc0 = c.first ? 0u : 1u;
        // weight gradient: contraction over the 128 sample rows, 16 rows (2 x 1024 bytes) per instruction
        if (OUT == 64) {
#pragma unroll
            for (int k = 0; k < 8; ++k) tc::mma_f16_ss(c.tmem_dw + dw_col, bt_desc_mn(z0 + k * 2048), bt_desc_mn(x0 + k * 2048), bt_idesc(64, IN, 1, 1), k > 0 ? 1u : acc0);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) tc::mma_f16_ss(c.tmem_dw + dw_col, bt_desc_mn(x0 + k * 2048), bt_desc_mn(z0 + k * 2048), bt_idesc(64, OUT, 1, 1), k > 0 ? 1u : acc0);
        }
        // input gradient: contraction over OUT; B = forward weight image W[OUT rows][IN cols] read MN-major (N = IN contiguous, K = OUT rows)
#pragma unroll
        for (int k = 0; k < OUT / 16; ++k) tc::mma_f16_ss(c.tmem_chain, tc::smem_desc_sw128(z0 + k * 32), bt_desc_mn(w0 + k * 2048), bt_idesc(128, IN, 0, 1), k > 0 ? 1u : 0u);
    });
}

// epilogues ------------------------------------------------------------------------------------------------------------------------
// forward hidden layer: accumulator (64 cols of my lane) -> ReLU -> fp16 -> tile t
__device__ __forceinline__ void bt_epi_relu(const BtWg &c, int t) {
    const uint32_t taddr = c.tmem_chain + (((c.row >> 5) * 32u) << 16);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + half * 32, r);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            bt_store_chunk(c, t, half * 4 + q, make_uint4(pack_relu_h2(r[8 * q], r[8 * q + 1]), pack_relu_h2(r[8 * q + 2], r[8 * q + 3]), pack_relu_h2(r[8 * q + 4], r[8 * q + 5]),
                                                          pack_relu_h2(r[8 * q + 6], r[8 * q + 7])));
    }
}
// backward hidden layer: dX (64 cols) * [activation > 0] -> fp16 -> dZ tile
__device__ __forceinline__ uint32_t bt_mask_h2(uint32_t a_bits, uint32_t b_bits, uint32_t act2) {
    const __half2 g = __floats2half2_rn(__uint_as_float(a_bits), __uint_as_float(b_bits));
    const __half2 m = __hgt2(*reinterpret_cast<const __half2 *>(&act2), __float2half2_rn(0.f));   // 1.0 / 0.0 per half
    const __half2 o = __hmul2(g, m);
    return *reinterpret_cast<const uint32_t *>(&o);
}
__device__ __forceinline__ void bt_epi_mask(const BtWg &c, int act_tile, int dz_tile) {
    const uint32_t taddr = c.tmem_chain + (((c.row >> 5) * 32u) << 16);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + half * 32, r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 a = bt_load_chunk(c, act_tile, half * 4 + q);
            bt_store_chunk(c, dz_tile, half * 4 + q, make_uint4(bt_mask_h2(r[8 * q], r[8 * q + 1], a.x), bt_mask_h2(r[8 * q + 2], r[8 * q + 3], a.y), bt_mask_h2(r[8 * q + 4], r[8 * q + 5], a.z),
                                                                bt_mask_h2(r[8 * q + 6], r[8 * q + 7], a.w)));
        }
    }
}
__device__ __forceinline__ void bt_read32(const BtWg &c, float *v) {
    const uint32_t taddr = c.tmem_chain + (((c.row >> 5) * 32u) << 16);
    uint32_t r[32];
    tc::tmem_ld32(taddr, r);
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}
__device__ __forceinline__ void bt_read16(const BtWg &c, float *v) {
    const uint32_t taddr = c.tmem_chain + (((c.row >> 5) * 32u) << 16);
    tc::tmem_ld16(taddr, v);
}
// 16 fp32 values -> chunks 0,1 of tile t (fp16)
__device__ __forceinline__ void bt_store16(const BtWg &c, int t, const float *v) {
    bt_store_chunk(c, t, 0, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
    bt_store_chunk(c, t, 1, make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15])));
}

struct BtLayout {   // tile indices and dW column offsets for (DH, CH)
    int enc, h0, cin, c0, dz, n_tiles;
    uint32_t dw_din, dw_dhid, dw_dout, dw_cin, dw_chid, dw_cout, dw_cols;
};
__host__ __device__ constexpr BtLayout bt_layout(int DH, int CH) {
    BtLayout L{};
    L.enc = 0; L.h0 = 1; L.cin = 1 + DH; L.c0 = 2 + DH; L.dz = 2 + DH + CH; L.n_tiles = 3 + DH + CH;
    uint32_t o = 0;
    L.dw_din = o; o += 32; L.dw_dhid = o; o += 64u * (DH - 1); L.dw_dout = o; o += 16; L.dw_cin = o; o += 32; L.dw_chid = o; o += 64u * (CH - 1); L.dw_cout = o; o += 16; L.dw_cols = o;
    return L;
}
template <int DH, int CH>
__host__ __device__ constexpr size_t bt_smem_bytes(uint32_t image_bytes) { return 1024 + image_bytes + 2 * (size_t)bt_layout(DH, CH).n_tiles * BT_TILE + 64; }

// flush one M=64 dW atom: row m of the accumulator lives in TMEM lane 32*(m/16) + m%16, i.e. the lower 16 lanes of warp m/16
// TRANSPOSED == false: D[m = out][n = in] -> dW[out][in] (row length IN);  true: D[m = in][n = out] -> dW[out][in]
template <int NCOLS, bool TRANSPOSED>
__device__ __forceinline__ void bt_flush(const BtWg &c, uint32_t dw_col, float *__restrict__ dW, int in_w) {
    const uint32_t warp = c.row >> 5, lane = c.row & 31;
    const uint32_t taddr = c.tmem_dw + dw_col + ((warp * 32u) << 16);
    const uint32_t m = warp * 16 + lane;
#pragma unroll
    for (int n0 = 0; n0 < NCOLS; n0 += 16) {
        float v[16];
        tc::tmem_ld16(taddr + n0, v);    // .sync.aligned: executed by the whole warp; only lanes < 16 hold accumulator rows
        if (lane < 16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int n = n0 + k;
                float *dst = TRANSPOSED ? dW + (size_t)n * in_w + m : dW + (size_t)m * in_w + n;
                atomicAdd(dst, v[k] * (1.f / BT_SCALE));
            }
        }
    }
}

template <int DH, int CH, int NP>
__global__ void __launch_bounds__(256, 1) ngp_field_bwd_tc_kernel(HashGridDev g, const __half2 *__restrict__ table, const uint8_t *__restrict__ cells, const void *__restrict__ weight_image,
                                                                  uint32_t image_bytes, const float *__restrict__ pts, int pts_stride, const float *__restrict__ dirs, int dirs_stride,
                                                                  const float4 *__restrict__ dl_draw, int n, const int32_t *__restrict__ n_dev, float *__restrict__ d_table,
                                                                  float *__restrict__ d_dens, float *__restrict__ d_color, int dbg, __half2 *__restrict__ genc_out) {
    // dbg (developer ablation, XRB_BWD_DBG): 1 = no hash-gradient scatter, 2 = no gather (encoding := position bits), 4 = no tensor-core rounds,
    // 8 = dL/d encoding written to genc_out (fp16 [n,32]) for a separate scatter kernel instead of the in-kernel atomics
    extern __shared__ uint8_t dyn_smem[];
    if (n_dev) n = min(n, max(*n_dev, 0));
    constexpr BtLayout T = bt_layout(DH, CH);
    static_assert(64 + T.dw_cols <= 256, "TMEM budget: 2 x (64 + dW columns) <= 512");
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    uint8_t *W = base, *tiles0 = base + image_bytes;
    uint64_t *bars = (uint64_t *)(tiles0 + 2 * (size_t)T.n_tiles * BT_TILE);
    uint32_t *tmem_slot = (uint32_t *)(bars + 3);
    const uint32_t warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) tc::mbar_init(bars + k, 1);
        tc::fence_mbar_init();
        tc::mbar_expect_tx(bars + 2, image_bytes);
        tc::tma_bulk_g2s(W, weight_image, image_bytes, bars + 2);
    }
    if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    tc::mbar_wait(bars + 2, 0);
    BtWg c;
    c.wg = threadIdx.x >> 7; c.row = threadIdx.x & 127;
    c.tiles = tiles0 + (size_t)c.wg * T.n_tiles * BT_TILE; c.W = W; c.mbar = bars + c.wg; c.phase = 0;
    const uint32_t tmem_base = *tmem_slot;
    c.tmem_chain = tmem_base + c.wg * 64; c.tmem_dw = tmem_base + 128 + c.wg * T.dw_cols; c.first = true;
    const WeightImageLayout L = weight_image_layout(DH, CH);

    const int n_tiles = (n + 127) / 128;
    for (int tile = blockIdx.x * 2 + c.wg; tile < n_tiles; tile += gridDim.x * 2) {
        const int i = tile * 128 + c.row;
        const bool valid = i < n;
        float px = 0.5f, py = 0.5f, pz = 0.5f, dx = 0.5f, dy = 0.5f, dz = 0.5f;
        float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);       // padding rows: zero dL/draw -> every dZ row of theirs is zero -> no contribution to any gradient
        if (valid) {
            const float *pp = pts + (size_t)i * pts_stride; px = pp[0]; py = pp[1]; pz = pp[2];
            const float *dd = dirs + (size_t)i * dirs_stride; dx = dd[0]; dy = dd[1]; dz = dd[2];
            gr = __ldg(dl_draw + i);
        }
        // ------------------------------------------------ forward recompute: every layer input parked as a tile
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 f = (dbg & 2) ? make_float2(px, py) : hash_level(table, cells, g, 4 * q + k, px, py, pz, plan_mode<NP>(4 * q + k));
                e[k] = pack_h2(f.x, f.y);
            }
            bt_store_chunk(c, T.enc, q, make_uint4(e[0], e[1], e[2], e[3]));
        }
        float genc[32];
        if (dbg & 4) {
#pragma unroll
            for (int k = 0; k < 32; ++k) genc[k] = gr.x * (float)k;
        } else {
        bt_fwd<32, 64>(c, T.enc, L.d_in);
        bt_epi_relu(c, T.h0);
#pragma unroll
        for (int h = 1; h < DH; ++h) { bt_fwd<64, 64>(c, T.h0 + h - 1, L.d_hid[h - 1]); bt_epi_relu(c, T.h0 + h); }
        bt_fwd<64, 16>(c, T.h0 + DH - 1, L.d_out);
        {
            float dout[16], sh[16], cin[32];
            bt_read16(c, dout);
            sh4(dx, dy, dz, sh);
#pragma unroll
            for (int k = 0; k < 15; ++k) cin[k] = dout[k + 1];        // fp16 rounding happens in the pack below (round(round(x)) == round(x))
#pragma unroll
            for (int k = 0; k < 16; ++k) cin[15 + k] = sh[k];
            cin[31] = 1.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                bt_store_chunk(c, T.cin, q, make_uint4(pack_h2(cin[8 * q], cin[8 * q + 1]), pack_h2(cin[8 * q + 2], cin[8 * q + 3]), pack_h2(cin[8 * q + 4], cin[8 * q + 5]), pack_h2(cin[8 * q + 6], cin[8 * q + 7])));
        }
        bt_fwd<32, 64>(c, T.cin, L.c_in);
        bt_epi_relu(c, T.c0);
#pragma unroll
        for (int h = 1; h < CH; ++h) { bt_fwd<64, 64>(c, T.c0 + h - 1, L.c_hid[h - 1]); bt_epi_relu(c, T.c0 + h); }
        // (the colour output layer itself is not needed: dL/draw is given)
        // ------------------------------------------------ backward: colour net
        {
            float gy[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) gy[k] = 0.f;
            gy[0] = gr.x * BT_SCALE; gy[1] = gr.y * BT_SCALE; gy[2] = gr.z * BT_SCALE;
            bt_store16(c, T.dz, gy);
        }
        bt_bwd<64, 16>(c, T.dz, T.c0 + CH - 1, L.c_out, T.dw_cout);
        bt_epi_mask(c, T.c0 + CH - 1, T.dz);
#pragma unroll
        for (int k = CH - 1; k >= 1; --k) {      // hidden layer k: input C[k-1], output C[k]
            bt_bwd<64, 64>(c, T.dz, T.c0 + k - 1, L.c_hid[k - 1], T.dw_chid + 64u * (k - 1));
            bt_epi_mask(c, T.c0 + k - 1, T.dz);
        }
        bt_bwd<32, 64>(c, T.dz, T.cin, L.c_in, T.dw_cin);
        {
            float gcin[32], gy[16];
            bt_read32(c, gcin);                     // dL/dcin (scaled)
            gy[0] = gr.w * BT_SCALE;                // dL/d density_out = (draw.w, dcin[0..14])   (hashnerf_mlp.py:69-78: raw[3] = density_out[0], colour input = density_out[1:])
#pragma unroll
            for (int k = 0; k < 15; ++k) gy[k + 1] = gcin[k];
            bt_store16(c, T.dz, gy);
        }
        // ------------------------------------------------ backward: density net
        bt_bwd<64, 16>(c, T.dz, T.h0 + DH - 1, L.d_out, T.dw_dout);
        bt_epi_mask(c, T.h0 + DH - 1, T.dz);
#pragma unroll
        for (int k = DH - 1; k >= 1; --k) {
            bt_bwd<64, 64>(c, T.dz, T.h0 + k - 1, L.d_hid[k - 1], T.dw_dhid + 64u * (k - 1));
            bt_epi_mask(c, T.h0 + k - 1, T.dz);
        }
        bt_bwd<32, 64>(c, T.dz, T.enc, L.d_in, T.dw_din);
        c.first = false;
        bt_read32(c, genc);                         // dL/d encoding (scaled)
        }
        if ((dbg & 8) && valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) genc_out[(size_t)i * 16 + q] = __floats2half2_rn(genc[2 * q], genc[2 * q + 1]);
        }
        // ------------------------------------------------ hash-table gradient scatter
        if (valid && !(dbg & 9)) {
            // The scatter is bound by the LSU's reduction issue rate (~1.3 cycles per lane and RED, measured: 450 of 880 us per 700 K samples with 128 two-float
            // REDs per sample), so x-adjacent corners that are also adjacent in memory go out as ONE 16-byte red.global.add.v4.f32:
            //   hashed level: entries of (gx, y, z) and (gx^1, y, z) are the two halves of an aligned 16-byte pair -> even gx: 4 v4 instead of 8 v2
            //   dense level : idx and idx + 1 -> one v4 when idx is even (and the pair does not wrap at the table end)
#pragma unroll 1
            for (int l = 0; l < 16; ++l) {
                const uint32_t hs = g.offset[l + 1] - g.offset[l], res = g.res[l];
                float2 *tl = reinterpret_cast<float2 *>(d_table) + g.offset[l];
                const float sc = g.scale[l];
                float qx = __fmaf_rn(sc, px, 0.5f), qy = __fmaf_rn(sc, py, 0.5f), qz = __fmaf_rn(sc, pz, 0.5f);
                int ixs, iys, izs;
                float fx = floor_small(qx, &ixs), fy = floor_small(qy, &iys), fz = floor_small(qz, &izs);
                const uint32_t ix = (uint32_t)ixs, iy = (uint32_t)iys, iz = (uint32_t)izs;
                fx = qx - fx; fy = qy - fy; fz = qz - fz;
                float g0 = 0.f, g1 = 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) if (q == l) { g0 = genc[2 * q]; g1 = genc[2 * q + 1]; }   // register array indexed by the rolled loop counter: select, never spill
                g0 *= (1.f / BT_SCALE); g1 *= (1.f / BT_SCALE);
                const bool hashed = (g.hashed_mask >> l) & 1u;
#pragma unroll
                for (int yz = 0; yz < 4; ++yz) {
                    const float wyz = ((yz & 1) ? fy : 1.f - fy) * ((yz & 2) ? fz : 1.f - fz);
                    const float w0 = (1.f - fx) * wyz, w1 = fx * wyz;
                    const uint32_t i0 = grid_index(ix, iy + (yz & 1), iz + (yz >> 1), hs, res), i1 = grid_index(ix + 1u, iy + (yz & 1), iz + (yz >> 1), hs, res);
                    const bool pair = hashed ? ((ix & 1u) == 0u) : (((i0 & 1u) == 0u) && i1 == i0 + 1u);   // hashed, even gx: i1 == i0 ^ 1; both cases: i0 even, i1 = i0 + 1... or i0 odd (hashed)
                    if (pair) {
                        const uint32_t lo = i0 & ~1u;                      // hashed level with an odd i0: the pair is (i1, i0) in memory order
                        const bool swap = (i0 & 1u) != 0u;
                        const float a0 = swap ? w1 : w0, a1 = swap ? w0 : w1;
                        atomicAdd(reinterpret_cast<float4 *>(tl + lo), make_float4(a0 * g0, a0 * g1, a1 * g0, a1 * g1));
                    } else {
                        atomicAdd(tl + i0, make_float2(w0 * g0, w0 * g1));
                        atomicAdd(tl + i1, make_float2(w1 * g0, w1 * g1));
                    }
                }
            }
        }
    }
    // ---------------------------------------------------- flush the weight-gradient accumulators (once per warpgroup)
    if (!c.first) {
        tc::tc_fence_after_sync();
        float *dd = d_dens, *dc = d_color;
        bt_flush<32, false>(c, T.dw_din, dd, 32);
#pragma unroll
        for (int k = 1; k < DH; ++k) bt_flush<64, false>(c, T.dw_dhid + 64u * (k - 1), dd + 64 * 32 + (k - 1) * 64 * 64, 64);
        bt_flush<16, true>(c, T.dw_dout, dd + 64 * 32 + (DH - 1) * 64 * 64, 64);
        bt_flush<32, false>(c, T.dw_cin, dc, 32);
#pragma unroll
        for (int k = 1; k < CH; ++k) bt_flush<64, false>(c, T.dw_chid + 64u * (k - 1), dc + 64 * 32 + (k - 1) * 64 * 64, 64);
        bt_flush<16, true>(c, T.dw_cout, dc + 64 * 32 + (CH - 1) * 64 * 64, 64);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

void launch_hashgrid_backward(const HashGridDev &g, const float *x, int x_stride, int n, const void *d_enc, float scale, float *d_table, cudaStream_t s);   // ngp_backward.cu

template <int DH, int CH, int NP>
static int launch_bwd_tc(const HashGridDev &g, const xrb_ngp_table *tab, const void *image, uint32_t image_bytes, const float *pts, int pts_stride, const float *dirs, int dirs_stride,
                         const float *dl_draw, int n, const int32_t *n_dev, float *d_table, float *d_dens, float *d_color, cudaStream_t s) {
    auto k = ngp_field_bwd_tc_kernel<DH, CH, NP>;
    const size_t smem = bt_smem_bytes<DH, CH>(image_bytes);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int n_tiles = (n + 127) / 128, grid = sms; if (grid > (n_tiles + 1) / 2) grid = (n_tiles + 1) / 2; if (grid < 1) grid = 1;
    static const int dbg = getenv("XRB_BWD_DBG") ? atoi(getenv("XRB_BWD_DBG")) : 0;
    static __half2 *genc_dbg = nullptr;
    if ((dbg & 8) && !genc_dbg) cudaMalloc(&genc_dbg, (size_t)(1 << 20) * 64);
    k<<<grid, 256, smem, s>>>(g, (const __half2 *)tab->table_fp16, (const uint8_t *)tab->cell_image, image, image_bytes, pts, pts_stride, dirs, dirs_stride, (const float4 *)dl_draw, n, n_dev,
                              d_table, d_dens, d_color, dbg, genc_dbg);
    if (dbg & 8) launch_hashgrid_backward(g, pts, pts_stride, n, genc_dbg, 1.f / BT_SCALE, d_table, s);
    return check_launch("ngp_mlp_backward_tc");
}

}  // namespace xrb

using namespace xrb;

extern "C" {

int xrb_ngp_mlp_backward_tc(const xrb_ngp_config *cfg, const xrb_ngp_table *table, const void *weight_image, const float *pts, int pts_stride, const float *dirs, int dirs_stride,
                            const float *dl_draw, int n, const int32_t *n_rows_dev, float *d_table, float *d_density, float *d_color, void *stream) {
    int e = check_cfg(cfg); if (e) return e;
    XRB_REQUIRE(n >= 0 && pts_stride >= 3 && dirs_stride >= 3, "ngp_mlp_backward_tc: bad size");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(table && weight_image && pts && dirs && dl_draw && d_table && d_density && d_color, "ngp_mlp_backward_tc: null pointer");
    XRB_REQUIRE(((uintptr_t)dl_draw & 15) == 0 && ((uintptr_t)d_table & 15) == 0 && ((uintptr_t)weight_image & 15) == 0, "ngp_mlp_backward_tc: dl_draw, d_table and the weight image must be 16-byte aligned");
    if (!((cfg->density_hidden == 1 && cfg->color_hidden == 1) || (cfg->density_hidden == 1 && cfg->color_hidden == 2))) {
        set_error("ngp_mlp_backward_tc: tensor-core backward is built for (density_hidden, color_hidden) = (1,1) or (1,2); use xrb_ngp_mlp_backward");
        return XRB_E_UNSUPPORTED;
    }
    HashGridDev g; e = table_setup(cfg, table, &g, "ngp_mlp_backward_tc"); if (e) return e;
    const uint32_t image_bytes = weight_image_layout(cfg->density_hidden, cfg->color_hidden).total;
    cudaStream_t s = (cudaStream_t)stream;
    const bool plan6 = plan_valid(g, table->n_packed_levels) && table->n_packed_levels == 6;
#define XRB_BWT(DH, CH)                                                                                                                                          \
    if (cfg->density_hidden == DH && cfg->color_hidden == CH)                                                                                                    \
        return plan6 ? launch_bwd_tc<DH, CH, 6>(g, table, weight_image, image_bytes, pts, pts_stride, dirs, dirs_stride, dl_draw, n, n_rows_dev, d_table, d_density, d_color, s) \
                     : launch_bwd_tc<DH, CH, 0>(g, table, weight_image, image_bytes, pts, pts_stride, dirs, dirs_stride, dl_draw, n, n_rows_dev, d_table, d_density, d_color, s);
    XRB_BWT(1, 1) XRB_BWT(1, 2)
#undef XRB_BWT
    return XRB_E_UNSUPPORTED;
}

}  // extern "C"
