// xrnerf_b200 — backward of the Instant-NGP field (HashNerfMLP.run_mlp) and the fused Adam step.
//
// One persistent kernel per training step replaces tcnn's backward (3 network/encoding backward launches + 2 `cat`
// backward kernels in the reference stack, hashnerf_mlp.py:55-79) :
//   per 128-sample tile (thread == sample):
//     1. recompute the forward (hash gather -> density net -> SH concat -> colour net), parking every layer INPUT as fp16 in
//        shared memory (needed both for the ReLU masks and as the X operand of the weight gradients);
//     2. walk the layers backwards: dX = W^T dY per thread (weights broadcast from shared memory), and after each layer a
//        block-level register-tiled GEMM  dW[o][i] += sum_s dY[s][o] * X[s][i]  over the tile (dY staged as scaled fp16);
//        the dW accumulators live in REGISTERS for the whole persistent loop and hit HBM once per CTA (atomicAdd);
//     3. scatter dL/d(encoding) into the fp32 hash-table gradient with vector atomics (red.global.add.v2.f32).
// Numeric contract: fp32 backward math on the fp16-rounded forward activations (oracle/tcnn_oracle.c
// oracle_ngp_mlp_backward); the dY staging uses fp16 with a fixed 2^12 scale (tcnn uses fp16 + loss scale 128).
// v1 runs the GEMM-shaped parts on CUDA cores; moving dX/dW onto tcgen05 tiles is the next step.
#include "ngp_field.cuh"
#include <cuda_bf16.h>

namespace xrb {

constexpr int BW_THREADS = 128;
constexpr float DY_SCALE = 4096.f;

struct BwSmem {
    __half *W;      // fp16 weights: density net then colour net (tcnn layout)
    __half *ENC;    // [128][32]
    __half *H[4];   // density hidden activations [128][64] each (up to 4)
    __half *CIN;    // [128][32]
    __half *C[4];   // colour hidden activations
    __half *DY;     // [128][64] staged, scaled dY of the layer being processed
};

// y[o] = sum_k W[o][k] x[k]  (x from shared fp16 row, fp32 accumulate), optional relu, fp16 rounding — forward semantics
template <int IN, int OUT, bool RELU>
__device__ __forceinline__ void fwd_layer_smem(const __half *__restrict__ W, const __half *__restrict__ xrow, float *y) {
    float x[IN];
#pragma unroll
    for (int k = 0; k < IN / 2; ++k) { float2 v = __half22float2(reinterpret_cast<const __half2 *>(xrow)[k]); x[2 * k] = v.x; x[2 * k + 1] = v.y; }
#pragma unroll 4
    for (int o = 0; o < OUT; ++o) {
        float s = 0.f;
        const __half2 *w2 = reinterpret_cast<const __half2 *>(W + (size_t)o * IN);
#pragma unroll
        for (int k = 0; k < IN / 2; ++k) { float2 w = __half22float2(w2[k]); s = fmaf(w.x, x[2 * k], s); s = fmaf(w.y, x[2 * k + 1], s); }
        y[o] = round_h(RELU ? fmaxf(s, 0.f) : s);
    }
}
__device__ __forceinline__ void store_row_h(__half *row, const float *v, int n) {
    for (int k = 0; k < n; k += 2) reinterpret_cast<__half2 *>(row)[k >> 1] = __floats2half2_rn(v[k], v[k + 1]);
}
// gx[k] = sum_o W[o][k] * gy[o]
template <int IN, int OUT>
__device__ __forceinline__ void bwd_layer_dx(const __half *__restrict__ W, const float *gy, float *gx) {
#pragma unroll
    for (int k = 0; k < IN; ++k) gx[k] = 0.f;
#pragma unroll 2
    for (int o = 0; o < OUT; ++o) {
        const float g = gy[o];
        const __half2 *w2 = reinterpret_cast<const __half2 *>(W + (size_t)o * IN);
#pragma unroll
        for (int k = 0; k < IN / 2; ++k) { float2 w = __half22float2(w2[k]); gx[2 * k] = fmaf(w.x, g, gx[2 * k]); gx[2 * k + 1] = fmaf(w.y, g, gx[2 * k + 1]); }
    }
}
// block GEMM: acc[TO][TI] += sum_s DY[s][o0+a] * X[s][i0+b]; thread tile TOxTI, 128 threads cover OUT x IN exactly
template <int IN, int OUT, int TO, int TI>
__device__ __forceinline__ void dw_accumulate(const __half *__restrict__ DY /*[128][64]*/, const __half *__restrict__ X /*[128][IN]*/, float *acc, int n_valid) {
    static_assert((OUT / TO) * (IN / TI) == BW_THREADS, "thread tiling must cover the matrix");
    const int to = threadIdx.x / (IN / TI), ti = threadIdx.x % (IN / TI);
    for (int s = 0; s < n_valid; ++s) {
        float dy[TO], x[TI];
#pragma unroll
        for (int a = 0; a < TO; a += 2) { float2 v = __half22float2(*reinterpret_cast<const __half2 *>(DY + (size_t)s * 64 + to * TO + a)); dy[a] = v.x; dy[a + 1] = v.y; }
#pragma unroll
        for (int b = 0; b < TI; b += 2) { float2 v = __half22float2(*reinterpret_cast<const __half2 *>(X + (size_t)s * IN + ti * TI + b)); x[b] = v.x; x[b + 1] = v.y; }
#pragma unroll
        for (int a = 0; a < TO; ++a)
#pragma unroll
            for (int b = 0; b < TI; ++b) acc[a * TI + b] = fmaf(dy[a], x[b], acc[a * TI + b]);
    }
}
template <int IN, int OUT, int TO, int TI>
__device__ __forceinline__ void dw_flush(const float *acc, float *__restrict__ dW /* [OUT][IN] */) {
    const int to = threadIdx.x / (IN / TI), ti = threadIdx.x % (IN / TI);
#pragma unroll
    for (int a = 0; a < TO; ++a)
#pragma unroll
        for (int b = 0; b < TI; ++b) atomicAdd(dW + (size_t)(to * TO + a) * IN + ti * TI + b, acc[a * TI + b] * (1.f / DY_SCALE));
}
__device__ __forceinline__ void stage_dy(__half *DY, const float *g, int n) {  // my row; n <= 64, rest untouched
    __half *row = DY + (size_t)threadIdx.x * 64;
    for (int k = 0; k < n; k += 2) reinterpret_cast<__half2 *>(row)[k >> 1] = __floats2half2_rn(g[k] * DY_SCALE, g[k + 1] * DY_SCALE);
}

// DH/CH fixed by template so that the register accumulators have static shapes
template <int DH, int CH>
__global__ void __launch_bounds__(BW_THREADS) ngp_field_bwd_kernel(HashGridDev g, const __half2 *__restrict__ table, const uint8_t *__restrict__ cells, const __half *__restrict__ dens_p, const __half *__restrict__ color_p,
                                                                   const float *__restrict__ pts, int pts_stride, const float *__restrict__ dirs, int dirs_stride,
                                                                   const float4 *__restrict__ dl_draw, int n, const int32_t *__restrict__ n_dev, float *__restrict__ d_table,
                                                                   float *__restrict__ d_dens, float *__restrict__ d_color) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    if (n_dev) n = min(n, max(*n_dev, 0));
    constexpr int ND = 64 * 32 + (DH - 1) * 64 * 64 + 16 * 64, NC = 64 * 32 + (CH - 1) * 64 * 64 + 16 * 64;
    __half *p = reinterpret_cast<__half *>(smem_raw);
    __half *W = p; p += ND + NC;
    __half *ENC = p; p += 128 * 32;
    __half *H[DH]; for (int k = 0; k < DH; ++k) { H[k] = p; p += 128 * 64; }
    __half *CIN = p; p += 128 * 32;
    __half *C[CH]; for (int k = 0; k < CH; ++k) { C[k] = p; p += 128 * 64; }
    __half *DY = p; p += 128 * 64;
    for (int k = threadIdx.x; k < ND; k += BW_THREADS) W[k] = dens_p[k];
    for (int k = threadIdx.x; k < NC; k += BW_THREADS) W[ND + k] = color_p[k];
    const __half *Wd = W, *Wc = W + ND;
    // register accumulators of the weight gradients (scaled by DY_SCALE)
    float a_d0[16] = {0}, a_dh[DH > 1 ? (DH - 1) * 32 : 1] = {0}, a_do[8] = {0}, a_c0[16] = {0}, a_ch[CH > 1 ? (CH - 1) * 32 : 1] = {0}, a_co[8] = {0};
    __syncthreads();

    const int n_tiles = (n + 127) / 128;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i = tile * 128 + threadIdx.x;
        const bool valid = i < n;
        const int n_valid = min(128, n - tile * 128);
        float px = 0.5f, py = 0.5f, pz = 0.5f, dx = 0.5f, dy = 0.5f, dz = 0.5f;
        float4 gr = make_float4(0, 0, 0, 0);
        if (valid) {
            const float *pp = pts + (size_t)i * pts_stride; px = pp[0]; py = pp[1]; pz = pp[2];
            const float *dd = dirs + (size_t)i * dirs_stride; dx = dd[0]; dy = dd[1]; dz = dd[2];
            gr = __ldg(dl_draw + i);
        }
        // ------------------------------------------------ forward recompute, layer inputs parked in shared memory
        {
            float enc[32];
#pragma unroll
            for (int l = 0; l < 16; ++l) { float2 f = hash_level(table, cells, g, l, px, py, pz); enc[2 * l] = f.x; enc[2 * l + 1] = f.y; }
            store_row_h(ENC + (size_t)threadIdx.x * 32, enc, 32);
        }
        float dout[16];
        {
            float h[64];
            fwd_layer_smem<32, 64, true>(Wd, ENC + (size_t)threadIdx.x * 32, h);
            store_row_h(H[0] + (size_t)threadIdx.x * 64, h, 64);
#pragma unroll
            for (int k = 1; k < DH; ++k) { fwd_layer_smem<64, 64, true>(Wd + 64 * 32 + (k - 1) * 64 * 64, H[k - 1] + (size_t)threadIdx.x * 64, h); store_row_h(H[k] + (size_t)threadIdx.x * 64, h, 64); }
            fwd_layer_smem<64, 16, false>(Wd + 64 * 32 + (DH - 1) * 64 * 64, H[DH - 1] + (size_t)threadIdx.x * 64, dout);
        }
        {
            float cin[32], sh[16];
            sh4(dx, dy, dz, sh);
#pragma unroll
            for (int k = 0; k < 15; ++k) cin[k] = dout[k + 1];
#pragma unroll
            for (int k = 0; k < 16; ++k) cin[15 + k] = sh[k];
            cin[31] = 1.0f;
            store_row_h(CIN + (size_t)threadIdx.x * 32, cin, 32);
            float h[64];
            fwd_layer_smem<32, 64, true>(Wc, CIN + (size_t)threadIdx.x * 32, h);
            store_row_h(C[0] + (size_t)threadIdx.x * 64, h, 64);
#pragma unroll
            for (int k = 1; k < CH; ++k) { fwd_layer_smem<64, 64, true>(Wc + 64 * 32 + (k - 1) * 64 * 64, C[k - 1] + (size_t)threadIdx.x * 64, h); store_row_h(C[k] + (size_t)threadIdx.x * 64, h, 64); }
        }
        // ------------------------------------------------ backward: colour net
        float gy[64], gx[64];
#pragma unroll
        for (int k = 0; k < 16; ++k) gy[k] = 0.f;
        gy[0] = gr.x; gy[1] = gr.y; gy[2] = gr.z;
        // output layer  [16 x 64]
        stage_dy(DY, gy, 16);
        __syncthreads();
        dw_accumulate<64, 16, 2, 4>(DY, C[CH - 1], a_co, n_valid);
        bwd_layer_dx<64, 16>(Wc + 64 * 32 + (CH - 1) * 64 * 64, gy, gx);
        __syncthreads();
#pragma unroll
        for (int k = CH - 1; k >= 1; --k) {  // hidden layer k: input C[k-1], output C[k]
            const __half *act = C[k] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<64, 64, 4, 8>(DY, C[k - 1], a_ch + (k - 1) * 32, n_valid);
            bwd_layer_dx<64, 64>(Wc + 64 * 32 + (k - 1) * 64 * 64, gy, gx);
            __syncthreads();
        }
        {
            const __half *act = C[0] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<32, 64, 4, 4>(DY, CIN, a_c0, n_valid);
            bwd_layer_dx<32, 64>(Wc, gy, gx);   // gx[0..31] = dL/dcin
            __syncthreads();
        }
        // ------------------------------------------------ backward: density net; dL/ddout = (draw.w, dcin[0..14])
        gy[0] = gr.w;
#pragma unroll
        for (int k = 0; k < 15; ++k) gy[k + 1] = gx[k];
        stage_dy(DY, gy, 16);
        __syncthreads();
        dw_accumulate<64, 16, 2, 4>(DY, H[DH - 1], a_do, n_valid);
        bwd_layer_dx<64, 16>(Wd + 64 * 32 + (DH - 1) * 64 * 64, gy, gx);
        __syncthreads();
#pragma unroll
        for (int k = DH - 1; k >= 1; --k) {
            const __half *act = H[k] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<64, 64, 4, 8>(DY, H[k - 1], a_dh + (k - 1) * 32, n_valid);
            bwd_layer_dx<64, 64>(Wd + 64 * 32 + (k - 1) * 64 * 64, gy, gx);
            __syncthreads();
        }
        {
            const __half *act = H[0] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<32, 64, 4, 4>(DY, ENC, a_d0, n_valid);
            bwd_layer_dx<32, 64>(Wd, gy, gx);   // gx[0..31] = dL/denc
            __syncthreads();
        }
        // ------------------------------------------------ hash-table gradient scatter
        if (valid) {
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
                const float g0 = gx[2 * l], g1 = gx[2 * l + 1];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float w = ((c & 1) ? fx : 1.f - fx) * ((c & 2) ? fy : 1.f - fy) * ((c & 4) ? fz : 1.f - fz);
                    uint32_t idx = grid_index(ix + (c & 1), iy + ((c >> 1) & 1), iz + ((c >> 2) & 1), hs, res);
                    atomicAdd(tl + idx, make_float2(w * g0, w * g1));
                }
            }
        }
    }
    // ---------------------------------------------------- flush weight-gradient accumulators (once per CTA)
    dw_flush<32, 64, 4, 4>(a_d0, d_dens);
#pragma unroll
    for (int k = 1; k < DH; ++k) dw_flush<64, 64, 4, 8>(a_dh + (k - 1) * 32, d_dens + 64 * 32 + (k - 1) * 64 * 64);
    dw_flush<64, 16, 2, 4>(a_do, d_dens + 64 * 32 + (DH - 1) * 64 * 64);
    dw_flush<32, 64, 4, 4>(a_c0, d_color);
#pragma unroll
    for (int k = 1; k < CH; ++k) dw_flush<64, 64, 4, 8>(a_ch + (k - 1) * 32, d_color + 64 * 32 + (k - 1) * 64 * 64);
    dw_flush<64, 16, 2, 4>(a_co, d_color + 64 * 32 + (CH - 1) * 64 * 64);
}

template <int DH, int CH>
static int launch_bwd(const HashGridDev &g, const void *table, const void *cells, const void *dens, const void *color, const float *pts, int pts_stride, const float *dirs, int dirs_stride,
                      const float *dl_draw, int n, const int32_t *n_dev, float *d_table, float *d_dens, float *d_color, cudaStream_t s) {
    constexpr int ND = 64 * 32 + (DH - 1) * 64 * 64 + 16 * 64, NC = 64 * 32 + (CH - 1) * 64 * 64 + 16 * 64;
    size_t smem = sizeof(__half) * ((size_t)ND + NC + 128 * 32 * 2 + (size_t)(DH + CH) * 128 * 64 + 128 * 64);
    auto k = ngp_field_bwd_kernel<DH, CH>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, BW_THREADS, smem); if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int n_tiles = (n + 127) / 128, grid = sms * per_sm; if (grid > n_tiles) grid = n_tiles; if (grid < 1) grid = 1;
    k<<<grid, BW_THREADS, smem, s>>>(g, (const __half2 *)table, (const uint8_t *)cells, (const __half *)dens, (const __half *)color, pts, pts_stride, dirs, dirs_stride, (const float4 *)dl_draw, n, n_dev, d_table,
                                     d_dens, d_color);
    return check_launch("ngp_mlp_backward");
}

// ---------------------------------------------------------------------------- fused Adam (+ fp16 shadow refresh)
template <typename G> __device__ __forceinline__ float grad_load(const G *g, int64_t i);
template <> __device__ __forceinline__ float grad_load<float>(const float *g, int64_t i) { return g[i]; }
template <> __device__ __forceinline__ float grad_load<__nv_bfloat16>(const __nv_bfloat16 *g, int64_t i) { return __bfloat162float(g[i]); }

template <typename G>
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, __half *__restrict__ p16, const G *__restrict__ grad, float *__restrict__ m, float *__restrict__ v, int64_t n,
                                                   float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt_inv, float grad_mul, float *__restrict__ ema, float ema_m) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float w = p[i];
        float gval = grad_load<G>(grad, i) * grad_mul + wd * w;       // torch.optim.Adam: weight_decay folded into the gradient
        float mi = b1 * m[i] + (1.f - b1) * gval;
        float vi = b2 * v[i] + (1.f - b2) * gval * gval;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) * bc2_sqrt_inv + eps;                // sqrt(v)/sqrt(bias_correction2) + eps
        w -= (lr / bc1) * (mi / denom);
        p[i] = w;
        if (p16) p16[i] = __float2half_rn(w);
        if (ema) ema[i] = ema[i] * (1.f - ema_m) + ema_m * w;        // EMAHook.after_train_iter: buffer.mul_(1 - momentum).add_(momentum, param)
    }
}

// ---------------------------------------------------------------------------- backward of the stand-alone tcnn-shaped modules (xrnerf_b200/tcnn.py autograd)
// tcnn.Encoding(HashGrid).backward: dL/d enc fp16[n,32] -> fp32 table gradient (+=). thread = (sample, level) like hashgrid_forward_kernel.
__global__ void __launch_bounds__(256) hashgrid_backward_kernel(HashGridDev g, const float *__restrict__ x, int x_stride, int n, const __half2 *__restrict__ d_enc, float scale,
                                                                float *__restrict__ d_table) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = t >> 4; int l = (int)(t & 15);
    if (i >= n) return;
    const float *p = x + (size_t)i * x_stride;
    const float2 gy = __half22float2(d_enc[(size_t)i * 16 + l]);
    const float g0 = gy.x * scale, g1 = gy.y * scale;
    if (g0 == 0.f && g1 == 0.f) return;
    const uint32_t hs = g.offset[l + 1] - g.offset[l], res = g.res[l];
    float2 *tl = reinterpret_cast<float2 *>(d_table) + g.offset[l];
    const float sc = g.scale[l];
    float qx = __fmaf_rn(sc, p[0], 0.5f), qy = __fmaf_rn(sc, p[1], 0.5f), qz = __fmaf_rn(sc, p[2], 0.5f);
    int ixs, iys, izs;
    float fx = floor_small(qx, &ixs), fy = floor_small(qy, &iys), fz = floor_small(qz, &izs);
    const uint32_t ix = (uint32_t)ixs, iy = (uint32_t)iys, iz = (uint32_t)izs;
    fx = qx - fx; fy = qy - fy; fz = qz - fz;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = ((c & 1) ? fx : 1.f - fx) * ((c & 2) ? fy : 1.f - fy) * ((c & 4) ? fz : 1.f - fz);
        uint32_t idx = grid_index(ix + (c & 1), iy + ((c >> 1) & 1), iz + ((c >> 2) & 1), hs, res);
        atomicAdd(tl + idx, make_float2(w * g0, w * g1));
    }
}

void launch_hashgrid_backward(const HashGridDev &g, const float *x, int x_stride, int n, const void *d_enc, float scale, float *d_table, cudaStream_t s) {
    int64_t threads = (int64_t)n * 16;
    hashgrid_backward_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(g, x, x_stride, n, (const __half2 *)d_enc, scale, d_table);
}

// tcnn.Network(FullyFusedMLP).backward: x fp16[n,32], dy fp16[n,16] -> dx fp16[n,32] (optional), fp32 parameter gradient (+=). Same structure as the density half of
// ngp_field_bwd_kernel: forward recompute with the layer inputs parked in shared memory, per-thread dX, block GEMM dW with register accumulators.
template <int NH>
__global__ void __launch_bounds__(BW_THREADS) tcnn_mlp_bwd_kernel(const __half *__restrict__ params, const __half *__restrict__ x, const __half *__restrict__ dy, int n,
                                                                  __half *__restrict__ dx, float *__restrict__ d_params) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int NPAR = 64 * 32 + (NH - 1) * 64 * 64 + 16 * 64;
    __half *p = reinterpret_cast<__half *>(smem_raw);
    __half *W = p; p += NPAR;
    __half *X = p; p += 128 * 32;
    __half *H[NH]; for (int k = 0; k < NH; ++k) { H[k] = p; p += 128 * 64; }
    __half *DY = p; p += 128 * 64;
    for (int k = threadIdx.x; k < NPAR; k += BW_THREADS) W[k] = params[k];
    float a_in[16] = {0}, a_hid[NH > 1 ? (NH - 1) * 32 : 1] = {0}, a_out[8] = {0};
    __syncthreads();
    const int n_tiles = (n + 127) / 128;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int i = tile * 128 + threadIdx.x;
        const bool valid = i < n;
        const int n_valid = min(128, n - tile * 128);
        {
            uint4 *dst = reinterpret_cast<uint4 *>(X + (size_t)threadIdx.x * 32);
            const uint4 *src = reinterpret_cast<const uint4 *>(x + (size_t)i * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = valid ? src[q] : make_uint4(0, 0, 0, 0);
        }
        float h[64];
        fwd_layer_smem<32, 64, true>(W, X + (size_t)threadIdx.x * 32, h);
        store_row_h(H[0] + (size_t)threadIdx.x * 64, h, 64);
#pragma unroll
        for (int k = 1; k < NH; ++k) { fwd_layer_smem<64, 64, true>(W + 64 * 32 + (k - 1) * 64 * 64, H[k - 1] + (size_t)threadIdx.x * 64, h); store_row_h(H[k] + (size_t)threadIdx.x * 64, h, 64); }
        float gy[64], gx[64];
#pragma unroll
        for (int k = 0; k < 16; ++k) gy[k] = valid ? __half2float(dy[(size_t)i * 16 + k]) : 0.f;
        stage_dy(DY, gy, 16);
        __syncthreads();
        dw_accumulate<64, 16, 2, 4>(DY, H[NH - 1], a_out, n_valid);
        bwd_layer_dx<64, 16>(W + 64 * 32 + (NH - 1) * 64 * 64, gy, gx);
        __syncthreads();
#pragma unroll
        for (int k = NH - 1; k >= 1; --k) {
            const __half *act = H[k] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<64, 64, 4, 8>(DY, H[k - 1], a_hid + (k - 1) * 32, n_valid);
            bwd_layer_dx<64, 64>(W + 64 * 32 + (k - 1) * 64 * 64, gy, gx);
            __syncthreads();
        }
        {
            const __half *act = H[0] + (size_t)threadIdx.x * 64;
#pragma unroll
            for (int q = 0; q < 64; ++q) gy[q] = __half2float(act[q]) > 0.f ? gx[q] : 0.f;
            stage_dy(DY, gy, 64);
            __syncthreads();
            dw_accumulate<32, 64, 4, 4>(DY, X, a_in, n_valid);
            bwd_layer_dx<32, 64>(W, gy, gx);
            __syncthreads();
        }
        if (dx && valid) store_row_h(dx + (size_t)i * 32, gx, 32);
    }
    dw_flush<32, 64, 4, 4>(a_in, d_params);
#pragma unroll
    for (int k = 1; k < NH; ++k) dw_flush<64, 64, 4, 8>(a_hid + (k - 1) * 32, d_params + 64 * 32 + (k - 1) * 64 * 64);
    dw_flush<64, 16, 2, 4>(a_out, d_params + 64 * 32 + (NH - 1) * 64 * 64);
}

template <int NH>
static int launch_tcnn_mlp_bwd(const void *params, const void *x, const void *dy, int n, void *dx, float *d_params, cudaStream_t s) {
    constexpr int NPAR = 64 * 32 + (NH - 1) * 64 * 64 + 16 * 64;
    size_t smem = sizeof(__half) * ((size_t)NPAR + 128 * 32 + (size_t)NH * 128 * 64 + 128 * 64);
    auto k = tcnn_mlp_bwd_kernel<NH>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, BW_THREADS, smem); if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int n_tiles = (n + 127) / 128, grid = sms * per_sm; if (grid > n_tiles) grid = n_tiles; if (grid < 1) grid = 1;
    k<<<grid, BW_THREADS, smem, s>>>((const __half *)params, (const __half *)x, (const __half *)dy, n, (__half *)dx, d_params);
    return check_launch("tcnn_mlp_backward");
}

// fp32 gradient -> bf16 (the wire format of the sharded data-parallel step: reduce-scatter in bf16, see xrnerf_b200/train.py)
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i + 3 < n; i += stride) {
        float4 v = *reinterpret_cast<const float4 *>(src + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o; o.x = *reinterpret_cast<uint32_t *>(&a); o.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(dst + i) = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) { int64_t t = (n & ~(int64_t)3) + threadIdx.x; if (t < n) dst[t] = __float2bfloat16_rn(src[t]); }
}

// 5 * HuberLoss(delta, 'sum') of HashNerfNetwork.train_step (networks/utils/metrics.py:8-16, hashnerf.py:39-44): gradient wrt rgb and the loss value in ONE pass
__global__ void __launch_bounds__(256) huber5_kernel(const float *__restrict__ rgb, const float *__restrict__ target, int64_t n, float delta, float *__restrict__ grad, float *__restrict__ loss_accum) {
    float local = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float diff = rgb[i] - target[i], rel = fabsf(diff);
        local += rel > delta ? rel - 0.5f * delta : (0.5f / delta) * rel * rel;
        grad[i] = 5.f * (rel > delta ? copysignf(1.f, diff) : diff / delta);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += part[k]; atomicAdd(loss_accum, 5.f * t); }
}

}  // namespace xrb

using namespace xrb;

extern "C" {

int xrb_ngp_huber5_grad(const float *rgb, const float *target, int64_t n_elements, float delta, float *grad_out, float *loss_accum, void *stream) {
    XRB_REQUIRE(n_elements >= 0 && delta > 0.f, "huber5_grad: bad arguments");
    if (n_elements == 0) return XRB_OK;
    XRB_REQUIRE(rgb && target && grad_out && loss_accum, "huber5_grad: null pointer");
    int64_t blocks = (n_elements + 255) / 256; if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
    huber5_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, target, n_elements, delta, grad_out, loss_accum);
    return check_launch("huber5_grad");
}

int xrb_pack_bf16(const float *src, void *dst_bf16, int64_t n, void *stream) {
    XRB_REQUIRE(n >= 0, "pack_bf16: negative size");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(src && dst_bf16, "pack_bf16: null pointer");
    XRB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 7) == 0, "pack_bf16: misaligned");
    int64_t blocks = (n / 4 + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16; if (blocks < 1) blocks = 1;
    pack_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16 *)dst_bf16, n);
    return check_launch("pack_bf16");
}

int xrb_tcnn_hashgrid_backward(const xrb_ngp_config *cfg, const float *x, int x_stride, int n, const void *d_enc_fp16, float grad_scale, float *d_table, void *stream) {
    int e = check_cfg(cfg); if (e) return e;
    XRB_REQUIRE(n >= 0 && x_stride >= 3, "hashgrid_backward: bad size");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(x && d_enc_fp16 && d_table, "hashgrid_backward: null pointer");
    XRB_REQUIRE(((uintptr_t)d_table & 7) == 0 && ((uintptr_t)d_enc_fp16 & 3) == 0, "hashgrid_backward: misaligned");
    HashGridDev g; hashgrid_build(cfg, &g);
    int64_t threads = (int64_t)n * 16;
    hashgrid_backward_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, x, x_stride, n, (const __half2 *)d_enc_fp16, grad_scale, d_table);
    return check_launch("hashgrid_backward");
}

int xrb_tcnn_mlp_backward(const void *params_fp16, const void *x_fp16, const void *dy_fp16, int n, int in_w, int width, int n_hidden, void *dx_fp16, float *d_params, void *stream) {
    XRB_REQUIRE(n >= 0, "mlp_backward: negative size");
    if (in_w != 32 || width != 64 || n_hidden < 1 || n_hidden > 4) { set_error("mlp_backward: only in=32 (padded), width=64, 1..4 hidden layers"); return XRB_E_UNSUPPORTED; }
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(params_fp16 && x_fp16 && dy_fp16 && d_params, "mlp_backward: null pointer");
    XRB_REQUIRE(((uintptr_t)x_fp16 & 15) == 0 && (!dx_fp16 || ((uintptr_t)dx_fp16 & 3) == 0), "mlp_backward: x must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    switch (n_hidden) {
        case 1: return launch_tcnn_mlp_bwd<1>(params_fp16, x_fp16, dy_fp16, n, dx_fp16, d_params, s);
        case 2: return launch_tcnn_mlp_bwd<2>(params_fp16, x_fp16, dy_fp16, n, dx_fp16, d_params, s);
        case 3: return launch_tcnn_mlp_bwd<3>(params_fp16, x_fp16, dy_fp16, n, dx_fp16, d_params, s);
        default: return launch_tcnn_mlp_bwd<4>(params_fp16, x_fp16, dy_fp16, n, dx_fp16, d_params, s);
    }
}

int xrb_ngp_mlp_backward(const xrb_ngp_config *cfg, const xrb_ngp_table *table, const void *density_fp16, const void *color_fp16, const float *pts, int pts_stride, const float *dirs,
                         int dirs_stride, const float *dl_draw, int n, float *d_table, float *d_density, float *d_color, void *stream) {
    int e = check_cfg(cfg); if (e) return e;
    XRB_REQUIRE(n >= 0 && pts_stride >= 3 && dirs_stride >= 3, "ngp_mlp_backward: bad size");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(table && density_fp16 && color_fp16 && pts && dirs && dl_draw && d_table && d_density && d_color, "ngp_mlp_backward: null pointer");
    XRB_REQUIRE(((uintptr_t)dl_draw & 15) == 0 && ((uintptr_t)d_table & 7) == 0, "ngp_mlp_backward: dl_draw must be 16-byte, d_table 8-byte aligned");
    HashGridDev g; e = table_setup(cfg, table, &g, "ngp_mlp_backward"); if (e) return e;
    cudaStream_t s = (cudaStream_t)stream;
#define XRB_BW(DH, CH) if (cfg->density_hidden == DH && cfg->color_hidden == CH) return launch_bwd<DH, CH>(g, table->table_fp16, table->cell_image, density_fp16, color_fp16, pts, pts_stride, dirs, dirs_stride, dl_draw, n, nullptr, d_table, d_density, d_color, s);
    XRB_BW(1, 1) XRB_BW(1, 2) XRB_BW(2, 2) XRB_BW(1, 3) XRB_BW(2, 3)
#undef XRB_BW
    set_error("ngp_mlp_backward: (density_hidden, color_hidden) must be one of (1,1) (1,2) (2,2) (1,3) (2,3)");
    return XRB_E_UNSUPPORTED;
}

int xrb_adam_step(float *param, void *param_fp16, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int step, float grad_div, void *stream) {
    return xrb_adam_ema_step(param, param_fp16, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_div, nullptr, 0.f, stream);
}

static int adam_launch(float *param, void *param_fp16, const void *grad, bool grad_bf16, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, float grad_div, float *ema, float ema_momentum, void *stream) {
    XRB_REQUIRE(n >= 0 && step >= 1 && grad_div != 0.f && ema_momentum >= 0.f && ema_momentum <= 1.f, "adam_step: bad arguments");
    if (n == 0) return XRB_OK;
    XRB_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    int64_t blocks = (n + 255) / 256; if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    if (grad_bf16)
        adam_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, (__half *)param_fp16, (const __nv_bfloat16 *)grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                                 (float)bc1, (float)(1.0 / sqrt(bc2)), 1.f / grad_div, ema, ema_momentum);
    else
        adam_kernel<float><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, (__half *)param_fp16, (const float *)grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                                         (float)(1.0 / sqrt(bc2)), 1.f / grad_div, ema, ema_momentum);
    return check_launch("adam_step");
}

int xrb_adam_ema_step(float *param, void *param_fp16, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, float grad_div, float *ema, float ema_momentum, void *stream) {
    return adam_launch(param, param_fp16, grad, false, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_div, ema, ema_momentum, stream);
}

int xrb_adam_ema_step_bf16grad(float *param, void *param_fp16, const void *grad_bf16, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int step, float grad_div, float *ema, float ema_momentum, void *stream) {
    return adam_launch(param, param_fp16, grad_bf16, true, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_div, ema, ema_momentum, stream);
}

}  // extern "C"
