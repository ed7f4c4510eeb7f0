// TEST INFRASTRUCTURE ONLY (oracle/): never imported, linked or executed by the product path.
//
// Host-execution shim that lets g++ compile the reference's own CUDA kernels
// (/root/reference/extensions/ngp_raymarch/src/*.cu) as plain C++ and run them
// on CPU threads. It defines the CUDA keywords away, emulates threadIdx/blockIdx,
// the handful of device intrinsics the reference kernels use, and a serial (or
// OpenMP-parallel) replacement for the <<<grid, block>>> launch syntax, which
// oracle/Makefile rewrites to cpu_launch(...) in a generated copy under oracle/_ref/.
//
// Nothing here restates reference arithmetic: the arithmetic is the reference's,
// compiled from where it lies. What the shim does decide:
//   * __expf -> expf (GPU __expf is ex2.approx based; parity tests allow for it)
//   * atomics -> __atomic builtins; with REF_SERIAL launch order == thread order,
//     which makes the reference's atomic-order-dependent buffer layout the
//     ray-order layout
//   * no FMA contraction (compile with -ffp-contract=off)
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cassert>
#include <algorithm>
#include <type_traits>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct ref_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local ref_dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

typedef void *cudaStream_t;
static inline int cudaDeviceSynchronize() { return 0; }
static inline int cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline void __syncthreads() {}

// ---- intrinsics used by the reference kernels
// glibc already declares an internal __expf; route the CUDA intrinsic name to our own symbol.
static inline float ref_fast_expf(float x) { return expf(x); }
#define __expf ref_fast_expf
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }

static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float *p, float v) {
    uint32_t *up = (uint32_t *)p; uint32_t old = __atomic_load_n(up, __ATOMIC_RELAXED);
    for (;;) { float nf = __uint_as_float(old) + v; uint32_t nu = __float_as_uint(nf);
        if (__atomic_compare_exchange_n(up, &old, nu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return __uint_as_float(old); }
}
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) {
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// CUDA's mixed-signedness min/max overloads (cuda/std math_functions.hpp semantics).
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

// Never executed (block_reduce is replaced by cpu_block_reduce below); must only parse.
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }

static inline bool ref_serial() { static int s = -1; if (s < 0) { const char *e = getenv("REF_SERIAL"); s = (!e || atoi(e) != 0) ? 1 : 0; } return s == 1; }

// Replacement for `kernel<<<grid, block, smem, stream>>>(args...)`.
template <typename K, typename... A>
static inline void cpu_launch(K kernel, uint32_t grid, uint32_t block, uint32_t /*smem*/, cudaStream_t, A... args) {
    if (ref_serial()) {
        blockDim.x = block; gridDim.x = grid;
        for (uint32_t b = 0; b < grid; ++b) { blockIdx.x = b; for (uint32_t t = 0; t < block; ++t) { threadIdx.x = t; kernel(args...); } }
    } else {
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t b = 0; b < (int64_t)grid; ++b) {
            blockDim.x = block; gridDim.x = grid; blockIdx.x = (unsigned)b;
            for (uint32_t t = 0; t < block; ++t) { threadIdx.x = t; kernel(args...); }
        }
    }
}

// Replacement for the reference's block_reduce<float,float,F> launch (raymarch_shared.h:657-745):
// float4 per thread -> xor-butterfly warp sum -> per-warp partials -> warp sum -> atomicAdd per block.
// Emulated with the same association order inside a block; blocks are added in index order.
template <typename F>
static inline void cpu_block_reduce(uint32_t n_vec, F fun, const float *input, float *output, uint32_t n_blocks, uint32_t threads) {
    for (uint32_t b = 0; b < n_blocks; ++b) {
        std::vector<float> v(threads, 0.f);
        for (uint32_t t = 0; t < threads; ++t) {
            uint32_t i = t + b * threads;
            if (i < n_vec) { const float *q = input + 4 * (size_t)i; v[t] = fun(q[0]) + fun(q[1]) + fun(q[2]) + fun(q[3]); }
        }
        auto butterfly = [](float *w) { for (int off = 16; off > 0; off /= 2) { float t[32]; for (int l = 0; l < 32; ++l) t[l] = w[l] + w[l ^ off]; memcpy(w, t, sizeof t); } };
        uint32_t n_warps = threads / 32; float sdata[32] = {0};
        for (uint32_t w = 0; w < n_warps; ++w) { butterfly(&v[w * 32]); sdata[w] = v[w * 32]; }
        float fin[32]; for (int l = 0; l < 32; ++l) fin[l] = ((uint32_t)l < n_warps) ? sdata[l] : 0.f;
        butterfly(fin);
        output[0] += fin[0];
    }
}
