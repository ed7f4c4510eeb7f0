// micro-benchmark: latencies of the tcgen05 / mbarrier protocol steps used by the fused MLP kernels (developer tool, not product)
#include "../../xrnerf_b200/csrc/tc.cuh"
#include <cstdio>
using namespace xrb;
__global__ void lat_kernel(long long *out) {
    extern __shared__ uint8_t dyn[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(base + 65536), *bar2 = bar + 1, *bar3 = bar + 2;
    uint32_t *slot = (uint32_t *)(bar + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar2, 1); tc::mbar_init(bar3, 4); tc::fence_mbar_init(); }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((uint32_t *)base)[i] = 0x3c003c00u;
    if (warp == 0) tc::tmem_alloc<256>(slot);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync(); __syncthreads(); tc::tc_fence_after_sync();
    uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        // (a) commit with nothing pending -> wait
        long long t0 = clock64(); tc::mma_commit(bar); tc::mbar_wait(bar, ph); ph ^= 1; long long t1 = clock64(); out[0] = t1 - t0;
        t0 = clock64(); tc::mma_commit(bar); tc::mbar_wait(bar, ph); ph ^= 1; t1 = clock64(); out[1] = t1 - t0;
        // (b) 4 MMAs (M128 N256 K16) + commit -> wait
        uint32_t a0 = tc::smem_u32(base), b0 = tc::smem_u32(base + 16384);
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            for (int k = 0; k < 4; ++k) tc::mma_f16_ss(tmem, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), tc::idesc_f16_m128(256), k);
            long long ti = clock64();
            tc::mma_commit(bar); tc::mbar_wait(bar, ph); ph ^= 1; t1 = clock64(); out[2 + 2 * rep] = ti - t0; out[3 + 2 * rep] = t1 - t0;
        }
        // (c) 16 MMAs N=256
        t0 = clock64();
        for (int q = 0; q < 4; ++q) for (int k = 0; k < 4; ++k) tc::mma_f16_ss(tmem, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), tc::idesc_f16_m128(256), 1);
        long long ti = clock64(); tc::mma_commit(bar); tc::mbar_wait(bar, ph); ph ^= 1; t1 = clock64(); out[8] = ti - t0; out[9] = t1 - t0;
        // (d) 32 MMAs N=128
        t0 = clock64();
        for (int q = 0; q < 8; ++q) for (int k = 0; k < 4; ++k) tc::mma_f16_ss(tmem, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), tc::idesc_f16_m128(128), 1);
        ti = clock64(); tc::mma_commit(bar); tc::mbar_wait(bar, ph); ph ^= 1; t1 = clock64(); out[10] = ti - t0; out[11] = t1 - t0;
        // (i) ISSUE cost of tcgen05.commit (no wait), then 8 commits back to back on 8 barriers, then their completion
        {
            uint64_t *bx = (uint64_t *)(base + 65536 + 256);
            for (int q = 0; q < 8; ++q) tc::mbar_init(bx + q, 1);
            tc::fence_mbar_init();
            long long c0 = clock64(); tc::mma_commit(bx); long long c1 = clock64(); out[21] = c1 - c0;
            tc::mbar_wait(bx, 0);
            c0 = clock64();
            for (int q = 1; q < 8; ++q) tc::mma_commit(bx + q);
            c1 = clock64(); out[22] = c1 - c0;
            for (int q = 1; q < 8; ++q) tc::mbar_wait(bx + q, 0);
            out[23] = clock64() - c0;
            // 4 MMAs + commit, 8 times back to back (issue time), then completion
            for (int q = 0; q < 8; ++q) tc::mbar_init(bx + q, 1);
            tc::fence_mbar_init();
            c0 = clock64();
            for (int q = 0; q < 8; ++q) {
                for (int k = 0; k < 4; ++k) tc::mma_f16_ss(tmem, tc::smem_desc_sw128(a0 + k * 32), tc::smem_desc_sw128(b0 + k * 32), tc::idesc_f16_m128(128), 1);
                tc::mma_commit(bx + q);
            }
            c1 = clock64(); out[24] = c1 - c0;
            for (int q = 0; q < 8; ++q) tc::mbar_wait(bx + q, 0);
            out[25] = clock64() - c0;
        }
        // (e) fence.proxy.async cost, tcgen05 fences
        t0 = clock64(); tc::fence_proxy_async_smem(); t1 = clock64(); out[12] = t1 - t0;
        t0 = clock64(); tc::tc_fence_before_sync(); tc::tc_fence_after_sync(); t1 = clock64(); out[13] = t1 - t0;
    }
    __syncthreads();
    // (f) cross-warp mbarrier hop: warp 1 lane 0 arrives, thread 0 waits: measure from a common start
    __shared__ long long ts[4];
    if (warp == 1 && lane == 0) { ts[0] = clock64(); asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar2)) : "memory"); }
    if (threadIdx.x == 0) { tc::mbar_wait(bar2, 0); ts[1] = clock64(); }
    __syncthreads();
    if (threadIdx.x == 0) out[14] = ts[1] - ts[0];
    // (g) tmem_ld32 latency, 4 warps
    if (warp < 4) { uint32_t r[32]; long long t0 = clock64(); tc::tmem_ld32(tmem + ((warp * 32u) << 16), r); long long t1 = clock64(); if (threadIdx.x == 0) { out[15] = t1 - t0; out[16] = r[0]; } }
    // (h) TMA bulk 16 KB from global (L2 hit after first) latency
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t *g = (const uint8_t *)out + 4096;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64(); tc::mbar_expect_tx(bar2, 16384); tc::tma_bulk_g2s(base + 32768, g, 16384, bar2); tc::mbar_wait(bar2, (rep + 1) & 1); long long t1 = clock64(); out[17 + rep] = t1 - t0;
        }
        long long t0 = clock64(); tc::mbar_expect_tx(bar2, 32768); tc::tma_bulk_g2s(base + 32768, g, 16384, bar2); tc::tma_bulk_g2s(base + 49152, g + 16384, 16384, bar2); tc::mbar_wait(bar2, 0); long long t1 = clock64(); out[20] = t1 - t0;
    }
    tc::tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}
int main() {
    long long *d; cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20);
    cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    lat_kernel<<<1, 192, 100000>>>(d); cudaError_t e = cudaDeviceSynchronize();
    long long h[40]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    const char *names[] = {"commit(empty)->wait #1", "commit(empty)->wait #2", "issue 4xMMA N256 (a)", "4xMMA N256 + commit -> done (a)", "issue (b)", "done (b)", "issue (c)", "done (c)", "issue 16xMMA N256", "16xMMA N256 done",
                           "issue 32xMMA N128", "32xMMA N128 done", "fence.proxy.async", "tcgen05 fences", "mbarrier arrive->wait hop", "tmem_ld32+wait", "(r0)", "TMA 16KB #1", "TMA 16KB #2", "TMA 16KB #3", "TMA 2x16KB", "issue 1 commit", "issue 7 commits", "7 commits landed", "issue 8x(4 MMA N128 + commit)", "8x(4 MMA N128 + commit) landed"};
    for (int i = 0; i < 26; ++i) printf("%-34s %lld cycles\n", names[i], h[i]);
    return 0;
}
