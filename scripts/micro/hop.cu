// micro-benchmark: cross-warp signalling latency: mbarrier (try_wait / test_wait / try_wait with hint) vs named barriers; ping-pong x1000
#include "../../xrnerf_b200/csrc/tc.cuh"
#include <cstdio>
using namespace xrb;
__device__ __forceinline__ void arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory"); return ok; }
__device__ __forceinline__ bool try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t hint) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity), "r"(hint) : "memory"); return ok; }
template <int MODE> __device__ __forceinline__ void waitm(uint64_t *b, uint32_t ph) {
    if (MODE == 0) { while (!tc::mbar_try_wait(b, ph)) {} }
    else if (MODE == 1) { while (!test_wait(b, ph)) {} }
    else { while (!try_wait_hint(b, ph, 20)) {} }
}
template <int MODE> __device__ void pingpong(uint64_t *ba, uint64_t *bb, int warp, int lane, long long *out, int slot, int N) {
    // warp 0 lane 0 <-> warp 4 lane 0 (different SMSP? warp%4: 0 and 0 -> same scheduler) ; use warp 5 for a different scheduler
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0 && lane == 0) { for (int i = 0; i < N; ++i) { arrive(ba); waitm<MODE>(bb, i & 1); } }
    if (warp == 5 && lane == 0) { for (int i = 0; i < N; ++i) { waitm<MODE>(ba, i & 1); arrive(bb); } }
    __syncthreads();
    if (threadIdx.x == 0) out[slot] = (clock64() - t0) / N;
}
__global__ void hop_kernel(long long *out) {
    __shared__ uint64_t bars[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(bars + i, 1); tc::fence_mbar_init(); }
    __syncthreads();
    pingpong<0>(bars + 0, bars + 1, warp, lane, out, 0, 1000);
    pingpong<1>(bars + 2, bars + 3, warp, lane, out, 1, 1000);
    pingpong<2>(bars + 4, bars + 5, warp, lane, out, 2, 1000);
    // named barriers: warp 0 and warp 5 ping-pong with bar.sync id 1/2, 64 threads
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) { for (int i = 0; i < 1000; ++i) { asm volatile("bar.arrive 1, 64;" ::: "memory"); asm volatile("bar.sync 2, 64;" ::: "memory"); } }
    if (warp == 5) { for (int i = 0; i < 1000; ++i) { asm volatile("bar.sync 1, 64;" ::: "memory"); asm volatile("bar.arrive 2, 64;" ::: "memory"); } }
    __syncthreads();
    if (threadIdx.x == 0) out[3] = (clock64() - t0) / 1000;
    // 128 threads arriving on one mbarrier (count 128) -> single waiter, x200
    __shared__ uint64_t big[2];
    if (threadIdx.x == 0) { tc::mbar_init(big, 128); tc::mbar_init(big + 1, 1); tc::fence_mbar_init(); }
    __syncthreads();
    t0 = clock64();
    if (warp < 4) { for (int i = 0; i < 200; ++i) { arrive(big); if (lane == 0) waitm<0>(big + 1, i & 1); __syncwarp(); } }
    if (warp == 5 && lane == 0) { for (int i = 0; i < 200; ++i) { waitm<0>(big, i & 1); arrive(big + 1); } }
    __syncthreads();
    if (threadIdx.x == 0) out[4] = (clock64() - t0) / 200;
}
int main() {
    long long *d; cudaMalloc(&d, 4096); cudaMemset(d, 0, 4096);
    hop_kernel<<<1, 192>>>(d); cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\nround trip (2 hops), cycles: try_wait=%lld test_wait=%lld try_wait_hint20=%lld named_bar=%lld\n128-arrive+reply round trip: %lld\n", cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4]);
    return 0;
}
