// xrnerf_b200 — measurement utility (not on any product path): the rate at which this GPU serves RANDOM small reads from an L2-resident table, i.e. the roof the
// hash-grid gather actually runs under. The multiresolution hash encoding reads 4-byte entries scattered over a 24.4 MB fp16 table: every entry costs one 32-byte L2
// sector (8x the useful bytes), so its ceiling is the L2 -> SM sector rate, not the HBM copy bandwidth the contract's roofline is quoted against. bench.py runs this
// next to the field kernel and reports both ("give the headline an honest roof").
#include "common.cuh"

namespace xrb {

template <int W> struct Rec;
template <> struct Rec<4> { using T = uint32_t; };
template <> struct Rec<8> { using T = uint2; };
template <> struct Rec<32> { struct __align__(32) T { uint32_t v[8]; }; };

template <int W>
__device__ __forceinline__ uint32_t ld_fold(const uint8_t *p);
template <> __device__ __forceinline__ uint32_t ld_fold<4>(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
template <> __device__ __forceinline__ uint32_t ld_fold<8>(const uint8_t *p) { uint2 v = __ldg(reinterpret_cast<const uint2 *>(p)); return v.x ^ v.y; }
template <> __device__ __forceinline__ uint32_t ld_fold<32>(const uint8_t *p) {
    uint32_t r[8];
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}

// every thread: `rounds` rounds of 8 independent loads at pseudo-random record indices (one LCG stream per thread: no two lanes of a warp share a sector on purpose)
template <int W>
__global__ void __launch_bounds__(256) micro_gather_kernel(const uint8_t *__restrict__ table, uint32_t n_records, int rounds, uint32_t *__restrict__ sink) {
    uint32_t s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int r = 0; r < rounds; ++r) {
        uint32_t idx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; idx[k] = (uint32_t)(((uint64_t)(s >> 4) * n_records) >> 28); }
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ld_fold<W>(table + (size_t)idx[k] * W);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k];
    }
    if (acc == 0x12345u) sink[0] = acc;   // keeps the loads alive
}

}  // namespace xrb

using namespace xrb;

extern "C" {

// Launches the random-gather probe: n_records records of record_bytes (4, 8 or 32) each, `rounds` rounds of 8 loads per thread, SMs x 8 CTAs x 256 threads.
// Returns in *loads_issued the number of loads the launch performs (time it with events on `stream`).
int xrb_micro_gather(const void *table, int64_t n_records, int record_bytes, int rounds, int64_t *loads_issued, void *sink, void *stream) {
    XRB_REQUIRE(table && sink && n_records > 0 && n_records < (1ll << 28) && rounds > 0, "micro_gather: bad arguments");
    XRB_REQUIRE(record_bytes == 4 || record_bytes == 8 || record_bytes == 32, "micro_gather: record_bytes must be 4, 8 or 32");
    XRB_REQUIRE(((uintptr_t)table & 31) == 0, "micro_gather: table must be 32-byte aligned");
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    cudaStream_t s = (cudaStream_t)stream;
    if (record_bytes == 4) micro_gather_kernel<4><<<grid, 256, 0, s>>>((const uint8_t *)table, (uint32_t)n_records, rounds, (uint32_t *)sink);
    else if (record_bytes == 8) micro_gather_kernel<8><<<grid, 256, 0, s>>>((const uint8_t *)table, (uint32_t)n_records, rounds, (uint32_t *)sink);
    else micro_gather_kernel<32><<<grid, 256, 0, s>>>((const uint8_t *)table, (uint32_t)n_records, rounds, (uint32_t *)sink);
    if (loads_issued) *loads_issued = (int64_t)grid * 256 * 8 * rounds;
    return check_launch("micro_gather");
}

}  // extern "C"
