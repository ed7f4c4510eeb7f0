// xrnerf_b200 — NerfMLP.run_mlp (/root/reference/xrnerf/models/mlps/nerf_mlp.py:70-94) on tcgen05, v3: TWO 128-row tiles in flight per SM.
//
// v2 (nerf_mlp_tc.cu) keeps one tile per CTA: every layer is MMA -> TMEM drain -> next MMA on the same rows, so the tensor pipe idles while
// the epilogue converts the accumulators (ncu: tensor pipe 22 % active, profiles/r01b_nerf_mlp_tc2_ncu.md). v3 runs two independent tile
// pipelines p = 0,1 on one SM; while pipeline 0 drains layer l, the tensor core works on pipeline 1's layer:
//
//   warps 0-15   four compute warpgroups WG(p, c): tile p, output-column half c. thread == row == TMEM lane. Per layer: tcgen05.ld 32 columns
//                at a time, + bias (packed half2), ReLU, fp16, STS.128 into the tile's swizzled H blocks (the next layer's A operand);
//   warps 16,17  producer of pipeline 0 / 1 (one lane): streams the layer weights as 16 KB slabs ([256 x 32] K-halves for the 256-wide layers) through the
//                weight ring — by default ONE ring of 4 slots shared by both pipelines and consumed in the global order P0.layer, P1.layer, ... — and
//                TMA-loads the tile's encoding blocks;
//   warps 18,19  MMA issuer of pipeline 0 / 1 (converged warp, one elected lane): per slab 2 x tcgen05.mma (M=128, N=256, K=16) (4 x N=128|16 for the
//                last two layers) into the pipeline's 256 TMEM columns; everything a layer reads is complete before its epilogue rewrites H.
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
#include <utility>

namespace xrb {

constexpr int N3_THREADS = 640;
constexpr int N3_LAYERS = 11;
constexpr int N3_MAX_KB = 6;
constexpr uint32_t N3_BLOCK = 16384;                 // one [128 x 64] fp16 operand block / ring slot
constexpr uint32_t N3_PIPE_A = 5 * N3_BLOCK;         // H0..H3 + AUX
constexpr int N3_RING = 2;
constexpr int N3_AUX = 4;                            // A block index of AUX
constexpr int N3_BIAS_LAYERS = 9 * 256 + 128 + 16;   // per-layer biases
constexpr int N3_BIAS_TOTAL = N3_BIAS_LAYERS + 257;  // + Wa[256] + ba
constexpr int N3_BIAS_H_OFF = (N3_BIAS_TOTAL + 7) & ~7;   // the fp16 copy of the per-layer biases follows the fp32 vector

// The layer schedule is a COMPILE-TIME table (template <MIP>, fully unrolled layer / K-block loops): the first v3 read it from a kernel-parameter
// struct with dynamic indices, and the single issuing thread spent 3.7 K cycles per layer in dependent indexed constant loads (LDC ~120 cycles
// each, timeline in profiles/r01c_nerf_mlp_tc3.md) — more than the 2.1 K cycles it takes to issue the layer's 32 MMAs.
struct N3L {
    int n_kb, N, n_halves, relu, alpha_dot, bias_off, commit_h3;
    int src[N3_MAX_KB];      // A block per K-block: 0-3 = H0..H3, 4 = AUX
    int wait_enc[N3_MAX_KB]; // issuer waits before K-block kb: 0 none, 1 = tile's first point block in AUX (+ second in H3, Mip), 2 = direction block, 3 = second point block in AUX
    int reload[N3_MAX_KB];   // AUX is dead after K-block kb; the producer refills it with: 0 nothing, 1 = direction block, 2 = second point block, 3 = next tile's first point block
};
template <bool MIP>
__host__ __device__ constexpr N3L n3_layer(int l) {
    N3L L{};
    L.N = 256; L.n_halves = 2; L.relu = 1;
    L.bias_off = l <= 9 ? 256 * l : 9 * 256 + 128;
    L.n_kb = 4;
    for (int k = 0; k < 4; ++k) L.src[k] = k;
    if (l == 0) {                                            // pts_linears.0
        L.src[0] = N3_AUX; L.wait_enc[0] = 1;
        if (MIP) { L.n_kb = 2; L.src[1] = 3; } else L.n_kb = 1;
    } else if (l == 5) {                                     // pts_linears.5 on cat([pts, h]): AUX first so that its refill starts early
        L.src[0] = N3_AUX; for (int k = 0; k < 4; ++k) L.src[1 + k] = k;
        if (MIP) { L.n_kb = 6; L.src[5] = N3_AUX; L.reload[0] = 2; L.wait_enc[5] = 3; L.reload[5] = 1; } else { L.n_kb = 5; L.reload[0] = 1; }
    } else if (l == 7) {                                     // pts_linears.7 (+ alpha dot in its epilogue)
        L.alpha_dot = 1;
    } else if (l == 8) {                                     // feature_linear
        L.relu = 0;
    } else if (l == 9) {                                     // views_linears.0 on cat([feature, dirs])
        L.N = 128; L.n_halves = 1; L.n_kb = 5; L.src[0] = N3_AUX; for (int k = 0; k < 4; ++k) L.src[1 + k] = k;
        L.wait_enc[0] = 2; L.reload[0] = 3; L.commit_h3 = 1;
    } else if (l == 10) {                                    // rgb_linear
        L.N = 16; L.n_halves = 1; L.relu = 0; L.n_kb = 2;
    }
    return L;
}

// per-layer tables packed into immediates: the K-block loops stay ROLLED (the fully unrolled issuer was 110 KB of SASS walked once per tile by a
// single thread) and index the schedule with shifts instead of loads
__host__ __device__ constexpr uint32_t n3_pack(const int *v, int bits) { uint32_t r = 0; for (int k = 0; k < N3_MAX_KB; ++k) r |= (uint32_t)v[k] << (bits * k); return r; }
__device__ __forceinline__ uint64_t n3_desc(uint32_t lo) { return ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (uint64_t)lo; }   // tc::smem_desc_sw128 split: hi word constant
__device__ __forceinline__ uint64_t n3_desc64(uint32_t lo) { return ((uint64_t)((512u >> 4) | (1u << 14) | (4u << 29)) << 32) | (uint64_t)lo; }   // K-major SWIZZLE_64B: 8-row atoms of 64-byte rows, 512 B apart
__device__ __forceinline__ uint32_t n3_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }

// ---- CTA-pair mode (template CL): the two CTAs of a cluster walk identical schedules; rank 0's producers load every weight slab ONCE and the TMA
// multicasts it into both CTAs' rings (same offsets), halving the weight bytes read out of L2 — the measured ceiling of the single-CTA version (8.6 TB/s).
__device__ __forceinline__ uint32_t n3_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void n3_cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void n3_tma_bulk_g2s_mc(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void n3_commit_mc(uint64_t *bar, uint16_t mask) {   // arrives on the barrier at this offset in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

__device__ long long n3_trace_buf[8][16];   // developer timeline (dbg bit4): [event][layer] clock64 of block 0 / pipeline 0 / its 3rd tile
__device__ __forceinline__ void n3_bar_arrive(uint32_t id, uint32_t n_threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }
__device__ __forceinline__ void n3_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory"); }

// barrier indices inside a pipeline's block of 16
enum { B_FULL = 0, B_EMPTY = 2, B_ACC = 5, B_E0 = 6, B_E1 = 7, B_E2 = 8, B_E3 = 9, B_AUXFREE = 10, B_H3FREE = 11, B_LEMPTY = 12, B_PER_PIPE = 16 };   // B_LEMPTY[2]: CTA-pair mode, rank 1: my own slot release (re-arms FULL)

struct N3Ctx {            // per-role constants of one pipeline
    uint64_t *b; uint64_t *bars0; uint8_t *A; uint8_t *ring; uint8_t *ring0; int p, dbg, crank; uint32_t tmem_p; volatile uint32_t *busy;   // busy[p]: issuer p is inside a layer's issue phase
};

// ---------------------------------------------------------------------------------------------------- producer (one lane)
struct N3Prod { uint32_t it, af; int pending, age; size_t off; };
// Ring barriers of the SHARED ring (MD == 2): 4 slots = the two pipelines' private slot pairs, back to back; slot s uses pipeline (s>>1)'s barrier pair (s&1).
__device__ __forceinline__ uint64_t *n3_sr_full(const N3Ctx &c, uint32_t slot) { return c.bars0 + (slot >> 1) * B_PER_PIPE + B_FULL + (slot & 1u); }
__device__ __forceinline__ uint64_t *n3_sr_empty(const N3Ctx &c, uint32_t slot) { return c.bars0 + (slot >> 1) * B_PER_PIPE + B_EMPTY + (slot & 1u); }

// MD == 0: private 2-slot rings. MD == 1: CTA pairs, rank 0 multicasts every slab into both CTAs' private rings. MD == 2: ONE 4-slot ring consumed in the
// global order  P0.layer0, P1.layer0, P0.layer1, ...  (every pipeline walks the same number of rounds, phantom tiles at the ragged end): the pipeline whose
// turn it is has 64 KB of weights in flight instead of 32 — measured: alone on the tensor pipe a pipeline's MMA phase is 3.9 K cycles with the weight TMA and
// 2.6 K without (2-slot ring x ~900-cycle refill) — and the turn order makes the two pipelines alternate instead of locking in phase.
template <bool MIP, int MD, int LX>
__device__ __forceinline__ void n3_produce_layer(const N3Ctx &c, N3Prod &st, const uint8_t *__restrict__ weight_image, const uint8_t *enc, const uint8_t *enc_next) {
    constexpr N3L L = n3_layer<MIP>(LX);
    constexpr bool CL = MD == 1, SR = MD == 2;
    constexpr uint32_t bytes = (uint32_t)(L.N / L.n_halves) * 128u;
    constexpr int aux_blocks = MIP ? 2 : 1;
    constexpr uint32_t RL = n3_pack(L.reload, 2);
    constexpr uint32_t cnt = (uint32_t)(L.n_kb * L.n_halves);
    uint32_t sr_it = 0;
    if (SR) {
        sr_it = st.it + (c.p ? cnt : 0u);                  // my slabs in the global order
        st.it += 2 * cnt;
        while (c.busy[3] != sr_it) {}                      // producers push in the global order too (a barrier wait must never be two ring rounds ahead)
    }
#pragma unroll 1
    for (int kb = 0; kb < L.n_kb; ++kb) {
#pragma unroll 1
        for (int h = 0; h < L.n_halves; ++h) {
            if (SR) {
                const uint32_t slot = sr_it % 4u, round = sr_it / 4u;
                if (round > 0) tc::mbar_wait(n3_sr_empty(c, slot), (round - 1) & 1);
                tc::mbar_expect_tx(n3_sr_full(c, slot), bytes);
                tc::tma_bulk_g2s(c.ring0 + (size_t)slot * N3_BLOCK, weight_image + st.off, bytes, n3_sr_full(c, slot));
                ++sr_it;
            } else {
                const uint32_t slot = st.it % N3_RING, round = st.it / N3_RING;
                if (CL) {
                    // rank 0 loads for both CTAs once BOTH have released the slot (EMPTY counts 2: my issuer's commit + the peer's multicast commit);
                    // rank 1 only re-arms its own FULL barrier when its own issuer has released the slot (the bytes arrive from rank 0's multicast)
                    if (round > 0) tc::mbar_wait(c.b + (c.crank ? B_LEMPTY : B_EMPTY) + slot, (round - 1) & 1);
                    tc::mbar_expect_tx(c.b + B_FULL + slot, bytes);
                    if (c.crank == 0) n3_tma_bulk_g2s_mc(c.ring + (size_t)slot * N3_BLOCK, weight_image + st.off, bytes, c.b + B_FULL + slot, (uint16_t)3);
                } else if (!(c.dbg & 8)) {      // (bit3, only with bits 0|1: no ring handshake at all)
                    if (round > 0) tc::mbar_wait(c.b + B_EMPTY + slot, (round - 1) & 1);
                    if (c.dbg & 1) n3_arrive(c.b + B_FULL + slot);
                    else { tc::mbar_expect_tx(c.b + B_FULL + slot, bytes); tc::tma_bulk_g2s(c.ring + (size_t)slot * N3_BLOCK, weight_image + st.off, bytes, c.b + B_FULL + slot); }
                }
                ++st.it;
            }
            st.off += bytes;
        }
        if (st.pending) {
            // AUX refill. MD 0/1: one K-block after the one that last read AUX (its ring slots have been released since, so its commit on AUXFREE, issued right
            // behind the slot release, has landed or is about to). MD 2: two K-blocks after it (4 slots): either way the wait does not stall the weight stream.
            if (!SR || st.age == 1) {
                if (st.pending != 3 || enc_next) {
                    tc::mbar_wait(c.b + B_AUXFREE, st.af & 1);
                    const uint8_t *src = st.pending == 1 ? enc + (size_t)aux_blocks * N3_BLOCK : st.pending == 2 ? enc + N3_BLOCK : enc_next;
                    uint64_t *eb = c.b + (st.pending == 1 ? B_E1 : st.pending == 2 ? B_E2 : B_E0);
                    tc::mbar_expect_tx(eb, N3_BLOCK);
                    tc::tma_bulk_g2s(c.A + N3_AUX * N3_BLOCK, src, N3_BLOCK, eb);
                }
                ++st.af;
                st.pending = 0;
            } else st.age = 1;
        }
        if ((RL >> (2 * kb)) & 3u) { st.pending = (int)((RL >> (2 * kb)) & 3u); st.age = 0; }
    }
    if (SR) c.busy[3] = sr_it;                             // the next producer in the global order may push
}
template <bool MIP, int MD, int... LS>
__device__ __forceinline__ void n3_produce_tile(std::integer_sequence<int, LS...>, const N3Ctx &c, N3Prod &st, const uint8_t *__restrict__ weight_image, const uint8_t *enc, const uint8_t *enc_next) {
    (n3_produce_layer<MIP, MD, LS>(c, st, weight_image, enc, enc_next), ...);
}

// ---------------------------------------------------------------------------------------------------- MMA issuer (whole warp walks, lane 0 issues)
// The issuer warp runs CONVERGED (all 32 lanes walk the schedule and wait on the barriers; one elected lane executes tcgen05.mma / commit): every
// operand is then provably warp-uniform and lives in uniform registers. Issued from `if (lane == 0)` code each MMA cost a 5 x R2UR.BROADCAST waterfall
// loop (~65 cycles per MMA, as long as an N=128 MMA executes), so the single issuing thread, not the tensor pipe, set the pace (timeline r01c).
__device__ __forceinline__ bool n3_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <bool MIP, int MD, int LX>
__device__ __forceinline__ void n3_issue_layer(const N3Ctx &c, uint32_t &it, uint32_t tcount) {
    constexpr N3L L = n3_layer<MIP>(LX);
    constexpr uint32_t idesc = tc::idesc_f16_m128((uint32_t)L.N);
    constexpr uint32_t SRC = n3_pack(L.src, 3), WE = n3_pack(L.wait_enc, 2), RL = n3_pack(L.reload, 2);
    constexpr bool CL = MD == 1, SR = MD == 2;
    constexpr uint32_t cnt = (uint32_t)(L.n_kb * L.n_halves);
    uint32_t sr_it = 0;
    if (SR) { sr_it = it + (c.p ? cnt : 0u); it += 2 * cnt; }      // my slabs in the shared ring's global order: pipeline 0's slabs of this layer, then pipeline 1's
    tc::named_bar_sync(5 + c.p, 288);                             // the layer's input rows are in H, the previous accumulator is drained (hardware barrier)
    const bool tr = (c.dbg & 16) && blockIdx.x == 0 && tcount == 2;     // both pipelines' issuers are traced: events 0/1 (pipeline 0), 4/7 (pipeline 1)
    // Tensor-pipe turn taking. Left alone the two pipelines LOCK IN PHASE (timeline r01c: both issuers start every layer within ~20 cycles of each
    // other, even when pipeline 1 is started 3000 cycles late): their MMA phases then share the pipe half/half (4.1 K cycles = 2 x 2048) and it idles
    // during both epilogues. So an issuer does not start a layer while the other one is issuing; pipeline 1 looks a little later than pipeline 0
    // raises its flag, which breaks the tie. The flag drops when the layer's last MMA has been ISSUED (the pipe still holds ~2 slabs of queued work,
    // so the hand-over leaves no gap). Neither side ever waits while the other cannot make progress: a waiter only waits on an issuing pipeline.
    if (c.dbg & 32) {   // measured: 2.68 ms with turn taking vs 2.35 ms without — the MMA phase is 4.1 K cycles even ALONE on the pipe (shared-memory bytes), so it stays an experiment
        if (c.p == 1) __nanosleep(64);
        else if (n3_elect_one()) c.busy[0] = 1;
        __syncwarp();
        while (c.busy[c.p ^ 1]) {}
        if (c.p == 1 && n3_elect_one()) c.busy[1] = 1;
        __syncwarp();
    }
    if (SR) { while (c.busy[2] != sr_it) {} }   // my turn: every earlier slab of the global order has been issued (and no parity wait can be two ring rounds ahead)
    if (tr && n3_elect_one()) n3_trace_buf[c.p ? 4 : 0][LX] = clock64();
    tc::tc_fence_after_sync();
    const uint32_t a_lo = n3_desc_lo(tc::smem_u32(c.A)), b_lo = n3_desc_lo(tc::smem_u32(SR ? c.ring0 : c.ring));
#pragma unroll 1
    for (int kb = 0; kb < L.n_kb; ++kb) {
        const uint32_t we = (WE >> (2 * kb)) & 3u;
        if (we) {
            if (we == 1) { tc::mbar_wait(c.b + B_E0, tcount & 1); if (MIP) tc::mbar_wait(c.b + B_E3, tcount & 1); }
            else tc::mbar_wait(c.b + (we == 2 ? B_E1 : B_E2), tcount & 1);
            tc::tc_fence_after_sync();
        }
        const uint32_t a0 = a_lo + ((SRC >> (3 * kb)) & 7u) * (N3_BLOCK >> 4);
#pragma unroll
        for (int h = 0; h < L.n_halves; ++h) {
            const uint32_t slot = SR ? sr_it % 4u : it % N3_RING, round = SR ? sr_it / 4u : it / N3_RING;
            if (SR) ++sr_it; else ++it;
            uint64_t *full = SR ? n3_sr_full(c, slot) : c.b + B_FULL + slot, *empty = SR ? n3_sr_empty(c, slot) : c.b + B_EMPTY + slot;
            if (!SR && (c.dbg & 8)) continue;
            tc::mbar_wait(full, round & 1);
            tc::tc_fence_after_sync();
            const uint32_t b0 = b_lo + slot * (N3_BLOCK >> 4);
            if (n3_elect_one()) {
                if (!(c.dbg & 2)) {
                    if (L.n_halves == 2) {
                        // N=256 layers: the slab is a K-HALF of the K-block, [256 x 32] fp16 in the 64-byte-swizzle layout: two N=256 MMAs per slab read every
                        // A element once (N=128 halves read A twice: 448 KB of shared-memory traffic per tile-layer against a 128 B/clk port = the 3.5 K-cycle
                        // tile-layers of the r01c timeline; this layout needs 384 KB)
#pragma unroll
                        for (int k = 0; k < 2; ++k) tc::mma_f16_ss(c.tmem_p, n3_desc(a0 + 2 * (2 * h + k)), n3_desc64(b0 + 2 * k), idesc, (kb | h | k) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc::mma_f16_ss(c.tmem_p, n3_desc(a0 + 2 * k), n3_desc(b0 + 2 * k), idesc, (kb | k) ? 1u : 0u);   // +32 bytes per K step of 16
                    }
                }
                if (CL && c.crank) { tc::mma_commit(c.b + B_LEMPTY + slot); n3_commit_mc(c.b + B_EMPTY + slot, (uint16_t)1); }   // my producer may re-arm; rank 0 may reload
                else tc::mma_commit(empty);
            }
            __syncwarp();
        }
        if (((RL >> (2 * kb)) & 3u) && n3_elect_one()) tc::mma_commit(c.b + B_AUXFREE);       // every MMA that reads AUX's current content has been issued
        __syncwarp();
    }
    if (n3_elect_one()) {
        if (L.commit_h3) tc::mma_commit(c.b + B_H3FREE);
        tc::mma_commit(c.b + B_ACC);                                  // accumulator of layer LX complete, H/AUX reads of the layer done
        c.busy[c.p] = 0;
        if (SR) c.busy[2] = sr_it;                                    // the next consumer in the global order may start
        if (tr) n3_trace_buf[c.p ? 7 : 1][LX] = clock64();
    }
    __syncwarp();
}
template <bool MIP, int MD, int... LS>
__device__ __forceinline__ void n3_issue_tile(std::integer_sequence<int, LS...>, const N3Ctx &c, uint32_t &it, uint32_t tcount) {
    (n3_issue_layer<MIP, MD, LS>(c, it, tcount), ...);
}

// ---------------------------------------------------------------------------------------------------- compute warpgroups
struct N3Comp {
    uint32_t acc_phase, taddr, r7, row_off, row; int cc, warp, lane; float alpha_acc; float *apart;
    const float *bias_g; float *raw; int64_t i; bool valid, tr;
};
template <bool MIP, int LX>
__device__ __forceinline__ void n3_epilogue_layer(const N3Ctx &c, N3Comp &s) {
    constexpr N3L L = n3_layer<MIP>(LX);
    constexpr bool last = LX == N3_LAYERS - 1;
    // ONE thread per pipeline polls the accumulator mbarrier (tcgen05.commit can only signal an mbarrier); the other 255 block in a hardware named
    // barrier. 16 polling lanes executed 4.5 M try_waits per SM and saturated the pipe that also converts fp32 -> fp16 (ncu: xu 110 %).
    if ((s.warp & 7) == 0 && s.lane == 0) { tc::mbar_wait(c.b + B_ACC, s.acc_phase); if (s.tr) n3_trace_buf[2][LX] = clock64(); }
    __syncwarp();
    tc::named_bar_sync(3 + c.p, 256);
    if (s.tr && threadIdx.x == 1) n3_trace_buf[3][LX] = clock64();
    tc::tc_fence_after_sync();
    s.acc_phase ^= 1;
    if (!last) {
        constexpr int cols = L.N >> 1;
        const int col0 = s.cc * cols;
        const __half2 zero2 = __float2half2_rn(0.f);
        if (!(c.dbg & 4)) {
#pragma unroll 1
            for (int ck = 0; ck < cols / 32; ++ck) {
                const int colb = col0 + ck * 32;
                uint32_t r[32];
                tc::tmem_ld32(s.taddr + colb, r);
                const uint4 *bh = reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(s.bias_g + N3_BIAS_H_OFF) + L.bias_off + colb);
                uint8_t *dst = c.A + (size_t)(colb >> 6) * N3_BLOCK + s.row_off;
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
                    *reinterpret_cast<uint4 *>(dst + (((cb + (uint32_t)q) ^ s.r7) << 4)) = *reinterpret_cast<uint4 *>(h);
                }
            }
        }
        if (L.alpha_dot) {   // alpha_linear on the fp16 output of pts_linears.7: partial dot over my columns, read back from my own row of H
#pragma unroll 1
            for (int ck = 0; ck < cols / 32; ++ck) {
                const int colb = col0 + ck * 32;
                const uint8_t *src = c.A + (size_t)(colb >> 6) * N3_BLOCK + s.row_off;
                const uint32_t cb = (uint32_t)(colb & 63) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 hv = *reinterpret_cast<const uint4 *>(src + (((cb + (uint32_t)q) ^ s.r7) << 4));
                    const __half2 *h = reinterpret_cast<const __half2 *>(&hv);
                    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(s.bias_g + N3_BIAS_LAYERS + colb + 8 * q)), w1 = __ldg(reinterpret_cast<const float4 *>(s.bias_g + N3_BIAS_LAYERS + colb + 8 * q + 4));
                    float2 f;
                    f = __half22float2(h[0]); s.alpha_acc = fmaf(f.x, w0.x, s.alpha_acc); s.alpha_acc = fmaf(f.y, w0.y, s.alpha_acc);
                    f = __half22float2(h[1]); s.alpha_acc = fmaf(f.x, w0.z, s.alpha_acc); s.alpha_acc = fmaf(f.y, w0.w, s.alpha_acc);
                    f = __half22float2(h[2]); s.alpha_acc = fmaf(f.x, w1.x, s.alpha_acc); s.alpha_acc = fmaf(f.y, w1.y, s.alpha_acc);
                    f = __half22float2(h[3]); s.alpha_acc = fmaf(f.x, w1.z, s.alpha_acc); s.alpha_acc = fmaf(f.y, w1.w, s.alpha_acc);
                }
            }
            if (s.cc == 1) s.apart[s.row] = s.alpha_acc;
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before_sync();
        __syncwarp();
        if (s.tr && (threadIdx.x == 1 || threadIdx.x == 224)) n3_trace_buf[threadIdx.x == 1 ? 5 : 6][LX] = clock64();
        n3_bar_arrive(5 + c.p, 288);
    } else {
        if (s.cc == 0) {
            float o16[16];
            tc::tmem_ld16(s.taddr, o16);
            tc::named_bar_sync(1 + c.p, 256);    // WG(p,1)'s partial alpha is in shared memory
            const float *bl = s.bias_g + L.bias_off;
            const float alpha = s.alpha_acc + s.apart[s.row] + __ldg(s.bias_g + N3_BIAS_TOTAL - 1);
            if (s.valid) reinterpret_cast<float4 *>(s.raw)[s.i] = make_float4(o16[0] + __ldg(bl), o16[1] + __ldg(bl + 1), o16[2] + __ldg(bl + 2), alpha);
        } else {
            tc::named_bar_sync(1 + c.p, 256);
        }
    }
}
template <bool MIP, int... LS>
__device__ __forceinline__ void n3_epilogue_tile(std::integer_sequence<int, LS...>, const N3Ctx &c, N3Comp &s) {
    (n3_epilogue_layer<MIP, LS>(c, s), ...);
}

template <bool MIP, int MD>
__global__ void __launch_bounds__(N3_THREADS, 1) nerf_mlp_tc3_kernel(int dbg, int stagger, const uint8_t *__restrict__ weight_image, const float *__restrict__ bias_g, const uint8_t *__restrict__ enc_image,
                                                                      int64_t n_rows, float *__restrict__ raw) {
    extern __shared__ uint8_t dyn_smem[];
    uint8_t *base = (uint8_t *)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023);
    uint8_t *ring_base = base + 2 * N3_PIPE_A;
    float *alpha_part = (float *)(ring_base + 2 * N3_RING * N3_BLOCK);          // [2][128]
    uint64_t *bars = (uint64_t *)(alpha_part + 256);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * B_PER_PIPE);
    volatile uint32_t *busy = tmem_slot + 1;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform by construction (uniform registers)
    using Layers = std::make_integer_sequence<int, N3_LAYERS>;

    if (threadIdx.x == 0) {
        busy[0] = 0; busy[1] = 0; busy[2] = 0; busy[3] = 0;
        for (int p = 0; p < 2; ++p) {
            uint64_t *b = bars + p * B_PER_PIPE;
            for (int s = 0; s < N3_RING; ++s) { tc::mbar_init(b + B_FULL + s, 1); tc::mbar_init(b + B_EMPTY + s, MD == 1 ? 2 : 1); tc::mbar_init(b + B_LEMPTY + s, 1); }
            tc::mbar_init(b + B_ACC, 1);
            tc::mbar_init(b + B_E0, 1); tc::mbar_init(b + B_E1, 1); tc::mbar_init(b + B_E2, 1); tc::mbar_init(b + B_E3, 1);
            tc::mbar_init(b + B_AUXFREE, 1); tc::mbar_init(b + B_H3FREE, 1);
        }
        tc::fence_mbar_init();
    }
    if (warp == 16) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before_sync();
    __syncthreads();
    if (MD == 1) n3_cluster_sync();                                   // the peer's barriers are initialised before anything of mine can signal them
    tc::tc_fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int64_t n_tiles = (n_rows + 127) / 128;
    constexpr uint32_t enc_tile_bytes = (uint32_t)((MIP ? 2 : 1) + 1) * N3_BLOCK;

    N3Ctx c;
    c.p = warp >= 16 ? (warp & 1) : (warp >> 3);                 // warps 16,18 / 0-7 -> pipeline 0; 17,19 / 8-15 -> pipeline 1
    c.b = bars + c.p * B_PER_PIPE; c.A = base + (size_t)c.p * N3_PIPE_A; c.ring = ring_base + (size_t)c.p * N3_RING * N3_BLOCK; c.dbg = dbg;
    c.tmem_p = tmem + (uint32_t)c.p * 256u; c.busy = busy; c.crank = MD == 1 ? (int)n3_cluster_rank() : 0; c.bars0 = bars; c.ring0 = ring_base;
    // tile of (round r, CTA, pipeline p) = (r * gridDim.x + blockIdx.x) * 2 + p. EVERY pipeline of the grid walks the same number of rounds; a tile index
    // past the end is a PHANTOM tile: it is computed on the last tile's encodings and writes nothing. (The two CTAs of a pair must consume the multicast
    // weight stream in lock step, so neither may stop early; without clusters the phantom work is at most one tile per pipeline.)
    const int64_t vcta = (int64_t)blockIdx.x * 2 + c.p, vstride = (int64_t)gridDim.x * 2;
    const int64_t rounds = (n_tiles + vstride - 1) / vstride;

    if (warp >= 18) {
        // ===================================================== MMA issuer of pipeline p
        // Both pipelines start together and would stay IN PHASE (sharing the tensor pipe half/half during their MMA phases and leaving it idle during
        // both epilogues: the r01c timeline shows 4.1 K-cycle MMA phases = 2 x 2048). The phase offset between them is neutrally stable, so pipeline 1
        // simply starts half a layer period late and the two alternate: one drains its accumulators while the other one's MMAs run.
        if (c.p == 1) { const long long t0 = clock64(); while (clock64() - t0 < (long long)stagger) {} }
        uint32_t it = 0, tcount = 0;
        for (int64_t r = 0; r < rounds; ++r, ++tcount) n3_issue_tile<MIP, MD>(Layers{}, c, it, tcount);
    } else if (warp >= 16) {
        // ===================================================== producer of pipeline p
        if (lane == 0) {
            N3Prod st{};
            uint32_t n = 0;
            for (int64_t r = 0; r < rounds; ++r, ++n) {
                const int64_t tile = min(vcta + r * vstride, n_tiles - 1), tile_next = min(vcta + (r + 1) * vstride, n_tiles - 1);   // phantom tiles re-read the last tile's encodings
                const uint8_t *enc = enc_image + (size_t)tile * enc_tile_bytes;
                const uint8_t *enc_next = r + 1 < rounds ? enc_image + (size_t)tile_next * enc_tile_bytes : nullptr;
                if (n == 0) { tc::mbar_expect_tx(c.b + B_E0, N3_BLOCK); tc::tma_bulk_g2s(c.A + N3_AUX * N3_BLOCK, enc, N3_BLOCK, c.b + B_E0); }   // later tiles: loaded behind the previous tile's direction block
                if (MIP) {
                    if (n > 0) tc::mbar_wait(c.b + B_H3FREE, (n - 1) & 1);          // views_linears.0 of the previous tile has read the feature block in H3
                    tc::mbar_expect_tx(c.b + B_E3, N3_BLOCK);
                    tc::tma_bulk_g2s(c.A + 3 * N3_BLOCK, enc + N3_BLOCK, N3_BLOCK, c.b + B_E3);
                }
                st.off = 0; st.pending = 0;
                n3_produce_tile<MIP, MD>(Layers{}, c, st, weight_image, enc, enc_next);
            }
        }
    } else {
        // ===================================================== compute warpgroups WG(p, cc): thread == row, cc == column half
        N3Comp s;
        s.cc = (warp >> 2) & 1; s.warp = warp; s.lane = lane;
        s.apart = alpha_part + c.p * 128;
        s.row = threadIdx.x & 127;
        s.taddr = c.tmem_p + (((uint32_t)(warp & 3) * 32u) << 16);
        s.r7 = s.row & 7u; s.row_off = (s.row >> 3) * 1024u + s.r7 * 128u;
        s.bias_g = bias_g; s.raw = raw;
        s.acc_phase = 0;
        uint32_t tcnt = 0;
        for (int64_t r = 0; r < rounds; ++r, ++tcnt) {
            const int64_t tile = vcta + r * vstride;
            s.tr = (dbg & 16) && blockIdx.x == 0 && c.p == 0 && tcnt == 2;
            s.i = tile * 128 + s.row;
            s.valid = s.i < n_rows;                                  // false for every row of a phantom tile
            tc::tc_fence_before_sync();
            __syncwarp();
            n3_bar_arrive(5 + c.p, 288);                             // my reads of the previous tile's accumulators are done (hardware barrier towards the issuer warp)
            s.alpha_acc = 0.f;
            n3_epilogue_tile<MIP>(Layers{}, c, s);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (MD == 1) n3_cluster_sync();                                   // nobody leaves while the peer may still signal one of its barriers
    if (warp == 16) tc::tmem_dealloc<512>(tmem);
}

}  // namespace xrb

using namespace xrb;

extern "C" {

// weight image / bias vector from xrnerf_b200.nerf_mlp.pack_nerf_mlp_v3 (same schedule, mirrored on the host)
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
    const bool mip = input_ch > 64;
    const int dbg = getenv("XRB_NM_DBG") ? atoi(getenv("XRB_NM_DBG")) : 0;   // attribution experiments: bit0 no weight TMA, bit1 no MMAs, bit2 no epilogue math, bit3 (with 0|1) no weight-ring handshake, bit4 timeline, bit5 tensor-pipe turn taking between the two issuers
    const int stagger = getenv("XRB_N3_STAGGER") ? atoi(getenv("XRB_N3_STAGGER")) : 0;   // cycles by which pipeline 1 trails pipeline 0 (see the kernel)
    constexpr size_t smem = 1024 + 2 * (size_t)N3_PIPE_A + 2 * (size_t)N3_RING * N3_BLOCK + 256 * sizeof(float) + 8 * (2 * B_PER_PIPE) + 32;
    static_assert(smem <= 232448, "v3 shared memory budget");
    int dev = 0, sms = NUM_SMS; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (n_rows + 127) / 128, pairs = (n_tiles + 1) / 2;
    int grid = (int)(pairs < sms ? pairs : sms);
    const bool cluster = getenv("XRB_N3_CLUSTER") ? atoi(getenv("XRB_N3_CLUSTER")) != 0 : false;   // CTA pairs with multicast weight loads (see the kernel)
    if (cluster) {
        grid = (grid + 1) & ~1;                                   // whole pairs; a CTA without real tiles walks phantom tiles
        if (grid > sms) grid = sms & ~1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(N3_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const uint8_t *wi = (const uint8_t *)weight_image, *ei = (const uint8_t *)enc_image;
        cudaError_t e;
        if (mip) {
            cudaFuncSetAttribute(nerf_mlp_tc3_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e = cudaLaunchKernelEx(&cfg, nerf_mlp_tc3_kernel<true, 1>, dbg, stagger, wi, bias, ei, n_rows, raw);
        } else {
            cudaFuncSetAttribute(nerf_mlp_tc3_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e = cudaLaunchKernelEx(&cfg, nerf_mlp_tc3_kernel<false, 1>, dbg, stagger, wi, bias, ei, n_rows, raw);
        }
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return -100; }
    } else {
        const int shared_ring = getenv("XRB_N3_SHARED_RING") ? atoi(getenv("XRB_N3_SHARED_RING")) : 1;   // MD 2, see n3_produce_layer
        auto launch = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, N3_THREADS, smem, (cudaStream_t)stream>>>(dbg, stagger, (const uint8_t *)weight_image, bias, (const uint8_t *)enc_image, n_rows, raw);
        };
        if (mip) { if (shared_ring) launch(nerf_mlp_tc3_kernel<true, 2>); else launch(nerf_mlp_tc3_kernel<true, 0>); }
        else { if (shared_ring) launch(nerf_mlp_tc3_kernel<false, 2>); else launch(nerf_mlp_tc3_kernel<false, 0>); }
    }
    return check_launch("nerf_mlp_forward_v3");
}

// developer tool (scripts/probe_v3.py): copies the timeline written under XRB_NM_DBG bit4 to the host (8 x 16 int64)
int xrb_internal_n3_trace(long long *out_host) {
    return cudaMemcpyFromSymbol(out_host, n3_trace_buf, sizeof(long long) * 8 * 16) == cudaSuccess ? XRB_OK : -100;
}

}  // extern "C"
