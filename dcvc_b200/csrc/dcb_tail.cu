// dcb_tail.cu — dc.3 -> ffn.0 -> ffn.2 (-> next dc.0) of a DepthConvBlock as one persistent CTA-pair kernel
// (see dcb_tail.cuh for the what and why).  Hand-written for sm_100a: TMA, tcgen05.mma cta_group::2 with the A operand
// from shared memory (phases 1, 3) or from tensor memory (phases 2, 4), tcgen05.ld / tcgen05.st epilogues.
//
// Per CTA (one SM), 19 warps:
//   warp 0        TMA producer: the tile's t2 k-blocks into the resident buffer P, then the weight k-blocks of all four
//                 GEMMs, in consumption order, through two rings of kbs x 8 KB stages (this CTA's 64 of a chunk's 128
//                 columns; one ring per MMA issuer)
//   warps 1, 2    warp 1 allocates TMEM; in the leader CTA (cluster rank 0) one thread of each issues the tcgen05.mma
//                 of the even / odd chunks (accumulator buffer 0 / 1, weight ring 0 / 1)
//   warps 3..18   epilogue: warp (b, h, q) drains lane quarter q / column half h of the chunks that land in accumulator
//                 buffer b (chunks alternate between the two 128-column buffers)
// TMEM (512 columns): [0, C/2) = O: the tile's o, later y, as packed fp16 — the A operand of phases 2 and 4 and the
//                     residual of phase 3;   [256, 384) and [384, 512) = the two fp32 accumulator buffers.
// smem: P [inner/64 x 16 KB] (t2, later t1', UMMA K-major SWIZZLE_128B) | weight rings | 16 x 2 KB store slabs | barriers.
//
// Ordering between the phases needs no grid-wide or CTA-wide barrier: every hand-over is an mbarrier, and "all earlier
// MMAs have completed" is implied by the tcgen05.commit that publishes a later chunk's accumulator — commits complete in
// issue order per issuing thread, and each issuer waits for the other's last commit of the previous phase (xs barriers)
// before its first MMA of a phase — which is what makes the in-place reuse of P and O safe:
//   P: TMA(t2) -> phase-1 MMAs -> phase-2 epilogue writes t1' -> phase-3 MMAs -> commit(p_empty) -> TMA(next t2)
//   O: phase-1 epilogue writes o -> phase-2 MMAs -> phase-3 epilogue reads o, writes y -> phase-4 MMAs -> next tile
#include "dcb_tail.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "ptx.cuh"
#include "pw_gemm_internal.cuh"

namespace dcvc {

static constexpr int DT_CHUNK_N = 128;                              // GEMM columns per accumulator chunk
static constexpr int DT_KB_BYTES = (DT_CHUNK_N / 2) * BLOCK_K * 2;  // 8 KB: this CTA's half of one weight k-block
static constexpr int DT_MAX_STAGES = 8;                             // (control block: 57 barriers in 496 bytes)
static constexpr int DT_MAX_KB = 8;                                 // K <= 512
static constexpr int DT_ACC_COL0 = 256;
static constexpr int DT_TMEM_COLS = 512;
static constexpr int DT_STAGING = EPI_WARPS * EPI_SLAB_BYTES;       // one 2 KB slab per epilogue warp
static constexpr int DT_THREADS = NUM_THREADS + 32;                 // producer + TWO MMA issuers + 16 epilogue warps
static constexpr int DT_EPI_WARP0 = 3;

// Cross-CTA signals of the pair.  Default: arrive with CTA-scope release + plain try_wait, as CUTLASS's 2-SM pipelines do;
// DCVC_B200_GEMM_DBG bit 16 (A/B measurements) switches back to cluster-scope release / acquire.
#define dt_arrive(addr) do { if (p.dbg & 16) mbar_arrive_cluster(addr); else mbar_arrive_remote(addr); } while (0)
#define dt_wait(bar, parity) do { if (p.dbg & 16) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); } while (0)

// timeline marks (tools/dcb_tail_trace.py): globaltimer into p.trace[slot], CTA `cta` only
__device__ __forceinline__ void dt_mark(const DcbTailParams& p, int cta, int slot)
{
    if (p.trace && static_cast<int>(blockIdx.x) == cta && slot < 2048) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[slot] = t;
    }
}

__global__ void __launch_bounds__(DT_THREADS, 1)   // 19 warps are allocated as 20: 96 registers per thread is the ceiling
dcb_tail_kernel(const __grid_constant__ DcbTailParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* P = smem;
    uint8_t* ring = P + p.p_bytes;
    uint8_t* staging = ring + p.stages * p.stage_bytes;
    uint8_t* ctrl = smem + SMEM_USABLE;
    uint64_t* p_full = reinterpret_cast<uint64_t*>(ctrl);   // [8]  leader: t2 k-block of both CTAs landed
    uint64_t* p_ready = p_full + DT_MAX_KB;                 // [8]  leader: t1' k-block written by both CTAs' epilogues
    uint64_t* b_full = p_ready + DT_MAX_KB;                 // [16] leader: both halves of a weight stage landed
    uint64_t* b_empty = b_full + DT_MAX_STAGES;             // [16] every CTA: ring slot consumed
    uint64_t* acc_full = b_empty + DT_MAX_STAGES;           // [2]  every CTA: accumulator chunk complete
    uint64_t* acc_empty = acc_full + 2;                     // [2]  leader: chunk drained by both CTAs' 16 warps
    uint64_t* p_empty = acc_empty + 2;                      // [1]  every CTA: phase-3 MMAs done, P may be reloaded
    uint64_t* o_ready = p_empty + 1;                        // [8]  leader: k-block of O holds the tile's o (both CTAs)
    uint64_t* y_ready = o_ready + DT_MAX_KB;                // [8]  leader: k-block of O holds the tile's y; its o is not read any more
    uint64_t* xs = y_ready + DT_MAX_KB;                     // [4]  leader: issuer X's MMAs of a phase are complete ([2 X + phase parity])
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ctrl + 496);
    static_assert((4 * DT_MAX_KB + 2 * DT_MAX_STAGES + 9) * 8 <= 496, "control block overflow");

    const int rank = static_cast<int>(cluster_ctarank());
    const int pair = static_cast<int>(blockIdx.x >> 1);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a);
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tm_w[i]);
        tma_prefetch_desc(&p.tm_y);
        tma_prefetch_desc(&p.tm_t);
        for (int i = 0; i < DT_MAX_KB; ++i) {
            mbar_init(&p_full[i], 1);
            mbar_init(&p_ready[i], 32);               // 2 chunks x 8 warps x 2 CTAs
            mbar_init(&o_ready[i], 8);                // 4 warps (lane quarters) x 2 CTAs
            mbar_init(&y_ready[i], 8);
        }
        for (int i = 0; i < DT_MAX_STAGES; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&acc_full[g], 1);
            mbar_init(&acc_empty[g], 16);             // 8 warps x 2 CTAs
        }
        mbar_init(p_empty, 2);                        // both MMA issuers' phase-3 MMAs
        for (int i = 0; i < 4; ++i) mbar_init(&xs[i], 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(tmem_ptr_smem, DT_TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / commit multicast
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();
    // programmatic dependent launch: everything above overlapped the previous kernel's tail.  The weights do not depend on it
    // either, so the producer fills the ring first and only then waits (before it touches t2); everybody else waits here
    if (warp != 0) griddep_wait();

    if (warp == 0) {
        if (!(p.dbg & 8) && elect_one_sync()) {
            // ------------------------------------------------------------ TMA producer (both CTAs)
            auto load_tile = [&](int T, int i) {
                mbar_wait(p_empty, static_cast<uint32_t>((i & 1) ^ 1));
                const int row0 = T * 256 + rank * BLOCK_M;
                for (int kb = 0; kb < p.nkb[0]; ++kb) {
                    if (rank == 0) mbar_expect_tx(&p_full[kb], 2 * A_STAGE_BYTES);
                    tma_load_2d_2sm(P + kb * A_STAGE_BYTES, &p.tm_a, mapa_u32(smem_u32(&p_full[kb]), 0), kb * BLOCK_K, row0);
                }
            };
            // Two rings of stages / 2 slots: the chunks alternate between them like they alternate between the two MMA issuers,
            // so every slot has ONE consumer that sees each of its phases (a consumer that skipped a use of a slot could not tell
            // the phase it waits for from the one two uses back: mbarrier waits know one parity bit)
            const int half = p.stages / 2;
            int sr[2] = { 0, 0 };
            uint32_t bphr[2] = { 0, 0 };
            uint32_t tc = 0;   // chunk counter
            int i = 0;
            int issued = 0;
            int T = pair;
            bool dep_done = false;   // griddepcontrol.wait + the first tile's t2: after the weights that need neither
            // weights that may go out before griddepcontrol.wait: the first chunk of each ring when a ring holds a whole chunk
            // (no slot is reused, so nothing waits for an MMA that itself waits for t2), else one ring's worth.  (Measured: 1, 2
            // or 4 stages ahead make no difference behind a warm L2, 92.1-92.3 us; the depth matters for a cold first launch.)
            const int pre_target = half >= p.nst[0] ? p.nst[0] * (p.nch[0] < 2 ? p.nch[0] : 2) : half;
            // the next tile's t2 goes out once the ring is full of phase-4 weights: by then phase 3 has been issued
            // completely, so the wait for p_empty is short and phase 4 starts on a full ring
            const int total4 = p.nch[3] * p.nst[3];
            const int trigger = total4 < p.stages ? total4 : p.stages;
            for (; T < p.tiles; T += p.num_pairs, ++i) {
                const int Tn = T + p.num_pairs;
                bool next_loaded = false;
                for (int ph = 0; ph < 4; ++ph) {
                    int cnt = 0;
                    const uint32_t tx = static_cast<uint32_t>(2 * p.kbs[ph] * DT_KB_BYTES);
                    for (int n = 0; n < p.nch[ph]; ++n, ++tc) {
                        const int nrow = n * DT_CHUNK_N + rank * (DT_CHUNK_N / 2);
                        const int r = static_cast<int>(tc & 1);
                        for (int st = 0; st < p.nst[ph]; ++st) {
                            const int s = r * half + sr[r];
                            mbar_wait(&b_empty[s], bphr[r] ^ 1);
                            if (rank == 0) mbar_expect_tx(&b_full[s], tx);
                            // a stage = kbs k-blocks of this CTA's 64 weight rows, one 2-D request each (rank-2 boxes stream
                            // through the TMA unit; a rank-3 box {64, 64, kbs} was measured ~25 % slower), ONE barrier round trip
                            const uint32_t bar = mapa_u32(smem_u32(&b_full[s]), 0);
                            for (int j = 0; j < p.kbs[ph]; ++j)
                                tma_load_2d_2sm(ring + s * p.stage_bytes + j * DT_KB_BYTES, &p.tm_w[ph], bar,
                                                (st * p.kbs[ph] + j) * BLOCK_K, nrow);
                            dt_mark(p, 0, 1024 + issued);
                            ++issued;
                            if (!dep_done && issued == pre_target) {
                                griddep_wait();
                                load_tile(T, 0);
                                dep_done = true;
                            }
                            if (++sr[r] == half) { sr[r] = 0; bphr[r] ^= 1; }
                            if (ph == 3 && !next_loaded && ++cnt == trigger) {
                                if (Tn < p.tiles) load_tile(Tn, i + 1);
                                next_loaded = true;
                            }
                        }
                    }
                }
                if (!dep_done) {   // (a tile with fewer stages than the ring: not a shape the planner accepts, but be safe)
                    griddep_wait();
                    load_tile(T, 0);
                    dep_done = true;
                }
                if (!next_loaded && Tn < p.tiles) load_tile(Tn, i + 1);
            }
        } else if (p.dbg & 8) {
            griddep_wait();
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        if (rank == 0 && elect_one_sync()) {
            // ------------------------------------------------------------ MMA issuers (leader CTA only)
            // TWO issuing threads: X = 0 (warp 1) takes the even chunks / accumulator buffer 0, X = 1 (warp 2) the odd ones.
            // One thread was measured at 1.44 us per 24-MMA chunk + 0.28 us between chunks (tensor pipe 39 % active in ncu)
            // while a cta_group::2 N = 128 MMA occupies the pipe for 64 cycles; with two threads the next chunk's MMAs are
            // queued while the first thread waits for its stage / accumulator barriers: 89.7 -> 86.5 us on the P8 tail
            // (same box, DCVC_B200_GEMM_DBG bit 32 = one issuer).
            // Ordering across the two: a tcgen05.commit only covers the MMAs of its own thread, so before a thread issues
            // the first MMA of a phase it waits until the OTHER thread's MMAs of the previous phase are complete (xs
            // barriers, two per thread, alternating by phase).  With that, "accumulator of a phase-(p+1) chunk complete"
            // still implies "every MMA of phase p is done", which the in-place reuse of P and O relies on.
            const int X = warp - 1;
            constexpr uint32_t idesc = make_idesc_f16_f32(BLOCK_M * 2, DT_CHUNK_N);
            const bool do_mma = !(p.dbg & 1);
            const bool no_acc = (p.dbg & 4) != 0;   // timing experiments: no accumulator hand-off (the epilogue warps idle)
            const bool no_tma = (p.dbg & 8) != 0;   // timing experiments: no loads (operands are whatever smem holds)
            const bool solo = (p.dbg & 32) != 0;    // timing experiments: one issuer does everything (the second one idles)
            uint32_t t = 0;
            const int half = p.stages / 2;          // two weight rings, one per chunk parity (see the producer)
            int sr[2] = { 0, 0 };
            uint32_t bphr[2] = { 0, 0 };
            int i = 0;
            uint32_t c = 0;                         // global phase counter
            for (int T = (solo && X == 1) ? p.tiles : pair; T < p.tiles; T += p.num_pairs, ++i) {
                const uint32_t tph = static_cast<uint32_t>(i & 1);
                for (int ph = 0; ph < 4; ++ph, ++c) {
                    if (c > 0 && !solo) {
                        // the other issuer's MMAs of the previous phase (phase counter c - 1)
                        mbar_wait(&xs[2 * (1 - X) + ((c - 1) & 1)], ((c - 1) >> 1) & 1);
                        tcgen05_fence_after();
                    }
                    const bool a_tmem = (ph & 1) != 0;
                    const int kbs = p.kbs[ph];
                    for (int n = 0; n < p.nch[ph]; ++n, ++t) {
                        const int g = static_cast<int>(t & 1);
                        if (!solo && g != X) continue;   // the other issuer's chunk (and the other ring)
                        if (!no_acc) {
                            dt_wait(&acc_empty[g], ((t >> 1) & 1) ^ 1);  // both CTAs drained this buffer
                            tcgen05_fence_after();
                        }
                        const uint32_t acc = tmem_base + DT_ACC_COL0 + g * DT_CHUNK_N;
                        dt_mark(p, 0, 512 + 4 * static_cast<int>(t));
                        const bool first = solo ? (n == 0) : (n < 2);   // this issuer's first chunk of the phase
                        int kb = 0;
                        for (int st = 0; st < p.nst[ph]; ++st) {
                            const int s = g * half + sr[g];
                            if (!no_tma) {
                                mbar_wait(&b_full[s], bphr[g]);
                                tcgen05_fence_after();
                            }
                            if (st == 0) dt_mark(p, 0, 512 + 4 * static_cast<int>(t) + 1);
                            const uint32_t b_addr = smem_u32(ring + s * p.stage_bytes);
                            for (int j = 0; j < kbs; ++j, ++kb) {
                                if (first && ph == 0 && !no_tma) { mbar_wait(&p_full[kb], tph); tcgen05_fence_after(); }
                                if (first && !no_acc) {
                                    // the first chunk of a phase takes its A k-blocks as the epilogue of the previous phase
                                    // finishes them (later chunks find everything there)
                                    if (ph == 1) { dt_wait(&o_ready[kb], tph); tcgen05_fence_after(); }
                                    if (ph == 2) { dt_wait(&p_ready[kb], tph); tcgen05_fence_after(); }
                                    if (ph == 3) { dt_wait(&y_ready[kb], tph); tcgen05_fence_after(); }
                                }
                                const uint64_t b_desc = make_kmajor_sw128_desc(b_addr + j * DT_KB_BYTES);
                                if (do_mma) {
                                    if (a_tmem) {
                                        const uint32_t a_t = tmem_base + kb * (BLOCK_K / 2);
#pragma unroll
                                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                            umma_f16_ts_2cta(acc, a_t + k * (UMMA_K / 2), b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                                    } else {
                                        const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(P + kb * A_STAGE_BYTES));
#pragma unroll
                                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                            umma_f16_ss_2cta(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                                    }
                                }
                            }
                            if (!no_tma) umma_commit_2cta_mc(&b_empty[s], 3);
                            if (++sr[g] == half) { sr[g] = 0; bphr[g] ^= 1; }
                        }
                        if (!no_acc) umma_commit_2cta_mc(&acc_full[g], 3);
                        dt_mark(p, 0, 512 + 4 * static_cast<int>(t) + 2);
                    }
                    // this issuer's MMAs of phase c (possibly none): complete -> xs[2 X + (c & 1)] (leader CTA only)
                    if (!solo) umma_commit_2cta_mc(&xs[2 * X + (c & 1)], 1);
                    if (ph == 2) {
                        umma_commit_2cta_mc(p_empty, 3);  // (one of two arrivals) P is free for the next tile's t2
                        if (solo) umma_commit_2cta_mc(p_empty, 3);
                        // Before the NEXT tile's phase-1 epilogue may overwrite O, the phase-3 epilogue must be through with it
                        // (it reads o and writes y in place).  With a phase 4 that is implied (its first chunk waits for every
                        // y k-block); without one the wait happens here — else next-tile writes race the readers, and barrier
                        // phases nobody consumed would swallow next-tile arrivals (the round-2 "several tiles per pair" bug)
                        if (!no_acc && p.nch[3] == 0) {
                            for (int kb = 0; kb < p.nkb[1]; ++kb) dt_wait(&y_ready[kb], tph);
                            tcgen05_fence_after();
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue (16 warps, both CTAs)
        // Two groups of 8 warps take alternate chunks (group b <- accumulator buffer b), so one group's TMEM drain overlaps
        // the other's arithmetic and signalling.  Warp (b, h, q) owns lane quarter q (32 pixel rows) and column half h of
        // its chunks: 64 accumulator columns = exactly one 64-channel k-block of O (phases 1, 3), drained as two 32-column
        // pieces; the accumulator goes back to the MMA warp right after the second tcgen05.ld.  Whatever a piece needs from
        // memory (bias, x) is requested before the wait it can hide behind.
        const int q = warp & 3;
        const int b = ((warp - DT_EPI_WARP0) >> 2) & 1;
        const int h = (warp - DT_EPI_WARP0) >> 3;
        uint8_t* slab = staging + (warp - DT_EPI_WARP0) * EPI_SLAB_BYTES;
        const uint32_t slab_u = smem_u32(slab);
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t acc = tmem_base + lane_off + DT_ACC_COL0 + b * DT_CHUNK_N + h * 64;   // this warp's 64 accumulator columns
        const uint32_t o_base = tmem_base + lane_off;                                        // O, this warp's lanes
        const uint32_t acc_empty_r = mapa_u32(smem_u32(&acc_empty[b]), 0);
        const uint32_t o_ready_r = mapa_u32(smem_u32(o_ready), 0);
        const uint32_t y_ready_r = mapa_u32(smem_u32(y_ready), 0);
        const uint32_t p_ready_r = mapa_u32(smem_u32(p_ready), 0);
        const uint32_t P_u = smem_u32(P);
        const uint32_t my_sw = static_cast<uint32_t>(lane * 64);
        const uint32_t my_x = static_cast<uint32_t>((lane >> 1) & 3);
        const uint16_t ONE = 0x3C00;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        const bool skip_body = (p.dbg & 2) != 0;
        const bool tr = b == 0 && h == 0 && q == 0 && lane == 0;

        auto hand_back = [&]() {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) dt_arrive(acc_empty_r);
        };
        auto slab_free = [&]() {  // the TMA store that last read this warp's slab has finished reading it
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
        };
        // coalesced fetch of a [32 rows][32 channels] piece of x into the slab (one LDGSTS = 8 rows x 64 B)
        auto fetch_x = [&](int row_base, int c0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t rj = static_cast<uint32_t>((lane >> 2) + 8 * j);
                long long gr = static_cast<long long>(row_base) + rj;
                gr = gr < p.M ? gr : p.M - 1;
                ldgsts16(slab_u + rj * 64 + ((static_cast<uint32_t>(lane & 3) ^ ((rj >> 1) & 3)) << 4),
                         p.x + gr * p.x_pitch + c0 + (lane & 3) * 8);
            }
            cp_async_commit();
        };
        auto add_slab = [&](float (&tv)[32]) {  // tv += this thread's row of the slab
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                const uint4 r = lds128(slab_u + my_sw + ((static_cast<uint32_t>(gq) ^ my_x) << 4));
                const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    tv[gq * 8 + 2 * e] = fma_f32_f16(static_cast<uint16_t>(w1[e] & 0xffffu), ONE, tv[gq * 8 + 2 * e]);
                    tv[gq * 8 + 2 * e + 1] = fma_f32_f16(static_cast<uint16_t>(w1[e] >> 16), ONE, tv[gq * 8 + 2 * e + 1]);
                }
            }
        };
        auto load_bias = [&](const __half* bias, int c0, uint4 (&cb)[4]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) cb[j] = bias ? __ldg(reinterpret_cast<const uint4*>(bias + c0) + j) : zero4;
        };
        auto add_bias = [&](const uint4 (&cb)[4], const uint32_t (&v)[32], float (&tv)[32]) {
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(cb);
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                tv[e] = fma_f32_f16(static_cast<uint16_t>(bw[e >> 1] & 0xffffu), ONE, __uint_as_float(v[e]));
                tv[e + 1] = fma_f32_f16(static_cast<uint16_t>(bw[e >> 1] >> 16), ONE, __uint_as_float(v[e + 1]));
            }
        };
        auto pack16 = [&](const float (&tv)[32], uint32_t (&w)[16]) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __half2 h2 = __floats2half2_rn(tv[2 * j], tv[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
        };
        // this thread's packed row -> its row of the slab -> one TMA store of [32 rows][32 channels]
        auto store_slab = [&](const CUtensorMap* tm, const uint32_t (&w)[16], int c0, int row_base) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
                sts128(slab_u + my_sw + ((static_cast<uint32_t>(gq) ^ my_x) << 4), make_uint4(w[4 * gq], w[4 * gq + 1], w[4 * gq + 2], w[4 * gq + 3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tm, slab, c0, row_base);
                tma_store_commit();
            }
        };
        // x * sigmoid(4 x) up to the factor 1/2: u = x tanh(2 x) + x  (wsilu = u / 2; the caller applies the exact 1/2 once)
        auto wsilu2 = [&](float x) -> float {
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x + x));
            return fmaf(x, th, x);
        };

        uint32_t t = 0;
        for (int T = (p.dbg & 4) ? p.tiles : pair; T < p.tiles; T += p.num_pairs) {
            const int row_w = T * 256 + rank * BLOCK_M + q * 32;  // first pixel row of this warp
            for (int ph = 0; ph < 4; ++ph) {
                const __half* bias = p.bias[ph];
                const bool need_x = (ph == 0) || (ph == 2 && p.shortcut);
                for (int n = 0; n < p.nch[ph]; ++n, ++t) {
                    if (static_cast<int>(t & 1) != b) continue;
                    const int cw = n * DT_CHUNK_N + h * 64;   // GEMM column of this warp's accumulator column 0
                    if (tr) dt_mark(p, 0, 8 * static_cast<int>(t));
                    // ---- requests of the first piece that do not depend on the accumulator
                    uint4 cb[4];
                    load_bias(bias, cw, cb);
                    if (!skip_body && need_x) {
                        slab_free();
                        fetch_x(row_w, cw);
                    }
                    if (tr) dt_mark(p, 0, 8 * static_cast<int>(t) + 5);
                    mbar_wait(&acc_full[b], (t >> 1) & 1);
                    tcgen05_fence_after();
                    if (tr) dt_mark(p, 0, 8 * static_cast<int>(t) + 1);
                    if (skip_body) {
                        hand_back();
                        if (lane == 0) {
                            if (ph == 0) dt_arrive(o_ready_r + (2 * n + h) * 8);
                            if (ph == 1) dt_arrive(p_ready_r + (n >> 1) * 8);
                            if (ph == 2) dt_arrive(y_ready_r + (2 * n + h) * 8);
                        }
                        continue;
                    }
#pragma unroll 1
                    for (int a = 0; a < 2; ++a) {
                        const int c0 = cw + a * 32;
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(acc + a * 32, v);
                        uint32_t ov[16];
                        if (ph == 2) tmem_ld_32x32b_x16(o_base + (c0 >> 1), ov);   // o: the residual of ffn.2
                        tmem_ld_wait();
                        if (a == 1) {
                            hand_back();
                            if (tr) dt_mark(p, 0, 8 * static_cast<int>(t) + 2);
                        }
                        float tv[32];
                        add_bias(cb, v, tv);
                        if (a == 0) load_bias(bias, cw + 32, cb);   // the second piece's bias, behind the first piece's arithmetic
                        if (ph == 0) {
                            // ---------------- o = acc + b3 + x  ->  O (TMEM, packed fp16)
                            cp_async_wait_all();
                            __syncwarp();
                            add_slab(tv);
                            __syncwarp();                          // every lane has read its slab row
                            if (a == 0) fetch_x(row_w, cw + 32);   // the second piece's x, behind the packing + tcgen05.st
                            uint32_t w[16];
                            pack16(tv, w);
                            tmem_st_32x32b_x16(o_base + (c0 >> 1), w);
                        } else if (ph == 1) {
                            // ---------------- t1' = fold4(wsilu(acc + bf0))  ->  P (smem, UMMA K-major SWIZZLE_128B)
                            uint4 o4;
                            uint32_t* ow = reinterpret_cast<uint32_t*>(&o4);
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float o2[2];
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int j = jj * 2 + hh;
                                    float s4 = wsilu2(tv[4 * j]);
                                    s4 += wsilu2(tv[4 * j + 1]);
                                    s4 += wsilu2(tv[4 * j + 2]);
                                    s4 += wsilu2(tv[4 * j + 3]);
                                    o2[hh] = 0.5f * s4;   // exact: the same bits as summing the four halves
                                }
                                const __half2 h2 = __floats2half2_rn(o2[0], o2[1]);
                                ow[jj] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                            // output channels 32 n + 16 h + 8 a .. + 8: k-block n / 2, 16-byte chunk 4 (n & 1) + 2 h + a of the row
                            const uint32_t row = static_cast<uint32_t>(q * 32 + lane);
                            sts128(P_u + (n >> 1) * A_STAGE_BYTES + sw128_offset(row, static_cast<uint32_t>((n & 1) * 4 + h * 2 + a)), o4);
                        } else if (ph == 2) {
                            // ---------------- y = (acc + bf2 + o [+ x]) [* q]  ->  O (in place) and global
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                tv[2 * j] = fma_f32_f16(static_cast<uint16_t>(ov[j] & 0xffffu), ONE, tv[2 * j]);
                                tv[2 * j + 1] = fma_f32_f16(static_cast<uint16_t>(ov[j] >> 16), ONE, tv[2 * j + 1]);
                            }
                            if (p.shortcut) {
                                cp_async_wait_all();
                                __syncwarp();
                                add_slab(tv);
                                __syncwarp();
                            }
                            if (p.qscale) {
                                uint4 cq[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) cq[j] = __ldg(reinterpret_cast<const uint4*>(p.qscale + c0) + j);
                                const uint32_t* qw = reinterpret_cast<const uint32_t*>(cq);
#pragma unroll
                                for (int e = 0; e < 32; e += 2) {
                                    const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qw[e >> 1]));
                                    tv[e] *= qf.x;
                                    tv[e + 1] *= qf.y;
                                }
                            }
                            uint32_t w[16];
                            pack16(tv, w);
                            tmem_st_32x32b_x16(o_base + (c0 >> 1), w);
                            if (!p.shortcut) slab_free();   // (with the shortcut the slab was freed before x was fetched into it)
                            store_slab(&p.tm_y, w, c0, row_w);
                            if (p.shortcut && a == 0) {
                                slab_free();
                                fetch_x(row_w, cw + 32);
                            }
                        } else {
                            // ---------------- t1n = wsilu(acc + b0n)  ->  global
#pragma unroll
                            for (int e = 0; e < 32; ++e) tv[e] = 0.5f * wsilu2(tv[e]);
                            uint32_t w[16];
                            pack16(tv, w);
                            slab_free();
                            store_slab(&p.tm_t, w, c0, row_w);
                        }
                    }
                    if (tr) dt_mark(p, 0, 8 * static_cast<int>(t) + 3);
                    // ---- the chunk's 64 columns are out: tell the MMA warp
                    if (ph == 0 || ph == 2) {
                        tmem_st_wait();
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) dt_arrive((ph == 0 ? o_ready_r : y_ready_r) + (2 * n + h) * 8);
                    } else if (ph == 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) dt_arrive(p_ready_r + (n >> 1) * 8);
                    }
                    if (tr) dt_mark(p, 0, 8 * static_cast<int>(t) + 4);
                }
            }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
    }

    __syncthreads();
    cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer can still signal its barriers / read its smem
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2cta(tmem_base, DT_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host

int dcb_tail_init()
{
    static bool done = false;
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(dcb_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) {
        gemm_set_error(std::string("cudaFuncSetAttribute(dcb_tail): ") + cudaGetErrorString(e));
        return 1;
    }
    done = true;
    return 0;
}

static int dt_max_pairs(int num_sms)
{
    static int cached = 0;
    if (cached) return cached;
    if (dcb_tail_init()) return num_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * 64, 1, 1);
    cfg.blockDim = dim3(DT_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, dcb_tail_kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms / 2;
    }
    cached = n;
    return n;
}

static bool view_ok(const ActView& v)
{
    return v.ptr && (v.pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && v.C <= v.pitch;
}

int dcb_tail_plan(DcbTailOp& op)
{
    op.planned = false;
    static const bool enabled = []() { const char* e = getenv("DCVC_B200_FUSE_TAIL"); return !(e && e[0] == '0'); }();
    if (!enabled) return 1;
    const int C = op.x.C, inner = op.t2.C;
    const int inner_n = op.t1n.ptr ? op.t1n.C : 0;
    const long long M = static_cast<long long>(op.y.W) * op.y.H;
    if (!view_ok(op.t2) || !view_ok(op.x) || !view_ok(op.y) || (op.t1n.ptr && !view_ok(op.t1n))) return 1;
    if (op.y.C != C || C % DT_CHUNK_N || C > 512) return 1;
    if (inner % 64 || inner > 512 || inner < 64) return 1;
    if (inner_n % DT_CHUNK_N || inner_n > 512) return 1;
    const ActView* vs[3] = { &op.t2, &op.x, &op.t1n };
    for (int i = 0; i < (op.t1n.ptr ? 3 : 2); ++i)
        if (static_cast<long long>(vs[i]->W) * vs[i]->H != M) return 1;
    if (M < 1 || M > (1LL << 30)) return 1;
    if (!op.w3 || !op.wf0 || !op.wf2 || (op.t1n.ptr && !op.w0n)) return 1;

    int num_sms = 148;
    {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) num_sms = v;
    }
    DcbTailParams& p = op.p;
    memset(&p, 0, sizeof(p));
    p.p_bytes = inner / 64 * A_STAGE_BYTES;
    p.M = static_cast<int>(M);
    p.C = C; p.inner = inner; p.inner_next = inner_n;
    p.nkb[0] = inner / 64; p.nch[0] = C / DT_CHUNK_N;
    p.nkb[1] = C / 64;     p.nch[1] = 4 * inner / DT_CHUNK_N;
    p.nkb[2] = inner / 64; p.nch[2] = C / DT_CHUNK_N;
    p.nkb[3] = C / 64;     p.nch[3] = inner_n / DT_CHUNK_N;
    p.bias[0] = op.b3; p.bias[1] = op.bf0; p.bias[2] = op.bf2; p.bias[3] = op.b0n;
    p.qscale = op.qscale;
    p.x = static_cast<const __half*>(op.x.ptr);
    p.x_pitch = op.x.pitch;
    p.shortcut = op.shortcut ? 1 : 0;
    p.tiles = static_cast<int>((M + 255) / 256);
    int max_pairs = dt_max_pairs(num_sms);
    if (const char* e = getenv("DCVC_B200_DT_MAXPAIRS")) {  // debugging: many tiles per pair on small problems
        const int v = atoi(e);
        if (v >= 1 && v < max_pairs) max_pairs = v;
    }
    p.num_pairs = p.tiles < max_pairs ? p.tiles : max_pairs;
    if (p.num_pairs < 1) return 1;

    if (encode_act_map(&p.tm_a, op.t2, false, true, true, BLOCK_M, 1)) return 2;
    const __half* ws[4] = { op.w3, op.wf0, op.wf2, op.w0n ? op.w0n : op.w3 };
    const int Ns[4] = { C, 4 * inner, C, inner_n ? inner_n : C };
    const int Ks[4] = { inner, C, inner, inner_n ? C : inner };
    // weight stages: kbs k-blocks of a chunk per ring slot and barrier round trip (measured with one k-block per slot:
    // ~0.2 us per slot whatever the ring depth — the per-slot protocol, not the bytes, set the pace)
    // ... as large as leaves four slots: the chunks alternate between two rings (one per MMA issuer) of stages / 2 slots
    int stages = 0;
    for (int cap = 4; cap >= 1; --cap) {
        int max_kbs = 1;
        for (int ph = 0; ph < 4; ++ph) {
            int kbs = 1;
            for (int c = cap; c >= 1; --c)
                if (p.nkb[ph] % c == 0) { kbs = c; break; }
            p.kbs[ph] = kbs;
            p.nst[ph] = p.nkb[ph] / kbs;
            if (p.nch[ph] > 0 && kbs > max_kbs) max_kbs = kbs;
        }
        p.stage_bytes = max_kbs * DT_KB_BYTES;
        stages = (SMEM_USABLE - p.p_bytes - DT_STAGING) / p.stage_bytes;
        if (stages > DT_MAX_STAGES) stages = DT_MAX_STAGES;
        stages &= ~1;
        if (stages >= 4) break;
    }
    if (stages < 2) return 1;
    p.stages = stages;
    for (int i = 0; i < 4; ++i) {
        uint64_t d2[2] = { static_cast<uint64_t>(Ks[i]), static_cast<uint64_t>(Ns[i]) };
        uint64_t s2[1] = { static_cast<uint64_t>(Ks[i]) * 2 };
        uint32_t b2[2] = { 64, DT_CHUNK_N / 2 };
        if (encode_map(&p.tm_w[i], ws[i], 2, d2, s2, b2)) return 2;
    }
    if (encode_act_map(&p.tm_y, op.y, false, true, true, 32, 1, 32)) return 2;
    if (encode_act_map(&p.tm_t, op.t1n.ptr ? op.t1n : op.y, false, true, true, 32, 1, 32)) return 2;
    if (const char* d = getenv("DCVC_B200_GEMM_DBG")) p.dbg = atoi(d);
    if (const char* d = getenv("DCVC_B200_GEMM_TRACE")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(d, nullptr, 0));
    op.grid = dim3(2 * p.num_pairs, 1, 1);
    op.planned = true;
    return 0;
}

int dcb_tail_launch(const DcbTailOp& op, cudaStream_t stream)
{
    if (!op.planned) { gemm_set_error("dcb_tail_launch: op not planned"); return 1; }
    if (dcb_tail_init()) return 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = op.grid;
    cfg.blockDim = dim3(DT_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (gemm_pdl_enabled() && op.pdl) ? 2 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, dcb_tail_kernel, op.p);
    if (e != cudaSuccess) {
        gemm_set_error(std::string("dcb_tail launch failed: ") + cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

}  // namespace dcvc
