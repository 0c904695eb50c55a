// pw_gemm_epilogue.cuh — the TMEM -> registers -> swizzled smem -> TMA-store epilogue shared by the pw_gemm kernels
// (pw_gemm.cu): bias, WSiLU, 4:1 chunk-add, up to two
// residuals, per-channel quant scale.  One call drains one [128 pixels][BLOCK_N] accumulator tile with the four
// warps of an epilogue group (one warp per TMEM lane quarter).
//
// Sixteen epilogue warps per CTA (four per SM sub-partition hide each other's TMEM / L1 / store latencies; measured
// with ncu on the previous 8-warp version: 27 % issue utilisation, 3.5 us to drain one 128x192 tile).  A warp owns
// 32 accumulator rows x half the columns of a tile, stages 32 output columns at a time in its own 2 KB slab and
// stores it with its own TMA store: no cross-warp barriers, no divergent branches, no integer divisions.
#pragma once
#include "ptx.cuh"
#include "pw_gemm.cuh"

namespace dcvc {


__device__ __forceinline__ float wsilu_f(float x)
{
    // x * sigmoid(4x) = 0.5 x (1 + tanh(2x))   (reference: src/layers/layers.py:106-111); one MUFU op
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(2.f * x));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}

// 16-byte residual load.  Keeps the default L1 allocation on purpose: a thread walks its own row 16 bytes at a
// time, so 7 of 8 loads of a 128-byte line are L1 hits (L1::no_allocate turned them into 8 L2 requests: 40.7 us
// instead of 25.8 us for the M=32640, N=K=384 shortcut GEMM)
__device__ __forceinline__ uint4 ld_stream16(const __half* p)
{
    return *reinterpret_cast<const uint4*>(p);
}

// pull one 16-byte piece (hence its 128-byte line) into L1
__device__ __forceinline__ void l1_touch(const __half* p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// d = a * b + c with fp16 a, b and fp32 c, d in one instruction (SASS FHFMA, .H0/.H1 operand selectors)
__device__ __forceinline__ float fma_f32_f16(uint16_t a, uint16_t b, float c)
{
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
    return c;
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// 16-byte global -> shared copy that bypasses registers and L1 (LDGSTS)
__device__ __forceinline__ void ldgsts16(uint32_t smem_addr, const void* gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

struct TileCoord {
    int n0, ox0, oy0, oc0, opx, opy;
};

__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) { return (__umulhi(x, f.mul) + x) >> f.shr; }

__device__ __forceinline__ void trace_mark(const PwGemmParams& p, int slot)
{
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[blockIdx.x * 64 + slot] = t;
    }
}

// per-thread state of an epilogue warp.  16 epilogue warps per CTA: warp (b, h, q) drains TMEM lane quarter q
// (accumulator rows 32q .. 32q+31) of column half h of every tile that lands in accumulator buffer b.
struct EpiWarp {
    int q;           // TMEM lane quarter this warp may touch (warp index & 3)
    int lane;
    int b;           // accumulator buffer served
    int h;           // column half served
    uint8_t* slab;   // this warp's staging slabs ([32 rows][64 B], SWIZZLE_64B), `slabs` of them (generic pointer: TMA source)
    int slabs;       // 1 or 2
    uint32_t cnt;    // store counter (slab ring)
};

static constexpr int EPI_SLAB_BYTES = 32 * 64;  // one [32 rows][32 fp16] store box

// Drains this warp's share of one accumulator tile: rows 32q..32q+31, accumulator columns [h*BLOCK_N/2, (h+1)*BLOCK_N/2).
// acc: TMEM address of the accumulator buffer including this warp's lane offset (column 0 of the tile).
// full_bar / full_parity: "accumulator complete" barrier (local).  The accumulator is handed back by one arrival
// per warp on the "accumulator drained" barrier: `empty_remote` != 0 is its shared::cluster address (CTA-pair
// kernel: the barrier lives in the leader CTA), otherwise `empty_local` is used.
// Every 32 output columns leave through the warp's own staging slab and its own TMA store (box = 32 rows x 32
// columns): no cross-warp barrier anywhere in the epilogue.
template <int BLOCK_N, bool CHUNK>
__device__ __forceinline__ void epilogue_tile(const PwGemmParams& p, const TileCoord& tc, uint32_t acc, uint64_t* full_bar,
                                              uint32_t full_parity, uint64_t* empty_local, uint32_t empty_remote,
                                              EpiWarp& w, int trace_slot)
{
    static_assert(!CHUNK || BLOCK_N == 256, "the 4:1 fold is built for 256-column tiles");
    constexpr int HALF = BLOCK_N / 2;  // accumulator columns per warp
    constexpr int NC = HALF / 32;      // 32-column chunks per warp
    const int lane = w.lane, q = w.q;
    const int row = q * 32 + lane;
    const int col0 = w.h * HALF;                      // first accumulator column of this warp
    const int ocol0 = CHUNK ? col0 / 4 : col0;        // first output column (relative to tc.oc0)
    const bool has_bias = p.bias != nullptr;
    const bool has_q = p.qscale != nullptr;
    const int n_res = p.n_res;
    const bool act = p.act == ACT_WSILU;
    const bool gdn = p.act >= ACT_GDN;
    const uint16_t ONE = 0x3C00;  // fp16 1.0: fma_f32_f16(h, ONE, x) == x + float(h) in one FHFMA

    auto hand_back = [&]() {
        // every tcgen05.ld of this warp for this tile has completed: hand the accumulator back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (empty_remote) mbar_arrive_cluster(empty_remote);
            else mbar_arrive(empty_local);
        }
    };
    // staging slab for the next store box: the store that last used it has finished reading it
    auto acquire_slab = [&]() -> uint8_t* {
        if (lane == 0) {
            if (w.slabs == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
        }
        __syncwarp();
        return w.slab + (w.slabs == 2 ? (w.cnt & 1) : 0) * EPI_SLAB_BYTES;
    };
    // all 32 rows of the box are written: one lane issues the TMA store of [32 rows][32 columns] at output column c0
    auto publish = [&](uint8_t* sbuf, int c0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (p.linear) tma_store_2d(&p.tm_c, sbuf, c0, tc.ox0 + q * 32);
            else tma_store_5d(&p.tm_c, sbuf, c0, tc.opx, tc.ox0, tc.opy, tc.oy0 + q * p.epi_rows_y);
            tma_store_commit();
        }
        ++w.cnt;
    };
    auto sw64 = [&](int chunk16) -> uint32_t {
        return static_cast<uint32_t>(lane * 64 + ((chunk16 ^ ((lane >> 1) & 3)) << 4));
    };

    // residual rows of this thread (same pixel grid as the output); rows past the edge are clamped to a valid
    // one — their results are clipped by the TMA store
    const __half* r1_row = nullptr;
    const __half* r2_row = nullptr;
    if (n_res > 0) {
        const int ry = static_cast<int>(fdiv(row, p.fd_bw));
        long long x = tc.ox0 + (row - ry * p.bw);
        long long y = tc.oy0 + ry;
        x = x < p.res_w ? x : p.res_w - 1;
        y = y < p.res_h ? y : p.res_h - 1;
        r1_row = p.r1 + (y * p.res_w + x) * p.r1_pitch + tc.oc0 + ocol0;
        r2_row = (n_res > 1) ? p.r2 + (y * p.res_w + x) * p.r2_pitch + tc.oc0 + ocol0 : r1_row;
    }
    const uint4* bias_v = reinterpret_cast<const uint4*>(p.bias + tc.n0 + col0);
    const uint4* q_v = reinterpret_cast<const uint4*>(p.qscale + tc.oc0 + ocol0);
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // warm L1 with this warp's bias / scale vectors while the accumulator is still being produced
    if (has_bias && lane * 8 < HALF) l1_touch(p.bias + tc.n0 + col0 + lane * 8);
    if (has_q && lane * 8 < (CHUNK ? HALF / 4 : HALF)) l1_touch(p.qscale + tc.oc0 + ocol0 + lane * 8);

    if constexpr (!CHUNK) {
        // ------------------------------------------------ plain tile: 32 accumulator columns -> 32 outputs -> one store
        // First residual: fetched warp-coalesced (one LDGSTS moves 8 rows x 64 B: lane l copies piece l & 3 of rows
        // (l >> 2) + 8 j) straight into the staging slab the chunk is stored from; each thread then picks up its own
        // row from there.  (One thread walking its own row with 16-byte loads — the previous scheme — makes every
        // load instruction touch 32 different lines: the LSU was the bottleneck of the shortcut GEMMs.)
        const __half* rr[4];
        if (n_res > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = q * 32 + (lane >> 2) + 8 * j;
                const int ry = static_cast<int>(fdiv(r, p.fd_bw));
                long long x = tc.ox0 + (r - ry * p.bw);
                long long y = tc.oy0 + ry;
                x = x < p.res_w ? x : p.res_w - 1;
                y = y < p.res_h ? y : p.res_h - 1;
                rr[j] = p.r1 + (y * p.res_w + x) * p.r1_pitch + tc.oc0 + ocol0 + (lane & 3) * 8;
            }
        }
        const uint32_t slab0 = smem_u32(w.slab);
        const uint32_t my_sw = static_cast<uint32_t>(lane * 64);
        const uint32_t my_x = static_cast<uint32_t>((lane >> 1) & 3);
        mbar_wait(full_bar, full_parity);
        tcgen05_fence_after();
        if (lane == 0 && q == 0 && w.h == 0 && trace_slot >= 0) trace_mark(p, trace_slot);
        if (p.dbg & 2) {  // micro-benchmark: drain nothing, just hand the accumulator back
            hand_back();
            return;
        }
        // per-phase SM-clock marks of the first tile's chunks (tools/gemm_trace.py): slots 16 + 8a + k
        const bool tr = p.trace && trace_slot == 7 && q == 0 && w.h == 0 && lane == 0;
        auto clk = [&](int a, int k) {
            if (tr && a < 4) p.trace[blockIdx.x * 64 + 16 + a * 8 + k] = static_cast<unsigned long long>(clock64());
        };
#pragma unroll 1
        for (int a = 0; a < NC; ++a) {
            clk(a, 0);
            uint32_t v[32];
            tmem_ld_32x32b_x32(acc + col0 + a * 32, v);
            // staging slab of this chunk: the store that last used it has finished reading it
            if (lane == 0) {
                if (w.slabs == 2) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
            }
            __syncwarp();
            const uint32_t sb = (w.slabs == 2 ? (w.cnt & 1) : 0) * EPI_SLAB_BYTES;
            const uint32_t sbuf = slab0 + sb;
            if (n_res > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t rj = static_cast<uint32_t>((lane >> 2) + 8 * j);
                    ldgsts16(sbuf + rj * 64 + ((static_cast<uint32_t>(lane & 3) ^ ((rj >> 1) & 3)) << 4), rr[j] + a * 32);
                }
                cp_async_commit();
            }
            uint4 cb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cb[j] = has_bias ? __ldg(bias_v + a * 4 + j) : zero4;
            tmem_ld_wait();
            clk(a, 1);
            if (a + 1 == NC) hand_back();
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(cb);
            float t[32];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                t[e] = fma_f32_f16(static_cast<uint16_t>(bw[e >> 1] & 0xffffu), ONE, __uint_as_float(v[e]));
                t[e + 1] = fma_f32_f16(static_cast<uint16_t>(bw[e >> 1] >> 16), ONE, __uint_as_float(v[e + 1]));
            }
            if (act) {
#pragma unroll
                for (int e = 0; e < 32; ++e) t[e] = wsilu_f(t[e]);
            }
            clk(a, 2);
            if (n_res > 0) {
                cp_async_wait_all();
                __syncwarp();
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const uint4 r = lds128(sbuf + my_sw + ((static_cast<uint32_t>(gq) ^ my_x) << 4));
                    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&r);
                    if (gdn) {
                        // divisive normalisation: the "residual" operand is x, the accumulator beta + gamma . x^2
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float xv = fma_f32_f16(static_cast<uint16_t>(e & 1 ? w1[e >> 1] >> 16 : w1[e >> 1] & 0xffffu), ONE, 0.f);
                            const float nv = t[gq * 8 + e];
                            t[gq * 8 + e] = xv * (p.act == ACT_GDN ? rsqrtf(nv) : sqrtf(nv));
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            t[gq * 8 + 2 * e] = fma_f32_f16(static_cast<uint16_t>(w1[e] & 0xffffu), ONE, t[gq * 8 + 2 * e]);
                            t[gq * 8 + 2 * e + 1] = fma_f32_f16(static_cast<uint16_t>(w1[e] >> 16), ONE, t[gq * 8 + 2 * e + 1]);
                        }
                    }
                }
            }
            if (n_res > 1) {
                uint4 c2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) c2[j] = ld_stream16(r2_row + a * 32 + j * 8);
                const uint32_t* w2 = reinterpret_cast<const uint32_t*>(c2);
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    t[e] = fma_f32_f16(static_cast<uint16_t>(w2[e >> 1] & 0xffffu), ONE, t[e]);
                    t[e + 1] = fma_f32_f16(static_cast<uint16_t>(w2[e >> 1] >> 16), ONE, t[e + 1]);
                }
            }
            if (has_q) {
                uint4 cq[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) cq[j] = __ldg(q_v + a * 4 + j);
                const uint32_t* qw = reinterpret_cast<const uint32_t*>(cq);
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qw[e >> 1]));
                    t[e] *= qf.x;
                    t[e + 1] *= qf.y;
                }
            }
            clk(a, 3);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                uint4 o;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const __half2 h2 = __floats2half2_rn(t[gq * 8 + jj * 2], t[gq * 8 + jj * 2 + 1]);
                    ow[jj] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                sts128(sbuf + my_sw + ((static_cast<uint32_t>(gq) ^ my_x) << 4), o);
            }
            clk(a, 4);
            publish(w.slab + sb, tc.oc0 + ocol0 + a * 32);
            clk(a, 5);
        }
    } else {
        // ------------------------------------------------ 4 -> 1 fold: 4 x 32 accumulator columns -> 32 outputs -> one store
        mbar_wait(full_bar, full_parity);
        tcgen05_fence_after();
        if (lane == 0 && q == 0 && w.h == 0 && trace_slot >= 0) trace_mark(p, trace_slot);
        if (p.dbg & 2) {
            hand_back();
            return;
        }
        uint8_t* sbuf = acquire_slab();
#pragma unroll 1
        for (int a = 0; a < NC; ++a) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(acc + col0 + a * 32, v);
            uint4 cb[4], cq = zero4, c1 = zero4, c2 = zero4;
#pragma unroll
            for (int j = 0; j < 4; ++j) cb[j] = has_bias ? __ldg(bias_v + a * 4 + j) : zero4;
            if (has_q) cq = __ldg(q_v + a);
            if (n_res > 0) c1 = ld_stream16(r1_row + a * 8);
            if (n_res > 1) c2 = ld_stream16(r2_row + a * 8);
            tmem_ld_wait();
            if (a + 1 == NC) hand_back();
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(cb);
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(&cq);
            const uint32_t* w1 = reinterpret_cast<const uint32_t*>(&c1);
            const uint32_t* w2 = reinterpret_cast<const uint32_t*>(&c2);
            uint4 o4;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float o[2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int j = jj * 2 + hh;  // output column of this chunk; folds acc columns 4j .. 4j+3
                    float s4 = 0.f;
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {
                        const uint32_t bword = bw[(4 * j + e) >> 1];
                        float t0 = fma_f32_f16(static_cast<uint16_t>(bword & 0xffffu), ONE, __uint_as_float(v[4 * j + e]));
                        float t1 = fma_f32_f16(static_cast<uint16_t>(bword >> 16), ONE, __uint_as_float(v[4 * j + e + 1]));
                        if (act) { t0 = wsilu_f(t0); t1 = wsilu_f(t1); }
                        s4 += t0;
                        s4 += t1;
                    }
                    o[hh] = s4;
                }
                if (n_res > 0) {
                    o[0] = fma_f32_f16(static_cast<uint16_t>(w1[jj] & 0xffffu), ONE, o[0]);
                    o[1] = fma_f32_f16(static_cast<uint16_t>(w1[jj] >> 16), ONE, o[1]);
                }
                if (n_res > 1) {
                    o[0] = fma_f32_f16(static_cast<uint16_t>(w2[jj] & 0xffffu), ONE, o[0]);
                    o[1] = fma_f32_f16(static_cast<uint16_t>(w2[jj] >> 16), ONE, o[1]);
                }
                if (has_q) {
                    const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qw[jj]));
                    o[0] *= qf.x;
                    o[1] *= qf.y;
                }
                const __half2 h2 = __floats2half2_rn(o[0], o[1]);
                ow[jj] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(sbuf + sw64(a)) = o4;
        }
        publish(sbuf, tc.oc0 + ocol0);
    }
}

}  // namespace dcvc
