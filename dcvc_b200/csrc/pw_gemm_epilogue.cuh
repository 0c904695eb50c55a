// pw_gemm_epilogue.cuh — the TMEM -> registers -> swizzled smem -> TMA-store epilogue shared by the pw_gemm kernels
// (streaming kernel and A-resident / CTA-pair kernel in pw_gemm.cu): bias, WSiLU, 4:1 chunk-add, up to two
// residuals, per-channel quant scale.  One call drains one [128 pixels][BLOCK_N] accumulator tile with the four
// warps of an epilogue group (one warp per TMEM lane quarter).
//
// Latency-bound by construction, so everything it waits for is requested one 32-column chunk ahead: the next
// tcgen05.ld, and bias / quant-scale / residual vectors as 16-byte global loads into registers.  One named barrier
// per store box; no divergent branches, no integer divisions.
#pragma once
#include "ptx.cuh"
#include "pw_gemm.cuh"

namespace dcvc {

static constexpr int EPI_SUB_TILE_BYTES = 128 * 64 * 2;  // one [128][64] fp16 store box

__device__ __forceinline__ float wsilu_f(float x)
{
    // x * sigmoid(4x) = 0.5 x (1 + tanh(2x))   (reference: src/layers/layers.py:106-111); one MUFU op
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(2.f * x));
    return 0.5f * x * (1.f + t);
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

// per-thread state of an epilogue warp
struct EpiWarp {
    int g;              // epilogue group <-> accumulator buffer
    int q;              // TMEM lane quarter this warp may touch
    int lane;
    int row;            // q * 32 + lane: the accumulator row (pixel of the tile) this thread owns
    bool issuer;        // the thread of the group that issues TMA stores
    bool two_bufs;      // two staging buffers per group
    uint32_t bar_id;    // named barrier of the group
    uint8_t* stage_g;   // staging buffers of the group
    uint32_t cnt;       // store-box counter of the group
};

// acc: TMEM address of the accumulator buffer including this warp's lane offset.
// full_bar / full_parity: "accumulator complete" barrier (local).  The accumulator is handed back by one arrival
// per warp on the "accumulator drained" barrier: `empty_remote` != 0 is its shared::cluster address (CTA-pair
// kernel: the barrier lives in the leader CTA), otherwise `empty_local` is used.
template <int BLOCK_N, bool CHUNK>
__device__ __forceinline__ void epilogue_tile(const PwGemmParams& p, const TileCoord& tc, uint32_t acc, uint64_t* full_bar,
                                              uint32_t full_parity, uint64_t* empty_local, uint32_t empty_remote,
                                              EpiWarp& w, int trace_slot)
{
    constexpr bool OUT32 = CHUNK && BLOCK_N == 128;  // 32-column store box (SWIZZLE_64B rows)
    constexpr int SUB_BYTES = OUT32 ? EPI_SUB_TILE_BYTES / 2 : EPI_SUB_TILE_BYTES;
    constexpr int NC = BLOCK_N / 32;  // accumulator chunks of 32 columns per tile
    const int row = w.row, lane = w.lane, q = w.q;
    const bool two_bufs = w.two_bufs;
    const bool issuer = w.issuer;
    const uint32_t bar_id = w.bar_id;
    uint8_t* stage_g = w.stage_g;
    const bool has_bias = p.bias != nullptr;
    const bool has_q = p.qscale != nullptr;
    const int n_res = p.n_res;
    const bool act = p.act == ACT_WSILU;
    const uint16_t ONE = 0x3C00;  // fp16 1.0: fma_f32_f16(h, ONE, x) == x + float(h) in one FHFMA

    auto hand_back = [&]() {
        // every tcgen05.ld of this tile has completed: hand the accumulator back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (empty_remote) mbar_arrive_cluster(empty_remote);
            else mbar_arrive(empty_local);
        }
    };
    auto sw_off = [&](int chunk) -> uint32_t {
        return OUT32 ? static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4))
                     : sw128_offset(row, chunk);
    };
    // publish a finished store box: every earlier store of this group has left its staging buffer (so the buffer
    // the NEXT box writes is free), all 128 rows are written, then one thread issues the TMA store
    auto publish = [&](uint8_t* sbuf, int c0) {
        fence_proxy_async_smem();
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        if (issuer) {
            if (p.linear) tma_store_2d(&p.tm_c, sbuf, c0, tc.ox0);
            else tma_store_5d(&p.tm_c, sbuf, c0, tc.opx, tc.ox0, tc.opy, tc.oy0);
            tma_store_commit();
            if (!two_bufs) tma_store_wait_read<0>();
        }
        if (!two_bufs) named_bar_sync(bar_id, 128);
        ++w.cnt;
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
        r1_row = p.r1 + (y * p.res_w + x) * p.r1_pitch + tc.oc0;
        r2_row = (n_res > 1) ? p.r2 + (y * p.res_w + x) * p.r2_pitch + tc.oc0 : r1_row;
    }
    const uint4* bias_v = reinterpret_cast<const uint4*>(p.bias + tc.n0);
    const uint4* q_v = reinterpret_cast<const uint4*>(p.qscale + tc.oc0);
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // warm L1 with this tile's bias / scale vectors while the accumulator is still being produced: the
    // per-chunk loads below are then warp-uniform L1 hits (their miss latency used to be exposed once per
    // 128-byte line, i.e. every other chunk)
    if (has_bias && lane * 8 < BLOCK_N) l1_touch(p.bias + tc.n0 + lane * 8);
    if (has_q && lane * 8 < (CHUNK ? BLOCK_N / 4 : BLOCK_N)) l1_touch(p.qscale + tc.oc0 + lane * 8);

    if constexpr (!CHUNK) {
        // ------------------------------------------------ plain tile: 32 accumulator columns -> 32 outputs
        uint4 n1[4];  // first residual of the NEXT chunk (L2 latency); bias / scale / second residual are
                      // requested at the top of their own chunk (warp-uniform L1 hits, resp. rarely used)
        auto prefetch = [&](int a) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n_res > 0) n1[j] = ld_stream16(r1_row + a * 32 + j * 8);
            }
        };
        prefetch(0);
        mbar_wait(full_bar, full_parity);
        tcgen05_fence_after();
        if (lane == 0 && q == 0 && trace_slot >= 0) trace_mark(p, trace_slot);
        if (p.dbg & 2) {  // micro-benchmark: drain nothing, just hand the accumulator back
            hand_back();
            return;
        }
        uint32_t vn[32];
        tmem_ld_32x32b_x32(acc, vn);
        uint8_t* sbuf = nullptr;
#pragma unroll 1
        for (int a = 0; a < NC; ++a) {
            uint4 cb[4], cq[4], c1[4], c2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cb[j] = has_bias ? __ldg(bias_v + a * 4 + j) : zero4;
                if (has_q) cq[j] = __ldg(q_v + a * 4 + j);
                if (n_res > 1) c2[j] = ld_stream16(r2_row + a * 32 + j * 8);
                c1[j] = n1[j];
            }
            uint32_t v[32];
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = vn[j];
            if (a + 1 < NC) {
                tmem_ld_32x32b_x32(acc + (a + 1) * 32, vn);
                prefetch(a + 1);
            } else {
                hand_back();
            }
            if ((a & 1) == 0) sbuf = stage_g + (two_bufs ? (w.cnt & 1) : 0) * SUB_BYTES;
            // one pass per epilogue term over the 32 columns: each optional term is a warp-uniform branch
            // around a short unrolled loop (keeps the kernel small: no per-combination code clones)
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(cb);
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(cq);
            const uint32_t* w1 = reinterpret_cast<const uint32_t*>(c1);
            const uint32_t* w2 = reinterpret_cast<const uint32_t*>(c2);
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
            if (n_res > 0) {
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    t[e] = fma_f32_f16(static_cast<uint16_t>(w1[e >> 1] & 0xffffu), ONE, t[e]);
                    t[e + 1] = fma_f32_f16(static_cast<uint16_t>(w1[e >> 1] >> 16), ONE, t[e + 1]);
                }
            }
            if (n_res > 1) {
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    t[e] = fma_f32_f16(static_cast<uint16_t>(w2[e >> 1] & 0xffffu), ONE, t[e]);
                    t[e + 1] = fma_f32_f16(static_cast<uint16_t>(w2[e >> 1] >> 16), ONE, t[e + 1]);
                }
            }
            if (has_q) {
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qw[e >> 1]));
                    t[e] *= qf.x;
                    t[e + 1] *= qf.y;
                }
            }
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                uint4 o;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const __half2 h = __floats2half2_rn(t[gq * 8 + jj * 2], t[gq * 8 + jj * 2 + 1]);
                    ow[jj] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(sbuf + sw128_offset(row, (a & 1) * 4 + gq)) = o;
            }
            if (a & 1) publish(sbuf, tc.oc0 + (a >> 1) * 64);
        }
    } else {
        // ------------------------------------------------ 4 -> 1 fold: 32 accumulator columns -> 8 outputs
        uint4 nb[4], nq, n1, n2;
        auto prefetch = [&](int a) {
#pragma unroll
            for (int j = 0; j < 4; ++j) nb[j] = has_bias ? __ldg(bias_v + a * 4 + j) : zero4;
            if (has_q) nq = __ldg(q_v + a);
            if (n_res > 0) n1 = ld_stream16(r1_row + a * 8);
            if (n_res > 1) n2 = ld_stream16(r2_row + a * 8);
        };
        prefetch(0);
        mbar_wait(full_bar, full_parity);
        tcgen05_fence_after();
        if (lane == 0 && q == 0 && trace_slot >= 0) trace_mark(p, trace_slot);
        if (p.dbg & 2) {
            hand_back();
            return;
        }
        uint32_t vn[32];
        tmem_ld_32x32b_x32(acc, vn);
        uint8_t* sbuf = stage_g + (two_bufs ? (w.cnt & 1) : 0) * SUB_BYTES;
#pragma unroll 1
        for (int a = 0; a < NC; ++a) {
            uint4 cb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cb[j] = nb[j];
            const uint4 cq = nq, c1 = n1, c2 = n2;
            uint32_t v[32];
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = vn[j];
            if (a + 1 < NC) {
                tmem_ld_32x32b_x32(acc + (a + 1) * 32, vn);
                prefetch(a + 1);
            } else {
                hand_back();
            }
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
                for (int h = 0; h < 2; ++h) {
                    const int j = jj * 2 + h;  // output column of this chunk; folds acc columns 4j .. 4j+3
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
                    o[h] = s4;
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
            *reinterpret_cast<uint4*>(sbuf + sw_off(a)) = o4;
        }
        publish(sbuf, tc.oc0);
    }
}

}  // namespace dcvc
