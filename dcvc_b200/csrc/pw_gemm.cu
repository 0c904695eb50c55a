// pw_gemm.cu — tcgen05 / TMEM / TMA implementation of the dense contractions of the
// DCVC-UF hot path (see pw_gemm.cuh).  Hand-written for sm_100a.
//
// One persistent CTA per SM (grid = min(#tiles, #SMs)), 18 warps, tiles dealt round-robin with the N tile fastest:
//   warp 0       TMA producer: per 64-wide k-block one box of activations (2-D [pixels][channels] for 1x1 ops; a 5-D
//                "tap" box of the 2x2-phase-split NHWC tensor for the conv kinds, OOB = zero padding) + one 2-D box of
//                weights, both SWIZZLE_128B, mbarrier complete_tx, through a ring of num_stages stages.
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (M = 128, N = BLOCK_N, K = 16); tcgen05.commit
//                releases smem stages / publishes the accumulator; two accumulator buffers in TMEM.
//   warps 2..17  epilogue (pw_gemm_epilogue.cuh): warp (b, h, q) drains lane quarter q / column half h of the tiles in
//                accumulator buffer b: tcgen05.ld, bias, WSiLU, 4:1 chunk-add, residual(s), per-channel quant scale,
//                fp16 pack into its own swizzled slab, its own TMA store (clips ragged edges).
// The pixel-local runs of DepthConvBlocks do not come through here any more: dcb_tail.cu fuses them.
#include "pw_gemm.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "ptx.cuh"
#include "pw_gemm_epilogue.cuh"
#include "pw_gemm_internal.cuh"

namespace dcvc {

template <int BLOCK_N>
__device__ __forceinline__ TileCoord tile_coord(const PwGemmParams& p, int tile)
{
    TileCoord t;
    const int mt = static_cast<int>(fdiv(tile, p.fd_n_tiles));
    const int nt = tile - mt * p.n_tiles;
    const int ty = static_cast<int>(fdiv(mt, p.fd_tiles_x));
    t.n0 = nt * BLOCK_N;
    t.ox0 = (mt - ty * p.tiles_x) * p.bw;
    t.oy0 = ty * p.bh;
    t.oc0 = p.chunk_add ? t.n0 / 4 : t.n0;
    t.opx = 0;
    t.opy = 0;
    if (p.phase_c > 0) {  // tconv: this N tile is one 2x2 phase of the upsampled image
        const int phase = static_cast<int>(fdiv(t.n0, p.fd_phase_c));
        t.oc0 = t.n0 - phase * p.phase_c;
        t.opx = phase & 1;
        t.opy = phase >> 1;
    }
    return t;
}

// Persistent kernel: grid = min(#tiles, #SMs), one CTA per SM, tiles assigned round-robin with the
// N tile fastest so that concurrently running CTAs share the same activation tile in L2.
template <int BLOCK_N, bool CHUNK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_kernel(const __grid_constant__ PwGemmParams p)
{
    using Cfg = TileCfg<BLOCK_N>;
    const int STAGES = p.num_stages;

    // run-time carve-up of the 227 KB: [STAGES x (A 16 KB | B BLOCK_N*128 B)] [staging] [control]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int a_stride = Cfg::STAGE_BYTES;
    uint8_t* a_base = smem;
    uint8_t* staging = a_base + STAGES * a_stride;
    uint8_t* ctrl = smem + SMEM_USABLE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);   // [MAX_STAGES]
    uint64_t* empty_bar = full_bar + MAX_STAGES;              // [MAX_STAGES]
    uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;         // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;             // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ctrl + 256);
    uint32_t* stage_clk = reinterpret_cast<uint32_t*>(ctrl + 272);  // [48] SM-clock marks of the first 24 k-blocks (trace)

    // tile schedule: tile id = pixel tile * n_tiles + N tile, dealt round-robin, so that concurrently running CTAs share
    // the same activation tile in L2
    auto tile_of = [&](int i) -> int {
        const int w = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        return w < p.total_tiles ? w : -1;
    };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) trace_mark(p, 0);  // entry

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a);
        tma_prefetch_desc(&p.tm_b);
        tma_prefetch_desc(&p.tm_c);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&tmem_full_bar[g], 1);
            mbar_init(&tmem_empty_bar[g], 8);  // one arrival per epilogue warp serving the buffer
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // programmatic dependent launch: everything above touched only this CTA's smem / TMEM and overlapped the tail
    // of the previous kernel; from here on its output is read
    griddep_launch_dependents();
    griddep_wait();
    if (threadIdx.x == 0) trace_mark(p, 1);  // setup done
    if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 64 + 15] = static_cast<uint32_t>(clock64());  // SM clock at mark 1

    if (warp == 0) {
        if (elect_one_sync()) {
            // ------------------------------------------------------------ TMA producer
            // stage / phase / tap counters are kept incrementally: this loop runs on ONE thread and its instruction
            // latency chain bounds how fast k-blocks can be requested (measured 0.32 us per k-block with the
            // straightforward it % STAGES, kb / kblk_per_tap formulation: three integer divisions per iteration)
            uint32_t it = 0;
            int s = 0;
            uint32_t ph = 0;
            const uint32_t stage_tx = Cfg::STAGE_BYTES;
            const int num_kblocks = p.num_kblocks;
            const int kblk_per_tap = p.kblk_per_tap;
            const bool lin = p.linear != 0;
            for (int i = 0;; ++i) {
                const int tile = tile_of(i);
                if (tile < 0) break;
                const TileCoord tc = tile_coord<BLOCK_N>(p, tile);
                int tap = 0, kc = 0;
                for (int kb = 0; kb < num_kblocks; ++kb, ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* a_dst = a_base + s * a_stride;
                    mbar_expect_tx(&full_bar[s], stage_tx);
                    if (lin) {
                        tma_load_2d(a_dst, &p.tm_a, &full_bar[s], kc * BLOCK_K, tc.ox0);
                    } else {
                        tma_load_5d(a_dst, &p.tm_a, &full_bar[s], kc * BLOCK_K, p.tap_px[tap],
                                    tc.ox0 + p.tap_dx[tap], p.tap_py[tap], tc.oy0 + p.tap_dy[tap]);
                    }
                    tma_load_2d(a_dst + A_STAGE_BYTES, &p.tm_b, &full_bar[s], kb * BLOCK_K, tc.n0);
                    if (it == 0) trace_mark(p, 2);  // first stage requested
                    if (p.trace && it < 24) stage_clk[it] = static_cast<uint32_t>(clock64());
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                    if (++kc == kblk_per_tap) { kc = 0; ++tap; }
                }
            }
            trace_mark(p, 3);  // last stage requested
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = make_idesc_f16_f32(BLOCK_M, BLOCK_N);
            uint32_t it = 0;
            int s = 0;
            uint32_t ph = 0;
            const int num_kblocks = p.num_kblocks;
            const bool do_mma = !(p.dbg & 1);
            int my_tiles = 0;
            while (tile_of(my_tiles) >= 0) ++my_tiles;
            for (int i = 0; i < my_tiles; ++i) {
                const int g = i & 1;
                const uint32_t u = static_cast<uint32_t>(i >> 1);
                mbar_wait(&tmem_empty_bar[g], (u & 1) ^ 1);  // epilogue drained this accumulator
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + g * Cfg::ACC_COLS;
                for (int kb = 0; kb < num_kblocks; ++kb, ++it) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    if (it == 0) trace_mark(p, 4);  // first stage landed
                    if (p.trace && it < 24) stage_clk[24 + it] = static_cast<uint32_t>(clock64());
                    const uint32_t a_addr = smem_u32(a_base + s * a_stride);
                    const uint32_t b_addr = a_addr + A_STAGE_BYTES;
                    const uint64_t a_desc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t b_desc = make_kmajor_sw128_desc(b_addr);
                    if (do_mma) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // advance 16 fp16 = 32 B inside the 128 B swizzle span: +2 in 16 B units
                            umma_f16_ss(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full_bar[g]);
                if (i == 0) trace_mark(p, 5);  // first tile issued
            }
            trace_mark(p, 6);  // last tile issued
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue (2 groups x 4 warps)
        EpiWarp ew;
        ew.q = warp & 3;
        ew.lane = lane;
        ew.b = ((warp - 2) >> 2) & 1;
        ew.h = (warp - 2) >> 3;
        ew.slabs = p.staging_bufs;
        ew.slab = staging + (warp - 2) * p.staging_bufs * EPI_SLAB_BYTES;
        ew.cnt = 0;
        const int g = ew.b;
        for (int i = g;; i += 2) {
            const int tile = tile_of(i);
            if (tile < 0) break;
            const TileCoord tc = tile_coord<BLOCK_N>(p, tile);
            const uint32_t u = static_cast<uint32_t>(i >> 1);
            const uint32_t acc = tmem_base + g * Cfg::ACC_COLS + (static_cast<uint32_t>(ew.q * 32) << 16);
            epilogue_tile<BLOCK_N, CHUNK>(p, tc, acc, &tmem_full_bar[g], u & 1, &tmem_empty_bar[g], 0u, ew,
                                          i == 0 ? 7 : (i == 1 ? 9 : 11));
            if (lane == 0 && ew.q == 0 && ew.h == 0) trace_mark(p, i == 0 ? 8 : (i == 1 ? 10 : 12));  // epilogue of tile i done
        }
        if (lane == 0) tma_store_wait_read<0>();
        if (lane == 0 && ew.q == 0 && ew.h == 0) trace_mark(p, 13 + g);  // group drained
        __syncwarp();
    }

    __syncthreads();
    if (p.trace && threadIdx.x < 48) p.trace[blockIdx.x * 64 + 16 + threadIdx.x] = stage_clk[threadIdx.x];
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host

static thread_local std::string g_err;
const char* gemm_last_error() { return g_err.c_str(); }
void gemm_set_error(const std::string& e) { g_err = e; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e =
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        g_err = "cuTensorMapEncodeTiled entry point not available";
        return 1;
    }
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof(buf),
                 "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu %llu", (int)r,
                 rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                 (unsigned long long)(rank > 2 ? gd[2] : 0),
                 (unsigned long long)(rank > 3 ? gd[3] : 0),
                 (unsigned long long)(rank > 4 ? gd[4] : 0));
        g_err = buf;
        return 1;
    }
    return 0;
}

// 5-D map of an NHWC view.  split2: expose the 2x2 pixel phases as dims 1 and 3.
int encode_act_map(CUtensorMap* m, const ActView& v, bool split2, bool linear, bool lin2d, int bw, int bh, int box_c)
{
    const uint64_t pb = static_cast<uint64_t>(v.pitch) * 2;
    uint64_t dims[5];
    uint64_t st[4];
    uint32_t box[5] = { static_cast<uint32_t>(box_c), 1, static_cast<uint32_t>(bw), 1, static_cast<uint32_t>(bh) };
    if (linear && !lin2d) {
        const uint64_t M = static_cast<uint64_t>(v.W) * v.H;
        dims[0] = v.C; dims[1] = 1; dims[2] = M; dims[3] = 1; dims[4] = 1;
        st[0] = pb; st[1] = pb; st[2] = M * pb; st[3] = M * pb;
    } else if (linear) {
        // 1x1: a genuine 2-D [pixels][channels] map
        const uint64_t M = static_cast<uint64_t>(v.W) * v.H;
        uint64_t d2[2] = { static_cast<uint64_t>(v.C), M };
        uint64_t s2[1] = { pb };
        uint32_t b2[2] = { static_cast<uint32_t>(box_c), static_cast<uint32_t>(bw * bh) };
        if (b2[0] > d2[0]) b2[0] = static_cast<uint32_t>(d2[0]);
        return encode_map(m, v.ptr, 2, d2, s2, b2,
                          box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    } else if (split2) {
        dims[0] = v.C; dims[1] = 2; dims[2] = v.W / 2; dims[3] = 2; dims[4] = v.H / 2;
        st[0] = pb; st[1] = 2 * pb; st[2] = static_cast<uint64_t>(v.W) * pb;
        st[3] = 2 * static_cast<uint64_t>(v.W) * pb;
    } else {
        dims[0] = v.C; dims[1] = 1; dims[2] = v.W; dims[3] = 1; dims[4] = v.H;
        st[0] = pb; st[1] = pb; st[2] = static_cast<uint64_t>(v.W) * pb;
        st[3] = static_cast<uint64_t>(v.W) * pb;
    }
    if (box[0] > dims[0]) box[0] = static_cast<uint32_t>(dims[0]);
    return encode_map(m, v.ptr, 5, dims, st, box,
                      box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

struct TilePlan {
    int bn = 0;
    int stages = 0;
    int staging_bufs = 2;
};

FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;  // ceil(log2 d)
    f.shr = s;
    f.mul = static_cast<uint32_t>(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
}

// Chooses the N tile and the pipeline depth.  Every tile streams its activation k-blocks and its weight k-blocks (hot in
// L2) through the ring; measured on B200 (round 1: tools/gemm_micro.py), weight-resident CTAs and clusters that share one
// multicast activation tile were both slower on every shape of the codec and are gone.
static TilePlan pick_tile_plan(int n_unit, int n_total, bool chunk_add, long long m_tiles, int num_kblocks, int num_sms)
{
    TilePlan best;
    const int cand_all[4] = { 256, 192, 128, 64 };
    double best_cost = 1e30;
    for (int ci = 0; ci < 4; ++ci) {
        const int bn = cand_all[ci];
        if (n_unit % bn) continue;
        if (chunk_add && bn != 256) continue;  // 4->1 fold: 128 accumulator columns per epilogue warp -> one 32-column store box
        const int n_tiles = n_total / bn;
        const int b_stage = bn * BLOCK_K * 2;
        const long long items = m_tiles * n_tiles;
        const int ctas = static_cast<int>(items < num_sms ? items : num_sms);
        if (ctas < 1) continue;
        const long long per_cta = (items + ctas - 1) / ctas;
        TilePlan t;
        t.bn = bn;
        t.staging_bufs = 2;
        t.stages = (SMEM_USABLE - EPI_GROUPS * t.staging_bufs * SUB_TILE_BYTES) / (A_STAGE_BYTES + b_stage);
        if (t.stages < 2) continue;
        if (t.stages > MAX_STAGES) t.stages = MAX_STAGES;
        // cost model (arbitrary time units per pixel tile of the whole problem): activation stream (one 16 KB k-block per
        // N tile), weight stream (hot in L2, ~3x cheaper per byte), tensor pipe, all scaled by the wave quantisation of the
        // persistent grid and a latency penalty for shallow pipelines
        const double act = static_cast<double>(n_tiles) * num_kblocks * A_STAGE_BYTES;
        const double wgt = static_cast<double>(n_tiles) * num_kblocks * b_stage / 3.0;
        const double mma = static_cast<double>(n_total) * num_kblocks * 64.0 * 128.0 / 1500.0;  // ~bytes-equivalent
        const double quant = static_cast<double>(per_cta) * ctas / static_cast<double>(items);
        const double depth = 1.0 + 1.5 / t.stages;
        const double cost = (act + wgt > mma ? act + wgt : mma) * quant * depth * (static_cast<double>(num_sms) / ctas);
        if (cost < best_cost) {
            best_cost = cost;
            best = t;
        }
    }
    return best;
}

int gemm_plan(GemmOp& op)
{
    PwGemmParams& p = op.p;
    memset(&p, 0, sizeof(p));
    const int C = op.in.C;
    if (C % 64 != 0) { g_err = "gemm_plan: input channels must be a multiple of 64"; return 1; }
    if (op.in.pitch % 8 || op.out.pitch % 8) { g_err = "gemm_plan: pitch must be a multiple of 8"; return 1; }
    if ((reinterpret_cast<uintptr_t>(op.in.ptr) & 15) || (reinterpret_cast<uintptr_t>(op.out.ptr) & 15)) {
        g_err = "gemm_plan: activation pointers must be 16-byte aligned";
        return 1;
    }
    int taps = 1;
    bool linear = false;
    switch (op.kind) {
    case GEMM_PW:
        taps = 1; linear = true;
        p.tap_px[0] = p.tap_py[0] = p.tap_dx[0] = p.tap_dy[0] = 0;
        if (op.out.W != op.in.W || op.out.H != op.in.H) { g_err = "gemm_plan: pw size mismatch"; return 1; }
        break;
    case GEMM_CONV3X3_S2:
        taps = 9;
        for (int ky = 0; ky < 3; ++ky) {
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                // input coord 2*o + k - 1  ->  (phase, shift) of the 2x2-split view
                p.tap_py[t] = (ky == 1) ? 0 : 1;
                p.tap_dy[t] = (ky == 0) ? -1 : 0;
                p.tap_px[t] = (kx == 1) ? 0 : 1;
                p.tap_dx[t] = (kx == 0) ? -1 : 0;
            }
        }
        break;
    case GEMM_CONV2X2_S2:
        taps = 4;
        for (int t = 0; t < 4; ++t) {
            p.tap_py[t] = t >> 1; p.tap_px[t] = t & 1; p.tap_dx[t] = p.tap_dy[t] = 0;
        }
        break;
    case GEMM_TCONV2X2:
        taps = 1;
        p.tap_px[0] = p.tap_py[0] = p.tap_dx[0] = p.tap_dy[0] = 0;
        break;
    case GEMM_CONV3X3_PS2:
        taps = 9;
        for (int ky = 0; ky < 3; ++ky) {
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                p.tap_px[t] = p.tap_py[t] = 0;   // unsplit view: the tap is a pure shift, OOB = zero padding
                p.tap_dx[t] = static_cast<int8_t>(kx - 1);
                p.tap_dy[t] = static_cast<int8_t>(ky - 1);
            }
        }
        break;
    default:
        g_err = "gemm_plan: bad kind";
        return 1;
    }
    if (op.kind == GEMM_CONV3X3_S2 || op.kind == GEMM_CONV2X2_S2) {
        if ((op.in.W & 1) || (op.in.H & 1) || op.out.W != op.in.W / 2 || op.out.H != op.in.H / 2) {
            g_err = "gemm_plan: stride-2 conv needs even input and out = in/2";
            return 1;
        }
    }
    const bool upsamples = (op.kind == GEMM_TCONV2X2 || op.kind == GEMM_CONV3X3_PS2);  // N = 4 phases x out.C
    if (upsamples) {
        if (op.out.W != op.in.W * 2 || op.out.H != op.in.H * 2 || op.N != op.out.C * 4) {
            g_err = "gemm_plan: tconv geometry mismatch";
            return 1;
        }
    }
    // pixel tile geometry (over the GEMM-M pixel grid = output grid, except tconv = input grid)
    const int gw = upsamples ? op.in.W : op.out.W;
    const int gh = upsamples ? op.in.H : op.out.H;
    long long m_tiles;
    int tiles_x, tiles_y;
    if (linear) {
        p.bw = 128; p.bh = 1;
        const long long M = static_cast<long long>(gw) * gh;
        tiles_x = static_cast<int>((M + 127) / 128); tiles_y = 1;
    } else {
        if (gw % 16 == 0) { p.bw = 16; p.bh = 8; } else { p.bw = 8; p.bh = 16; }
        tiles_x = (gw + p.bw - 1) / p.bw;
        tiles_y = (gh + p.bh - 1) / p.bh;
    }
    m_tiles = static_cast<long long>(tiles_x) * tiles_y;

    const int n_unit = upsamples ? op.out.C : op.N;
    int num_sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            int v = 0;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) num_sms = v;
        }
    }
    {
        // geometry checks shared by both kernels
        const int out_c = op.chunk_add ? op.N / 4 : (upsamples ? op.N / 4 : op.N);
        if (op.out.C != out_c) { g_err = "gemm_plan: out.C does not match N"; return 1; }
        if (op.res2.ptr && !op.res1.ptr) { g_err = "gemm_plan: res2 without res1"; return 1; }
        if (op.act >= ACT_GDN && (!op.res1.ptr || op.res2.ptr || op.chunk_add || op.kind != GEMM_PW)) {
            g_err = "gemm_plan: GDN / IGDN is a 1x1 op with exactly one residual operand (x) and no chunk-add";
            return 1;
        }
        if (op.res1.ptr) {
            if (upsamples) { g_err = "gemm_plan: tconv takes no residual"; return 1; }
            const ActView* rs[2] = { &op.res1, &op.res2 };
            for (int i = 0; i < (op.res2.ptr ? 2 : 1); ++i) {
                if (rs[i]->W != op.out.W || rs[i]->H != op.out.H || rs[i]->C != op.out.C || (rs[i]->pitch % 8) ||
                    (reinterpret_cast<uintptr_t>(rs[i]->ptr) & 15)) {
                    g_err = "gemm_plan: residual geometry mismatch";
                    return 1;
                }
            }
        }
    }
    const TilePlan tp = pick_tile_plan(n_unit, op.N, op.chunk_add != 0, m_tiles, taps * C / 64, num_sms);
    const int bn = tp.bn;
    if (bn == 0 || op.N % bn != 0) { g_err = "gemm_plan: unsupported N"; return 1; }
    const int out_c_expected = op.chunk_add ? op.N / 4 : (upsamples ? op.N / 4 : op.N);
    if (op.out.C != out_c_expected) { g_err = "gemm_plan: out.C does not match N"; return 1; }
    op.block_n = bn;

    p.num_kblocks = taps * C / 64;
    p.kblk_per_tap = C / 64;
    p.act = op.act;
    p.chunk_add = op.chunk_add;
    p.bias = op.bias;
    p.qscale = op.qscale;
    p.phase_c = upsamples ? op.out.C : 0;
    p.n_res = (op.res1.ptr ? 1 : 0) + (op.res2.ptr ? 1 : 0);
    if (op.res2.ptr && !op.res1.ptr) { g_err = "gemm_plan: res2 without res1"; return 1; }

    const bool in_split = (op.kind == GEMM_CONV3X3_S2 || op.kind == GEMM_CONV2X2_S2);
    const bool lin2d = linear;
    if (encode_act_map(&p.tm_a, op.in, in_split, linear, lin2d, p.bw, p.bh)) return 1;
    {
        const uint64_t Ktot = static_cast<uint64_t>(taps) * C;
        uint64_t dims[2] = { Ktot, static_cast<uint64_t>(op.N) };
        uint64_t st[1] = { Ktot * 2 };
        uint32_t box[2] = { 64, static_cast<uint32_t>(bn) };
        if (encode_map(&p.tm_b, op.weight, 2, dims, st, box)) return 1;
    }
    const bool out_split = upsamples;
    // store box of one epilogue warp: 32 pixels (rows 32q .. 32q+31 of the tile) x 32 channels, SWIZZLE_64B
    p.epi_rows_y = linear ? 1 : 32 / p.bw;
    if (encode_act_map(&p.tm_c, op.out, out_split, linear, lin2d, linear ? 32 : p.bw, linear ? 1 : 32 / p.bw, 32)) return 1;
    if (op.res1.ptr) {
        if (upsamples) { g_err = "gemm_plan: tconv takes no residual"; return 1; }
        const ActView* rs[2] = { &op.res1, &op.res2 };
        for (int i = 0; i < p.n_res; ++i) {
            if (rs[i]->W != op.out.W || rs[i]->H != op.out.H || rs[i]->C != op.out.C || (rs[i]->pitch % 8) ||
                (reinterpret_cast<uintptr_t>(rs[i]->ptr) & 15)) {
                g_err = "gemm_plan: residual geometry mismatch";
                return 1;
            }
        }
        p.r1 = static_cast<const __half*>(op.res1.ptr);
        p.r1_pitch = op.res1.pitch;
        if (op.res2.ptr) {
            p.r2 = static_cast<const __half*>(op.res2.ptr);
            p.r2_pitch = op.res2.pitch;
        }
        if (linear) {
            p.res_w = static_cast<int>(static_cast<long long>(op.out.W) * op.out.H);
            p.res_h = 1;
        } else {
            p.res_w = op.out.W;
            p.res_h = op.out.H;
        }
    }
    p.n_tiles = op.N / bn;
    p.tiles_x = tiles_x;
    p.total_tiles = static_cast<int>(m_tiles) * p.n_tiles;
    const int grid_ctas = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    p.m_tiles = static_cast<int>(m_tiles);
    p.linear = lin2d ? 1 : 0;
    p.fd_n_tiles = make_fastdiv(p.n_tiles);
    p.fd_tiles_x = make_fastdiv(p.tiles_x);
    p.fd_phase_c = make_fastdiv(p.phase_c > 0 ? p.phase_c : 1);
    p.fd_bw = make_fastdiv(p.bw);
    op.grid = dim3(grid_ctas, 1, 1);
    op.smem = SMEM_TOTAL;
    p.num_stages = tp.stages;
    p.staging_bufs = tp.staging_bufs;
    if (const char* d = getenv("DCVC_B200_GEMM_DBG")) p.dbg = atoi(d);
    if (const char* d = getenv("DCVC_B200_GEMM_TRACE")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(d, nullptr, 0));
    op.stages = p.num_stages;
    if (p.num_stages < 2) { g_err = "gemm_plan: pipeline does not fit"; return 1; }
    if (!op.chunk_add && bn % 64 != 0) { g_err = "gemm_plan: BLOCK_N must be a multiple of 64"; return 1; }
    if (op.chunk_add && bn != 256) { g_err = "gemm_plan: chunk-add needs BLOCK_N 256"; return 1; }
    op.planned = true;
    return 0;
}

template <int BN, bool CHUNK>
static cudaError_t set_attr()
{
    return cudaFuncSetAttribute(pw_gemm_kernel<BN, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
}

int gemm_init()
{
    static bool done = false;
    if (done) return 0;
    cudaError_t e = set_attr<64, false>();
    if (e == cudaSuccess) e = set_attr<128, false>();
    if (e == cudaSuccess) e = set_attr<192, false>();
    if (e == cudaSuccess) e = set_attr<256, false>();
    if (e == cudaSuccess) e = set_attr<256, true>();
    if (e != cudaSuccess) {
        g_err = std::string("cudaFuncSetAttribute(pw_gemm): ") + cudaGetErrorString(e);
        return 1;
    }
    done = true;
    return 0;
}

// DCVC_B200_PDL=0 turns programmatic dependent launch off (debugging)
static const bool g_pdl = []() { const char* e = getenv("DCVC_B200_PDL"); return !(e && e[0] == '0'); }();
bool gemm_pdl_enabled() { return g_pdl; }

template <int BN, bool CHUNK>
static cudaError_t launch_bn(const GemmOp& op, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = op.grid;
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = op.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && op.pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, pw_gemm_kernel<BN, CHUNK>, op.p);
}

int gemm_launch(const GemmOp& op, cudaStream_t stream)
{
    if (!op.planned) { g_err = "gemm_launch: op not planned"; return 1; }
    if (gemm_init()) return 1;
    cudaError_t e;
    switch (op.block_n + (op.chunk_add ? 1 : 0)) {
    case 64: e = launch_bn<64, false>(op, stream); break;
    case 128: e = launch_bn<128, false>(op, stream); break;
    case 192: e = launch_bn<192, false>(op, stream); break;
    case 256: e = launch_bn<256, false>(op, stream); break;
    case 257: e = launch_bn<256, true>(op, stream); break;
    default: g_err = "gemm_launch: bad block_n"; return 1;
    }
    if (e != cudaSuccess) {
        g_err = std::string("pw_gemm launch failed: ") + cudaGetErrorString(e);
        return 1;
    }
    return 0;
}

}  // namespace dcvc
