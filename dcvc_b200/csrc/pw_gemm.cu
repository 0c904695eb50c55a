// pw_gemm.cu — tcgen05 / TMEM / TMA implementation of the dense contractions of the
// DCVC-UF hot path (see pw_gemm.cuh).  Hand-written for sm_100a.
//
// CTA = 128 output pixels x BLOCK_N GEMM columns, 6 warps:
//   warp 0      TMA producer: per 64-wide k-block one 5-D box of activations (a "tap" of the
//               2x2-phase-split NHWC tensor; OOB = zero padding) + one 2-D box of weights,
//               both SWIZZLE_128B, mbarrier complete_tx; afterwards the residual tile.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BLOCK_N, K=16),
//               tcgen05.commit releases smem stages / publishes the accumulator.
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns), bias, WSiLU, 4:1 chunk-add,
//               residual(s), per-channel quant scale, fp16 pack into a swizzled staging tile
//               (aliased over the drained pipeline stages), TMA store (clips ragged edges).
// Two CTAs are co-resident per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
#include "pw_gemm.cuh"

#include <stdio.h>
#include <string.h>

#include <string>

#include "ptx.cuh"

namespace dcvc {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;
static constexpr int UMMA_K = 16;
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
static constexpr int SUB_TILE_BYTES = BLOCK_M * 64 * 2;      // one [128][64] fp16 store box
static constexpr int NUM_THREADS = 192;

template <int BLOCK_N>
struct TileCfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BLOCK_N >= 192) ? 2 : (BLOCK_N == 128 ? 3 : 4);
    static constexpr int TMEM_COLS = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128 ? 128 : 256);
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    // + barriers (2*STAGES + 2) * 8, tmem ptr, bias tile, 1 KB alignment slack
    static constexpr int SMEM_BYTES = PIPE_BYTES + 256 + BLOCK_N * 2 + 1024;
};

__device__ __forceinline__ float wsilu_f(float x)
{
    // x * sigmoid(4x)  (reference: src/layers/layers.py:106-111)
    return __fdividef(x, 1.f + __expf(-4.f * x));
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_kernel(const __grid_constant__ PwGemmParams p)
{
    using Cfg = TileCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::PIPE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* res_full_bar = tmem_full_bar + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_full_bar + 1);
    __half* bias_s = reinterpret_cast<__half*>(smem + Cfg::PIPE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int nt = blockIdx.x;          // N tile
    const int ox0 = blockIdx.y * p.bw;  // pixel tile origin
    const int oy0 = blockIdx.z * p.bh;
    const int n0 = nt * BLOCK_N;

    const int out_cols = p.chunk_add ? BLOCK_N / 4 : BLOCK_N;
    const int n_sub = (out_cols + 63) / 64;
    // output channel origin / 2x2 phase (tconv stores phase (opy, opx) of the upsampled image)
    int oc0 = p.chunk_add ? n0 / 4 : n0;
    int opx = 0, opy = 0;
    if (p.phase_c > 0) {
        const int phase = n0 / p.phase_c;
        oc0 = n0 - phase * p.phase_c;
        opx = phase & 1;
        opy = phase >> 1;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a);
        tma_prefetch_desc(&p.tm_b);
        tma_prefetch_desc(&p.tm_c);
        if (p.n_res > 0) tma_prefetch_desc(&p.tm_r1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(res_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
            bias_s[i] = p.bias ? p.bias[n0 + i] : __float2half(0.f);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int tap = kb / p.kblk_per_tap;
                const int kc = kb - tap * p.kblk_per_tap;
                uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                tma_load_5d(a_dst, &p.tm_a, &full_bar[s], kc * BLOCK_K, p.tap_px[tap],
                            ox0 + p.tap_dx[tap], p.tap_py[tap], oy0 + p.tap_dy[tap]);
                tma_load_2d(b_dst, &p.tm_b, &full_bar[s], kb * BLOCK_K, n0);
            }
            if (p.n_res > 0) {
                // staging tile aliases the pipeline stages: wait until every MMA has drained
                mbar_wait(tmem_full_bar, 0);
                mbar_expect_tx(res_full_bar, n_sub * SUB_TILE_BYTES);
                for (int j = 0; j < n_sub; ++j) {
                    tma_load_5d(smem + j * SUB_TILE_BYTES, &p.tm_r1, res_full_bar, oc0 + j * 64,
                                opx, ox0, opy, oy0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = make_idesc_f16_f32(BLOCK_M, BLOCK_N);
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t a_desc = make_kmajor_sw128_desc(a_addr);
                const uint64_t b_desc = make_kmajor_sw128_desc(a_addr + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    // advance 16 fp16 = 32 B inside the 128 B swizzle span: +2 in 16 B units
                    umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full_bar);
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        if (p.n_res > 0) mbar_wait(res_full_bar, 0);

        const __half* qs = p.qscale;
        // second residual: read straight from global (rare: ResidualBlock* with shortcut)
        const __half* r2_row = nullptr;
        if (p.n_res > 1) {
            // tm_r2 is unused by TMA; its first 16 bytes carry {ptr, pitch, W, H} (see gemm_plan)
            const uint64_t* raw = reinterpret_cast<const uint64_t*>(&p.tm_r2);
            const __half* base = reinterpret_cast<const __half*>(raw[0]);
            const int pitch = static_cast<int>(raw[1] & 0xffffffffu);
            const int W = static_cast<int>(raw[2] & 0xffffffffu);
            const int H = static_cast<int>(raw[2] >> 32);
            const int bx = row % p.bw;
            const int by = row / p.bw;
            const long long x = ox0 + bx;
            const long long y = oy0 + by;
            if (x < W && y < H) r2_row = base + (y * W + x) * pitch;
        }

#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
            tmem_ld_wait();
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float t = __uint_as_float(v[j]) + __half2float(bias_s[c0 + j]);
                x[j] = (p.act == ACT_WSILU) ? wsilu_f(t) : t;
            }
            if (p.chunk_add) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = (x[4 * j] + x[4 * j + 1]) + (x[4 * j + 2] + x[4 * j + 3]);
                }
                const int oc = c0 >> 2;  // output column inside the tile
                const int sub = oc >> 6;
                const int chunk = (oc & 63) >> 3;
                uint8_t* dst = smem + sub * SUB_TILE_BYTES + sw128_offset(row, chunk);
                if (p.n_res > 0) {
                    const uint4 r = *reinterpret_cast<const uint4*>(dst);
                    const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(rh[j]);
                        o[2 * j] += f.x;
                        o[2 * j + 1] += f.y;
                    }
                }
                if (r2_row) {
                    const uint4 r = *reinterpret_cast<const uint4*>(r2_row + oc0 + oc);
                    const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(rh[j]);
                        o[2 * j] += f.x;
                        o[2 * j + 1] += f.y;
                    }
                }
                if (qs) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] *= __half2float(qs[oc0 + oc + j]);
                }
                uint4 w;
                __half2* wh = reinterpret_cast<__half2*>(&w);
#pragma unroll
                for (int j = 0; j < 4; ++j) wh[j] = __floats2half2_rn(o[2 * j], o[2 * j + 1]);
                *reinterpret_cast<uint4*>(dst) = w;
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int oc = c0 + g * 8;
                    const int sub = oc >> 6;
                    const int chunk = (oc & 63) >> 3;
                    uint8_t* dst = smem + sub * SUB_TILE_BYTES + sw128_offset(row, chunk);
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = x[g * 8 + j];
                    if (p.n_res > 0) {
                        const uint4 r = *reinterpret_cast<const uint4*>(dst);
                        const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(rh[j]);
                            o[2 * j] += f.x;
                            o[2 * j + 1] += f.y;
                        }
                    }
                    if (r2_row) {
                        const uint4 r = *reinterpret_cast<const uint4*>(r2_row + oc0 + oc);
                        const __half2* rh = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(rh[j]);
                            o[2 * j] += f.x;
                            o[2 * j + 1] += f.y;
                        }
                    }
                    if (qs) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] *= __half2float(qs[oc0 + oc + j]);
                    }
                    uint4 w;
                    __half2* wh = reinterpret_cast<__half2*>(&w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) wh[j] = __floats2half2_rn(o[2 * j], o[2 * j + 1]);
                    *reinterpret_cast<uint4*>(dst) = w;
                }
            }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (warp == 2 && lane == 0) {
            for (int j = 0; j < n_sub; ++j) {
                tma_store_5d(&p.tm_c, smem + j * SUB_TILE_BYTES, oc0 + j * 64, opx, ox0, opy, oy0);
            }
            tma_store_commit();
            tma_store_wait_read0();
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host

static thread_local std::string g_err;
const char* gemm_last_error() { return g_err.c_str(); }

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

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box)
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
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
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
static int encode_act_map(CUtensorMap* m, const ActView& v, bool split2, bool linear, int bw, int bh)
{
    const uint64_t pb = static_cast<uint64_t>(v.pitch) * 2;
    uint64_t dims[5];
    uint64_t st[4];
    uint32_t box[5] = { 64, 1, static_cast<uint32_t>(bw), 1, static_cast<uint32_t>(bh) };
    if (linear) {
        const uint64_t M = static_cast<uint64_t>(v.W) * v.H;
        dims[0] = v.C; dims[1] = 1; dims[2] = M; dims[3] = 1; dims[4] = 1;
        st[0] = pb; st[1] = pb; st[2] = M * pb; st[3] = M * pb;
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
    return encode_map(m, v.ptr, 5, dims, st, box);
}

static int pick_block_n(int n_unit, bool chunk_add, long long m_tiles)
{
    if (chunk_add) return 256;
    int bn;
    if (n_unit % 256 == 0) bn = 256;
    else if (n_unit % 192 == 0) bn = 192;
    else if (n_unit % 128 == 0) bn = 128;
    else if (n_unit % 64 == 0) bn = 64;
    else return 0;
    // small problems: prefer more CTAs over wider tiles (148 SMs x 2 resident CTAs)
    while (bn > 64 && m_tiles * (n_unit / bn) < 148 && (bn % 2 == 0) && (n_unit % (bn / 2) == 0) &&
           ((bn / 2) % 64 == 0)) {
        bn /= 2;
    }
    return bn;
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
    if (op.kind == GEMM_TCONV2X2) {
        if (op.out.W != op.in.W * 2 || op.out.H != op.in.H * 2 || op.N != op.out.C * 4) {
            g_err = "gemm_plan: tconv geometry mismatch";
            return 1;
        }
    }
    // pixel tile geometry (over the GEMM-M pixel grid = output grid, except tconv = input grid)
    const int gw = (op.kind == GEMM_TCONV2X2) ? op.in.W : op.out.W;
    const int gh = (op.kind == GEMM_TCONV2X2) ? op.in.H : op.out.H;
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

    const int n_unit = (op.kind == GEMM_TCONV2X2) ? op.out.C : op.N;
    const int bn = pick_block_n(n_unit, op.chunk_add != 0, m_tiles);
    if (bn == 0 || op.N % bn != 0) { g_err = "gemm_plan: unsupported N"; return 1; }
    const int out_c_expected = op.chunk_add ? op.N / 4 : (op.kind == GEMM_TCONV2X2 ? op.N / 4 : op.N);
    if (op.out.C != out_c_expected) { g_err = "gemm_plan: out.C does not match N"; return 1; }
    op.block_n = bn;

    p.num_kblocks = taps * C / 64;
    p.kblk_per_tap = C / 64;
    p.act = op.act;
    p.chunk_add = op.chunk_add;
    p.bias = op.bias;
    p.qscale = op.qscale;
    p.phase_c = (op.kind == GEMM_TCONV2X2) ? op.out.C : 0;
    p.n_res = (op.res1.ptr ? 1 : 0) + (op.res2.ptr ? 1 : 0);
    if (op.res2.ptr && !op.res1.ptr) { g_err = "gemm_plan: res2 without res1"; return 1; }

    const bool in_split = (op.kind == GEMM_CONV3X3_S2 || op.kind == GEMM_CONV2X2_S2);
    if (encode_act_map(&p.tm_a, op.in, in_split, linear, p.bw, p.bh)) return 1;
    {
        const uint64_t Ktot = static_cast<uint64_t>(taps) * C;
        uint64_t dims[2] = { Ktot, static_cast<uint64_t>(op.N) };
        uint64_t st[1] = { Ktot * 2 };
        uint32_t box[2] = { 64, static_cast<uint32_t>(bn) };
        if (encode_map(&p.tm_b, op.weight, 2, dims, st, box)) return 1;
    }
    const bool out_split = (op.kind == GEMM_TCONV2X2);
    if (encode_act_map(&p.tm_c, op.out, out_split, linear, p.bw, p.bh)) return 1;
    if (op.res1.ptr) {
        if (op.res1.W != op.out.W || op.res1.H != op.out.H || op.res1.C != op.out.C) {
            g_err = "gemm_plan: residual geometry mismatch";
            return 1;
        }
        if (encode_act_map(&p.tm_r1, op.res1, out_split, linear, p.bw, p.bh)) return 1;
    }
    if (op.res2.ptr) {
        if (op.kind != GEMM_PW) { g_err = "gemm_plan: res2 only for 1x1"; return 1; }
        uint64_t* raw = reinterpret_cast<uint64_t*>(&p.tm_r2);
        raw[0] = reinterpret_cast<uint64_t>(op.res2.ptr);
        raw[1] = static_cast<uint64_t>(op.res2.pitch);
        // linear tiling: the epilogue sees a (W = M, H = 1) strip
        const uint64_t M = static_cast<uint64_t>(op.res2.W) * op.res2.H;
        raw[2] = M | (1ull << 32);
    }

    op.grid = dim3(op.N / bn, tiles_x, tiles_y);
    switch (bn) {
    case 64: op.smem = TileCfg<64>::SMEM_BYTES; op.stages = TileCfg<64>::STAGES; break;
    case 128: op.smem = TileCfg<128>::SMEM_BYTES; op.stages = TileCfg<128>::STAGES; break;
    case 192: op.smem = TileCfg<192>::SMEM_BYTES; op.stages = TileCfg<192>::STAGES; break;
    case 256: op.smem = TileCfg<256>::SMEM_BYTES; op.stages = TileCfg<256>::STAGES; break;
    }
    // staging tile (aliased over the pipeline stages) must fit
    const int out_cols = op.chunk_add ? bn / 4 : bn;
    const size_t staging = static_cast<size_t>((out_cols + 63) / 64) * SUB_TILE_BYTES;
    if (staging > op.smem - 1024 - 256 - bn * 2) { g_err = "gemm_plan: staging does not fit"; return 1; }
    op.planned = true;
    return 0;
}

template <int BN>
static cudaError_t set_attr()
{
    return cudaFuncSetAttribute(pw_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                TileCfg<BN>::SMEM_BYTES);
}

int gemm_init()
{
    static bool done = false;
    if (done) return 0;
    cudaError_t e = set_attr<64>();
    if (e == cudaSuccess) e = set_attr<128>();
    if (e == cudaSuccess) e = set_attr<192>();
    if (e == cudaSuccess) e = set_attr<256>();
    if (e != cudaSuccess) {
        g_err = std::string("cudaFuncSetAttribute(pw_gemm): ") + cudaGetErrorString(e);
        return 1;
    }
    done = true;
    return 0;
}

template <int BN>
static cudaError_t launch_bn(const GemmOp& op, cudaStream_t stream)
{
    pw_gemm_kernel<BN><<<op.grid, NUM_THREADS, op.smem, stream>>>(op.p);
    return cudaGetLastError();
}

int gemm_launch(const GemmOp& op, cudaStream_t stream)
{
    if (!op.planned) { g_err = "gemm_launch: op not planned"; return 1; }
    if (gemm_init()) return 1;
    cudaError_t e;
    switch (op.block_n) {
    case 64: e = launch_bn<64>(op, stream); break;
    case 128: e = launch_bn<128>(op, stream); break;
    case 192: e = launch_bn<192>(op, stream); break;
    case 256: e = launch_bn<256>(op, stream); break;
    default: g_err = "gemm_launch: bad block_n"; return 1;
    }
    if (e != cudaSuccess) {
        g_err = std::string("pw_gemm launch failed: ") + cudaGetErrorString(e);
        return 1;
    }
    return 0;
}

}  // namespace dcvc
