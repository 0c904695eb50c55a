// pw_gemm_ares.cu — the A-resident pw_gemm kernel for the 1x1 contractions (K <= 512), optionally on CTA pairs
// (tcgen05 cta_group::2).  Hand-written for sm_100a.
//
// Why: on B200 the L2 -> SM fabric moves distinct lines only about as fast as HBM does (measured ~7 TB/s
// chip-wide, tools/gemm_micro.py DCVC_B200_GEMM_DBG=3), so the streaming kernel, which re-reads the activation
// tile once per N tile and the weight tile once per pixel tile, is ingest-bound on every shape of the codec
// (M = 32640, N = K = 384: 125 MB through the fabric for 9.6 GFLOP).  Here
//   * a work item = (pixel super tile, group of N tiles); its activation tile [128 px][K] is loaded ONCE and
//     stays in smem for every N tile of the group (per k-block barriers: the MMAs start when k-block 0 lands,
//     and k-block kb is released for the next item as soon as the last N tile is done with it);
//   * only the weight k-blocks stream (ring of STAGES [BLOCK_N / CTAS][64] boxes);
//   * with CTAS == 2 two CTAs of a cluster (one TPC) form a pair: tcgen05.mma.cta_group::2 computes a
//     [256 px][BLOCK_N] tile from each CTA's own 128 activation rows and each CTA's HALF of the weight k-block,
//     so every SM ingests half the weight bytes; the accumulator halves live in each CTA's own TMEM and each CTA
//     runs its own epilogue (pw_gemm_epilogue.cuh).  The leader CTA (rank 0) owns the "full" barriers (the
//     peer's TMA loads credit them remotely) and issues every MMA; tcgen05.commit multicasts the "empty"
//     arrivals to both CTAs; the peer's epilogue warps hand accumulators back by remote mbarrier arrives.
// Warp roles as in pw_gemm.cu: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 epilogue
// (2 groups x 4 warps, one group per TMEM accumulator buffer).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "ptx.cuh"
#include "pw_gemm_internal.cuh"

namespace dcvc {

static constexpr int ARES_MAX_KB = 8;  // K <= 512

template <int BLOCK_N, bool CHUNK, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_ares_kernel(const __grid_constant__ PwGemmParams p)
{
    using Cfg = TileCfg<BLOCK_N>;
    constexpr int B_STAGE = (BLOCK_N / CTAS) * BLOCK_K * 2;  // this CTA's share of one weight k-block
    constexpr bool PAIR = CTAS == 2;
    const int STAGES = p.num_stages;
    const int nkb = p.num_kblocks;

    // carve-up of the 227 KB: [A: nkb x 16 KB] [B ring: STAGES x B_STAGE] [staging] ... [control]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* b_base = smem + nkb * A_STAGE_BYTES;
    uint8_t* staging = b_base + STAGES * B_STAGE;
    uint8_t* ctrl = smem + SMEM_USABLE;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);  // [8]   leader: both CTAs' A k-block landed
    uint64_t* a_empty = a_full + ARES_MAX_KB;              // [8]   every CTA: k-block free for the next item
    uint64_t* b_full = a_empty + ARES_MAX_KB;              // [8]   leader: both halves of the weight k-block landed
    uint64_t* b_empty = b_full + MAX_STAGES;               // [8]   every CTA: ring slot free
    uint64_t* tmem_full_bar = b_empty + MAX_STAGES;        // [2]   every CTA: accumulator complete
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]   leader: accumulator drained by all epilogue warps
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ctrl + 384);

    const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
    const int grp = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int num_grps = p.num_clusters;
    const int items = p.m_tiles * p.n_groups;
    const int tpg = p.tiles_per_group;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) trace_mark(p, 0);  // entry

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a);
        tma_prefetch_desc(&p.tm_b);
        tma_prefetch_desc(&p.tm_c);
        for (int i = 0; i < ARES_MAX_KB; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&tmem_full_bar[g], 1);
            mbar_init(&tmem_empty_bar[g], 8 * CTAS);  // one arrival per epilogue warp serving the buffer, both CTAs
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (PAIR) {
            tmem_alloc_2cta(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish_2cta();
        } else {
            tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / commit multicast
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();
    griddep_wait();
    if (threadIdx.x == 0) trace_mark(p, 1);  // setup done

    if (warp == 0) {
        if (elect_one_sync()) {
            // ------------------------------------------------------------ TMA producer (both CTAs)
            int bs = 0;
            uint32_t bph = 0, aph = 0;
            bool first = true;
            for (int item = grp; item < items; item += num_grps) {
                const int st = static_cast<int>(fdiv(item, p.fd_n_groups));
                const int ng = item - st * p.n_groups;
                const int row0 = st * (BLOCK_M * CTAS) + rank * BLOCK_M;
                for (int j = 0; j < tpg; ++j) {
                    const int nrow = (ng * tpg + j) * BLOCK_N + rank * (BLOCK_N / CTAS);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (j == 0) {
                            mbar_wait(&a_empty[kb], aph ^ 1);
                            if (PAIR) {
                                if (rank == 0) mbar_expect_tx(&a_full[kb], CTAS * A_STAGE_BYTES);
                                tma_load_2d_2sm(smem + kb * A_STAGE_BYTES, &p.tm_a, mapa_u32(smem_u32(&a_full[kb]), 0),
                                                kb * BLOCK_K, row0);
                            } else {
                                mbar_expect_tx(&a_full[kb], A_STAGE_BYTES);
                                tma_load_2d(smem + kb * A_STAGE_BYTES, &p.tm_a, &a_full[kb], kb * BLOCK_K, row0);
                            }
                        }
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(&b_full[bs], CTAS * B_STAGE);
                            tma_load_2d_2sm(b_base + bs * B_STAGE, &p.tm_b, mapa_u32(smem_u32(&b_full[bs]), 0),
                                            kb * BLOCK_K, nrow);
                        } else {
                            mbar_expect_tx(&b_full[bs], B_STAGE);
                            tma_load_2d(b_base + bs * B_STAGE, &p.tm_b, &b_full[bs], kb * BLOCK_K, nrow);
                        }
                        if (first) { trace_mark(p, 2); first = false; }  // first stage requested
                        if (++bs == STAGES) { bs = 0; bph ^= 1; }
                    }
                }
                aph ^= 1;
            }
            trace_mark(p, 3);  // last stage requested
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0 && elect_one_sync()) {
            // ------------------------------------------------------------ MMA issuer (leader CTA only)
            constexpr uint32_t idesc = make_idesc_f16_f32(BLOCK_M * CTAS, BLOCK_N);
            const bool do_mma = !(p.dbg & 1);
            int bs = 0;
            uint32_t bph = 0, aph = 0, t = 0;
            for (int item = grp; item < items; item += num_grps) {
                for (int j = 0; j < tpg; ++j, ++t) {
                    const int g = t & 1;
                    const uint32_t u = t >> 1;
                    if (PAIR) mbar_wait_cluster(&tmem_empty_bar[g], (u & 1) ^ 1);  // both CTAs drained this accumulator
                    else mbar_wait(&tmem_empty_bar[g], (u & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t acc = tmem_base + g * Cfg::ACC_COLS;
                    const bool last = (j == tpg - 1);
                    for (int kb = 0; kb < nkb; ++kb) {
                        if (j == 0) mbar_wait(&a_full[kb], aph);
                        mbar_wait(&b_full[bs], bph);
                        tcgen05_fence_after();
                        if (t == 0 && kb == 0) trace_mark(p, 4);  // first stage landed
                        const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem + kb * A_STAGE_BYTES));
                        const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(b_base + bs * B_STAGE));
                        if (do_mma) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                // advance 16 fp16 = 32 B inside the 128 B swizzle span: +2 in 16 B units
                                if (PAIR) umma_f16_ss_2cta(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                                else umma_f16_ss(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                        if (PAIR) {
                            umma_commit_2cta_mc(&b_empty[bs], 3);
                            if (last) umma_commit_2cta_mc(&a_empty[kb], 3);
                        } else {
                            umma_commit(&b_empty[bs]);
                            if (last) umma_commit(&a_empty[kb]);
                        }
                        if (++bs == STAGES) { bs = 0; bph ^= 1; }
                    }
                    if (PAIR) umma_commit_2cta_mc(&tmem_full_bar[g], 3);
                    else umma_commit(&tmem_full_bar[g]);
                    if (t == 0) trace_mark(p, 5);  // first tile issued
                }
                aph ^= 1;
            }
            trace_mark(p, 6);  // last tile issued
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue (2 groups x 4 warps, both CTAs)
        EpiWarp ew;
        ew.q = warp & 3;
        ew.lane = lane;
        ew.b = ((warp - 2) >> 2) & 1;
        ew.h = (warp - 2) >> 3;
        ew.slabs = p.staging_bufs;
        ew.slab = staging + (warp - 2) * p.staging_bufs * EPI_SLAB_BYTES;
        ew.cnt = 0;
        const int g = ew.b;
        const uint32_t empty_remote = PAIR ? mapa_u32(smem_u32(&tmem_empty_bar[g]), 0) : 0u;
        const uint32_t acc = tmem_base + g * Cfg::ACC_COLS + (static_cast<uint32_t>(ew.q * 32) << 16);
        uint32_t t = 0;
        for (int item = grp; item < items; item += num_grps) {
            const int st = static_cast<int>(fdiv(item, p.fd_n_groups));
            const int ng = item - st * p.n_groups;
            TileCoord tc;
            tc.ox0 = st * (BLOCK_M * CTAS) + rank * BLOCK_M;
            tc.oy0 = 0;
            tc.opx = 0;
            tc.opy = 0;
            for (int j = 0; j < tpg; ++j, ++t) {
                if (static_cast<int>(t & 1) != g) continue;
                tc.n0 = (ng * tpg + j) * BLOCK_N;
                tc.oc0 = CHUNK ? tc.n0 / 4 : tc.n0;
                epilogue_tile<BLOCK_N, CHUNK>(p, tc, acc, &tmem_full_bar[g], (t >> 1) & 1, &tmem_empty_bar[g], empty_remote,
                                              ew, t == 0 ? 7 : (t == 1 ? 9 : 11));
                if (lane == 0 && ew.q == 0 && ew.h == 0) trace_mark(p, t == 0 ? 8 : (t == 1 ? 10 : 12));
            }
        }
        if (lane == 0) tma_store_wait_read<0>();
        if (lane == 0 && ew.q == 0 && ew.h == 0) trace_mark(p, 13 + g);  // group drained
        __syncwarp();
    }

    __syncthreads();
    if (PAIR) cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer can still signal its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host

template <int BN, bool CHUNK, int CTAS>
static cudaError_t ares_set_attr()
{
    return cudaFuncSetAttribute(pw_gemm_ares_kernel<BN, CHUNK, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
}

template <int CTAS>
static cudaError_t ares_set_attr_all()
{
    cudaError_t e = ares_set_attr<64, false, CTAS>();
    if (e == cudaSuccess) e = ares_set_attr<128, false, CTAS>();
    if (e == cudaSuccess) e = ares_set_attr<192, false, CTAS>();
    if (e == cudaSuccess) e = ares_set_attr<256, false, CTAS>();
    if (e == cudaSuccess) e = ares_set_attr<256, true, CTAS>();
    return e;
}

int ares_init()
{
    static bool done = false;
    if (done) return 0;
    cudaError_t e = ares_set_attr_all<1>();
    if (e == cudaSuccess) e = ares_set_attr_all<2>();
    if (e != cudaSuccess) {
        gemm_set_error(std::string("cudaFuncSetAttribute(pw_gemm_ares): ") + cudaGetErrorString(e));
        return 1;
    }
    done = true;
    return 0;
}

// co-resident CTA pairs (a pair needs both SMs of one TPC)
static int max_active_pairs(int num_sms)
{
    static int cached = 0;
    if (cached) return cached;
    if (ares_init()) return num_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * 64, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, pw_gemm_ares_kernel<256, false, 2>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms / 2;
    }
    cached = n;
    return n;
}

int ares_plan(GemmOp& op, int num_sms)
{
    // opt-outs (debugging / A-B measurements): DCVC_B200_GEMM_ARES=0 keeps every op on the streaming kernel,
    // DCVC_B200_GEMM_PAIR=0 runs the A-resident kernel on single CTAs
    static const int env_ares = []() { const char* e = getenv("DCVC_B200_GEMM_ARES"); return e ? atoi(e) : 0; }();
    static const int env_pair = []() { const char* e = getenv("DCVC_B200_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    if (!env_ares) return 1;
    if (op.kind != GEMM_PW) return 1;
    const int C = op.in.C;
    const int nkb = C / 64;
    if (nkb > ARES_MAX_KB) return 1;
    const int ctas = env_pair ? 2 : 1;
    // N tile: the largest of 256 / 192 / 128 / 64 that divides N
    int bn = 0;
    const int cand[4] = { 256, 192, 128, 64 };
    int force_bn = 0;
    if (const char* f = getenv("DCVC_B200_GEMM_BN")) force_bn = atoi(f);
    for (int i = 0; i < 4 && !bn; ++i) {
        if (op.N % cand[i]) continue;
        if (op.chunk_add && cand[i] != 256) continue;
        if (force_bn && cand[i] != force_bn) continue;
        bn = cand[i];
    }
    if (!bn) return 1;
    const int n_tiles = op.N / bn;
    const long long M = static_cast<long long>(op.out.W) * op.out.H;
    const int rows = BLOCK_M * ctas;
    const int s_tiles = static_cast<int>((M + rows - 1) / rows);
    const int max_grps = ctas == 2 ? max_active_pairs(num_sms) : num_sms;

    // smem carve-up
    const int sub_bytes = SUB_TILE_BYTES;
    const int b_stage = bn / ctas * BLOCK_K * 2;
    int staging_bufs = 2;
    int stages = (SMEM_USABLE - nkb * A_STAGE_BYTES - EPI_GROUPS * staging_bufs * sub_bytes) / b_stage;
    if (stages < 4) {
        staging_bufs = 1;
        stages = (SMEM_USABLE - nkb * A_STAGE_BYTES - EPI_GROUPS * staging_bufs * sub_bytes) / b_stage;
    }
    if (const char* f = getenv("DCVC_B200_GEMM_STAGING")) {
        staging_bufs = atoi(f) == 1 ? 1 : 2;
        stages = (SMEM_USABLE - nkb * A_STAGE_BYTES - EPI_GROUPS * staging_bufs * sub_bytes) / b_stage;
    }
    if (stages < 3) return 1;
    if (stages > MAX_STAGES) stages = MAX_STAGES;

    // N groups: a work item keeps its activation tile for n_tiles / n_groups N tiles.  More groups = more items to
    // spread over the SMs (small M) at the price of re-reading the activation tile once per group.
    int best_ng = 1;
    double best_cost = 1e30;
    int force_ng = 0;
    if (const char* f = getenv("DCVC_B200_GEMM_NGROUPS")) force_ng = atoi(f);
    for (int ng = 1; ng <= n_tiles; ++ng) {
        if (n_tiles % ng) continue;
        if (force_ng && ng != force_ng) continue;
        const long long it = static_cast<long long>(s_tiles) * ng;
        const long long grps = it < max_grps ? it : max_grps;
        const long long rounds = (it + grps - 1) / grps;
        const double t_a = nkb * A_STAGE_BYTES / 40.0;                        // clocks to ingest the activation tile
        const double t_n_mma = nkb * 4.0 * (bn / 2.0);                        // clocks of tensor pipe per N tile
        const double t_n_ld = nkb * static_cast<double>(b_stage) / 40.0;      // clocks to ingest its weights
        const double t_n = t_n_mma > t_n_ld ? t_n_mma : t_n_ld;
        const double cost = rounds * (t_a + (n_tiles / ng) * t_n) + 1500.0;   // + fixed prologue / drain
        if (cost < best_cost) {
            best_cost = cost;
            best_ng = ng;
        }
    }

    PwGemmParams& p = op.p;
    p.ares_ctas = ctas;
    p.bw = BLOCK_M;
    p.bh = 1;
    p.num_kblocks = nkb;
    p.kblk_per_tap = nkb;
    p.act = op.act;
    p.chunk_add = op.chunk_add;
    p.bias = op.bias;
    p.qscale = op.qscale;
    p.phase_c = 0;
    p.n_res = (op.res1.ptr ? 1 : 0) + (op.res2.ptr ? 1 : 0);
    if (encode_act_map(&p.tm_a, op.in, false, true, true, BLOCK_M, 1)) return 2;
    {
        uint64_t dims[2] = { static_cast<uint64_t>(C), static_cast<uint64_t>(op.N) };
        uint64_t st[1] = { static_cast<uint64_t>(C) * 2 };
        uint32_t box[2] = { 64, static_cast<uint32_t>(bn / ctas) };
        if (encode_map(&p.tm_b, op.weight, 2, dims, st, box)) return 2;
    }
    p.epi_rows_y = 1;
    if (encode_act_map(&p.tm_c, op.out, false, true, true, 32, 1, 32)) return 2;
    if (op.res1.ptr) {
        p.r1 = static_cast<const __half*>(op.res1.ptr);
        p.r1_pitch = op.res1.pitch;
        if (op.res2.ptr) {
            p.r2 = static_cast<const __half*>(op.res2.ptr);
            p.r2_pitch = op.res2.pitch;
        }
        p.res_w = static_cast<int>(M);
        p.res_h = 1;
    }
    p.n_tiles = n_tiles;
    p.n_groups = best_ng;
    p.tiles_per_group = n_tiles / best_ng;
    p.m_tiles = s_tiles;
    p.tiles_x = s_tiles;
    p.total_tiles = s_tiles * n_tiles;
    p.cluster = ctas;
    {
        const long long it = static_cast<long long>(s_tiles) * best_ng;
        p.num_clusters = static_cast<int>(it < max_grps ? it : max_grps);
    }
    p.linear = 1;
    p.b_resident = 0;
    p.num_stages = stages;
    p.staging_bufs = staging_bufs;
    p.fd_n_tiles = make_fastdiv(p.n_tiles);
    p.fd_n_groups = make_fastdiv(p.n_groups);
    p.fd_tiles_x = make_fastdiv(p.tiles_x);
    p.fd_phase_c = make_fastdiv(1);
    p.fd_bw = make_fastdiv(p.bw);
    if (const char* d = getenv("DCVC_B200_GEMM_DBG")) p.dbg = atoi(d);
    if (const char* d = getenv("DCVC_B200_GEMM_TRACE")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(d, nullptr, 0));
    op.block_n = bn;
    op.ares = ctas;
    op.stages = stages;
    op.grid = dim3(p.num_clusters * ctas, 1, 1);
    op.smem = SMEM_TOTAL;
    op.planned = true;
    return 0;
}

template <int BN, bool CHUNK, int CTAS>
static cudaError_t ares_launch_bn(const GemmOp& op, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = op.grid;
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = op.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (gemm_pdl_enabled() && op.pdl) ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, pw_gemm_ares_kernel<BN, CHUNK, CTAS>, op.p);
}

template <int CTAS>
static cudaError_t ares_launch_ctas(const GemmOp& op, cudaStream_t stream)
{
    switch (op.block_n + (op.chunk_add ? 1 : 0)) {
    case 64: return ares_launch_bn<64, false, CTAS>(op, stream);
    case 128: return ares_launch_bn<128, false, CTAS>(op, stream);
    case 192: return ares_launch_bn<192, false, CTAS>(op, stream);
    case 256: return ares_launch_bn<256, false, CTAS>(op, stream);
    case 257: return ares_launch_bn<256, true, CTAS>(op, stream);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t ares_launch(const GemmOp& op, cudaStream_t stream)
{
    return op.ares == 2 ? ares_launch_ctas<2>(op, stream) : ares_launch_ctas<1>(op, stream);
}

}  // namespace dcvc
