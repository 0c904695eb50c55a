// pw_gemm_internal.cuh — constants and host helpers shared by the pw_gemm translation units
// (pw_gemm.cu: the per-op kernel + planner; dcb_tail.cu: the fused DepthConvBlock tail).
#pragma once
#include <string>

#include "pw_gemm.cuh"
#include "pw_gemm_epilogue.cuh"

namespace dcvc {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;
static constexpr int UMMA_K = 16;
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
static constexpr int EPI_WARPS = 16;                          // 2 accumulator buffers x 2 column halves x 4 lane quarters
static constexpr int SUB_TILE_BYTES = EPI_WARPS / 2 * EPI_SLAB_BYTES;  // staging per slab index and accumulator buffer (16 KB)
static constexpr int EPI_GROUPS = 2;                          // one per TMEM accumulator buffer
static constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;       // TMA warp + MMA warp + 16 epilogue warps

static constexpr int MAX_STAGES = 8;
static constexpr int SMEM_TOTAL = 232448;   // 227 KB: the whole SM, one persistent CTA per SM
// control block at the end of the carve-up: barriers, tmem pointer
static constexpr int CTRL_BYTES = 512;
static constexpr int SMEM_USABLE = SMEM_TOTAL - 1024 /*alignment slack*/ - CTRL_BYTES;

template <int BLOCK_N>
struct TileCfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int ACC_COLS = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128 ? 128 : 256);
    static constexpr int TMEM_COLS = 2 * ACC_COLS;  // double-buffered accumulator
};

// ---- host helpers (pw_gemm.cu)
void gemm_set_error(const std::string& e);
int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);
// 5-D (or, for 1x1 ops, 2-D) tensor map of an NHWC view; split2 exposes the 2x2 pixel phases as dims 1 and 3
int encode_act_map(CUtensorMap* m, const ActView& v, bool split2, bool linear, bool lin2d, int bw, int bh, int box_c = 64);
FastDiv make_fastdiv(uint32_t d);
bool gemm_pdl_enabled();

}  // namespace dcvc
