// pw_gemm: the dense-contraction kernel family of the DCVC-UF hot path.
//
// One kernel covers every GEMM-shaped op of the reference's CUTLASS directory
// (SURVEY.md §2.2): conv1x1_bias{,_wsilu,_shortcut,_shortcut2,_with_quant,
// _shortcut_with_quant,_wsilu_chunk_add}, conv_bias (3x3/s2, 2x2/s2) and
// transposed_conv (2x2/s2), as  Y[pix, n] = epi( sum_taps X[pix+tap, :] . W[n, tap, :] ).
//
// Layout: activations fp16 NHWC with an arbitrary channel pitch (cat-buffer slices),
// weights fp16 [N][taps*C] K-major, fp32 accumulation in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcvc {

// NHWC activation view: element (y, x, c) lives at ptr[(y*W + x)*pitch + c].
struct ActView {
    const void* ptr = nullptr;
    int C = 0;      // channels visible through this view
    int pitch = 0;  // channel pitch in elements (>= C, multiple of 8)
    int W = 0;
    int H = 0;
};

enum GemmKind : int {
    GEMM_PW = 0,         // 1x1, stride 1
    GEMM_CONV3X3_S2 = 1, // 3x3, stride 2, pad 1     (reference conv_bias, image_model.py:62)
    GEMM_CONV2X2_S2 = 2, // pixel_unshuffle(2)+1x1   (reference conv_bias, layers_proxy.cpp:260-266)
    GEMM_TCONV2X2 = 3,   // 1x1 + pixel_shuffle(2)   (reference transposed_conv, layers_proxy.cpp:314-323)
    GEMM_CONV3X3_PS2 = 4, // 3x3, stride 1, pad 1 + pixel_shuffle(2): the HT-L SubpelConv2x (video_model_ht.py:41,
                          // layers_proxy.cpp:271-275)
};

// ACT_GDN / ACT_IGDN: out = res1 * rsqrt(acc + bias) resp. res1 * sqrt(acc + bias) — generalised divisive normalisation
// (DCVC-family/DCVC/src/layers/gdn.py:52-67) as an epilogue of the 1x1 GEMM over x^2; res1 carries x
enum GemmAct : int { ACT_NONE = 0, ACT_WSILU = 1, ACT_GDN = 2, ACT_IGDN = 3 };

// x / d for x < 2^31 as mulhi + add + shift (the kernel's scheduling loops run on single threads: a hardware
// integer division is ~40 dependent instructions there)
struct FastDiv {
    uint32_t mul, shr;
};

struct alignas(64) PwGemmParams {
    CUtensorMap tm_a;
    CUtensorMap tm_b;
    CUtensorMap tm_c;
    const __half* bias;    // [N] (GEMM columns, i.e. before chunk-add) or nullptr
    const __half* qscale;  // [N_out] per-output-channel multiplier or nullptr
    const __half* r1;      // residual operands: same pixel grid as the output, own channel pitch
    const __half* r2;
    int r1_pitch, r2_pitch;
    int res_w, res_h;      // output pixel grid seen by the tile scheduler (linear tiling: W = M, H = 1)
    int n_tiles;           // N tiles
    int tiles_x;           // pixel tiles along x
    int total_tiles;       // n_tiles * tiles_x * tiles_y
    int m_tiles;           // tiles_x * tiles_y
    int num_stages;        // pipeline depth (A + B stages)
    int staging_bufs;      // staging slabs per epilogue warp (1 or 2)
    int linear;            // 1x1: tm_a / tm_c are plain 2-D [pixels][channels] maps (a 5-D box costs the TMA unit ~4
                           // cycles per row to walk, measured 0.25 us per 128-row k-block; 2-D rows stream)
    int dbg;               // micro-benchmark switches (env DCVC_B200_GEMM_DBG): 1 = no MMA, 2 = no epilogue body
    unsigned long long* trace;  // env DCVC_B200_GEMM_TRACE=<device address>: 16 globaltimer slots per CTA (tools/gemm_trace.py)
    int num_kblocks;       // taps * C / 64
    int kblk_per_tap;      // C / 64
    int bw, bh;            // pixel tile, bw*bh == 128
    int epi_rows_y;        // 32 / bw: image rows covered by one epilogue warp's 32 pixels (5-D tiles)
    int act;               // GemmAct
    int chunk_add;         // 1: out[:, j] = sum_{i<4} act(acc[:, 4j+i])
    int n_res;             // 0,1,2 residual operands (same geometry as the output)
    int phase_c;           // tconv: output channels per 2x2 phase (0: not a tconv)
    FastDiv fd_n_tiles, fd_tiles_x, fd_phase_c, fd_bw;
    int8_t tap_px[9];
    int8_t tap_py[9];
    int8_t tap_dx[9];
    int8_t tap_dy[9];
};

struct GemmOp {
    int kind = GEMM_PW;
    ActView in, out, res1, res2;
    const __half* weight = nullptr;  // packed [N][Ktot], Ktot = taps * in.C
    const __half* bias = nullptr;
    const __half* qscale = nullptr;
    int N = 0;          // GEMM columns (tconv: 4*Cout, chunk-add: 4*C')
    int act = ACT_NONE;
    int chunk_add = 0;
    bool pdl = true;    // launch with the programmatic-dependent-launch attribute (when DCVC_B200_PDL allows it at all)
    // ---- derived by gemm_plan()
    PwGemmParams p;
    dim3 grid;
    int block_n = 0;
    int stages = 0;
    size_t smem = 0;
    bool planned = false;
};

// Fills op.p / grid / block_n; returns cudaSuccess or an error (message in gemm_last_error()).
int gemm_plan(GemmOp& op);
int gemm_launch(const GemmOp& op, cudaStream_t stream);
const char* gemm_last_error();
// one-time kernel attribute setup (dynamic smem opt-in); call outside stream capture
int gemm_init();

}  // namespace dcvc
