// dcb_tail: the pixel-local run of a DepthConvBlock as ONE persistent kernel on CTA pairs.
//
// A DepthConvBlock (reference: src/layers/layers.py:152-159 == layers_proxy.cpp:71-101) is
//     t1 = wsilu(dc.0(x));  t2 = dw3x3(t1);  o = dc.3(t2) + x;  t1' = chunk_add(wsilu(ffn.0(o)));
//     y  = ffn.2(t1') + o (+ x) (* q)
// Only the depthwise 3x3 needs neighbouring pixels.  Everything behind it — dc.3 -> ffn.0 -> ffn.2 and, when the next
// op of the network is another block's dc.0, that GEMM too — is pixel-local, so one CTA can run a 128-pixel tile
// through all four GEMMs without the intermediates (o, t1') ever leaving the SM:
//     phase 1  o   = t2 . W3^T + b3 + x              A = t2 tile in smem (TMA),            out -> TMEM (fp16, packed)
//     phase 2  t1' = fold4(wsilu(o . Wf0^T + bf0))   A = o in TMEM (tcgen05.mma A-from-TMEM), out -> smem (UMMA layout)
//     phase 3  y   = (t1' . Wf2^T + bf2 + o [+ x]) [* q]   A = t1' in smem, residual o from TMEM, out -> TMEM + global
//     phase 4  t1n = wsilu(y . W0n^T + b0n)          A = y in TMEM,                        out -> global
// Two CTAs of one TPC form a pair (tcgen05 cta_group::2, M = 256): each holds its own 128 pixel rows and HALF of every
// weight tile, so the weights — the only operand that still streams from L2, 1-2 MB per tile — are ingested once per 256
// pixels.  Measured motivation (profiles/r2_gemm_steady_state.md): the per-op kernels are bound by L2 -> SM ingest
// (the loads alone of the N = K = 384 GEMM take 11.8 of its 17.7 us) and by ~6 us of fill / drain per launch.
#pragma once
#include "pw_gemm.cuh"

namespace dcvc {

struct alignas(64) DcbTailParams {
    CUtensorMap tm_a;      // t2   [M][inner]    load box {64 ch, 128 rows}, SWIZZLE_128B
    CUtensorMap tm_w[4];   // W_i  [N_i][K_i]    load box {64 k, 64 rows}, SWIZZLE_128B (this CTA's half of a 128-column chunk)
    CUtensorMap tm_y;      // y    [M][C]        store box {32 ch, 32 rows}, SWIZZLE_64B
    CUtensorMap tm_t;      // t1n  [M][inner_n]  store box {32 ch, 32 rows}, SWIZZLE_64B
    const __half* bias[4]; // per GEMM column (nullptr: none)
    const __half* qscale;  // [C] or nullptr
    const __half* x;       // block input (residual of dc.3, and of ffn.2 when `shortcut`)
    int x_pitch;
    int shortcut;
    int M;                 // pixels
    int C, inner, inner_next;
    int nkb[4];            // k-blocks (64 channels) per phase
    int nch[4];            // 128-column chunks per phase (nch[3] == 0: no phase 4)
    int tiles;             // ceil(M / 256)
    int num_pairs;         // gridDim.x / 2
    int kbs[4];            // k-blocks per weight stage (one ring slot / barrier round trip, kbs TMA requests), per phase
    int nst[4];            // stages per chunk = nkb / kbs
    int stage_bytes;       // ring slot size = max kbs * 8 KB
    int stages;            // weight ring depth
    int p_bytes;           // bytes of the resident activation buffer (inner / 64 * 16 KB)
    int dbg;
    unsigned long long* trace;   // env DCVC_B200_GEMM_TRACE=<device address of 2048 u64>: timeline of CTA 0 (tools/dcb_tail_trace.py)
};

struct DcbTailOp {
    ActView t2, x, y, t1n;         // t1n.ptr == nullptr: no phase 4
    const __half *w3 = nullptr, *b3 = nullptr, *wf0 = nullptr, *bf0 = nullptr, *wf2 = nullptr, *bf2 = nullptr;
    const __half *w0n = nullptr, *b0n = nullptr;
    const __half* qscale = nullptr;
    bool shortcut = false;
    bool pdl = true;
    // ---- derived
    DcbTailParams p;
    dim3 grid;
    bool planned = false;
};

// 0: planned; 1: shape not eligible (caller keeps the per-op kernels); 2: error (gemm_last_error())
int dcb_tail_plan(DcbTailOp& op);
int dcb_tail_launch(const DcbTailOp& op, cudaStream_t stream);
int dcb_tail_init();

}  // namespace dcvc
