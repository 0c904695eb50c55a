// Elementwise / layout / entropy-parameter kernels of the DCVC-UF hot path (sm_100a).
// HBM-bound byte and half work: 16-byte vector accesses, one warp per latent pixel for the
// entropy-parameter path (warp-ballot counts instead of the reference's 3-kernel block scan).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_gemm.cuh"

namespace dcvc {

// depthwise 3x3, pad 1, no bias (bias is folded into the next 1x1: layers_proxy.cpp:175-178).
// w: [9][C] fp16 (tap-major).
int launch_dw3x3(const ActView& in, const ActView& out, const __half* w, cudaStream_t s, bool pdl = true);

// replicate-pad to x8 + pixel_unshuffle(8): x[1,Cs,H,W] (element strides sc,sh,sw) -> [H8][W8][Cs*64]
// (reference pad_and_unshuffle_8_kernel, cat_and_pad.cu:7-51)
int launch_unshuffle8_pad(const __half* x, int Cs, int H, int W, long long sc, long long sh,
                          long long sw, const ActView& out, cudaStream_t s);

// pixel_shuffle(8) + clamp(+-0.5): [H8][W8][Cs*64] -> NHWC [H8*8][W8*8][Cs]
// (reference pixel_shuffle_8_out_3_kernel<true>, shuffle.cu:75-121)
int launch_shuffle8_clamp(const ActView& in, __half* out, int Cs, int clamp, cudaStream_t s);

// out(y,x,:) = in(min(y,in.H-1), min(x,in.W-1), :)  — replicate_pad / slice (cat_and_pad.cu:53-110)
int launch_pad_crop(const ActView& in, const ActView& out, cudaStream_t s);

// out = half(in * q[c])  (multiply_with_broadcast, stream.cu:484-546)
int launch_scale_channels(const ActView& in, const __half* q, const ActView& out, cudaStream_t s);

// z -> clamp(round_half_away(z), -64, 63) as fp16 and int8 (round_z_kernel, stream.cu:862-884)
int launch_round_z(const __half* z, __half* z_hat, int8_t* z_i8, long long n, cudaStream_t s);
int launch_int8_to_half(const int8_t* x, __half* out, long long n, cudaStream_t s);

struct EntropyStepArgs {
    // latent geometry: [H][W][C], C = 4 * G channels in 4 groups
    int H = 0, W = 0, G = 64;
    int step = 0;             // 0..3  (mask_k of common_model.py:174-195)
    const __half* y = nullptr;        // [H][W][y_pitch]   analysis output (encoder only)
    int y_pitch = 0;
    const __half* q_enc = nullptr;    // [C] per-channel scale applied to y first (or nullptr)
    const __half* scales = nullptr;   // [H][W][p_pitch]
    const __half* means = nullptr;    // [H][W][p_pitch]
    int p_pitch = 0;
    __half* y_hat_acc = nullptr;      // [H][W][acc_pitch] running y_hat_so_far
    int acc_pitch = 0;
    float skip_thres = 0.f;
    const uint8_t* scale_lut = nullptr;  // 65536-entry fp16-bits -> table index LUT
    // --- chunk-codec (HT-S / LD) variants: per-element quantiser step, means in their own buffer,
    //     symbols built once over all 4*G channels after the four means-refinement steps
    const __half* q_div = nullptr;    // [H][W][q_pitch]: y *= 1 / max(q_div, 0.5) first (stream.cu:422-443)
    int q_pitch = 0;
    int m_pitch = 0;                  // pitch of `means` (0: same as p_pitch)
    int8_t* yq_dense = nullptr;       // [H*W][4G] quantised latents, written (enc) / read (dec restore)
    int full = 0;                     // 1: index/compaction kernels run over all 4*G channels of a pixel
    int ng = 4;                       // channel groups per pixel: 4 (4-step mask) or 2 (two-step checkerboard of the LD model, C = 2*G)
    // outputs, per pixel row of G entries
    int16_t* sym_raw = nullptr;       // encoder: (sym << 8) + idx, uncompacted [H*W][G]
    uint8_t* idx_raw = nullptr;       // decoder: idx, uncompacted [H*W][G]
    int32_t* counts = nullptr;        // [H*W] number of coded (non-skipped) entries
};

// Parameter block of the entropy kernels as it is passed to the device (built from EntropyStepArgs by the launchers;
// declared here because the host-side kernel emulation of tests/cpp reads the same block).
struct EntropyDev {
    const __half* qdiv; int q_pitch; int m_pitch; int8_t* yq; int full; int ng;
    int H, W, G, step;
    const __half* y; int y_pitch;
    const __half* q_enc;
    const __half* scales; const __half* means; int p_pitch;
    __half* acc; int acc_pitch;
    __half thres;
    const uint8_t* lut;
    int16_t* sym_raw; uint8_t* idx_raw; int32_t* counts;
};

// encoder step: process_with_mask + 4->1 fold + build_index_enc + per-pixel count
// (stream.cu:548-630, 931-949, 130-161)
int launch_entropy_enc_step(const EntropyStepArgs& a, cudaStream_t s);
// decoder step part 1: 4->1 fold of scales + build_index_dec + per-pixel count (stream.cu:896-915, 89-115)
int launch_entropy_dec_index(const EntropyStepArgs& a, cudaStream_t s);
// exclusive scan of per-pixel counts -> offsets[n+1]; total also written to *total
int launch_scan_counts(const int32_t* counts, int32_t* offsets, int32_t* total, int n, cudaStream_t s);
// stream compaction (conditional_index_step3, stream.cu:261-282): keep entries whose scale > skip
int launch_compact_i16(const EntropyStepArgs& a, const int32_t* offsets, int16_t* out, cudaStream_t s);
int launch_compact_u8(const EntropyStepArgs& a, const int32_t* offsets, uint8_t* out, cudaStream_t s);
// decoder step part 2: scatter decoded symbols + restore y_hat (stream.cu:359-383, 756-818)
int launch_entropy_dec_restore(const EntropyStepArgs& a, const int32_t* offsets,
                               const int8_t* decoded, cudaStream_t s);

// chunk codecs: symbols of the whole latent from the dense y_q (build_index_enc_cuda on the full tensor,
// dmc_hts_proxy.cpp:559-562): a.full = 1, a.G = channels per pixel
int launch_entropy_build_symbols_full(const EntropyStepArgs& a, cudaStream_t s);
// conditional_recover_with_type_conversion (stream.cu:359-383) into the dense int8 y_q
int launch_entropy_recover_dense(const EntropyStepArgs& a, const int32_t* offsets, const int8_t* decoded, cudaStream_t s);
// restore_y / restore_y_and_add_inplace from the dense y_q (stream.cu:685-754)
int launch_entropy_restore_dense(const EntropyStepArgs& a, cudaStream_t s);
// out = half(in * max(q, 0.5)) elementwise (the final multiply of stream.cu:605-616 / :717-725)
int launch_mul_clamp_min(const ActView& in, const ActView& q, const ActView& out, cudaStream_t s);

// Host: fp16-bits -> Gaussian scale-table index, built with the reference's half arithmetic
// (scale_to_index, stream.cu:77-87 + def_const.h:6-12).
void build_scale_lut(uint8_t* lut65536);

}  // namespace dcvc
