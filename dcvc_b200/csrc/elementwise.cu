// elementwise.cu — see elementwise.cuh.  All kernels: fp16 NHWC, 16-byte vector accesses on the
// channel axis, grid sized from the element count (multiples of full 256-thread blocks).
#include "elementwise.cuh"

#include <stdlib.h>
#include <string.h>

#include <math.h>
#include <stdio.h>

namespace dcvc {

#define DCVC_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) {                            \
            fprintf(stderr, "dcvc kernel launch failed at %s:%d: %s\n", __FILE__, __LINE__, \
                    cudaGetErrorString(e__));                \
            return 1;                                        \
        }                                                    \
    } while (0)

static inline int blocks_for(long long n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------- dw3x3
// One thread = 8 channels x a vertical strip of DW_TY output rows: every input row of the strip is
// loaded once per thread (3 x 16 B, the x-1 / x+1 neighbours hit L1 because adjacent threads of the
// CTA load them as their centre) and scattered into the <= 3 output rows it contributes to.  HBM/L2
// traffic per output drops from ~3 input rows to (DW_TY + 2) / DW_TY.
static constexpr int DW_TY_DEFAULT = 4;

// d = a * b + c with fp16 a, b and fp32 c, d in one instruction (FHFMA)
__device__ __forceinline__ float fma_f32_f16(uint16_t a, uint16_t b, float c)
{
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
    return c;
}

__device__ __forceinline__ void dw_load_row(const __half* in, int in_pitch, int W, int H, int x, int yy, int c,
                                            uint4 (&q)[3])
{
    q[0] = q[1] = q[2] = make_uint4(0, 0, 0, 0);
    if (yy < 0 || yy >= H) return;
    const __half* rowp = in + (static_cast<long long>(yy) * W + x) * in_pitch + c;
    if (x > 0) q[0] = *reinterpret_cast<const uint4*>(rowp - in_pitch);
    q[1] = *reinterpret_cast<const uint4*>(rowp);
    if (x + 1 < W) q[2] = *reinterpret_cast<const uint4*>(rowp + in_pitch);
}

template <int DW_TY>
__global__ void __launch_bounds__(128)
dw3x3_kernel(const __half* __restrict__ in, int in_pitch, __half* __restrict__ out, int out_pitch,
             const __half* __restrict__ w, int C, int W, int H)
{
    // programmatic dependent launch (the kernel sits between two GEMMs of a block): the next kernel may start its
    // prologue now; this one waits here for the GEMM whose output it reads
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int cg_n = C >> 3;
    const int row_items = W * cg_n;  // (x, channel-group) pairs of one row
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int strips = (H + DW_TY - 1) / DW_TY;
    if (tid >= static_cast<long long>(row_items) * strips) return;
    const int item = static_cast<int>(tid % row_items);
    const int strip = static_cast<int>(tid / row_items);
    const int cg = item % cg_n;
    const int x = item / cg_n;
    const int c = cg << 3;
    const int y0 = strip * DW_TY;

    // all DW_TY + 2 input rows are requested up front (18 independent 16-byte loads in flight per thread)
    uint4 rows[DW_TY + 2][3];
#pragma unroll
    for (int r = 0; r < DW_TY + 2; ++r) dw_load_row(in, in_pitch, W, H, x, y0 + r - 1, c, rows[r]);
    uint4 wk[9];  // packed fp16 taps, converted on use
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = __ldg(reinterpret_cast<const uint4*>(w + k * C + c));

    // The kernel was instruction-bound with per-element conversions + fp32 FMAs (~28 instr / output element).
    // sm_100 has a mixed-precision FMA (PTX fma.rn.f32.f16 -> SASS FHFMA, with .H0/.H1 operand selectors): fp16
    // operands straight out of the packed registers, fp32 accumulate — 9 instructions per output element and
    // bit-identical to fmaf(half2float(v), half2float(k), acc), i.e. the reference's fp32 accumulation.
#pragma unroll
    for (int r = 0; r < DW_TY; ++r) {
        const int y = y0 + r;
        if (y >= H) break;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const uint32_t* vh = reinterpret_cast<const uint32_t*>(&rows[r + ky][kx]);
                const uint32_t* kh = reinterpret_cast<const uint32_t*>(&wk[ky * 3 + kx]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[2 * i] = fma_f32_f16(static_cast<uint16_t>(vh[i] & 0xffffu), static_cast<uint16_t>(kh[i] & 0xffffu), acc[2 * i]);
                    acc[2 * i + 1] = fma_f32_f16(static_cast<uint16_t>(vh[i] >> 16), static_cast<uint16_t>(kh[i] >> 16), acc[2 * i + 1]);
                }
            }
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
        *reinterpret_cast<uint4*>(out + (static_cast<long long>(y) * W + x) * out_pitch + c) = o;
    }
}

template <int DW_TY>
static int launch_dw3x3_ty(const ActView& in, const ActView& out, const __half* w, cudaStream_t s, bool pdl)
{
    const long long total = static_cast<long long>(in.W) * (in.C / 8) * ((in.H + DW_TY - 1) / DW_TY);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks_for(total, 128), 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const char* e = getenv("DCVC_B200_PDL");
    cfg.numAttrs = ((e && e[0] == '0') || !pdl) ? 0 : 1;
    cudaLaunchKernelEx(&cfg, dw3x3_kernel<DW_TY>, static_cast<const __half*>(in.ptr), in.pitch,
                       static_cast<__half*>(const_cast<void*>(out.ptr)), out.pitch, w, in.C, in.W, in.H);
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_dw3x3(const ActView& in, const ActView& out, const __half* w, cudaStream_t s, bool pdl)
{
    // rows per thread: 4 (2 and 8 were measured in round 1 and are gone)
    return launch_dw3x3_ty<DW_TY_DEFAULT>(in, out, w, s, pdl);
}

// ------------------------------------------------------------------------------- unshuffle8
__global__ void __launch_bounds__(256)
unshuffle8_pad_kernel(const __half* __restrict__ x, int Cs, int H, int W, long long sc,
                      long long sh, long long sw, __half* __restrict__ out, int out_pitch, int W8,
                      int H8)
{
    // one thread: 8 consecutive output channels = 8 horizontally adjacent source pixels
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int per_pix = Cs * 8;
    const long long total = static_cast<long long>(W8) * H8 * per_pix;
    if (tid >= total) return;
    const int r = static_cast<int>(tid % per_pix);
    const long long pix = tid / per_pix;
    const int cs = r >> 3;
    const int dy = r & 7;
    const int w8 = static_cast<int>(pix % W8);
    const int h8 = static_cast<int>(pix / W8);
    const int sy = min(h8 * 8 + dy, H - 1);
    __half v[8];
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
        const int sx = min(w8 * 8 + dx, W - 1);
        v[dx] = x[cs * sc + sy * sh + sx * sw];
    }
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cs * 64 + dy * 8) =
        *reinterpret_cast<const uint4*>(v);
}

int launch_unshuffle8_pad(const __half* x, int Cs, int H, int W, long long sc, long long sh,
                          long long sw, const ActView& out, cudaStream_t s)
{
    const long long total = static_cast<long long>(out.W) * out.H * Cs * 8;
    unshuffle8_pad_kernel<<<blocks_for(total, 256), 256, 0, s>>>(
        x, Cs, H, W, sc, sh, sw, static_cast<__half*>(const_cast<void*>(out.ptr)), out.pitch,
        out.W, out.H);
    DCVC_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- shuffle8
template <int CS>
__global__ void __launch_bounds__(256)
shuffle8_clamp_kernel(const __half* __restrict__ in, int in_pitch, int W8, int H8,
                      __half* __restrict__ out, int clamp)
{
    // one thread: source pixel (h8, w8), row dy of its 8x8 block -> 8 px x CS channels, contiguous
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(W8) * H8 * 8;
    if (tid >= total) return;
    const int dy = static_cast<int>(tid & 7);
    const long long pix = tid >> 3;
    const int w8 = static_cast<int>(pix % W8);
    const int h8 = static_cast<int>(pix / W8);
    __half o[8 * CS];
    const __half lo = __float2half(-0.5f);
    const __half hi = __float2half(0.5f);
#pragma unroll
    for (int cs = 0; cs < CS; ++cs) {
        const uint4 v = *reinterpret_cast<const uint4*>(in + pix * in_pitch + cs * 64 + dy * 8);
        const __half* vh = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int dx = 0; dx < 8; ++dx) {
            __half t = vh[dx];
            if (clamp) t = __hmin(__hmax(t, lo), hi);
            o[dx * CS + cs] = t;
        }
    }
    const long long Wo = static_cast<long long>(W8) * 8;
    __half* dst = out + ((static_cast<long long>(h8) * 8 + dy) * Wo + static_cast<long long>(w8) * 8) * CS;
    if ((CS * 16) % 16 == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < CS; ++i) {
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(o)[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8 * CS; ++i) dst[i] = o[i];
    }
}

int launch_shuffle8_clamp(const ActView& in, __half* out, int Cs, int clamp, cudaStream_t s)
{
    const long long total = static_cast<long long>(in.W) * in.H * 8;
    const __half* ip = static_cast<const __half*>(in.ptr);
    if (Cs == 3) {
        shuffle8_clamp_kernel<3><<<blocks_for(total, 256), 256, 0, s>>>(ip, in.pitch, in.W, in.H, out, clamp);
    } else if (Cs == 1) {
        shuffle8_clamp_kernel<1><<<blocks_for(total, 256), 256, 0, s>>>(ip, in.pitch, in.W, in.H, out, clamp);
    } else {
        fprintf(stderr, "shuffle8_clamp: unsupported Cs=%d\n", Cs);
        return 1;
    }
    DCVC_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- pad/crop, scale
__global__ void __launch_bounds__(256)
pad_crop_kernel(const __half* __restrict__ in, int in_pitch, int Wi, int Hi,
                __half* __restrict__ out, int out_pitch, int Wo, int Ho, int C)
{
    const int cg_n = C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(Wo) * Ho * cg_n;
    if (tid >= total) return;
    const int cg = static_cast<int>(tid % cg_n);
    const long long pix = tid / cg_n;
    const int x = static_cast<int>(pix % Wo);
    const int y = static_cast<int>(pix / Wo);
    const int sx = min(x, Wi - 1);
    const int sy = min(y, Hi - 1);
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cg * 8) = *reinterpret_cast<const uint4*>(
        in + (static_cast<long long>(sy) * Wi + sx) * in_pitch + cg * 8);
}

int launch_pad_crop(const ActView& in, const ActView& out, cudaStream_t s)
{
    const long long total = static_cast<long long>(out.W) * out.H * (out.C / 8);
    pad_crop_kernel<<<blocks_for(total, 256), 256, 0, s>>>(
        static_cast<const __half*>(in.ptr), in.pitch, in.W, in.H,
        static_cast<__half*>(const_cast<void*>(out.ptr)), out.pitch, out.W, out.H, out.C);
    DCVC_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
scale_channels_kernel(const __half* __restrict__ in, int in_pitch, const __half* __restrict__ q,
                      __half* __restrict__ out, int out_pitch, long long npix, int C)
{
    const int cg_n = C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= npix * cg_n) return;
    const int cg = static_cast<int>(tid % cg_n);
    const long long pix = tid / cg_n;
    const uint4 v = *reinterpret_cast<const uint4*>(in + pix * in_pitch + cg * 8);
    const uint4 k = __ldg(reinterpret_cast<const uint4*>(q + cg * 8));
    const __half2* vh = reinterpret_cast<const __half2*>(&v);
    const __half2* kh = reinterpret_cast<const __half2*>(&k);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __hmul2(vh[i], kh[i]);
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cg * 8) = o;
}

int launch_scale_channels(const ActView& in, const __half* q, const ActView& out, cudaStream_t s)
{
    const long long npix = static_cast<long long>(in.W) * in.H;
    scale_channels_kernel<<<blocks_for(npix * (in.C / 8), 256), 256, 0, s>>>(
        static_cast<const __half*>(in.ptr), in.pitch, q,
        static_cast<__half*>(const_cast<void*>(out.ptr)), out.pitch, npix, in.C);
    DCVC_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- z path
__device__ __forceinline__ float round_half_away(float v)
{
    // CUDA round(at::Half) == roundf: ties away from zero (stream.cu:587-588, 873-874)
    return roundf(v);
}

__global__ void __launch_bounds__(256)
round_z_kernel(const __half* __restrict__ z, __half* __restrict__ z_hat, int8_t* __restrict__ z_i8,
               long long n)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = round_half_away(__half2float(z[i]));
    v = fminf(fmaxf(v, -64.f), 63.f);
    z_hat[i] = __float2half_rn(v);
    z_i8[i] = static_cast<int8_t>(v);
}

int launch_round_z(const __half* z, __half* z_hat, int8_t* z_i8, long long n, cudaStream_t s)
{
    round_z_kernel<<<blocks_for(n, 256), 256, 0, s>>>(z, z_hat, z_i8, n);
    DCVC_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
int8_to_half_kernel(const int8_t* __restrict__ x, __half* __restrict__ out, long long n)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __int2half_rn(static_cast<int>(x[i]));
}

int launch_int8_to_half(const int8_t* x, __half* out, long long n, cudaStream_t s)
{
    int8_to_half_kernel<<<blocks_for(n, 256), 256, 0, s>>>(x, out, n);
    DCVC_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- entropy path
// One warp per latent pixel.  The active channel group of step k at 2x2 phase p = (h&1)*2+(w&1)
// is p ^ {0,3,2,1}[k]  (get_mask_4x, common_model.py:174-195; dmci_proxy.cpp:678-699).
__device__ __forceinline__ int active_group(int step, int h, int w)
{
    const int p = ((h & 1) << 1) | (w & 1);
    const int x = (step == 0) ? 0 : (step == 1 ? 3 : (step == 2 ? 2 : 1));
    return p ^ x;
}

// ng == 2: the two-step checkerboard of the low-delay model (get_mask_2x, common_model.py:157-172): the first
// channel half is coded on the even checkerboard phase in step 0 and on the odd one in step 1, the second half the
// other way round -> the active half of pixel (h, w) in step k is ((h + w) & 1) ^ k.
__device__ __forceinline__ int active_group_n(int ng, int step, int h, int w)
{
    return ng == 2 ? (((h + w) & 1) ^ step) : active_group(step, h, w);
}

static EntropyDev to_dev(const EntropyStepArgs& a)
{
    EntropyDev d;
    d.qdiv = a.q_div; d.q_pitch = a.q_pitch; d.m_pitch = a.m_pitch ? a.m_pitch : a.p_pitch;
    d.yq = a.yq_dense; d.full = a.full; d.ng = a.ng == 2 ? 2 : 4;
    d.H = a.H; d.W = a.W; d.G = a.G; d.step = a.step;
    d.y = a.y; d.y_pitch = a.y_pitch; d.q_enc = a.q_enc;
    d.scales = a.scales; d.means = a.means; d.p_pitch = a.p_pitch;
    d.acc = a.y_hat_acc; d.acc_pitch = a.acc_pitch;
    d.thres = __float2half_rn(a.skip_thres);
    d.lut = a.scale_lut;
    d.sym_raw = a.sym_raw; d.idx_raw = a.idx_raw; d.counts = a.counts;
    return d;
}

__global__ void __launch_bounds__(256)
entropy_enc_step_kernel(const EntropyDev d)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long npix = static_cast<long long>(d.H) * d.W;
    if (pix >= npix) return;
    const int w = static_cast<int>(pix % d.W);
    const int h = static_cast<int>(pix / d.W);
    const int g = active_group_n(d.ng, d.step, h, w);
    int count = 0;
    for (int c = lane * 2; c < d.G; c += 64) {
        const int ch = g * d.G + c;
        __half2 yv = *reinterpret_cast<const __half2*>(d.y + pix * d.y_pitch + ch);
        if (d.q_enc) yv = __hmul2_rn(yv, *reinterpret_cast<const __half2*>(d.q_enc + ch));
        if (d.qdiv) {
            // y * hrcp(max(q, 0.5)): reciprocal rounded to half, then a half multiply (stream.cu:437-440)
            const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(d.qdiv + pix * d.q_pitch + ch));
            const __half2 rc = __floats2half2_rn(1.0f / fmaxf(qf.x, 0.5f), 1.0f / fmaxf(qf.y, 0.5f));
            yv = __hmul2_rn(yv, rc);
        }
        const __half2 mv = *reinterpret_cast<const __half2*>(d.means + pix * d.m_pitch + ch);
        const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + ch);
        const __half2 res = __hsub2_rn(yv, mv);  // _rn: never contracted into an fma with the multiply above
        float q0 = round_half_away(__low2float(res));
        float q1 = round_half_away(__high2float(res));
        const bool c0 = __hgt(__low2half(sv), d.thres);
        const bool c1 = __hgt(__high2half(sv), d.thres);
        if (!c0) q0 = 0.f;
        if (!c1) q1 = 0.f;
        q0 = fminf(fmaxf(q0, -128.f), 127.f);
        q1 = fminf(fmaxf(q1, -128.f), 127.f);
        const __half2 yq = __floats2half2_rn(q0, q1);
        const __half2 yh = __hadd2_rn(yq, mv);
        *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + ch) = yh;
        if (d.yq) {
            char2 qq;
            qq.x = static_cast<signed char>(q0);
            qq.y = static_cast<signed char>(q1);
            *reinterpret_cast<char2*>(d.yq + pix * (d.ng * d.G) + ch) = qq;
        }
        if (d.sym_raw) {
            const int i0 = d.lut[__half_as_ushort(__low2half(sv))];
            const int i1 = d.lut[__half_as_ushort(__high2half(sv))];
            short2 sym;
            sym.x = static_cast<short>((static_cast<int>(q0) << 8) + i0);
            sym.y = static_cast<short>((static_cast<int>(q1) << 8) + i1);
            *reinterpret_cast<short2*>(d.sym_raw + pix * d.G + c) = sym;
        }
        count += (c0 ? 1 : 0) + (c1 ? 1 : 0);
    }
    if (d.step == 0) {
        // y_hat_so_far.copy_(y_hat): the three inactive groups start at zero (dmci_proxy.cpp:344)
        const __half2 z = __floats2half2_rn(0.f, 0.f);
        for (int c = lane * 2; c < d.ng * d.G; c += 64) {
            if (c / d.G != g) *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + c) = z;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0 && d.counts) d.counts[pix] = count;
}

__global__ void __launch_bounds__(256)
entropy_dec_index_kernel(const EntropyDev d)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long npix = static_cast<long long>(d.H) * d.W;
    if (pix >= npix) return;
    const int w = static_cast<int>(pix % d.W);
    const int h = static_cast<int>(pix / d.W);
    const int g = d.full ? 0 : active_group_n(d.ng, d.step, h, w);
    int count = 0;
    for (int c = lane * 2; c < d.G; c += 64) {
        const int ch = g * d.G + c;
        const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + ch);
        const bool c0 = __hgt(__low2half(sv), d.thres);
        const bool c1 = __hgt(__high2half(sv), d.thres);
        uchar2 idx;
        idx.x = d.lut[__half_as_ushort(__low2half(sv))];
        idx.y = d.lut[__half_as_ushort(__high2half(sv))];
        *reinterpret_cast<uchar2*>(d.idx_raw + pix * d.G + c) = idx;
        count += (c0 ? 1 : 0) + (c1 ? 1 : 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0) d.counts[pix] = count;
}

// rank of this lane's two adjacent entries (2*lane, 2*lane+1) among the kept entries of the warp
__device__ __forceinline__ void pair_ranks(bool c0, bool c1, int lane, int& r0, int& r1, int& tot)
{
    const unsigned b0 = __ballot_sync(0xffffffffu, c0);
    const unsigned b1 = __ballot_sync(0xffffffffu, c1);
    const unsigned lt = (1u << lane) - 1u;
    r0 = __popc(b0 & lt) + __popc(b1 & lt);
    r1 = r0 + (c0 ? 1 : 0);
    tot = __popc(b0) + __popc(b1);
}

template <typename T>
__global__ void __launch_bounds__(256)
compact_kernel(const EntropyDev d, const T* __restrict__ raw, const int32_t* __restrict__ offsets,
               T* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long npix = static_cast<long long>(d.H) * d.W;
    if (pix >= npix) return;
    const int w = static_cast<int>(pix % d.W);
    const int h = static_cast<int>(pix / d.W);
    const int g = d.full ? 0 : active_group_n(d.ng, d.step, h, w);
    int base = offsets[pix];
    for (int c0i = 0; c0i < d.G; c0i += 64) {
        const int c = c0i + lane * 2;
        bool k0 = false, k1 = false;
        T v0 = 0, v1 = 0;
        if (c < d.G) {
            const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + g * d.G + c);
            k0 = __hgt(__low2half(sv), d.thres);
            k1 = __hgt(__high2half(sv), d.thres);
            v0 = raw[pix * d.G + c];
            v1 = raw[pix * d.G + c + 1];
        }
        int r0, r1, tot;
        pair_ranks(k0, k1, lane, r0, r1, tot);
        if (k0) out[base + r0] = v0;
        if (k1) out[base + r1] = v1;
        base += tot;
    }
}

__global__ void __launch_bounds__(256)
entropy_dec_restore_kernel(const EntropyDev d, const int32_t* __restrict__ offsets,
                           const int8_t* __restrict__ decoded)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long npix = static_cast<long long>(d.H) * d.W;
    if (pix >= npix) return;
    const int w = static_cast<int>(pix % d.W);
    const int h = static_cast<int>(pix / d.W);
    const int g = active_group_n(d.ng, d.step, h, w);
    int base = offsets[pix];
    for (int c0i = 0; c0i < d.G; c0i += 64) {
        const int c = c0i + lane * 2;
        bool k0 = false, k1 = false;
        __half2 mv = __floats2half2_rn(0.f, 0.f);
        const int ch = g * d.G + c;
        if (c < d.G) {
            const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + ch);
            k0 = __hgt(__low2half(sv), d.thres);
            k1 = __hgt(__high2half(sv), d.thres);
            mv = *reinterpret_cast<const __half2*>(d.means + pix * d.p_pitch + ch);
        }
        int r0, r1, tot;
        pair_ranks(k0, k1, lane, r0, r1, tot);
        if (c < d.G) {
            const float q0 = k0 ? static_cast<float>(decoded[base + r0]) : 0.f;
            const float q1 = k1 ? static_cast<float>(decoded[base + r1]) : 0.f;
            const __half2 yh = __hadd2_rn(__floats2half2_rn(q0, q1), mv);
            *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + ch) = yh;
        }
        base += tot;
    }
    if (d.step == 0) {
        const __half2 z = __floats2half2_rn(0.f, 0.f);
        for (int c = lane * 2; c < d.ng * d.G; c += 64) {
            if (c / d.G != g) *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + c) = z;
        }
    }
}

// ---- chunk-codec kernels: symbols over the whole latent, dense y_q
__global__ void __launch_bounds__(256)
entropy_build_symbols_full_kernel(const EntropyDev d)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (pix >= static_cast<long long>(d.H) * d.W) return;
    int count = 0;
    for (int c = lane * 2; c < d.G; c += 64) {
        const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + c);
        const char2 qq = *reinterpret_cast<const char2*>(d.yq + pix * d.G + c);
        const int i0 = d.lut[__half_as_ushort(__low2half(sv))];
        const int i1 = d.lut[__half_as_ushort(__high2half(sv))];
        short2 sym;
        sym.x = static_cast<short>((static_cast<int>(qq.x) << 8) + i0);
        sym.y = static_cast<short>((static_cast<int>(qq.y) << 8) + i1);
        *reinterpret_cast<short2*>(d.sym_raw + pix * d.G + c) = sym;
        count += (__hgt(__low2half(sv), d.thres) ? 1 : 0) + (__hgt(__high2half(sv), d.thres) ? 1 : 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0) d.counts[pix] = count;
}

__global__ void __launch_bounds__(256)
entropy_recover_dense_kernel(const EntropyDev d, const int32_t* __restrict__ offsets,
                             const int8_t* __restrict__ decoded)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (pix >= static_cast<long long>(d.H) * d.W) return;
    int base = offsets[pix];
    for (int c0i = 0; c0i < d.G; c0i += 64) {
        const int c = c0i + lane * 2;
        bool k0 = false, k1 = false;
        if (c < d.G) {
            const __half2 sv = *reinterpret_cast<const __half2*>(d.scales + pix * d.p_pitch + c);
            k0 = __hgt(__low2half(sv), d.thres);
            k1 = __hgt(__high2half(sv), d.thres);
        }
        int r0, r1, tot;
        pair_ranks(k0, k1, lane, r0, r1, tot);
        if (c < d.G) {
            char2 qq;
            qq.x = k0 ? decoded[base + r0] : 0;
            qq.y = k1 ? decoded[base + r1] : 0;
            *reinterpret_cast<char2*>(d.yq + pix * d.G + c) = qq;
        }
        base += tot;
    }
}

__global__ void __launch_bounds__(256)
entropy_restore_dense_kernel(const EntropyDev d)
{
    const int lane = threadIdx.x & 31;
    const long long pix = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (pix >= static_cast<long long>(d.H) * d.W) return;
    const int w = static_cast<int>(pix % d.W);
    const int h = static_cast<int>(pix / d.W);
    const int g = active_group_n(d.ng, d.step, h, w);
    for (int c = lane * 2; c < d.G; c += 64) {
        const int ch = g * d.G + c;
        const char2 qq = *reinterpret_cast<const char2*>(d.yq + pix * (d.ng * d.G) + ch);
        const __half2 mv = *reinterpret_cast<const __half2*>(d.means + pix * d.m_pitch + ch);
        const __half2 yh = __hadd2_rn(__floats2half2_rn(static_cast<float>(qq.x), static_cast<float>(qq.y)), mv);
        *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + ch) = yh;
    }
    if (d.step == 0) {
        const __half2 z = __floats2half2_rn(0.f, 0.f);
        for (int c = lane * 2; c < d.ng * d.G; c += 64) {
            if (c / d.G != g) *reinterpret_cast<__half2*>(d.acc + pix * d.acc_pitch + c) = z;
        }
    }
}

__global__ void __launch_bounds__(256)
mul_clamp_min_kernel(const __half* __restrict__ in, int in_pitch, const __half* __restrict__ q, int q_pitch,
                     __half* __restrict__ out, int out_pitch, long long npix, int C)
{
    const int cg_n = C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= npix * cg_n) return;
    const int cg = static_cast<int>(tid % cg_n);
    const long long pix = tid / cg_n;
    const uint4 v = *reinterpret_cast<const uint4*>(in + pix * in_pitch + cg * 8);
    const uint4 k = *reinterpret_cast<const uint4*>(q + pix * q_pitch + cg * 8);
    const __half2* vh = reinterpret_cast<const __half2*>(&v);
    const __half2* kh = reinterpret_cast<const __half2*>(&k);
    const __half2 half_ = __floats2half2_rn(0.5f, 0.5f);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __hmul2_rn(vh[i], __hmax2(kh[i], half_));
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cg * 8) = o;
}

static inline int warp_grid(long long npix) { return static_cast<int>((npix * 32 + 255) / 256); }

int launch_entropy_enc_step(const EntropyStepArgs& a, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_enc_step_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a));
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_entropy_dec_index(const EntropyStepArgs& a, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_dec_index_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a));
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_compact_i16(const EntropyStepArgs& a, const int32_t* offsets, int16_t* out, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    compact_kernel<int16_t><<<warp_grid(npix), 256, 0, s>>>(to_dev(a), a.sym_raw, offsets, out);
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_compact_u8(const EntropyStepArgs& a, const int32_t* offsets, uint8_t* out, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    compact_kernel<uint8_t><<<warp_grid(npix), 256, 0, s>>>(to_dev(a), a.idx_raw, offsets, out);
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_entropy_dec_restore(const EntropyStepArgs& a, const int32_t* offsets,
                               const int8_t* decoded, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_dec_restore_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a), offsets, decoded);
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_entropy_build_symbols_full(const EntropyStepArgs& a, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_build_symbols_full_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a));
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_entropy_recover_dense(const EntropyStepArgs& a, const int32_t* offsets, const int8_t* decoded, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_recover_dense_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a), offsets, decoded);
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_entropy_restore_dense(const EntropyStepArgs& a, cudaStream_t s)
{
    const long long npix = static_cast<long long>(a.H) * a.W;
    entropy_restore_dense_kernel<<<warp_grid(npix), 256, 0, s>>>(to_dev(a));
    DCVC_LAUNCH_CHECK();
    return 0;
}

int launch_mul_clamp_min(const ActView& in, const ActView& q, const ActView& out, cudaStream_t s)
{
    const long long npix = static_cast<long long>(in.W) * in.H;
    mul_clamp_min_kernel<<<blocks_for(npix * (in.C / 8), 256), 256, 0, s>>>(
        static_cast<const __half*>(in.ptr), in.pitch, static_cast<const __half*>(q.ptr), q.pitch,
        static_cast<__half*>(const_cast<void*>(out.ptr)), out.pitch, npix, in.C);
    DCVC_LAUNCH_CHECK();
    return 0;
}

// single-CTA exclusive scan (n <= 1024 * 64): every thread owns PER consecutive counts, read once with 128-bit loads (a warp
// reads PER * 128 contiguous bytes), kept in registers across the block scan, written back as 128-bit stores.  It sits on the
// critical path of every prior step (count -> scan -> compact -> host), so its latency, not its bandwidth, is what matters.
template <int PER>
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t* __restrict__ counts, int32_t* __restrict__ offsets,
                   int32_t* __restrict__ total, int n, bool vec)   // vec: both buffers 16-byte aligned
{
    __shared__ int warp_sums[32];
    const int tid = threadIdx.x;
    const int begin = tid * PER;
    int c[PER];
#pragma unroll
    for (int j = 0; j < PER; j += 4) {
        const int i = begin + j;
        if (vec && i + 3 < n) {
            const int4 v = *reinterpret_cast<const int4*>(counts + i);
            c[j] = v.x; c[j + 1] = v.y; c[j + 2] = v.z; c[j + 3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) c[j + e] = (i + e < n) ? counts[i + e] : 0;
        }
    }
    int local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) local += c[j];
    // inclusive warp scan
    int v = local;
    const int lane = tid & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        int ws = warp_sums[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, ws, o);
            if (tid >= o) ws += t;
        }
        warp_sums[tid] = ws;
    }
    __syncthreads();
    int run = v - local + ((tid >> 5) > 0 ? warp_sums[(tid >> 5) - 1] : 0);
#pragma unroll
    for (int j = 0; j < PER; j += 4) {
        const int i = begin + j;
        int4 o;
        o.x = run; run += c[j];
        o.y = run; run += c[j + 1];
        o.z = run; run += c[j + 2];
        o.w = run; run += c[j + 3];
        if (vec && i + 3 < n) {
            *reinterpret_cast<int4*>(offsets + i) = o;
        } else {
            if (i < n) offsets[i] = o.x;
            if (i + 1 < n) offsets[i + 1] = o.y;
            if (i + 2 < n) offsets[i + 2] = o.z;
            if (i + 3 < n) offsets[i + 3] = o.w;
        }
    }
    if (tid == 1023) {
        offsets[n] = warp_sums[31];
        *total = warp_sums[31];
    }
}

int launch_scan_counts(const int32_t* counts, int32_t* offsets, int32_t* total, int n, cudaStream_t s)
{
    if (n > 1024 * 64) {
        fprintf(stderr, "scan_counts: n=%d too large\n", n);
        return 1;
    }
    const bool vec = ((reinterpret_cast<uintptr_t>(counts) | reinterpret_cast<uintptr_t>(offsets)) & 15) == 0;
    if (n <= 1024 * 8) scan_counts_kernel<8><<<1, 1024, 0, s>>>(counts, offsets, total, n, vec);
    else if (n <= 1024 * 16) scan_counts_kernel<16><<<1, 1024, 0, s>>>(counts, offsets, total, n, vec);
    else if (n <= 1024 * 32) scan_counts_kernel<32><<<1, 1024, 0, s>>>(counts, offsets, total, n, vec);
    else scan_counts_kernel<64><<<1, 1024, 0, s>>>(counts, offsets, total, n, vec);
    DCVC_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------- host LUT
void build_scale_lut(uint8_t* lut)
{
    // Reference (stream.cu:77-87) in half arithmetic:
    //   s = min(max(s, 0.11), 16); s = hlog(s) - LOG_SCALE_MIN; s = s * LOG_SCALE_STEP_RECIP;
    //   idx = __half2uint_rd(s)
    const float log_min_f = -2.2073f;
    const float log_max_f = 2.7726f;
    const float recip_f = 1.f / ((log_max_f - log_min_f) / 127.f);
    const __half h_min = __float2half_rn(0.11f);
    const __half h_max = __float2half_rn(16.f);
    const __half h_logmin = __float2half_rn(log_min_f);
    const __half h_recip = __float2half_rn(recip_f);
    for (int b = 0; b < 65536; ++b) {
        const __half_raw raw = { static_cast<unsigned short>(b) };
        float s = __half2float(__half(raw));
        const float fmin_ = __half2float(h_min);
        const float fmax_ = __half2float(h_max);
        if (!(s >= fmin_)) s = fmin_;  // also maps NaN / negatives to the minimum
        if (s > fmax_) s = fmax_;
        const __half l = __float2half_rn(logf(s));
        const __half d = __float2half_rn(__half2float(l) - __half2float(h_logmin));
        const __half m = __float2half_rn(__half2float(d) * __half2float(h_recip));
        float f = floorf(__half2float(m));
        if (f < 0.f) f = 0.f;
        if (f > 127.f) f = 127.f;
        lut[b] = static_cast<uint8_t>(f);
    }
}

}  // namespace dcvc
