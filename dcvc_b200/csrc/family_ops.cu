// family_ops.cu — ops of the older DCVC-family codecs that BASELINE.json's north_star names but DCVC-UF itself does
// not use: bilinear backward warp (motion compensation of DCVC-DC/FM) and the elementwise
// square in front of the GDN GEMM.  NHWC fp16 like every other kernel of the library.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcvc_b200.h"

namespace dcvc {

// one thread = one pixel x 8 channels (16 bytes)
__global__ void __launch_bounds__(256)
square_kernel(const __half* __restrict__ in, int in_pitch, __half* __restrict__ out, int out_pitch, long long npix, int C)
{
    const int cg_n = C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= npix * cg_n) return;
    const int cg = static_cast<int>(tid % cg_n);
    const long long pix = tid / cg_n;
    uint4 v = *reinterpret_cast<const uint4*>(in + pix * in_pitch + cg * 8);
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hmul2(h[i], h[i]);
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cg * 8) = v;
}

// Bilinear backward warp, border clamp.  Arithmetic of the reference's __half kernel (block_mc_kernel.cu:33-71):
// positions and the four weights in fp32, weights rounded to half, then r = fma(im_a, w_a, 0), fma(im_b, w_b, r),
// fma(im_c, w_c, r), fma(im_d, w_d, r) in half, a = (y0, x0), b = (y1, x0), c = (y0, x1), d = (y1, x1).
// One thread = one output pixel x 8 channels: four 16-byte gathers (the channel groups of a pixel are adjacent
// threads, so each neighbour pixel is read as one contiguous run), HFMA2 on packed pairs.
__global__ void __launch_bounds__(256)
warp_bilinear_kernel(const __half* __restrict__ im, int im_pitch, const __half* __restrict__ flow, long long fc,
                     long long fh, long long fw, __half* __restrict__ out, int out_pitch, int W, int H, int C)
{
    const int cg_n = C >> 3;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (tid >= static_cast<long long>(W) * H * cg_n) return;
    const int cg = static_cast<int>(tid % cg_n);
    const long long pix = tid / cg_n;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>(pix / W);
    const float x_off = __half2float(flow[h * fh + w * fw]);
    const float y_off = __half2float(flow[fc + h * fh + w * fw]);
    float x_pos = x_off + static_cast<float>(w);
    float y_pos = y_off + static_cast<float>(h);
    x_pos = fminf(fmaxf(x_pos, 0.f), static_cast<float>(W - 1));
    y_pos = fminf(fmaxf(y_pos, 0.f), static_cast<float>(H - 1));
    const int x0 = __float2int_rd(x_pos);
    const int x1 = min(x0 + 1, W - 1);
    const int y0 = __float2int_rd(y_pos);
    const int y1 = min(y0 + 1, H - 1);
    const float w_r = x_pos - static_cast<float>(x0);
    const float w_l = 1.f - w_r;
    const float w_b = y_pos - static_cast<float>(y0);
    const float w_t = 1.f - w_b;
    const __half2 wa = __half2half2(__float2half_rn(w_l * w_t));
    const __half2 wb = __half2half2(__float2half_rn(w_l * w_b));
    const __half2 wc = __half2half2(__float2half_rn(w_r * w_t));
    const __half2 wd = __half2half2(__float2half_rn(w_r * w_b));
    const uint4 va = *reinterpret_cast<const uint4*>(im + (static_cast<long long>(y0) * W + x0) * im_pitch + cg * 8);
    const uint4 vb = *reinterpret_cast<const uint4*>(im + (static_cast<long long>(y1) * W + x0) * im_pitch + cg * 8);
    const uint4 vc = *reinterpret_cast<const uint4*>(im + (static_cast<long long>(y0) * W + x1) * im_pitch + cg * 8);
    const uint4 vd = *reinterpret_cast<const uint4*>(im + (static_cast<long long>(y1) * W + x1) * im_pitch + cg * 8);
    const __half2* a = reinterpret_cast<const __half2*>(&va);
    const __half2* b = reinterpret_cast<const __half2*>(&vb);
    const __half2* c = reinterpret_cast<const __half2*>(&vc);
    const __half2* d = reinterpret_cast<const __half2*>(&vd);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
    const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 r = __hfma2(a[i], wa, zero);
        r = __hfma2(b[i], wb, r);
        r = __hfma2(c[i], wc, r);
        r = __hfma2(d[i], wd, r);
        oh[i] = r;
    }
    *reinterpret_cast<uint4*>(out + pix * out_pitch + cg * 8) = o;
}

}  // namespace dcvc

extern "C" {

static int fam_check(const dcvc_view* a, const dcvc_view* b)
{
    if (!a || !b || !a->ptr || !b->ptr) return 1;
    if (a->C != b->C || a->W != b->W || a->H != b->H || (a->C % 8) || (a->pitch % 8) || (b->pitch % 8)) return 1;
    if ((reinterpret_cast<uintptr_t>(a->ptr) & 15) || (reinterpret_cast<uintptr_t>(b->ptr) & 15)) return 1;
    return 0;
}

int dcvc_op_square(const dcvc_view* in, const dcvc_view* out, void* stream)
{
    if (fam_check(in, out)) return 1;
    const long long npix = static_cast<long long>(in->W) * in->H;
    const long long n = npix * (in->C / 8);
    dcvc::square_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(in->ptr), in->pitch, static_cast<__half*>(const_cast<void*>(out->ptr)), out->pitch, npix, in->C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int dcvc_op_warp_bilinear(const dcvc_view* im, const void* flow, int64_t fc, int64_t fh, int64_t fw, const dcvc_view* out,
                          void* stream)
{
    if (fam_check(im, out) || !flow) return 1;
    const long long n = static_cast<long long>(im->W) * im->H * (im->C / 8);
    dcvc::warp_bilinear_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(im->ptr), im->pitch, static_cast<const __half*>(flow), fc, fh, fw,
        static_cast<__half*>(const_cast<void*>(out->ptr)), out->pitch, im->W, im->H, im->C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // extern "C"
