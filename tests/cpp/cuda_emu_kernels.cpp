// cuda_emu_kernels.cpp — TEST INFRASTRUCTURE ONLY (tests/test_host_dry_run.py).
//
// Host-side restatements of WHAT each kernel of libdcvc_b200 computes, driven from the kernel's own parameter block
// as the library passed it to cudaLaunchKernel (intercepted by cuda_dry_shim.cpp).  This is not a CPU path of the
// product: it only exists inside the test's LD_PRELOAD shim, reads the *device-side* argument structs (PwGemmParams
// with the shim's transparent tensor maps, EntropyDev, raw pointers + pitches) and lets the CPU test tier check the
// host half of the codecs — operand wiring, weight packing, tap tables, tensor-map geometry, buffer aliasing, segment
// order, the rANS hand-off — against the oracles without a GPU.  fp32 accumulation in natural order, results rounded
// to fp16 where the kernels round; agreement with the device is therefore close, not bit-exact.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../dcvc_b200/csrc/elementwise.cuh"
#include "../../dcvc_b200/csrc/pw_gemm.cuh"

using namespace dcvc;

// the shim's transparent tensor map (cuda_dry_shim.cpp fake_encode_tiled writes it into the 128-byte CUtensorMap)
struct FakeMap {
    uint64_t magic;
    const uint8_t* ptr;
    uint32_t rank, swizzle;
    uint64_t dims[5];
    uint64_t strides[4];   // bytes, dims 1..4
    uint32_t box[5];
};
static_assert(sizeof(FakeMap) <= sizeof(CUtensorMap), "fake map must fit");
static constexpr uint64_t kMapMagic = 0x44525950414d5831ull;

static inline float h2f(__half h) { return __half2float(h); }
static inline __half f2h(float f) { return __float2half_rn(f); }
static inline float round_away(float v) { return roundf(v); }

static const FakeMap* map_of(const CUtensorMap& m)
{
    const FakeMap* f = reinterpret_cast<const FakeMap*>(&m);
    return f->magic == kMapMagic ? f : nullptr;
}

// address of element (c, i1, i2, i3, i4) or nullptr when out of bounds (TMA: zero fill on load, dropped on store)
static const uint8_t* map_at(const FakeMap* m, long long c, long long i1, long long i2, long long i3, long long i4)
{
    const long long idx[5] = { c, i1, i2, i3, i4 };
    for (uint32_t d = 0; d < m->rank; ++d)
        if (idx[d] < 0 || idx[d] >= static_cast<long long>(m->dims[d])) return nullptr;
    const uint8_t* p = m->ptr + c * 2;
    for (uint32_t d = 1; d < m->rank; ++d) p += idx[d] * static_cast<long long>(m->strides[d - 1]);
    return p;
}

// Access reports for the shim's lane race check (cuda_dry_shim.cpp): kernels that may run in parallel graph branches
// report the byte ranges they read and write (conservative: first to last byte of a view, pitch gaps included).
// A kernel without reports in a multi-lane graph fails that graph's launch.
extern "C" { void (*emu_access_hook)(const void* lo, const void* hi, int is_write) = nullptr; int emu_reported = 0; }
static void report(const void* lo, size_t bytes, int is_write)
{
    emu_reported = 1;
    if (emu_access_hook && lo && bytes) emu_access_hook(lo, static_cast<const uint8_t*>(lo) + bytes, is_write);
}
static size_t map_extent(const FakeMap* m)
{
    size_t e = m->dims[0] * 2;
    for (uint32_t d = 1; d < m->rank; ++d) e += (m->dims[d] - 1) * m->strides[d - 1];
    return e;
}

static inline float wsilu(float x)
{
    const float h = 0.5f * x;
    return fmaf(h, tanhf(2.f * x), h);
}

// ------------------------------------------------------------------------------------------------ pw_gemm
static int emu_pw_gemm(void** args)
{
    const PwGemmParams& p = *static_cast<const PwGemmParams*>(args[0]);
    const FakeMap *A = map_of(p.tm_a), *B = map_of(p.tm_b), *Cm = map_of(p.tm_c);
    if (!A || !B || !Cm) { fprintf(stderr, "emu pw_gemm: tensor map without the shim's magic\n"); return 1; }
    const int N = static_cast<int>(B->dims[1]);
    const int Ktot = static_cast<int>(B->dims[0]);
    const int Cin = p.kblk_per_tap * 64;
    const int taps = p.num_kblocks / p.kblk_per_tap;
    if (taps * Cin != Ktot || B->strides[0] != static_cast<uint64_t>(Ktot) * 2) { fprintf(stderr, "emu pw_gemm: weight map geometry\n"); return 1; }
    if (static_cast<int>(A->dims[0]) != Cin) { fprintf(stderr, "emu pw_gemm: A channels %d != %d\n", (int)A->dims[0], Cin); return 1; }
    const bool up = p.phase_c > 0;
    const bool two_d = A->rank == 2;
    long long gw, gh;
    if (two_d) { gw = static_cast<long long>(A->dims[1]); gh = 1; }
    else if (up) { gw = static_cast<long long>(A->dims[2]); gh = static_cast<long long>(A->dims[4]); }
    else { gw = static_cast<long long>(Cm->dims[2]); gh = static_cast<long long>(Cm->dims[4]); }
    const int n_out = p.chunk_add ? N / 4 : N;
    const int out_c = up ? p.phase_c : n_out;
    if (static_cast<int>(Cm->dims[0]) != out_c) { fprintf(stderr, "emu pw_gemm: C channels %d != %d\n", (int)Cm->dims[0], out_c); return 1; }
    // hazards the sequential restatement would hide: a kernel whose tiles run concurrently must not write where another
    // tile still reads.  Spatial taps / pixel-shuffle outputs may not overlap the input at all; a 1x1 may only overlap
    // it pixel for pixel (same base and pitch), and then must not widen or narrow the row
    {
        auto extent = [](const FakeMap* m) {
            size_t e = m->dims[0] * 2;
            for (uint32_t d = 1; d < m->rank; ++d) e += (m->dims[d] - 1) * m->strides[d - 1];
            return e;
        };
        const uint8_t *a0 = A->ptr, *a1 = A->ptr + extent(A), *c0 = Cm->ptr, *c1 = Cm->ptr + extent(Cm);
        const bool overlap = a0 < c1 && c0 < a1;
        if (overlap) {
            const bool same_rows = !up && taps == 1 && A->rank == Cm->rank && A->strides[A->rank == 2 ? 0 : 1] == Cm->strides[Cm->rank == 2 ? 0 : 1];
            // channel windows of one cat buffer (same pitch, disjoint channel ranges) are fine for any kind
            const size_t pitch = A->strides[A->rank == 2 ? 0 : 1];
            const size_t a_off = static_cast<size_t>(a0 - (a0 < c0 ? a0 : c0)) % pitch, c_off = static_cast<size_t>(c0 - (a0 < c0 ? a0 : c0)) % pitch;
            const bool disjoint_channels = pitch == Cm->strides[Cm->rank == 2 ? 0 : 1] &&
                                           (a_off + A->dims[0] * 2 <= c_off || c_off + Cm->dims[0] * 2 <= a_off) &&
                                           A->dims[0] * 2 + Cm->dims[0] * 2 <= pitch;
            if (!disjoint_channels && !(same_rows && a0 == c0)) {
                fprintf(stderr, "emu pw_gemm: output overlaps the input (taps %d, upsample %d)\n", taps, up ? 1 : 0);
                return 1;
            }
        }
    }
    report(A->ptr, map_extent(A), 0);
    report(B->ptr, map_extent(B), 0);
    report(Cm->ptr, map_extent(Cm), 1);
    {
        const size_t px = static_cast<size_t>(gw) * gh;
        if (p.n_res > 0 && px) report(p.r1, ((px - 1) * p.r1_pitch + n_out) * 2, 0);
        if (p.n_res > 1 && px) report(p.r2, ((px - 1) * p.r2_pitch + n_out) * 2, 0);
    }
    // weights as float [N][Ktot]
    std::vector<float> W(static_cast<size_t>(N) * Ktot);
    {
        const __half* w = reinterpret_cast<const __half*>(B->ptr);
        for (size_t i = 0; i < W.size(); ++i) W[i] = h2f(w[i]);
    }
    std::vector<float> a(static_cast<size_t>(Ktot)), acc(static_cast<size_t>(N)), o(static_cast<size_t>(n_out));
    std::vector<uint8_t> have(static_cast<size_t>(taps));
    for (long long gy = 0; gy < gh; ++gy) {
        for (long long gx = 0; gx < gw; ++gx) {
            for (int t = 0; t < taps; ++t) {
                const uint8_t* src = two_d ? map_at(A, 0, gx, 0, 0, 0)
                                           : map_at(A, 0, p.tap_px[t], gx + p.tap_dx[t], p.tap_py[t], gy + p.tap_dy[t]);
                have[t] = src != nullptr;
                if (src) {
                    const __half* s = reinterpret_cast<const __half*>(src);
                    for (int c = 0; c < Cin; ++c) a[static_cast<size_t>(t) * Cin + c] = h2f(s[c]);
                }
            }
            for (int n = 0; n < N; ++n) {
                const float* w = W.data() + static_cast<size_t>(n) * Ktot;
                float s = 0.f;
                for (int t = 0; t < taps; ++t) {
                    if (!have[t]) continue;
                    const float* wt = w + static_cast<size_t>(t) * Cin;
                    const float* at = a.data() + static_cast<size_t>(t) * Cin;
                    float st = 0.f;
                    for (int c = 0; c < Cin; ++c) st += at[c] * wt[c];
                    s += st;
                }
                if (p.bias) s += h2f(p.bias[n]);
                if (p.act == ACT_WSILU) s = wsilu(s);
                acc[n] = s;
            }
            if (p.chunk_add) for (int j = 0; j < n_out; ++j) o[j] = ((acc[4 * j] + acc[4 * j + 1]) + acc[4 * j + 2]) + acc[4 * j + 3];
            else for (int j = 0; j < n_out; ++j) o[j] = acc[j];
            const long long pix = gy * p.res_w + gx;
            for (int j = 0; j < n_out; ++j) {
                float v = o[j];
                if (p.n_res > 0) {
                    const float r = h2f(p.r1[pix * p.r1_pitch + j]);
                    if (p.act == ACT_GDN) v = r * (1.f / sqrtf(v));
                    else if (p.act == ACT_IGDN) v = r * sqrtf(v);
                    else v += r;
                }
                if (p.n_res > 1) v += h2f(p.r2[pix * p.r2_pitch + j]);
                if (p.qscale) v *= h2f(p.qscale[j]);
                const uint8_t* dst;
                if (two_d) dst = map_at(Cm, j, gx, 0, 0, 0);
                else if (up) { const int ph = j / p.phase_c; dst = map_at(Cm, j - ph * p.phase_c, ph & 1, gx, ph >> 1, gy); }
                else dst = map_at(Cm, j, 0, gx, 0, gy);
                if (dst) *reinterpret_cast<__half*>(const_cast<uint8_t*>(dst)) = f2h(v);
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ elementwise
template <class T> static T arg(void** args, int i) { return *static_cast<T*>(args[i]); }

static int emu_dw3x3(void** g)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1);
    __half* out = arg<__half*>(g, 2); const int op = arg<int>(g, 3);
    const __half* w = arg<const __half*>(g, 4); const int C = arg<int>(g, 5), W = arg<int>(g, 6), H = arg<int>(g, 7);
    if (in == out) { fprintf(stderr, "emu dw3x3: in-place\n"); return 1; }
    if (W > 0 && H > 0) {
        const size_t px = static_cast<size_t>(W) * H;
        report(in, ((px - 1) * ip + C) * 2, 0);
        report(w, static_cast<size_t>(9) * C * 2, 0);
        report(out, ((px - 1) * op + C) * 2, 1);
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        const int yy = y + ky - 1, xx = x + kx - 1;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        s = fmaf(h2f(in[(static_cast<long long>(yy) * W + xx) * ip + c]), h2f(w[(ky * 3 + kx) * C + c]), s);
                    }
                out[(static_cast<long long>(y) * W + x) * op + c] = f2h(s);
            }
    return 0;
}

static int emu_unshuffle8(void** g)
{
    const __half* x = arg<const __half*>(g, 0);
    const int Cs = arg<int>(g, 1), H = arg<int>(g, 2), W = arg<int>(g, 3);
    const long long sc = arg<long long>(g, 4), sh = arg<long long>(g, 5), sw = arg<long long>(g, 6);
    __half* out = arg<__half*>(g, 7); const int op = arg<int>(g, 8), W8 = arg<int>(g, 9), H8 = arg<int>(g, 10);
    for (int h8 = 0; h8 < H8; ++h8)
        for (int w8 = 0; w8 < W8; ++w8)
            for (int cs = 0; cs < Cs; ++cs)
                for (int dy = 0; dy < 8; ++dy)
                    for (int dx = 0; dx < 8; ++dx) {
                        const int sy = h8 * 8 + dy < H ? h8 * 8 + dy : H - 1, sx = w8 * 8 + dx < W ? w8 * 8 + dx : W - 1;
                        out[(static_cast<long long>(h8) * W8 + w8) * op + cs * 64 + dy * 8 + dx] = x[cs * sc + sy * sh + sx * sw];
                    }
    return 0;
}

static int emu_shuffle8(void** g, int CS)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1), W8 = arg<int>(g, 2), H8 = arg<int>(g, 3);
    __half* out = arg<__half*>(g, 4); const int clamp = arg<int>(g, 5);
    const long long Wo = static_cast<long long>(W8) * 8;
    for (int h8 = 0; h8 < H8; ++h8)
        for (int w8 = 0; w8 < W8; ++w8)
            for (int cs = 0; cs < CS; ++cs)
                for (int dy = 0; dy < 8; ++dy)
                    for (int dx = 0; dx < 8; ++dx) {
                        float t = h2f(in[(static_cast<long long>(h8) * W8 + w8) * ip + cs * 64 + dy * 8 + dx]);
                        if (clamp) t = fminf(fmaxf(t, -0.5f), 0.5f);
                        out[((static_cast<long long>(h8) * 8 + dy) * Wo + w8 * 8 + dx) * CS + cs] = f2h(t);
                    }
    return 0;
}

static int emu_pad_crop(void** g)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1), Wi = arg<int>(g, 2), Hi = arg<int>(g, 3);
    __half* out = arg<__half*>(g, 4); const int op = arg<int>(g, 5), Wo = arg<int>(g, 6), Ho = arg<int>(g, 7), C = arg<int>(g, 8);
    for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Wo; ++x) {
            const int sx = x < Wi ? x : Wi - 1, sy = y < Hi ? y : Hi - 1;
            memcpy(out + (static_cast<long long>(y) * Wo + x) * op, in + (static_cast<long long>(sy) * Wi + sx) * ip, static_cast<size_t>(C) * 2);
        }
    return 0;
}

static int emu_scale_channels(void** g)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1); const __half* q = arg<const __half*>(g, 2);
    __half* out = arg<__half*>(g, 3); const int op = arg<int>(g, 4); const long long npix = arg<long long>(g, 5); const int C = arg<int>(g, 6);
    for (long long p = 0; p < npix; ++p)
        for (int c = 0; c < C; ++c) out[p * op + c] = f2h(h2f(in[p * ip + c]) * h2f(q[c]));
    return 0;
}

static int emu_mul_clamp_min(void** g)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1); const __half* q = arg<const __half*>(g, 2); const int qp = arg<int>(g, 3);
    __half* out = arg<__half*>(g, 4); const int op = arg<int>(g, 5); const long long npix = arg<long long>(g, 6); const int C = arg<int>(g, 7);
    for (long long p = 0; p < npix; ++p)
        for (int c = 0; c < C; ++c) out[p * op + c] = f2h(h2f(in[p * ip + c]) * fmaxf(h2f(q[p * qp + c]), 0.5f));
    return 0;
}

static int emu_round_z(void** g)
{
    const __half* z = arg<const __half*>(g, 0); __half* zh = arg<__half*>(g, 1); int8_t* zi = arg<int8_t*>(g, 2); const long long n = arg<long long>(g, 3);
    for (long long i = 0; i < n; ++i) {
        const float v = fminf(fmaxf(round_away(h2f(z[i])), -64.f), 63.f);
        zh[i] = f2h(v);
        zi[i] = static_cast<int8_t>(v);
    }
    return 0;
}

static int emu_int8_to_half(void** g)
{
    const int8_t* x = arg<const int8_t*>(g, 0); __half* o = arg<__half*>(g, 1); const long long n = arg<long long>(g, 2);
    for (long long i = 0; i < n; ++i) o[i] = f2h(static_cast<float>(x[i]));
    return 0;
}

static int emu_scan(void** g)
{
    const int32_t* counts = arg<const int32_t*>(g, 0); int32_t* offs = arg<int32_t*>(g, 1); int32_t* total = arg<int32_t*>(g, 2); const int n = arg<int>(g, 3);
    int run = 0;
    for (int i = 0; i < n; ++i) { offs[i] = run; run += counts[i]; }
    offs[n] = run;
    *total = run;
    return 0;
}

// ---- frame IO (csrc/frame_io.cu): every fp16 step rounded like the kernels (integer results are exact)
static inline __half norm_u8(uint8_t v) { return f2h(h2f(f2h(static_cast<float>(v) / 255.0f)) - 0.5f); }
static inline __half to_255(__half h) { return f2h(fminf(fmaxf(h2f(f2h(h2f(h) * 255.0f)), 0.f), 255.f)); }

static int emu_yuv420_to_frame(void** g)
{
    const uint8_t *yp = arg<const uint8_t*>(g, 0), *up = arg<const uint8_t*>(g, 1), *vp = arg<const uint8_t*>(g, 2);
    const int H = arg<int>(g, 3), W = arg<int>(g, 4);
    __half* x = arg<__half*>(g, 5);
    const long long sc = arg<long long>(g, 6), sh = arg<long long>(g, 7), sw = arg<long long>(g, 8);
    const int Wc = W >> 1;
    for (int yy = 0; yy < H; ++yy)
        for (int xx = 0; xx < W; ++xx) {
            __half* dst = x + yy * sh + xx * sw;
            dst[0] = norm_u8(yp[static_cast<long long>(yy) * W + xx]);
            dst[sc] = norm_u8(up[static_cast<long long>(yy >> 1) * Wc + (xx >> 1)]);
            dst[2 * sc] = norm_u8(vp[static_cast<long long>(yy >> 1) * Wc + (xx >> 1)]);
        }
    return 0;
}

static int emu_frame_to_yuv420(void** g)
{
    const __half* x = arg<const __half*>(g, 0);
    const long long sc = arg<long long>(g, 1), sh = arg<long long>(g, 2), sw = arg<long long>(g, 3);
    const int H = arg<int>(g, 4), W = arg<int>(g, 5);
    uint8_t *yp = arg<uint8_t*>(g, 6), *up = arg<uint8_t*>(g, 7), *vp = arg<uint8_t*>(g, 8);
    const int Wc = W >> 1;
    for (int by = 0; by < (H >> 1); ++by)
        for (int bx = 0; bx < Wc; ++bx) {
            float su = 0.f, sv = 0.f;
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 2; ++c) {
                    const int yy = 2 * by + r, xx = 2 * bx + c;
                    const __half* src = x + yy * sh + xx * sw;
                    yp[static_cast<long long>(yy) * W + xx] = static_cast<uint8_t>(rintf(h2f(to_255(f2h(h2f(src[0]) + 0.5f)))));
                    su += h2f(f2h(h2f(src[sc]) + 0.5f));
                    sv += h2f(f2h(h2f(src[2 * sc]) + 0.5f));
                }
            up[static_cast<long long>(by) * Wc + bx] = static_cast<uint8_t>(h2f(to_255(f2h(su / 4.f))));
            vp[static_cast<long long>(by) * Wc + bx] = static_cast<uint8_t>(h2f(to_255(f2h(sv / 4.f))));
        }
    return 0;
}

static int emu_sse_u8(void** g)
{
    const uint8_t *a = arg<const uint8_t*>(g, 0), *b = arg<const uint8_t*>(g, 1);
    const long long n = arg<long long>(g, 2);
    unsigned long long* out = arg<unsigned long long*>(g, 3);
    unsigned long long acc = 0;
    for (long long i = 0; i < n; ++i) { const int d = static_cast<int>(a[i]) - static_cast<int>(b[i]); acc += static_cast<unsigned long long>(d * d); }
    *out += acc;
    return 0;
}

// ---- DCVC-family ops (csrc/family_ops.cu)
static int emu_square(void** g)
{
    const __half* in = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1);
    __half* out = arg<__half*>(g, 2); const int op = arg<int>(g, 3); const long long npix = arg<long long>(g, 4); const int C = arg<int>(g, 5);
    for (long long p = 0; p < npix; ++p)
        for (int c = 0; c < C; ++c) out[p * op + c] = f2h(h2f(in[p * ip + c]) * h2f(in[p * ip + c]));
    return 0;
}

// one rounding, like the device's HFMA: the product of two halfs and the sum are exact in double for all but
// astronomically separated exponents
static inline __half hfma(__half a, __half b, __half c)
{
    return __double2half(static_cast<double>(h2f(a)) * static_cast<double>(h2f(b)) + static_cast<double>(h2f(c)));
}

static int emu_warp_bilinear(void** g)
{
    const __half* im = arg<const __half*>(g, 0); const int ip = arg<int>(g, 1); const __half* flow = arg<const __half*>(g, 2);
    const long long fc = arg<long long>(g, 3), fh = arg<long long>(g, 4), fw = arg<long long>(g, 5);
    __half* out = arg<__half*>(g, 6); const int op = arg<int>(g, 7), W = arg<int>(g, 8), H = arg<int>(g, 9), C = arg<int>(g, 10);
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
            float x_pos = h2f(flow[h * fh + w * fw]) + static_cast<float>(w);
            float y_pos = h2f(flow[fc + h * fh + w * fw]) + static_cast<float>(h);
            x_pos = fminf(fmaxf(x_pos, 0.f), static_cast<float>(W - 1));
            y_pos = fminf(fmaxf(y_pos, 0.f), static_cast<float>(H - 1));
            const int x0 = static_cast<int>(floorf(x_pos)), y0 = static_cast<int>(floorf(y_pos));
            const int x1 = x0 + 1 < W ? x0 + 1 : W - 1, y1 = y0 + 1 < H ? y0 + 1 : H - 1;
            const float w_r = x_pos - static_cast<float>(x0), w_l = 1.f - w_r, w_b = y_pos - static_cast<float>(y0), w_t = 1.f - w_b;
            const __half wa = f2h(w_l * w_t), wb = f2h(w_l * w_b), wc = f2h(w_r * w_t), wd = f2h(w_r * w_b);
            for (int c = 0; c < C; ++c) {
                __half r = hfma(im[(static_cast<long long>(y0) * W + x0) * ip + c], wa, f2h(0.f));
                r = hfma(im[(static_cast<long long>(y1) * W + x0) * ip + c], wb, r);
                r = hfma(im[(static_cast<long long>(y0) * W + x1) * ip + c], wc, r);
                r = hfma(im[(static_cast<long long>(y1) * W + x1) * ip + c], wd, r);
                out[(static_cast<long long>(h) * W + w) * op + c] = r;
            }
        }
    return 0;
}

// ---- entropy-parameter kernels: one latent pixel at a time, channels in ascending order (the order the warp-ballot
//      compaction produces)
static inline int active_group(const EntropyDev& d, int h, int w)
{
    if (d.ng == 2) return ((h + w) & 1) ^ d.step;
    const int p = ((h & 1) << 1) | (w & 1);
    const int x = (d.step == 0) ? 0 : (d.step == 1 ? 3 : (d.step == 2 ? 2 : 1));
    return p ^ x;
}
static inline bool kept(const EntropyDev& d, __half s) { return h2f(s) > h2f(d.thres); }
static inline uint16_t hbits(__half h) { uint16_t b; memcpy(&b, &h, 2); return b; }

static int emu_enc_step(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        const int w = static_cast<int>(pix % d.W), h = static_cast<int>(pix / d.W);
        const int grp = active_group(d, h, w);
        int count = 0;
        for (int c = 0; c < d.G; ++c) {
            const int ch = grp * d.G + c;
            __half yv = d.y[pix * d.y_pitch + ch];
            if (d.q_enc) yv = f2h(h2f(yv) * h2f(d.q_enc[ch]));
            if (d.qdiv) {
                const __half rc = f2h(1.0f / fmaxf(h2f(d.qdiv[pix * d.q_pitch + ch]), 0.5f));
                yv = f2h(h2f(yv) * h2f(rc));
            }
            const __half mv = d.means[pix * d.m_pitch + ch];
            const __half sv = d.scales[pix * d.p_pitch + ch];
            const __half res = f2h(h2f(yv) - h2f(mv));
            float q = round_away(h2f(res));
            const bool k = kept(d, sv);
            if (!k) q = 0.f;
            q = fminf(fmaxf(q, -128.f), 127.f);
            d.acc[pix * d.acc_pitch + ch] = f2h(h2f(f2h(q)) + h2f(mv));
            if (d.yq) d.yq[pix * (d.ng * d.G) + ch] = static_cast<int8_t>(q);
            if (d.sym_raw) d.sym_raw[pix * d.G + c] = static_cast<int16_t>((static_cast<int>(q) << 8) + d.lut[hbits(sv)]);
            count += k ? 1 : 0;
        }
        if (d.step == 0)
            for (int c = 0; c < d.ng * d.G; ++c)
                if (c / d.G != grp) d.acc[pix * d.acc_pitch + c] = f2h(0.f);
        if (d.counts) d.counts[pix] = count;
    }
    return 0;
}

static int emu_dec_index(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        const int w = static_cast<int>(pix % d.W), h = static_cast<int>(pix / d.W);
        const int grp = d.full ? 0 : active_group(d, h, w);
        int count = 0;
        for (int c = 0; c < d.G; ++c) {
            const __half sv = d.scales[pix * d.p_pitch + grp * d.G + c];
            d.idx_raw[pix * d.G + c] = d.lut[hbits(sv)];
            count += kept(d, sv) ? 1 : 0;
        }
        d.counts[pix] = count;
    }
    return 0;
}

template <class T>
static int emu_compact(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    const T* raw = arg<const T*>(g, 1); const int32_t* offs = arg<const int32_t*>(g, 2); T* out = arg<T*>(g, 3);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        const int w = static_cast<int>(pix % d.W), h = static_cast<int>(pix / d.W);
        const int grp = d.full ? 0 : active_group(d, h, w);
        int base = offs[pix];
        for (int c = 0; c < d.G; ++c)
            if (kept(d, d.scales[pix * d.p_pitch + grp * d.G + c])) out[base++] = raw[pix * d.G + c];
    }
    return 0;
}

static int emu_dec_restore(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    const int32_t* offs = arg<const int32_t*>(g, 1); const int8_t* dec = arg<const int8_t*>(g, 2);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        const int w = static_cast<int>(pix % d.W), h = static_cast<int>(pix / d.W);
        const int grp = active_group(d, h, w);
        int base = offs[pix];
        for (int c = 0; c < d.G; ++c) {
            const int ch = grp * d.G + c;
            const bool k = kept(d, d.scales[pix * d.p_pitch + ch]);
            const float q = k ? static_cast<float>(dec[base++]) : 0.f;
            d.acc[pix * d.acc_pitch + ch] = f2h(h2f(f2h(q)) + h2f(d.means[pix * d.p_pitch + ch]));
        }
        if (d.step == 0)
            for (int c = 0; c < d.ng * d.G; ++c)
                if (c / d.G != grp) d.acc[pix * d.acc_pitch + c] = f2h(0.f);
    }
    return 0;
}

static int emu_build_symbols_full(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        int count = 0;
        for (int c = 0; c < d.G; ++c) {
            const __half sv = d.scales[pix * d.p_pitch + c];
            d.sym_raw[pix * d.G + c] = static_cast<int16_t>((static_cast<int>(d.yq[pix * d.G + c]) << 8) + d.lut[hbits(sv)]);
            count += kept(d, sv) ? 1 : 0;
        }
        d.counts[pix] = count;
    }
    return 0;
}

static int emu_recover_dense(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    const int32_t* offs = arg<const int32_t*>(g, 1); const int8_t* dec = arg<const int8_t*>(g, 2);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        int base = offs[pix];
        for (int c = 0; c < d.G; ++c)
            d.yq[pix * d.G + c] = kept(d, d.scales[pix * d.p_pitch + c]) ? dec[base++] : 0;
    }
    return 0;
}

static int emu_restore_dense(void** g)
{
    const EntropyDev& d = *static_cast<const EntropyDev*>(g[0]);
    for (long long pix = 0; pix < static_cast<long long>(d.H) * d.W; ++pix) {
        const int w = static_cast<int>(pix % d.W), h = static_cast<int>(pix / d.W);
        const int grp = active_group(d, h, w);
        for (int c = 0; c < d.G; ++c) {
            const int ch = grp * d.G + c;
            const float q = static_cast<float>(d.yq[pix * (d.ng * d.G) + ch]);
            d.acc[pix * d.acc_pitch + ch] = f2h(h2f(f2h(q)) + h2f(d.means[pix * d.m_pitch + ch]));
        }
        if (d.step == 0)
            for (int c = 0; c < d.ng * d.G; ++c)
                if (c / d.G != grp) d.acc[pix * d.acc_pitch + c] = f2h(0.f);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ dispatch
struct KernelEntry {
    const char* key;                 // substring of the mangled device function name
    int n_args;
    int sizes[12];
    int (*fn)(void**);
};
static int emu_shuffle8_3(void** g) { return emu_shuffle8(g, 3); }
static int emu_shuffle8_1(void** g) { return emu_shuffle8(g, 1); }

static const KernelEntry kTable[] = {
    { "pw_gemm_kernel", 1, { static_cast<int>(sizeof(PwGemmParams)) }, emu_pw_gemm },
    { "dw3x3_kernel", 8, { 8, 4, 8, 4, 8, 4, 4, 4 }, emu_dw3x3 },
    { "unshuffle8_pad_kernel", 11, { 8, 4, 4, 4, 8, 8, 8, 8, 4, 4, 4 }, emu_unshuffle8 },
    { "shuffle8_clamp_kernelILi3E", 6, { 8, 4, 4, 4, 8, 4 }, emu_shuffle8_3 },
    { "shuffle8_clamp_kernelILi1E", 6, { 8, 4, 4, 4, 8, 4 }, emu_shuffle8_1 },
    { "pad_crop_kernel", 9, { 8, 4, 4, 4, 8, 4, 4, 4, 4 }, emu_pad_crop },
    { "scale_channels_kernel", 7, { 8, 4, 8, 8, 4, 8, 4 }, emu_scale_channels },
    { "mul_clamp_min_kernel", 8, { 8, 4, 8, 4, 8, 4, 8, 4 }, emu_mul_clamp_min },
    { "round_z_kernel", 4, { 8, 8, 8, 8 }, emu_round_z },
    { "int8_to_half_kernel", 3, { 8, 8, 8 }, emu_int8_to_half },
    { "scan_counts_kernel", 5, { 8, 8, 8, 4, 1 }, emu_scan },
    { "entropy_enc_step_kernel", 1, { static_cast<int>(sizeof(EntropyDev)) }, emu_enc_step },
    { "entropy_dec_index_kernel", 1, { static_cast<int>(sizeof(EntropyDev)) }, emu_dec_index },
    { "compact_kernelIs", 4, { static_cast<int>(sizeof(EntropyDev)), 8, 8, 8 }, emu_compact<int16_t> },
    { "compact_kernelIh", 4, { static_cast<int>(sizeof(EntropyDev)), 8, 8, 8 }, emu_compact<uint8_t> },
    { "entropy_dec_restore_kernel", 3, { static_cast<int>(sizeof(EntropyDev)), 8, 8 }, emu_dec_restore },
    { "entropy_build_symbols_full_kernel", 1, { static_cast<int>(sizeof(EntropyDev)) }, emu_build_symbols_full },
    { "entropy_recover_dense_kernel", 3, { static_cast<int>(sizeof(EntropyDev)), 8, 8 }, emu_recover_dense },
    { "entropy_restore_dense_kernel", 1, { static_cast<int>(sizeof(EntropyDev)) }, emu_restore_dense },
    { "yuv420_to_frame_kernel", 9, { 8, 8, 8, 4, 4, 8, 8, 8, 8 }, emu_yuv420_to_frame },
    { "frame_to_yuv420_kernel", 9, { 8, 8, 8, 8, 4, 4, 8, 8, 8 }, emu_frame_to_yuv420 },
    { "sse_u8_kernel", 4, { 8, 8, 8, 8 }, emu_sse_u8 },
    { "square_kernel", 6, { 8, 4, 8, 4, 8, 4 }, emu_square },
    { "warp_bilinear_kernel", 11, { 8, 4, 8, 8, 8, 8, 8, 4, 4, 4, 4 }, emu_warp_bilinear },
};

extern "C" {

// index into the table for a mangled kernel name, or -1
int emu_lookup(const char* name)
{
    for (size_t i = 0; i < sizeof(kTable) / sizeof(kTable[0]); ++i)
        if (strstr(name, kTable[i].key)) return static_cast<int>(i);
    return -1;
}
int emu_num_args(int k) { return kTable[k].n_args; }
int emu_arg_size(int k, int i) { return kTable[k].sizes[i]; }
int emu_run(int k, void** args) { return kTable[k].fn(args); }
unsigned long long emu_map_magic() { return kMapMagic; }

}  // extern "C"
