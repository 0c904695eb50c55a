// codec_htl.cu — DCVC-UF HT-L chunk codec (8 frames per chunk, the large high-throughput model) behind the C ABI.
//
// Device-validated in round 2 (tests/test_htl_gpu.py, tests/test_reference_surface_gpu.py).
//
// B200-native counterpart of src/layers/extensions/inference/dmc_htl_proxy.{h,cpp}: the state machine and "cat"
// buffers of the HT-S codec (codec_hts.cu; dmc_htl_proxy.cpp:583-594, 700-703, 744-751) with the full-width networks of
// video_model_ht.py's non-HTS branches and the four-step prior WITH scale updates of the intra codec: one symbol run per
// step on the encoder side (:629-693), four index / decode / restore round trips on the decoder side (:764-890).
#include "codec_common.cuh"

namespace dcvc {

namespace {
constexpr int kG = 8;  // frames per chunk (video_model_ht.py:16)
constexpr int kSrcI = 192, kSrcD = 1536, kY = 256, kZ = 128, kD = 512, kM = 512, kRecon = 256;
constexpr int kP = 3 * kY;  // prior fusion width: (q_dec | scales | means)
}  // namespace

class HtlCodec : public CodecBase {
public:
    explicit HtlCodec(int device) : CodecBase(device) {}
    ~HtlCodec() override { clear_plan(); }

    void finalize(float skip_thres) override;
    int debug_fetch(const char* name, void* dst, int64_t max_bytes, int64_t* written) override;
    int arena_overflow_blocks() const override { return arena_.overflow_blocks(); }

    void add_ref(const void* frame, int H, int W, int64_t sc, int64_t sh, int64_t sw, int apply, cudaStream_t stream);
    void compress(const void* x, int H, int W, int64_t sc, int64_t sh, int64_t sw, int qp, int reset, int pad_b,
                  int pad_r, cudaStream_t stream, const uint8_t** bs, int32_t* len, int32_t* ec);
    void decompress(const uint8_t* bs, int len, int qp, int height, int width, int ec, int reset,
                    cudaStream_t stream, void* const* x_hat_out);

private:
    void plan(int height, int width);
    void clear_plan();
    void stage_qp(int qp, cudaStream_t stream);
    ActView chain(Segment& s, Level& L, ActView in, const DcbW* blocks, int n, const __half* q_last, const ActView* out);
    void build_spatial_prior(Segment& s, int k);

    // weights
    DcbW fa_i_[3], fa_m_[10], fe_[2], enc_[7], henc0_, henc1_, henc2_, hdec0_, hdec1_, hdec2_, tpe_, pf_[3], spa_[3],
        sp_[3], dec_[11], rh_[8][5];
    ConvW enc_down_, henc1_down_, henc2_down_, hdec0_up_, hdec1_up_, tpe_down_, pf3_, red_, sp3_, dec_up_, rh_out_[8];
    const __half *q_encoder_all_ = nullptr, *q_decoder_all_ = nullptr, *q_feature_all_ = nullptr;

    // plan
    int H8_ = 0, W8_ = 0, H16_ = 0, W16_ = 0, H16p_ = 0, W16p_ = 0, H64_ = 0, W64_ = 0;
    Arena arena_;
    Level l8_, l16_, l32_, l64_;
    __half *cat_enc_ = nullptr, *cat_fam_ = nullptr, *feature_i_ = nullptr, *temporal_in_ = nullptr, *head_out_ = nullptr;
    __half *y_ = nullptr, *ypad_ = nullptr, *hyp_p_ = nullptr, *cat_pf_ = nullptr, *common_ = nullptr, *cat_sp_ = nullptr,
           *sp_out_ = nullptr, *yhat_ = nullptr, *zhat_ = nullptr;
    int8_t *z_i8_ = nullptr, *decoded_ = nullptr;
    __half *q_enc_ = nullptr, *q_dec_ = nullptr, *q_feat_ = nullptr;
    int16_t *sym_raw_ = nullptr, *sym_c_[4] = { nullptr, nullptr, nullptr, nullptr };
    uint8_t *idx_raw_ = nullptr, *idx_c_ = nullptr;
    int32_t *counts_ = nullptr, *offsets_[4] = { nullptr, nullptr, nullptr, nullptr }, *totals_ = nullptr;
    int32_t* h_totals_ = nullptr;
    int16_t* h_sym_[4] = { nullptr, nullptr, nullptr, nullptr };
    uint8_t* h_idx_ = nullptr;
    int8_t *h_decoded_ = nullptr, *h_z_ = nullptr;
    size_t quarter_ = 0;
    bool memory_has_value_ = false;

    Segment s_enc0_, s_decoder_, s_reset_head_, s_fa_i_, s_fa_m_, s_fe_, s_temporal_, s_dec1_, s_dec_step_[4], s_recon_;
};

// =============================================================================== parameters
void HtlCodec::finalize(float skip_thres)
{
    finalize_begin(skip_thres);
    clear_plan();
    auto dcbs = [&](DcbW* dst, const std::string& prefix, int n) {
        for (int i = 0; i < n; ++i) dst[i] = load_dcb(prefix + std::to_string(i) + ".");
    };
    dcbs(fa_i_, "feature_adaptor_i.conv.", 3);
    dcbs(fa_m_, "feature_adaptor_m.conv.", 10);
    dcbs(fe_, "feature_extractor.conv.", 2);
    dcbs(enc_, "encoder.conv1.", 7);
    enc_down_ = load_conv("encoder.down.", DCVC_GEMM_CONV3X3_S2);
    henc0_ = load_dcb("hyper_encoder.conv.0.");
    henc1_down_ = load_conv("hyper_encoder.conv.1.down.", DCVC_GEMM_CONV2X2_S2);
    henc1_ = load_dcb("hyper_encoder.conv.1.conv.");
    henc2_down_ = load_conv("hyper_encoder.conv.2.down.", DCVC_GEMM_CONV2X2_S2);
    henc2_ = load_dcb("hyper_encoder.conv.2.conv.");
    hdec0_up_ = load_conv("hyper_decoder.conv.0.up.conv.0.", DCVC_GEMM_TCONV2X2);   // with bias (force_bias)
    hdec0_ = load_dcb("hyper_decoder.conv.0.conv.");
    hdec1_up_ = load_conv("hyper_decoder.conv.1.up.conv.0.", DCVC_GEMM_TCONV2X2);
    hdec1_ = load_dcb("hyper_decoder.conv.1.conv.");
    hdec2_ = load_dcb("hyper_decoder.conv.2.");
    tpe_down_ = load_conv("temporal_prior_encoder.conv.down.", DCVC_GEMM_CONV2X2_S2);
    tpe_ = load_dcb("temporal_prior_encoder.conv.conv.");
    dcbs(pf_, "y_prior_fusion.conv.", 3);
    pf3_ = load_conv("y_prior_fusion.conv.3.", DCVC_GEMM_PW);
    red_ = load_conv("y_spatial_prior_reduction.", DCVC_GEMM_PW);
    for (int i = 0; i < 3; ++i) spa_[i] = load_dcb("y_spatial_prior_adaptor_" + std::to_string(i + 1) + ".");
    dcbs(sp_, "y_spatial_prior.conv.", 3);
    sp3_ = load_conv("y_spatial_prior.conv.3.", DCVC_GEMM_PW);
    dec_up_ = load_conv("decoder.up.conv.0.", DCVC_GEMM_CONV3X3_PS2);               // 3x3 SubpelConv2x with bias
    dcbs(dec_, "decoder.conv1.", 11);
    for (int i = 0; i < kG; ++i) {
        const std::string p = "recon_head.conv." + std::to_string(i) + ".";
        for (int j = 0; j < 5; ++j) rh_[i][j] = load_dcb(p + std::to_string(j) + ".");
        rh_out_[i] = load_conv(p + "5.", DCVC_GEMM_PW);
    }
    q_encoder_all_ = upload_param("q_encoder");
    q_decoder_all_ = upload_param("q_decoder");
    q_feature_all_ = upload_param("q_feature");
    finalize_end();
}

// =============================================================================== plan
void HtlCodec::clear_plan()
{
    Segment* segs[] = { &s_enc0_, &s_decoder_, &s_reset_head_, &s_fa_i_, &s_fa_m_, &s_fe_, &s_temporal_, &s_dec1_,
                        &s_dec_step_[0], &s_dec_step_[1], &s_dec_step_[2], &s_dec_step_[3], &s_recon_ };
    for (Segment* s : segs) s->reset();
    if (h_totals_) { cudaFreeHost(h_totals_); h_totals_ = nullptr; }
    for (int k = 0; k < 4; ++k) if (h_sym_[k]) { cudaFreeHost(h_sym_[k]); h_sym_[k] = nullptr; }
    if (h_idx_) { cudaFreeHost(h_idx_); h_idx_ = nullptr; }
    if (h_decoded_) { cudaFreeHost(h_decoded_); h_decoded_ = nullptr; }
    if (h_z_) { cudaFreeHost(h_z_); h_z_ = nullptr; }
    arena_.release();
    H8_ = W8_ = 0;
    memory_has_value_ = false;
}

ActView HtlCodec::chain(Segment& s, Level& L, ActView in, const DcbW* blocks, int n, const __half* q_last, const ActView* out)
{
    ActView t = in;
    for (int i = 0; i < n; ++i) {
        const bool last = (i == n - 1);
        const bool external = (t.ptr != L.A && t.ptr != L.B);
        ActView first_out = make_view(L.B, blocks[i].c, blocks[i].c, t.W, t.H);
        const ActView* o = last ? out : nullptr;
        if (!o && external && !blocks[i].adaptor) o = &first_out;
        t = dcb(s, L, t, blocks[i], false, last ? q_last : nullptr, o);
    }
    return t;
}

void HtlCodec::build_spatial_prior(Segment& s, int k)
{
    // adaptor_k(cat(y_hat_so_far, reduced)) -> 3 blocks -> 1x1 -> (scales | means) of step k  (dmc_htl_proxy.cpp:645-648)
    Level L = l16_;
    const ActView cat = make_view(cat_sp_, 2 * kY, 2 * kY, W16_, H16_);
    ActView t = dcb(s, L, cat, spa_[k - 1], false, nullptr, nullptr);
    for (int i = 0; i < 3; ++i) t = dcb(s, L, t, sp_[i], false, nullptr, nullptr);
    conv1x1(s, t, make_view(sp_out_, 2 * kY, 2 * kY, W16_, H16_), sp3_);
}

void HtlCodec::plan(int height, int width)
{
    const int H = round_up(height, 16), W = round_up(width, 16);
    const int H8 = H / 8, W8 = W / 8;
    if (H8 == H8_ && W8 == W8_) return;
    if (!finalized_) throw std::runtime_error("set_param/finalize_params must be called first");
    clear_plan();
    H8_ = H8; W8_ = W8;
    H16_ = H8 / 2; W16_ = W8 / 2;
    H16p_ = round_up(H16_, 4); W16p_ = round_up(W16_, 4);
    const int H32 = H16p_ / 2, W32 = W16p_ / 2;
    H64_ = H16p_ / 4; W64_ = W16p_ / 4;
    const size_t p8 = static_cast<size_t>(H8) * W8, p16 = static_cast<size_t>(H16_) * W16_;
    const size_t p16p = static_cast<size_t>(H16p_) * W16p_, p32 = static_cast<size_t>(H32) * W32;
    const size_t p64 = static_cast<size_t>(H64_) * W64_;
    quarter_ = p16 * (kY / 4);
    const bool padded = (H16p_ != H16_) || (W16p_ != W16_);

    size_t bytes = p8 * 2 * (2048 + 1024 + kSrcI + kM + kSrcI * kG + 4 * kD);
    bytes += p16p * 2 * (4 * kP + 2 * kY + kP + kP + 2 * kY + 2 * kY + kY + kY);
    bytes += p32 * 2 * 4 * kY + p64 * 2 * 4 * kY + p64 * kZ * 3;
    bytes += quarter_ * (2 + 1 + 4 * 2 + 1 + 1) + p16 * 4 * 6 + (2u << 20) + 64 * 4096;
    arena_.reserve(bytes);
    dbg_base_ = arena_.base();
    dbg_bytes_ = bytes;

    cat_enc_ = arena_.halves(p8 * 2048);
    cat_fam_ = arena_.halves(p8 * 1024);
    feature_i_ = arena_.halves(p8 * kSrcI);
    temporal_in_ = arena_.halves(p8 * kM);
    head_out_ = arena_.halves(p8 * kSrcI * kG);
    l8_.H = H8; l8_.W = W8;
    l8_.A = arena_.halves(p8 * kD); l8_.B = arena_.halves(p8 * kD);
    l8_.T1 = arena_.halves(p8 * kD); l8_.T2 = arena_.halves(p8 * kD);
    l16_.H = H16p_; l16_.W = W16p_;
    l16_.A = arena_.halves(p16p * kP); l16_.B = arena_.halves(p16p * kP);
    l16_.T1 = arena_.halves(p16p * kP); l16_.T2 = arena_.halves(p16p * kP);
    l32_.H = H32; l32_.W = W32;
    l32_.A = arena_.halves(p32 * kY); l32_.B = arena_.halves(p32 * kY);
    l32_.T1 = arena_.halves(p32 * kY); l32_.T2 = arena_.halves(p32 * kY);
    l64_.H = H64_; l64_.W = W64_;
    l64_.A = arena_.halves(p64 * kY); l64_.B = arena_.halves(p64 * kY);
    l64_.T1 = arena_.halves(p64 * kY); l64_.T2 = arena_.halves(p64 * kY);
    y_ = arena_.halves(p16 * kY);
    ypad_ = padded ? arena_.halves(p16p * kY) : y_;
    cat_pf_ = arena_.halves(p16 * kP);
    hyp_p_ = padded ? arena_.halves(p16p * kY) : nullptr;
    common_ = arena_.halves(p16 * kP);
    cat_sp_ = arena_.halves(p16 * 2 * kY);
    sp_out_ = arena_.halves(p16 * 2 * kY);
    yhat_ = arena_.halves(p16 * kY);
    zhat_ = arena_.halves(p64 * kZ);
    z_i8_ = static_cast<int8_t*>(arena_.alloc(p64 * kZ));
    decoded_ = static_cast<int8_t*>(arena_.alloc(quarter_));
    sym_raw_ = static_cast<int16_t*>(arena_.alloc(quarter_ * 2));
    idx_raw_ = static_cast<uint8_t*>(arena_.alloc(quarter_));
    idx_c_ = static_cast<uint8_t*>(arena_.alloc(quarter_));
    counts_ = static_cast<int32_t*>(arena_.alloc(p16 * 4));
    totals_ = static_cast<int32_t*>(arena_.alloc(64));
    for (int k = 0; k < 4; ++k) {
        offsets_[k] = static_cast<int32_t*>(arena_.alloc((p16 + 1) * 4));
        sym_c_[k] = static_cast<int16_t*>(arena_.alloc(quarter_ * 2));
    }
    q_enc_ = arena_.halves(kD); q_dec_ = arena_.halves(kD); q_feat_ = arena_.halves(kD);
    CK(cudaMallocHost(&h_totals_, 64));
    for (int k = 0; k < 4; ++k) CK(cudaMallocHost(&h_sym_[k], quarter_ * 2));
    CK(cudaMallocHost(&h_idx_, quarter_));
    CK(cudaMallocHost(&h_decoded_, quarter_));
    CK(cudaMallocHost(&h_z_, p64 * kZ));

    // views of the "cat" buffers (same aliasing as the HT-S codec)
    const ActView v_cat_enc = make_view(cat_enc_, 2048, 2048, W8, H8);
    const ActView v_up_out = make_view(cat_enc_ + 1024, kD, 2048, W8, H8);
    const ActView v_cat_dec = make_view(cat_enc_ + 1024, 1024, 2048, W8, H8);
    const ActView v_ctx = make_view(cat_enc_ + 1536, kD, 2048, W8, H8);
    const ActView v_cat_fam = make_view(cat_fam_, 1024, 1024, W8, H8);
    const ActView v_memory = make_view(cat_fam_, kM, 1024, W8, H8);
    const ActView v_feature_p = make_view(cat_fam_ + 512, kD, 1024, W8, H8);
    const ActView v_feature_i = make_view(feature_i_, kSrcI, kSrcI, W8, H8);
    const ActView v_temporal_in = make_view(temporal_in_, kM, kM, W8, H8);
    const ActView v_y = make_view(y_, kY, kY, W16_, H16_);
    const ActView v_ypad = make_view(ypad_, kY, kY, W16p_, H16p_);
    const ActView v_hyper = make_view(cat_pf_, kY, kP, W16_, H16_);
    const ActView v_temporal = make_view(cat_pf_ + kY, 2 * kY, kP, W16_, H16_);
    const ActView v_cat_pf = make_view(cat_pf_, kP, kP, W16_, H16_);
    const ActView v_common = make_view(common_, kP, kP, W16_, H16_);
    const ActView v_qdec = make_view(common_, kY, kP, W16_, H16_);
    const ActView v_acc = make_view(cat_sp_, kY, 2 * kY, W16_, H16_);
    const ActView v_reduced = make_view(cat_sp_ + kY, kY, 2 * kY, W16_, H16_);
    const ActView v_yhat = make_view(yhat_, kY, kY, W16_, H16_);
    const int npix16 = static_cast<int>(p16);

    // scales / means of step k: the prior fusion output for k = 0, the spatial prior output afterwards
    auto step_args = [&](int k) {
        EntropyStepArgs a;
        a.H = H16_; a.W = W16_; a.G = kY / 4; a.step = k;
        if (k == 0) { a.scales = common_ + kY; a.means = common_ + 2 * kY; a.p_pitch = kP; a.m_pitch = kP; }
        else { a.scales = sp_out_; a.means = sp_out_ + kY; a.p_pitch = 2 * kY; a.m_pitch = 2 * kY; }
        a.y_hat_acc = cat_sp_; a.acc_pitch = 2 * kY;
        a.skip_thres = skip_thres_; a.scale_lut = lut_;
        return a;
    };
    // z_hat -> hyper params (cropped) into cat_pf[0:256], prior fusion -> common, reduction -> cat_sp[256:512]
    auto build_params = [&](Segment& s) {
        ActView u32 = make_view(l32_.A, kY, kY, W32, H32);
        add_gemm(s, GEMM_TCONV2X2, make_view(zhat_, kZ, kZ, W64_, H64_), u32, hdec0_up_.w, hdec0_up_.b, 4 * kY, ACT_NONE, 0, nullptr, nullptr, nullptr);
        u32 = dcb(s, l32_, u32, hdec0_, true, nullptr, nullptr);
        ActView u16 = make_view(l16_.A, kY, kY, W16p_, H16p_);
        add_gemm(s, GEMM_TCONV2X2, u32, u16, hdec1_up_.w, hdec1_up_.b, 4 * kY, ACT_NONE, 0, nullptr, nullptr, nullptr);
        u16 = dcb(s, l16_, u16, hdec1_, true, nullptr, nullptr);
        if (padded) {
            const ActView hp = make_view(hyp_p_, kY, kY, W16p_, H16p_);
            dcb(s, l16_, u16, hdec2_, false, nullptr, &hp);
            s.elem([hp, v_hyper](cudaStream_t st) { return launch_pad_crop(hp, v_hyper, st); });
        } else {
            dcb(s, l16_, u16, hdec2_, false, nullptr, &v_hyper);
        }
        Level L = l16_;
        ActView t = chain(s, L, v_cat_pf, pf_, 3, nullptr, nullptr);
        conv1x1(s, t, v_common, pf3_);
        conv1x1(s, v_common, v_reduced, red_);
    };
    auto build_temporal = [&](Segment& s) {
        // temporal prior: (memory * q_feature) -> 2x2/s2 -> block (shortcut) -> cat_pf[256:768]  (dmc_htl_proxy.cpp:619-620)
        const __half* qf = q_feat_;
        s.elem([v_memory, qf, v_temporal_in](cudaStream_t st) { return launch_scale_channels(v_memory, qf, v_temporal_in, st); });
        Level L = l16_;
        ActView d = make_view(L.A, 2 * kY, 2 * kY, W16_, H16_);
        add_gemm(s, GEMM_CONV2X2_S2, v_temporal_in, d, tpe_down_.w, tpe_down_.b, 2 * kY, ACT_NONE, 0, nullptr, nullptr, nullptr);
        dcb(s, L, d, tpe_, true, nullptr, &v_temporal);
    };

    // ------------------------------------------------------------------ feature memory / context segments
    chain(s_fa_i_, l8_, v_feature_i, fa_i_, 3, nullptr, &v_memory);
    chain(s_fa_m_, l8_, v_cat_fam, fa_m_, 10, nullptr, &v_memory);
    chain(s_fe_, l8_, v_memory, fe_, 2, nullptr, &v_ctx);
    build_temporal(s_temporal_);

    // ------------------------------------------------------------------ enc_0 (dmc_htl_proxy.cpp:609-694)
    {
        Segment& s = s_enc0_;
        ActView t = chain(s, l8_, v_cat_enc, enc_, 7, q_enc_, nullptr);
        add_gemm(s, GEMM_CONV3X3_S2, t, v_y, enc_down_.w, enc_down_.b, kY, ACT_NONE, 0, nullptr, nullptr, nullptr);
        if (padded) s.elem([v_y, v_ypad](cudaStream_t st) { return launch_pad_crop(v_y, v_ypad, st); });
        Level L16 = l16_;
        ActView h = chain(s, L16, v_ypad, &henc0_, 1, nullptr, nullptr);
        ActView d32 = make_view(l32_.A, kY, kY, W32, H32);
        add_gemm(s, GEMM_CONV2X2_S2, h, d32, henc1_down_.w, henc1_down_.b, kY, ACT_NONE, 0, nullptr, nullptr, nullptr);
        d32 = dcb(s, l32_, d32, henc1_, true, nullptr, nullptr);
        ActView d64 = make_view(l64_.A, kZ, kZ, W64_, H64_);
        add_gemm(s, GEMM_CONV2X2_S2, d32, d64, henc2_down_.w, henc2_down_.b, kZ, ACT_NONE, 0, nullptr, nullptr, nullptr);
        d64 = dcb(s, l64_, d64, henc2_, true, nullptr, nullptr);
        {
            const __half* z = static_cast<const __half*>(d64.ptr);
            __half* zh = zhat_;
            int8_t* zi = z_i8_;
            const long long n = static_cast<long long>(p64) * kZ;
            s.elem([z, zh, zi, n](cudaStream_t st) { return launch_round_z(z, zh, zi, n, st); });
        }
        build_temporal(s);
        build_params(s);
        for (int k = 0; k < 4; ++k) {
            if (k > 0) build_spatial_prior(s, k);
            EntropyStepArgs a = step_args(k);
            a.y = y_; a.y_pitch = kY;
            a.q_div = common_; a.q_pitch = kP;          // y / clamp_min(q_dec, .5) first (:625)
            a.sym_raw = sym_raw_; a.counts = counts_;
            int32_t* offs = offsets_[k];
            int32_t* tot = totals_ + k;
            int16_t* dst = sym_c_[k];
            s.elem([a](cudaStream_t st) { return launch_entropy_enc_step(a, st); });
            s.elem([a, offs, tot, npix16](cudaStream_t st) { return launch_scan_counts(a.counts, offs, tot, npix16, st); });
            s.elem([a, offs, dst](cudaStream_t st) { return launch_compact_i16(a, offs, dst, st); });
        }
        // add_and_multiply_with_clamp_min_inplace(y_hat_3, y_hat_so_far, q_dec)  (:693)
        s.elem([v_acc, v_qdec, v_yhat](cudaStream_t st) { return launch_mul_clamp_min(v_acc, v_qdec, v_yhat, st); });
    }
    // ------------------------------------------------------------------ synthesis: y_hat -> feature_p (Decoder, :41-60)
    {
        Segment& s = s_decoder_;
        add_gemm(s, GEMM_CONV3X3_PS2, v_yhat, v_up_out, dec_up_.w, dec_up_.b, 4 * kD, ACT_NONE, 0, nullptr, nullptr, nullptr);
        chain(s, l8_, v_cat_dec, dec_, 11, q_dec_, &v_feature_p);
    }
    // reset: recon head 7 without the shuffle -> feature_i  (video_model_ht.py:266-267)
    {
        Segment& s = s_reset_head_;
        ActView t = chain(s, l8_, v_feature_p, rh_[kG - 1], 5, nullptr, nullptr);
        conv1x1(s, t, v_feature_i, rh_out_[kG - 1]);
    }
    // ------------------------------------------------------------------ decoder segments (dmc_htl_proxy.cpp:764-890)
    {
        Segment& s = s_dec1_;
        const int8_t* zi = z_i8_;
        __half* zh = zhat_;
        const long long n = static_cast<long long>(p64) * kZ;
        s.elem([zi, zh, n](cudaStream_t st) { return launch_int8_to_half(zi, zh, n, st); });
        build_params(s);
        chain(s, l8_, v_memory, fe_, 2, nullptr, &v_ctx);   // context, inside dec_1 as in the reference (:775)
    }
    for (int k = 0; k < 4; ++k) {
        // segment producing the indexes of step k; for k > 0 it first restores step k-1 and predicts (scales, means)
        Segment& s = (k == 0) ? s_dec1_ : s_dec_step_[k];
        if (k > 0) {
            const EntropyStepArgs r = step_args(k - 1);
            const int32_t* offs = offsets_[k - 1];
            const int8_t* dec = decoded_;
            s.elem([r, offs, dec](cudaStream_t st) { return launch_entropy_dec_restore(r, offs, dec, st); });
            build_spatial_prior(s, k);
        }
        EntropyStepArgs a = step_args(k);
        a.idx_raw = idx_raw_; a.counts = counts_;
        int32_t* offs = offsets_[k];
        int32_t* tot = totals_ + k;
        uint8_t* dst = idx_c_;
        s.elem([a](cudaStream_t st) { return launch_entropy_dec_index(a, st); });
        s.elem([a, offs, tot, npix16](cudaStream_t st) { return launch_scan_counts(a.counts, offs, tot, npix16, st); });
        s.elem([a, offs, dst](cudaStream_t st) { return launch_compact_u8(a, offs, dst, st); });
    }
    {
        // dec_5: restore step 3, final multiply; the synthesis segments follow (:880-888)
        Segment& s = s_dec_step_[0];
        const EntropyStepArgs r = step_args(3);
        const int32_t* offs = offsets_[3];
        const int8_t* dec = decoded_;
        s.elem([r, offs, dec](cudaStream_t st) { return launch_entropy_dec_restore(r, offs, dec, st); });
        s.elem([v_acc, v_qdec, v_yhat](cudaStream_t st) { return launch_mul_clamp_min(v_acc, v_qdec, v_yhat, st); });
    }
    {
        // recon head: 8 x (5 blocks + 1x1); head i lands in head_out_[i], head 7 in feature_i (the reset reference)
        Segment& s = s_recon_;
        for (int i = 0; i < kG; ++i) {
            ActView t = chain(s, l8_, v_feature_p, rh_[i], 5, nullptr, nullptr);
            const ActView ho = (i == kG - 1) ? v_feature_i : make_view(head_out_ + static_cast<size_t>(i) * p8 * kSrcI, kSrcI, kSrcI, W8, H8);
            conv1x1(s, t, ho, rh_out_[i]);
        }
    }
    Segment* segs[] = { &s_enc0_, &s_decoder_, &s_reset_head_, &s_fa_i_, &s_fa_m_, &s_fe_, &s_temporal_, &s_dec1_,
                        &s_dec_step_[0], &s_dec_step_[1], &s_dec_step_[2], &s_dec_step_[3], &s_recon_ };
    for (Segment* s : segs) s->seal();
}

void HtlCodec::stage_qp(int qp, cudaStream_t stream)
{
    if (qp < 0 || qp >= kQpNum) throw std::runtime_error("qp out of range");
    const size_t off = static_cast<size_t>(qp) * kD;
    CK(cudaMemcpyAsync(q_enc_, q_encoder_all_ + off, kD * 2, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(q_dec_, q_decoder_all_ + off, kD * 2, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(q_feat_, q_feature_all_ + off, kD * 2, cudaMemcpyDeviceToDevice, stream));
}

// =============================================================================== reference feature
void HtlCodec::add_ref(const void* frame, int H, int W, int64_t sc, int64_t sh, int64_t sw, int apply, cudaStream_t stream)
{
    CK(cudaSetDevice(device_));
    plan(H, W);
    StreamHop hop(this, stream);
    stream = hop.run;
    if (launch_unshuffle8_pad(static_cast<const __half*>(frame), 3, H, W, sc, sh, sw, make_view(feature_i_, kSrcI, kSrcI, W8_, H8_), stream))
        throw std::runtime_error("unshuffle8 launch failed");
    ++launches;
    if (apply) {
        run(s_fa_i_, stream);
        run(s_fe_, stream);
    }
    memory_has_value_ = apply != 0;
}

// =============================================================================== compress
void HtlCodec::compress(const void* x, int H, int W, int64_t sc, int64_t sh, int64_t sw, int qp, int reset, int pad_b,
                        int pad_r, cudaStream_t stream, const uint8_t** bs, int32_t* len, int32_t* ec)
{
    CK(cudaSetDevice(device_));
    if ((H + pad_b) % 16 || (W + pad_r) % 16) throw std::runtime_error("padded size must be a multiple of 16");
    plan(H + pad_b, W + pad_r);
    StreamHop hop(this, stream);
    stream = hop.run;
    stage_qp(qp, stream);
    tev_n_ = 0;
    tick(stream);
    if (launch_unshuffle8_pad(static_cast<const __half*>(x), 3 * kG, H, W, sc, sh, sw, make_view(cat_enc_, kSrcD, 2048, W8_, H8_), stream))
        throw std::runtime_error("unshuffle8_pad launch failed");
    ++launches;
    run(s_enc0_, stream);
    CK(cudaEventRecord(ev_y_, stream));
    CK(cudaStreamWaitEvent(copy_stream_, ev_y_, 0));
    const size_t nz = static_cast<size_t>(H64_) * W64_ * kZ;
    CK(cudaMemcpyAsync(h_totals_, totals_, 16, cudaMemcpyDeviceToHost, copy_stream_));
    CK(cudaMemcpyAsync(h_z_, z_i8_, nz, cudaMemcpyDeviceToHost, copy_stream_));

    // enc_1: synthesis + the memory / context of the NEXT chunk (dmc_htl_proxy.cpp:700-703)
    run(s_decoder_, stream);
    if (reset) {
        run(s_reset_head_, stream);
        run(s_fa_i_, stream);
    } else {
        run(s_fa_m_, stream);
    }
    run(s_fe_, stream);
    tock(stream);

    CK(cudaStreamSynchronize(copy_stream_));
    int total = 0;
    for (int k = 0; k < 4; ++k) {
        const int n = h_totals_[k];
        if (n < 0 || static_cast<size_t>(n) > quarter_) throw std::runtime_error("corrupt symbol count");
        total += n;
        if (n) CK(cudaMemcpyAsync(h_sym_[k], sym_c_[k], static_cast<size_t>(n) * 2, cudaMemcpyDeviceToHost, copy_stream_));
    }
    CK(cudaStreamSynchronize(copy_stream_));
    const int n_par = RansCodec::ec_parallel_for(total);
    std::vector<EncodeJob> jobs;
    for (int k = 3; k >= 0; --k) {
        EncodeJob j;
        j.kind = EncodeJob::Y; j.y = h_sym_[k]; j.size = h_totals_[k];
        jobs.push_back(j);
    }
    EncodeJob jz;
    jz.kind = EncodeJob::Z; jz.z = h_z_; jz.size = static_cast<int>(nz);
    jz.cdf_offset = qp * kZ; jz.ch = kZ;
    jobs.push_back(jz);
    rans_.encode(jobs, n_par, bitstream_);
    *bs = bitstream_.data();
    *len = static_cast<int32_t>(bitstream_.size());
    *ec = n_par;
}

// =============================================================================== decompress
void HtlCodec::decompress(const uint8_t* bs, int len, int qp, int height, int width, int ec, int reset,
                          cudaStream_t stream, void* const* x_hat_out)
{
    CK(cudaSetDevice(device_));
    plan(height, width);
    StreamHop hop(this, stream);
    stream = hop.run;
    stage_qp(qp, stream);
    const int zh = (height + 63) / 64, zw = (width + 63) / 64;
    if (zh != H64_ || zw != W64_) throw std::runtime_error("z geometry mismatch");
    const int nz = kZ * zh * zw;
    tev_n_ = 0;
    // dec_0 runs on the GPU while the CPU decodes z (dmc_htl_proxy.cpp:741-751)
    tick(stream);
    run(memory_has_value_ ? s_fa_m_ : s_fa_i_, stream);
    run(s_temporal_, stream);
    tock(stream);
    rans_.set_stream(bs, len, ec);
    rans_.decode_z(h_z_, nz, qp * kZ, kZ);
    CK(cudaMemcpyAsync(z_i8_, h_z_, nz, cudaMemcpyHostToDevice, stream));
    for (int k = 0; k < 4; ++k) {
        tick(stream);
        run(k == 0 ? s_dec1_ : s_dec_step_[k], stream);
        tock(stream);
        CK(cudaMemcpyAsync(h_totals_ + k, totals_ + k, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        const int n = h_totals_[k];
        if (n < 0 || static_cast<size_t>(n) > quarter_) throw std::runtime_error("corrupt index count");
        if (n) {
            CK(cudaMemcpyAsync(h_idx_, idx_c_, n, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            rans_.decode_y(h_decoded_, h_idx_, n);
            CK(cudaMemcpyAsync(decoded_, h_decoded_, n, cudaMemcpyHostToDevice, stream));
        }
    }
    tick(stream);
    run(s_dec_step_[0], stream);   // restore step 3 + final multiply
    run(s_decoder_, stream);
    run(s_recon_, stream);
    for (int i = 0; i < kG; ++i) {
        const __half* src = (i == kG - 1) ? feature_i_ : head_out_ + static_cast<size_t>(i) * H8_ * W8_ * kSrcI;
        if (launch_shuffle8_clamp(make_view(src, kSrcI, kSrcI, W8_, H8_), static_cast<__half*>(x_hat_out[i]), 3, 1, stream))
            throw std::runtime_error("shuffle8_clamp launch failed");
        ++launches;
    }
    tock(stream);
    memory_has_value_ = !reset;  // host-side state (dmc_htl_proxy.cpp:890)
}

int HtlCodec::debug_fetch(const char* name, void* dst, int64_t max_bytes, int64_t* written)
{
    struct Tap { const char* n; const void* p; size_t bytes; };
    const size_t p8 = static_cast<size_t>(H8_) * W8_, p16 = static_cast<size_t>(H16_) * W16_;
    const Tap taps[] = {
        { "y", y_, p16 * kY * 2 }, { "y_hat", yhat_, p16 * kY * 2 }, { "common", common_, p16 * kP * 2 },
        { "cat_fam", cat_fam_, p8 * 1024 * 2 }, { "cat_enc", cat_enc_, p8 * 2048 * 2 },
        { "feature_i", feature_i_, p8 * kSrcI * 2 }, { "z_i8", z_i8_, static_cast<size_t>(H64_) * W64_ * kZ },
        { "totals", totals_, 16 }, { "sym0", sym_c_[0], quarter_ * 2 }, { "sym1", sym_c_[1], quarter_ * 2 },
        { "sym2", sym_c_[2], quarter_ * 2 }, { "sym3", sym_c_[3], quarter_ * 2 },
    };
    for (const Tap& t : taps) {
        if (strcmp(t.n, name) == 0) {
            if (!t.p) throw std::runtime_error("debug_fetch: buffer not allocated yet");
            const size_t n = std::min<size_t>(t.bytes, static_cast<size_t>(max_bytes));
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(dst, t.p, n, cudaMemcpyDeviceToHost));
            *written = static_cast<int64_t>(n);
            return 0;
        }
    }
    throw std::runtime_error(std::string("debug_fetch: unknown buffer '") + name + "'");
}

// ------------------------------------------------------------------------------- glue for codec.cu
CodecBase* make_htl_codec(int device) { return new HtlCodec(device); }

int htl_add_ref(CodecBase* c, const void* frame, int H, int W, int64_t sc, int64_t sh, int64_t sw, int apply, cudaStream_t stream)
{
    static_cast<HtlCodec*>(c)->add_ref(frame, H, W, sc, sh, sw, apply, stream);
    return 0;
}

int htl_compress(CodecBase* c, const void* x, int H, int W, int64_t sc, int64_t sh, int64_t sw, int qp, int reset,
                 int pad_b, int pad_r, cudaStream_t stream, const uint8_t** bs, int32_t* len, int32_t* ec)
{
    static_cast<HtlCodec*>(c)->compress(x, H, W, sc, sh, sw, qp, reset, pad_b, pad_r, stream, bs, len, ec);
    return 0;
}

int htl_decompress(CodecBase* c, const uint8_t* bs, int len, int qp, int height, int width, int ec, int reset,
                   cudaStream_t stream, void* const* x_hat_out)
{
    static_cast<HtlCodec*>(c)->decompress(bs, len, qp, height, width, ec, reset, stream, x_hat_out);
    return 0;
}

}  // namespace dcvc
