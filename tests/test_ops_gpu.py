"""Op-level parity: every kernel family called through the C ABI vs the per-op oracle
(oracle/ops_ref.py = the reference's own at:: compositions).  fp16 storage, fp32 accumulate:
tolerance = a few fp16 ulps of the output magnitude (written per test)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nhwc(x_nchw):  # [1,C,H,W] float -> fp16 [H,W,C] contiguous cuda
    return x_nchw[0].permute(1, 2, 0).contiguous().half().cuda()


def _nchw(x_hwc):  # fp16 [H,W,C] -> float [1,C,H,W]
    return x_hwc.float().permute(2, 0, 1).unsqueeze(0)


def _rand(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).half().float()


def _close(got, ref, rel=4e-3, abs_=2e-3):
    got = got.float().cpu()
    ref = ref.float().cpu()
    err = (got - ref).abs()
    tol = abs_ + rel * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{bad} / {err.numel()} mismatches, max err {err.max().item():.5f}"


# (H, W, Cin, N, act, chunk, nres, q)
PW_CASES = [
    (16, 16, 64, 64, False, False, 0, False),
    (16, 24, 128, 128, True, False, 0, False),
    (17, 30, 128, 512, False, False, 0, False),     # ragged M = 510 (P64 of 1080p)
    (32, 32, 192, 384, False, False, 0, False),     # adaptor 192 -> 384
    (32, 32, 384, 384, False, False, 1, False),     # + shortcut
    (32, 32, 384, 1536, True, True, 0, False),      # FFN expand + wsilu + chunk-add
    (32, 40, 384, 384, False, False, 1, True),      # shortcut + quant
    (16, 16, 128, 128, False, False, 2, False),     # shortcut2
    (16, 16, 512, 2048, True, True, 0, False),
    (68, 120, 512, 512, False, False, 1, False),    # P16 of 1080p
    (16, 16, 256, 1024, True, True, 1, False),      # chunk-add + residual (fused variant)
    (68, 120, 768, 3072, True, True, 0, False),     # HT-S prior-fusion FFN: N divisible by 192 must not pick a 48-col fold
    (34, 60, 768, 768, True, False, 0, False),
]


@pytest.mark.parametrize("H,W,Cin,N,act,chunk,nres,q", PW_CASES)
def test_pw_gemm(H, W, Cin, N, act, chunk, nres, q):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(H * 1000 + W + Cin + N)
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, N, Cin, 1, 1, scale=Cin ** -0.5)
    b = _rand(gen, N, scale=0.1)
    Co = N // 4 if chunk else N
    r1 = _rand(gen, 1, Co, H, W) if nres >= 1 else None
    r2 = _rand(gen, 1, Co, H, W) if nres >= 2 else None
    qs = (torch.rand(Co, generator=gen) + 0.5).half().float() if q else None
    ref = ops_ref.conv1x1(x, w, b, act=act, chunk_add=chunk, res1=r1, res2=r2, q=qs)

    wp = ops.pack_weight(ops.GEMM_PW, w)
    out = torch.zeros(H, W, Co, dtype=torch.float16, device="cuda")
    ops.gemm(ops.GEMM_PW, _nhwc(x), wp, N, out, bias=b.half().cuda(),
             act=ops.ACT_WSILU if act else ops.ACT_NONE, chunk_add=chunk,
             res1=_nhwc(r1) if r1 is not None else None,
             res2=_nhwc(r2) if r2 is not None else None,
             qscale=qs.half().cuda() if qs is not None else None)
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


def test_pw_gemm_pitched_slices():
    """input = channel slice of a wider cat buffer, output written into a slice (row pitch != C)"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(7)
    H, W = 20, 28
    cat = _rand(gen, 1, 512, H, W)
    w = _rand(gen, 256, 256, 1, 1, scale=1 / 16)
    b = _rand(gen, 256, scale=0.1)
    ref = ops_ref.conv1x1(cat[:, 256:], w, b)
    buf_in = _nhwc(cat)
    buf_out = torch.zeros(H, W, 512, dtype=torch.float16, device="cuda")
    ops.gemm(ops.GEMM_PW, buf_in[..., 256:], ops.pack_weight(ops.GEMM_PW, w), 256, buf_out[..., 256:],
             bias=b.half().cuda())
    torch.cuda.synchronize()
    _close(_nchw(buf_out[..., 256:]), ref)
    assert buf_out[..., :256].abs().max().item() == 0


@pytest.mark.parametrize("H,W,Cin,Cout", [(32, 32, 384, 256), (34, 60, 128, 128), (136, 240, 384, 256)])
def test_conv3x3_s2(H, W, Cin, Cout):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(11 + H)
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    b = _rand(gen, Cout, scale=0.1)
    ref = ops_ref.conv3x3_s2(x, w, b)
    out = torch.zeros(H // 2, W // 2, Cout, dtype=torch.float16, device="cuda")
    ops.gemm(ops.GEMM_CONV3X3_S2, _nhwc(x), ops.pack_weight(ops.GEMM_CONV3X3_S2, w), Cout, out, bias=b.half().cuda())
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


@pytest.mark.parametrize("H,W,Cin,Cout", [(16, 16, 128, 128), (68, 120, 128, 128), (34, 60, 128, 128)])
def test_conv2x2_s2(H, W, Cin, Cout):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(13 + H)
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, Cout, Cin * 4, 1, 1, scale=(4 * Cin) ** -0.5)
    b = _rand(gen, Cout, scale=0.1)
    ref = ops_ref.conv2x2_s2(x, w, b)
    out = torch.zeros(H // 2, W // 2, Cout, dtype=torch.float16, device="cuda")
    ops.gemm(ops.GEMM_CONV2X2_S2, _nhwc(x), ops.pack_weight(ops.GEMM_CONV2X2_S2, w), Cout, out, bias=b.half().cuda())
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


@pytest.mark.parametrize("H,W,Cin,Cout", [(16, 16, 256, 384), (17, 30, 128, 128), (68, 120, 256, 384)])
def test_tconv2x2(H, W, Cin, Cout):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(17 + H)
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, Cout * 4, Cin, 1, 1, scale=Cin ** -0.5)
    ref = ops_ref.tconv2x2(x, w)
    out = torch.zeros(H * 2, W * 2, Cout, dtype=torch.float16, device="cuda")
    ops.gemm(ops.GEMM_TCONV2X2, _nhwc(x), ops.pack_weight(ops.GEMM_TCONV2X2, w), Cout * 4, out)
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


@pytest.mark.parametrize("H,W,C", [(16, 16, 64), (17, 30, 128), (68, 120, 512), (136, 240, 384)])
def test_dw3x3(H, W, C):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(19 + C)
    x = _rand(gen, 1, C, H, W)
    w = _rand(gen, C, 1, 3, 3, scale=1 / 3)
    ref = ops_ref.dw3x3(x, w)
    w9c = w.view(C, 9).t().contiguous().half().cuda()
    out = torch.zeros(H, W, C, dtype=torch.float16, device="cuda")
    ops.dw3x3(_nhwc(x), w9c, out)
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


@pytest.mark.parametrize("H,W", [(256, 256), (1080, 1920), (100, 70)])
def test_unshuffle_shuffle_roundtrip(H, W):
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(23)
    x = (_rand(gen, 1, 3, H, W) * 0.4).half().float()
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    ref = ops_ref.unshuffle8_pad(x, Hp - H, Wp - W)
    xc = x.half().cuda().contiguous(memory_format=torch.channels_last)
    out = torch.zeros(Hp // 8, Wp // 8, 192, dtype=torch.float16, device="cuda")
    ops.unshuffle8_pad(xc, out)
    torch.cuda.synchronize()
    assert torch.equal(_nchw(out).cpu(), ref)          # pure data movement: bit exact
    back = torch.zeros(Hp, Wp, 3, dtype=torch.float16, device="cuda")
    ops.shuffle8_clamp(out, back, clamp=True)
    torch.cuda.synchronize()
    ref_back = ops_ref.shuffle8_clamp(ref, True)
    assert torch.equal(_nchw(back).cpu(), ref_back)


def test_pad_crop_scale_roundz():
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(29)
    x = _rand(gen, 1, 256, 135, 240)
    xin = _nhwc(x)
    out = torch.zeros(136, 240, 256, dtype=torch.float16, device="cuda")
    ops.pad_crop(xin, out)
    ref = torch.nn.functional.pad(x, (0, 0, 0, 1), mode="replicate")
    assert torch.equal(_nchw(out).cpu(), ref)
    crop = torch.zeros(135, 240, 256, dtype=torch.float16, device="cuda")
    ops.pad_crop(out, crop)
    assert torch.equal(_nchw(crop).cpu(), x)
    q = (torch.rand(256, generator=gen) + 0.5).half()
    sc = torch.zeros_like(xin)
    ops.scale_channels(xin, q.cuda(), sc)
    ref_sc = (x * q.float().view(1, -1, 1, 1)).half().float()
    assert torch.equal(_nchw(sc).cpu(), ref_sc)
    z = (_rand(gen, 4 * 8 * 128) * 30).half().cuda()
    zh = torch.zeros_like(z)
    zi = torch.zeros(z.numel(), dtype=torch.int8, device="cuda")
    ops.round_z(z, zh, zi)
    ref_z = torch.clamp(ops_ref.round_half_away(z.float().cpu()), -64, 63)
    assert torch.equal(zh.float().cpu(), ref_z)
    assert torch.equal(zi.cpu().float(), ref_z)
    back = torch.zeros_like(z)
    ops.int8_to_half(zi, back)
    assert torch.equal(back.float().cpu(), ref_z)


def test_scale_lut_matches_oracle():
    from dcvc_b200 import ops
    from oracle import ops_ref
    assert np.array_equal(ops.scale_index_lut(), ops_ref.scale_index_lut())


@pytest.mark.parametrize("H,W", [(16, 16), (17, 30), (68, 120)])
@pytest.mark.parametrize("skip", [0.0, 0.15])
def test_entropy_steps_bit_exact(H, W, skip):
    """integer/half work: bit-exact against the numpy restatement"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    rng = np.random.default_rng(H * 31 + W)
    C_, G = 256, 64
    lut = ops_ref.scale_index_lut()
    y = (rng.standard_normal((H, W, C_)) * 4).astype(np.float16)
    q_enc = (rng.random(C_) + 0.5).astype(np.float16)
    params = np.concatenate([np.exp(rng.standard_normal((H, W, C_)) * 1.5 - 1.0),
                             rng.standard_normal((H, W, C_))], axis=2).astype(np.float16)
    yt = torch.from_numpy(y).cuda()
    pt = torch.from_numpy(params).cuda()
    scales_t, means_t = pt[..., :C_], pt[..., C_:]
    cat = torch.full((H, W, 2 * C_), 7.0, dtype=torch.float16, device="cuda")
    acc = cat[..., :C_]
    cat_d = torch.full((H, W, 2 * C_), 7.0, dtype=torch.float16, device="cuda")
    acc_d = cat_d[..., :C_]
    b = ops.EntropyBuffers(H, W, G)
    acc_ref = np.zeros((H, W, C_), dtype=np.float16)
    for step in range(4):
        sym = ops.entropy_enc_step(b, step, yt, torch.from_numpy(q_enc).cuda(), scales_t, means_t, acc, skip)
        y_hat, sym_ref, y_q = ops_ref.entropy_enc_step_np(step, y, q_enc, params[..., :C_], params[..., C_:], skip, lut)
        acc_ref = (acc_ref.astype(np.float32) + y_hat.astype(np.float32)).astype(np.float16)
        got = sym.cpu().numpy()
        if not np.array_equal(got, sym_ref):
            n = min(len(got), len(sym_ref))
            bad = np.flatnonzero(got[:n] != sym_ref[:n])
            raise AssertionError(f"symbols differ at step {step}: len {len(got)} vs {len(sym_ref)}, {len(bad)} mismatches, "
                                 f"first {bad[:6]}, got {got[bad[:6]]}, ref {sym_ref[bad[:6]]}")
        if step == 3:
            assert np.array_equal(acc.cpu().numpy().view(np.uint16), acc_ref.view(np.uint16))
        # decoder side of the same step
        idx = ops.entropy_dec_index(b, step, scales_t, skip)
        idx_ref, keep = ops_ref.entropy_dec_index_np(step, params[..., :C_], skip, lut)
        assert np.array_equal(idx.cpu().numpy(), idx_ref)
        assert np.array_equal(idx_ref, (sym_ref & 0xff).astype(np.uint8))
        decoded = (sym_ref >> 8).astype(np.int8)
        ops.entropy_dec_restore(b, step, scales_t, means_t, acc_d, skip, torch.from_numpy(decoded).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(acc_d.cpu().numpy().view(np.uint16), acc_ref.view(np.uint16))
    assert torch.equal(cat[..., C_:], torch.full_like(cat[..., C_:], 7.0))


# ------------------------------------------------------------------------------------------ frame IO (§8 f2)
@pytest.mark.parametrize("H,W", [(16, 16), (34, 50), (256, 256), (1080, 1920)])
def test_yuv420_to_frame_bit_exact(H, W):
    from dcvc_b200 import frame_io
    from oracle import ops_ref
    rng = np.random.default_rng(H + W)
    y = rng.integers(0, 256, (H, W), dtype=np.uint8)
    u = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
    ref = ops_ref.yuv420_to_frame(y, u, v)
    got = frame_io.yuv420_to_frame(torch.from_numpy(y).cuda(), torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda())
    torch.cuda.synchronize()
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got.cpu(), ref)
    # into channels 3..5 of a stacked chunk input, NCHW-contiguous (generic-stride path)
    chunk = torch.zeros((1, 24, H, W), dtype=torch.float16, device="cuda")
    frame_io.yuv420_to_frame(torch.from_numpy(y).cuda(), torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda(), out=chunk, channel=3)
    torch.cuda.synchronize()
    assert torch.equal(chunk[:, 3:6].cpu(), ref) and chunk[:, :3].abs().max().item() == 0


@pytest.mark.parametrize("H,W,Hp,Wp,cl", [(16, 16, 16, 16, True), (34, 50, 48, 64, True), (34, 50, 48, 64, False),
                                          (1080, 1920, 1088, 1920, True)])
def test_frame_to_yuv420_bit_exact(H, W, Hp, Wp, cl):
    from dcvc_b200 import frame_io
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(H * 7 + W)
    # reconstruction-like values in and slightly beyond [-0.5, 0.5], including exact .5/255 ties
    x = (torch.rand(1, 3, Hp, Wp, generator=gen) * 1.1 - 0.55).half()
    x[0, :, 0, :8] = torch.tensor([-0.5, 0.5, 0.0, 1.5 / 255 - 0.5, 2.5 / 255 - 0.5, -0.6, 0.6, 0.25]).half()
    ry, ru, rv = ops_ref.frame_to_yuv420(x, H, W)
    xd = x.cuda()
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    y, u, v = frame_io.frame_to_yuv420(xd, H, W)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), ry) and np.array_equal(u.cpu().numpy(), ru) and np.array_equal(v.cpu().numpy(), rv)
    # PSNR numerator on the device == numpy
    src = torch.randint(0, 256, (H, W), dtype=torch.uint8, generator=gen)
    sse = frame_io.sse_u8(y, src.cuda())
    torch.cuda.synchronize()
    want = int(((ry.astype(np.int64) - src.numpy().astype(np.int64)) ** 2).sum())
    assert int(sse.item()) == want


# ---------------------------------------------------------------- DCVC-family ops named by north_star (§8 f4)
@pytest.mark.parametrize("H,W,C", [(16, 24, 64), (68, 120, 128), (135, 240, 64)])
def test_warp_bilinear(H, W, C):
    """bit-exact vs the numpy restatement of the reference's __half kernel; within fp16 rounding of the reference's
    PyTorch fallback (grid_sample, fp32)"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(H + W + C)
    im = _rand(gen, 1, C, H, W)
    flow = (torch.randn(1, 2, H, W, generator=gen) * 6).half()
    flow[0, :, 0, :4] = torch.tensor([[0.0, -100.0, 100.0, 0.5], [0.0, 100.0, -100.0, 0.25]]).half()  # exact / far OOB
    out = torch.zeros(H, W, C, dtype=torch.float16, device="cuda")
    ops.warp_bilinear(_nhwc(im), flow[0].cuda().contiguous(), out)
    torch.cuda.synchronize()
    got = out.permute(2, 0, 1).cpu().numpy()
    ref_h = ops_ref.warp_bilinear_half(im[0].half().numpy(), flow[0].numpy())
    assert np.array_equal(got.view(np.uint16), ref_h.view(np.uint16))
    ref_t = ops_ref.torch_warp(im, flow.float())
    _close(torch.from_numpy(got.astype(np.float32)).unsqueeze(0), ref_t, rel=4e-3, abs_=4e-3)


@pytest.mark.parametrize("H,W,C,inverse", [(16, 16, 64, False), (34, 60, 128, False), (34, 60, 192, True)])
def test_gdn(H, W, C, inverse):
    """GDN / IGDN = square + one pw_gemm with the rsqrt / sqrt epilogue, vs GDN.forward in fp32
    (tolerance: fp16 storage of x^2 and of the output, rsqrt.approx)"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(C + H)
    x = _rand(gen, 1, C, H, W)
    gamma = (0.1 * torch.eye(C) + 0.01 * torch.rand(C, C, generator=gen)).half().float()
    beta = (1.0 + 0.1 * torch.rand(C, generator=gen)).half().float()
    ref = ops_ref.gdn(x, gamma, beta, inverse)
    out = torch.zeros(H, W, C, dtype=torch.float16, device="cuda")
    ops.gdn(_nhwc(x), gamma, beta, out, inverse=inverse)
    torch.cuda.synchronize()
    _close(_nchw(out), ref, rel=6e-3, abs_=2e-3)


# ---------------------------------------------------------------- the GEMM kinds HT-L adds


@pytest.mark.parametrize("H,W,Cin,Cout", [(16, 16, 64, 64), (17, 30, 128, 128), (68, 120, 256, 512)])
def test_conv3x3_ps2(H, W, Cin, Cout):
    """3x3 / stride 1 / pad 1 conv + bias + pixel_shuffle(2) as a 9-tap pw_gemm with phase-major columns"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(23 + H)
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, Cout * 4, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    b = _rand(gen, Cout * 4, scale=0.1)
    ref = ops_ref.conv3x3_ps2(x, w, b)
    out = torch.zeros(H * 2, W * 2, Cout, dtype=torch.float16, device="cuda")
    b_packed = b.view(Cout, 4).t().contiguous().view(-1)    # GEMM column ph * Cout + co <- channel co * 4 + ph
    ops.gemm(ops.GEMM_CONV3X3_PS2, _nhwc(x), ops.pack_weight(ops.GEMM_CONV3X3_PS2, w), Cout * 4, out,
             bias=b_packed.half().cuda())
    torch.cuda.synchronize()
    _close(_nchw(out), ref)


def test_tconv2x2_with_bias():
    """ResidualBlockUpsample(force_bias=True) of the HT-L hyper decoder: 1x1 conv + bias + pixel_shuffle(2)"""
    from dcvc_b200 import ops
    from oracle import ops_ref
    gen = torch.Generator().manual_seed(29)
    H, W, Cin, Cout = 17, 30, 128, 256
    x = _rand(gen, 1, Cin, H, W)
    w = _rand(gen, Cout * 4, Cin, 1, 1, scale=Cin ** -0.5)
    b = _rand(gen, Cout * 4, scale=0.1)
    ref = torch.nn.functional.pixel_shuffle(torch.nn.functional.conv2d(x, w, b), 2)
    out = torch.zeros(H * 2, W * 2, Cout, dtype=torch.float16, device="cuda")
    b_packed = b.view(Cout, 4).t().contiguous().view(-1)
    ops.gemm(ops.GEMM_TCONV2X2, _nhwc(x), ops.pack_weight(ops.GEMM_TCONV2X2, w), Cout * 4, out, bias=b_packed.half().cuda())
    torch.cuda.synchronize()
    _close(_nchw(out), ref)
