"""Low-delay codec (DCVC-UF LD, one frame per call) on the GPU through the reference-facing API: intra frame -> P frames
with carried feature memory, the call sequence of test_video.py:223-238 (encoder) and :312-317 (decoder)."""
import os

import numpy as np
import pytest
import torch

from util_frames import psnr, synth_frame

pytestmark = pytest.mark.gpu

# First device run of the capture-lane switches happens under tools/round2_first_call.sh (which sets this), not in the
# driver's unattended round-end run: concurrent persistent kernels are the one kind of change that could hang a box on
# a first run, and a hung box there would take the bench tier with it.  Remove the gate once they have run green.
SKIP = 0.15


@pytest.fixture(scope="module")
def nets():
    from dcvc_b200.model import DMCI, DMCLD
    i_net = DMCI.synthetic(0)
    i_net.update(SKIP)
    p_net = DMCLD.synthetic(2)
    p_net.update(SKIP)
    return i_net.half().to("cuda"), p_net.half().to("cuda")


def _run(i_net, p_net, h, w, n_frames, qp_i, qp_p, reset_at, seed=300):
    frames = [synth_frame(h, w, seed + c) for c in range(1 + n_frames)]
    pad_r, pad_b = i_net.get_padding_size(h, w, 16)
    sps = {"height": h, "width": w}
    streams = []
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, qp_i, pad_b, pad_r)
    streams.append(("I", enc["bit_stream"], enc["ec_parallel"], 0))
    p_net.clear_dpb()
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    for c in range(n_frames):
        x = frames[1 + c].half().cuda().contiguous(memory_format=torch.channels_last)
        reset = 1 if c in reset_at else 0
        e = p_net.compress(x, qp_p, reset, pad_b, pad_r)
        streams.append(("P", e["bit_stream"], e["ec_parallel"], reset))
    torch.cuda.synchronize()
    enc_state = p_net.proxy.debug_fetch("cat_fam", np.float16).copy()   # encoder-side memory | feature_p
    recon = []
    for kind, bs, ec, reset in streams:
        if kind == "I":
            d = i_net.decompress(bs, sps, qp_i, ec)
            p_net.clear_dpb()
            p_net.add_ref_feature_from_frame(d["x_hat"], False)
        else:
            d = p_net.decompress(bs, sps, qp_p, ec, reset)
        recon.append(d["x_hat"].clone())
    torch.cuda.synchronize()
    dec_state = p_net.proxy.debug_fetch("cat_fam", np.float16).copy()
    return frames, streams, recon, enc_state, dec_state


@pytest.mark.parametrize("h,w,n_frames,reset_at", [(64, 64, 4, (1,)), (200, 328, 3, ()), (1080, 1920, 3, (1,)),
                                                   (2160, 3840, 2, ())])
def test_frame_roundtrip_state_consistency(nets, h, w, n_frames, reset_at):
    """after decoding everything the decoder holds the feature_p the encoder holds (bit for bit): the two state
    machines (eager memory update in compress, lazy in decompress) agree, and every frame decodes to a sane picture"""
    i_net, p_net = nets
    frames, streams, recon, enc_state, dec_state = _run(i_net, p_net, h, w, n_frames, 30, 25, reset_at)
    C2 = enc_state.size // 2
    e = enc_state.reshape(-1, 512)[:, 256:]
    d = dec_state.reshape(-1, 512)[:, 256:]
    assert np.array_equal(e.view(np.uint16), d.view(np.uint16)), "decoder feature_p differs from the encoder's"
    for c in range(n_frames):
        x_hat = recon[1 + c].float().cpu()[:, :, :h, :w]
        assert torch.isfinite(x_hat).all() and x_hat.abs().max() <= 0.5
        assert psnr(x_hat, frames[1 + c]) > 8.0
        assert len(streams[1 + c][1]) > 0


def test_ld_against_cpu_oracle(nets):
    """sequence vs the fp16-emulating CPU restatement of the reference proxy (oracle/ld_oracle.py): rate within 2 %,
    PSNR of every decoded frame within 0.1 dB (fp16 tie flips), same state machine."""
    from dcvc_b200.spec import ld_spec, synth_state_dict
    from oracle.ld_oracle import LdOracle
    i_net, p_net = nets
    h, w, n_frames = 128, 192, 3
    frames, streams, recon, _, _ = _run(i_net, p_net, h, w, n_frames, 30, 25, (1,), seed=700)
    oe = LdOracle(synth_state_dict(ld_spec(), 2), SKIP, True, threads=8)
    od = LdOracle(synth_state_dict(ld_spec(), 2), SKIP, True, threads=8)
    x_hat0 = recon[0].float().cpu()   # condition both sides on the GPU's intra reconstruction
    oe.add_ref_feature_from_frame(x_hat0, True)
    od.add_ref_feature_from_frame(x_hat0, False)
    for c in range(n_frames):
        reset = c == 1
        e = oe.compress(frames[1 + c], 25, reset, 0, 0)
        d = od.decompress(e["bit_stream"], 25, h, w, e["ec_parallel"], reset)
        n_gpu, n_ref = len(streams[1 + c][1]), len(e["bit_stream"])
        assert abs(n_gpu - n_ref) <= 0.02 * n_ref + 8, (c, n_gpu, n_ref)
        p_gpu = psnr(recon[1 + c].float().cpu()[:, :, :h, :w], frames[1 + c])
        p_ref = psnr(d["x_hat"][:, :, :h, :w], frames[1 + c])
        assert abs(p_gpu - p_ref) <= 0.1, (c, p_gpu, p_ref)


def test_ld_stream_bit_identical_to_reference_coder(nets):
    from oracle.build_ref import import_ref_shim
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not available")
    i_net, p_net = nets
    h, w = 256, 256
    x0 = synth_frame(h, w, 41).half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, 30, 0, 0)
    p_net.clear_dpb()
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    e = p_net.compress(synth_frame(h, w, 42).half().cuda().contiguous(memory_format=torch.channels_last), 40, 0, 0, 0)
    total = int(p_net.proxy.debug_fetch("total", np.int32)[0])
    sym = p_net.proxy.debug_fetch("sym", np.int16)[:total]
    z = p_net.proxy.debug_fetch("z_i8", np.int8)
    zc, zl, yc, yl = p_net._cdf
    r = ref.RansEncoder()
    r.set_cdf(zc, zl, 0)
    r.set_cdf(yc, yl, 1)
    r.reset()
    r.set_entropy_coder_parallel(e["ec_parallel"])
    r.encode_y(np.ascontiguousarray(sym))
    r.encode_z(z, 40 * 128, 128)
    r.flush()
    assert np.asarray(r.get_encoded_stream()).tobytes() == e["bit_stream"]


@pytest.mark.timeout(300, method="thread")
def test_fused_block_tails_change_no_bit(nets, monkeypatch):
    """DCVC_B200_FUSE_TAIL=0 (per-op kernels) against the default fused DepthConvBlock tails: streams, decoded frames and
    the carried state equal bit for bit."""
    from dcvc_b200.model import DMCLD
    i_net, p_net = nets
    h, w = 136, 200
    _, streams0, recon0, _, dec0 = _run(i_net, p_net, h, w, 3, 20, 33, (1,))
    monkeypatch.setenv("DCVC_B200_FUSE_TAIL", "0")         # read when the codec finalises its parameters
    p2 = DMCLD.synthetic(2)
    p2.update(SKIP)
    p2 = p2.half().to("cuda")
    _, streams1, recon1, _, dec1 = _run(i_net, p2, h, w, 3, 20, 33, (1,))
    for a, b in zip(streams0, streams1):
        assert np.array_equal(np.asarray(a[1]), np.asarray(b[1])) and a[2] == b[2]
    for a, b in zip(recon0, recon1):
        assert torch.equal(a, b)
    assert np.array_equal(np.asarray(dec0).view(np.uint16), np.asarray(dec1).view(np.uint16))
