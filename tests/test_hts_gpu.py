"""HT-S chunk codec (configs[2-4]) on the GPU through the reference-facing API: intra frame -> 8-frame
chunks with carried feature memory, exactly the call sequence of test_video.py:223-238 (encoder) and
:312-317 (decoder)."""
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
    from dcvc_b200.model import DMC, DMCI
    i_net = DMCI.synthetic(0)
    i_net.update(SKIP)
    p_net = DMC.synthetic(1)
    p_net.update(SKIP)
    return i_net.half().to("cuda"), p_net.half().to("cuda")


def _sequence(h, w, n_chunks, seed):
    frames = [synth_frame(h, w, seed)]
    for c in range(n_chunks):
        frames.append(synth_frame(h, w, seed + 1 + c, channels=24))
    return frames


def _run(i_net, p_net, h, w, n_chunks, qp_i, qp_p, reset_at, seed=300):
    """encode then decode (separate encoder / decoder proxies are the same objects here, as in the reference
    script, so the encoder-side and decoder-side state machines both run)."""
    frames = _sequence(h, w, n_chunks, seed)
    pad_r, pad_b = i_net.get_padding_size(h, w, 16)
    sps = {"height": h, "width": w}
    streams = []
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, qp_i, pad_b, pad_r)
    streams.append(("I", enc["bit_stream"], enc["ec_parallel"], 0))
    p_net.clear_dpb()
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    for c in range(n_chunks):
        x = frames[1 + c].half().cuda().contiguous(memory_format=torch.channels_last)
        reset = 1 if c in reset_at else 0
        e = p_net.compress(x, qp_p, reset, pad_b, pad_r)
        streams.append(("P", e["bit_stream"], e["ec_parallel"], reset))
    torch.cuda.synchronize()
    enc_feature = p_net.proxy.debug_fetch("cat_fam", np.float16).copy()   # encoder-side memory|feature_p
    # ---- decode
    recon = []
    for kind, bs, ec, reset in streams:
        if kind == "I":
            d = i_net.decompress(bs, sps, qp_i, ec)
            p_net.clear_dpb()
            p_net.add_ref_feature_from_frame(d["x_hat"], False)
            recon.append(d["x_hat"].clone())
        else:
            d = p_net.decompress(bs, sps, qp_p, ec, reset)
            recon.append([t.clone() for t in d["x_hat"]])
    torch.cuda.synchronize()
    dec_feature = p_net.proxy.debug_fetch("cat_fam", np.float16).copy()
    return frames, streams, recon, enc_feature, dec_feature


@pytest.mark.parametrize("h,w,n_chunks,reset_at", [(64, 64, 3, (1,)), (200, 328, 2, ()), (1080, 1920, 3, (1,)),
                                                   (2160, 3840, 1, ())])   # 4K: configs[4]; hyper path padded 135 -> 136 rows
def test_chunk_roundtrip_state_consistency(nets, h, w, n_chunks, reset_at):
    """size-independent property: after decoding the stream, the decoder holds bit-identical feature_p
    (the state the next chunk conditions on) to what the encoder derived — i.e. no encoder/decoder drift —
    and the reconstruction follows the source."""
    i_net, p_net = nets
    frames, streams, recon, enc_f, dec_f = _run(i_net, p_net, h, w, n_chunks, 30, 25, reset_at)
    C = 1024
    ef = enc_f.reshape(-1, C)[:, 512:]
    df = dec_f.reshape(-1, C)[:, 512:]
    assert np.array_equal(ef.view(np.uint16), df.view(np.uint16)), "decoder feature_p drifted from the encoder's"
    for c in range(n_chunks):
        assert len(recon[1 + c]) == 8
        for f in range(8):
            xh = recon[1 + c][f]
            assert xh.shape == (1, 3, (h + 15) // 16 * 16, (w + 15) // 16 * 16)
            assert xh.abs().max().item() <= 0.5
        src = frames[1 + c][:, 0:3]
        assert psnr(recon[1 + c][0].float().cpu()[:, :, :h, :w], src) > 6.0
    assert all(len(s[1]) > 4 for s in streams)


def test_chunk_encode_is_deterministic(nets):
    """the same chunk encoded twice from the same reference state gives the same bytes and the same carried state
    (guards against tile-overlap races between GEMM CTAs: the transform path has no atomics)."""
    i_net, p_net = nets
    h, w = 544, 960
    frames = _sequence(h, w, 1, 77)
    pad_r, pad_b = i_net.get_padding_size(h, w, 16)
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    x1 = frames[1].half().cuda().contiguous(memory_format=torch.channels_last)
    xh = i_net.compress(x0, 30, pad_b, pad_r)["x_hat"].clone()
    outs = []
    for _ in range(3):
        p_net.clear_dpb()
        p_net.add_ref_feature_from_frame(xh)
        e = p_net.compress(x1, 25, 0, pad_b, pad_r)
        torch.cuda.synchronize()
        outs.append((bytes(e["bit_stream"]), p_net.proxy.debug_fetch("cat_fam", np.float16).copy()))
    for bs, st in outs[1:]:
        assert bs == outs[0][0]
        assert np.array_equal(st.view(np.uint16), outs[0][1].view(np.uint16))


def test_hts_against_cpu_oracle(nets):
    """64x64... 128x128 sequence vs the fp16-emulating CPU restatement of the reference proxy: rate within 2 %,
    PSNR of every decoded frame within 0.1 dB (fp16 tie flips), same state machine."""
    from dcvc_b200.spec import dmci_spec, hts_spec, synth_state_dict
    from oracle.dmci_oracle import DmciOracle
    from oracle.hts_oracle import HtsOracle
    i_net, p_net = nets
    h, w, n_chunks = 128, 128, 2
    frames, streams, recon, _, _ = _run(i_net, p_net, h, w, n_chunks, 30, 25, (1,), seed=700)
    oi = DmciOracle(synth_state_dict(dmci_spec(), 0), SKIP, True, threads=8)
    oe = HtsOracle(synth_state_dict(hts_spec(), 1), SKIP, True, threads=8)
    od = HtsOracle(synth_state_dict(hts_spec(), 1), SKIP, True, threads=8)
    e0 = oi.compress(frames[0], 30, 0, 0)
    # condition both sides on the GPU's intra reconstruction so the comparison isolates the chunk codec
    x_hat0 = recon[0].float().cpu()
    oe.add_ref_feature_from_frame(x_hat0, True)
    od.add_ref_feature_from_frame(x_hat0, False)
    for c in range(n_chunks):
        reset = c == 1
        e = oe.compress(frames[1 + c], 25, reset, 0, 0)
        d = od.decompress(e["bit_stream"], 25, h, w, e["ec_parallel"], reset)
        n_gpu, n_ref = len(streams[1 + c][1]), len(e["bit_stream"])
        assert abs(n_gpu - n_ref) <= 0.02 * n_ref + 8, (c, n_gpu, n_ref)
        for f in range(8):
            src = frames[1 + c][:, 3 * f:3 * f + 3]
            p_gpu = psnr(recon[1 + c][f].float().cpu(), src)
            p_ref = psnr(d["x_hat"][f], src)
            assert abs(p_gpu - p_ref) <= 0.1, (c, f, p_gpu, p_ref)
    assert len(e0["bit_stream"]) > 0


def test_hts_stream_bit_identical_to_reference_coder(nets):
    from oracle.build_ref import import_ref_shim
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not available")
    i_net, p_net = nets
    h, w = 256, 256
    frames, streams, _, _, _ = _run(i_net, p_net, h, w, 1, 30, 40, ())
    # re-encode the last chunk's symbols with the reference coder
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, 30, 0, 0)
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    e = p_net.compress(frames[1].half().cuda().contiguous(memory_format=torch.channels_last), 40, 0, 0, 0)
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
@pytest.mark.parametrize("switch,value", [("DCVC_B200_HEAD_LANES", "1"), ("DCVC_B200_HEAD_LANES", "4"), ("DCVC_B200_FUSE_TAIL", "0")])
def test_capture_lanes_and_fusion_bit_identical(nets, switch, value, monkeypatch):
    """DCVC_B200_HEAD_LANES (default 2): the four recon-head pairs run as parallel branches of the recon graph, each on its
    own scratch level, so one persistent kernel's tail overlaps another branch's work.  DCVC_B200_FUSE_TAIL (default on):
    the fused DepthConvBlock tail instead of the per-op kernels.  Neither changes a bit: streams, every decoded frame and
    the carried state must equal the default run (the CPU tier checks under emulation that the captures fork / join and
    that no branch races another)."""
    from dcvc_b200.model import DMC
    i_net, p_net = nets
    h, w = 136, 200                                        # ragged: pads to 144 x 208
    _, streams0, recon0, _, dec0 = _run(i_net, p_net, h, w, 2, 20, 33, ())
    monkeypatch.setenv(switch, value)                      # read when the codec finalises its parameters / plans a resolution
    p2 = DMC.synthetic(1)
    p2.update(SKIP)
    p2 = p2.half().to("cuda")
    _, streams1, recon1, enc1, dec1 = _run(i_net, p2, h, w, 2, 20, 33, ())
    for a, b in zip(streams0, streams1):
        assert np.array_equal(np.asarray(a[1]), np.asarray(b[1])) and a[2] == b[2]
    for a, b in zip(recon0[1:], recon1[1:]):
        for fa, fb in zip(a, b):
            assert torch.equal(fa, fb)
    assert np.array_equal(dec0.view(np.uint16), dec1.view(np.uint16))
    # feature_p half of the state: encoder and decoder agree (the memory half is updated lazily on the decoder side)
    assert np.array_equal(enc1.reshape(-1, 1024)[:, 512:].view(np.uint16), dec1.reshape(-1, 1024)[:, 512:].view(np.uint16))


def test_p_unit_after_clear_dpb_is_refused(nets):
    """the reference forgets its reference feature in clear_dpb (video_model_ht.py:364-367) and fails on the next P unit;
    the mirror must not silently code with the previous GOP's state either"""
    i_net, p_net = nets
    h, w = 64, 64
    pad_r, pad_b = i_net.get_padding_size(h, w, 16)
    x0 = synth_frame(h, w, 1).half().cuda().contiguous(memory_format=torch.channels_last)
    x1 = synth_frame(h, w, 2, channels=24).half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, 30, pad_b, pad_r)
    p_net.clear_dpb()
    with pytest.raises(RuntimeError, match="add_ref_feature_from_frame"):
        p_net.compress(x1, 30, 0, pad_b, pad_r)
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    assert len(p_net.compress(x1, 30, 0, pad_b, pad_r)["bit_stream"]) > 8
