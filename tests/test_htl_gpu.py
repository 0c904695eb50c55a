"""HT-L chunk codec (the large high-throughput model) on the GPU through the reference-facing API.

Same call sequence as the HT-S tests (test_video.py:223-238 encoder, :312-317 decoder).  First device run: round 2
(tools/r2_call1.sh), green."""

import numpy as np
import pytest
import torch

from util_frames import psnr, synth_frame

pytestmark = pytest.mark.gpu
SKIP = 0.15


@pytest.fixture(scope="module")
def nets():
    from dcvc_b200.model import DMCHTL, DMCI
    i_net = DMCI.synthetic(0)
    i_net.update(SKIP)
    p_net = DMCHTL.synthetic(3)
    p_net.update(SKIP)
    return i_net.half().to("cuda"), p_net.half().to("cuda")


def _sequence(h, w, n_chunks, seed):
    frames = [synth_frame(h, w, seed)]
    for c in range(n_chunks):
        frames.append(synth_frame(h, w, seed + 1 + c, channels=24))
    return frames


def _run(i_net, p_net, h, w, n_chunks, qp_i, qp_p, reset_at, seed=300):
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
    enc_feature = p_net.proxy.debug_fetch("cat_fam", np.float16).copy()
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


@pytest.mark.parametrize("h,w,n_chunks,reset_at", [(64, 64, 3, (1,)), (200, 328, 2, ()), (1080, 1920, 2, (1,))])
def test_chunk_roundtrip_state_consistency(nets, h, w, n_chunks, reset_at):
    """after decoding the stream the decoder holds the feature_p the encoder derived, bit for bit"""
    i_net, p_net = nets
    frames, streams, recon, enc_f, dec_f = _run(i_net, p_net, h, w, n_chunks, 30, 25, reset_at)
    ef = enc_f.reshape(-1, 1024)[:, 512:]
    df = dec_f.reshape(-1, 1024)[:, 512:]
    assert np.array_equal(ef.view(np.uint16), df.view(np.uint16)), "decoder feature_p drifted from the encoder's"
    for c in range(n_chunks):
        assert len(recon[1 + c]) == 8
        for f in range(8):
            xh = recon[1 + c][f]
            assert xh.shape == (1, 3, (h + 15) // 16 * 16, (w + 15) // 16 * 16)
            assert torch.isfinite(xh).all() and xh.abs().max().item() <= 0.5
    assert all(len(s[1]) > 4 for s in streams)


def test_htl_against_cpu_oracle(nets):
    """sequence vs the fp16-emulating CPU restatement of the reference proxy (oracle/htl_oracle.py): rate within 2 %,
    PSNR of every decoded frame within 0.1 dB (fp16 tie flips), same state machine."""
    from dcvc_b200.spec import htl_spec, synth_state_dict
    from oracle.htl_oracle import HtlOracle
    i_net, p_net = nets
    h, w, n_chunks = 128, 128, 2
    frames, streams, recon, _, _ = _run(i_net, p_net, h, w, n_chunks, 30, 25, (1,), seed=700)
    oe = HtlOracle(synth_state_dict(htl_spec(), 3), SKIP, True, threads=8)
    od = HtlOracle(synth_state_dict(htl_spec(), 3), SKIP, True, threads=8)
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


def test_htl_stream_bit_identical_to_reference_coder(nets):
    """the bytes are what the reference's own rANS coder makes of the same four symbol runs (step 3 first) and z"""
    from oracle.build_ref import import_ref_shim
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not available")
    i_net, p_net = nets
    h, w = 256, 256
    frames = _sequence(h, w, 1, 300)
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, 30, 0, 0)
    p_net.clear_dpb()
    p_net.add_ref_feature_from_frame(enc["x_hat"])
    e = p_net.compress(frames[1].half().cuda().contiguous(memory_format=torch.channels_last), 40, 0, 0, 0)
    totals = p_net.proxy.debug_fetch("totals", np.int32)[:4]
    z = p_net.proxy.debug_fetch("z_i8", np.int8)
    zc, zl, yc, yl = p_net._cdf
    r = ref.RansEncoder()
    r.set_cdf(zc, zl, 0)
    r.set_cdf(yc, yl, 1)
    r.reset()
    r.set_entropy_coder_parallel(e["ec_parallel"])
    for k in (3, 2, 1, 0):
        r.encode_y(np.ascontiguousarray(p_net.proxy.debug_fetch(f"sym{k}", np.int16)[:int(totals[k])]))
    r.encode_z(z, 40 * 128, 128)
    r.flush()
    assert np.asarray(r.get_encoded_stream()).tobytes() == e["bit_stream"]
