"""Sequence driver on the device (callers either side of the path): 8-bit YUV 4:2:0 planes -> container
-> planes through dcvc_b200.sequence with the device frame IO, for the HT-S (8 pictures per chunk) and LD (1 picture per
call) models.  The driver must reproduce, byte for byte and plane for plane, what the same pictures give when the
reference-facing model API is driven by hand in the order of test_video.py:204-372."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SKIP = 0.15
H, W = 72, 104            # even, not a multiple of 16: the padding path of compress / the crop of frame_to_yuv420


def _planes(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    base = rng.random((H + 8, W + 8)).astype(np.float32)
    for t in range(n):
        img = np.roll(base, (t, t // 2), axis=(0, 1))[4:4 + H, 4:4 + W]
        y = torch.from_numpy(np.clip(img * 200 + 20, 0, 255).astype(np.uint8)).cuda()
        u = torch.from_numpy(np.clip(img[::2, ::2] * 90 + 80, 0, 255).astype(np.uint8)).cuda().contiguous()
        v = torch.from_numpy(np.clip(img[1::2, 1::2] * 70 + 100, 0, 255).astype(np.uint8)).cuda().contiguous()
        out.append((y.contiguous(), u, v))
    return out


@pytest.fixture(scope="module")
def nets():
    from dcvc_b200.model import DMC, DMCI, DMCLD
    nets = {}
    i_net = DMCI.synthetic(0)
    i_net.update(SKIP)
    nets["i"] = i_net.half().to("cuda")
    for name, cls, seed in (("hts", DMC, 1), ("ld", DMCLD, 2)):
        m = cls.synthetic(seed)
        m.update(SKIP)
        nets[name] = m.half().to("cuda")
    return nets


@pytest.mark.parametrize("which,delay,n_frames,reset_interval", [("hts", 8, 19, 16), ("ld", 1, 6, 4)])
def test_sequence_driver_matches_the_hand_driven_api(nets, which, delay, n_frames, reset_interval):
    from dcvc_b200 import stream
    from dcvc_b200.frame_io import frame_to_yuv420, yuv420_to_frame
    from dcvc_b200.sequence import SequenceDecoder, SequenceEncoder, frame_schedule
    i_net, p_net = nets["i"], nets[which]
    frames = _planes(n_frames, 5)
    enc = SequenceEncoder(i_net, p_net, H, W, qp_i=30, qp_p=25, frame_delay=delay, reset_interval=reset_interval)
    data = enc.encode(frames)
    dec = [tuple(p.clone() for p in planes) for planes in SequenceDecoder(i_net, p_net, frame_delay=delay).decode(data, n_frames)]
    torch.cuda.synchronize()
    assert len(dec) == n_frames and sum(enc.bits) == 8 * len(data)
    for (y, u, v) in dec:
        assert y.shape == (H, W) and u.shape == (H // 2, W // 2) and v.shape == u.shape and y.dtype == torch.uint8

    # ---- the same pictures through the model API by hand
    pad_r, pad_b = i_net.get_padding_size(H, W, 16)
    sps = {"sps_id": 0, "height": H, "width": W}
    out = io.BytesIO()
    stream.write_sps(out, sps)
    for unit in frame_schedule(n_frames, delay, -1, reset_interval):
        group = frames[unit.first:unit.first + unit.count]
        group = group + [group[-1]] * ((1 if unit.is_intra else delay) - len(group))
        x = torch.cat([yuv420_to_frame(*g) for g in group], dim=1).contiguous(memory_format=torch.channels_last)
        if unit.is_intra:
            e = i_net.compress(x, 30, pad_b, pad_r)
            p_net.clear_dpb()
            p_net.add_ref_feature_from_frame(e["x_hat"])
            stream.write_ip(out, True, 0, 30, e["ec_parallel"], 0, e["bit_stream"])
        else:
            e = p_net.compress(x, 25, unit.reset_feature_memory, pad_b, pad_r)
            stream.write_ip(out, False, 0, 25, e["ec_parallel"], unit.reset_feature_memory, e["bit_stream"])
    assert out.getvalue() == data, "the driver's stream differs from the hand-driven one"

    f = io.BytesIO(data)
    hand = []
    h = stream.read_header(f)
    assert h["nal_type"] == stream.NalType.NAL_SPS
    sps = stream.read_sps_remaining(f, h["sps_id"])
    while len(hand) < n_frames:
        h = stream.read_header(f)
        qp, ec, reset, bs = stream.read_ip_remaining(f)
        if h["nal_type"] == stream.NalType.NAL_I:
            d = i_net.decompress(bs, sps, qp, ec)
            p_net.clear_dpb()
            p_net.add_ref_feature_from_frame(d["x_hat"], False)
            hats = [d["x_hat"]]
        else:
            d = p_net.decompress(bs, sps, qp, ec, reset)
            hats = d["x_hat"] if isinstance(d["x_hat"], (list, tuple)) else [d["x_hat"]]
            hats = hats[: n_frames - len(hand)]
        for xh in hats:
            hand.append(tuple(p.clone() for p in frame_to_yuv420(xh, H, W)))
    torch.cuda.synchronize()
    for a, b in zip(dec, hand):
        assert all(torch.equal(p, q) for p, q in zip(a, b))
