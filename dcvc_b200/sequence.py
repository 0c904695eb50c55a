"""Sequence-level driver: the loop the reference's `test_video.py` runs around the codec proxies — which pictures are intra,
how P pictures are grouped into chunks, when the feature memory is reset, and how every coded unit is framed in the
bitstream container (reference: test_video.py:204-266 encoder loop, :281-372 decoder loop; SURVEY.md §8 f2 "callers
either side of the path").  Frames enter and leave as 8-bit YUV 4:2:0 planes; the 4:2:0 <-> model-tensor conversion is
the device frame IO of dcvc_b200.frame_io unless other converters are injected (the CPU tests inject the oracle's).

    enc = SequenceEncoder(i_net, p_net, height, width, qp_i, qp_p, frame_delay=8)
    data = enc.encode(frames)                       # frames: iterable of (y, u, v) uint8 planes
    for y, u, v in SequenceDecoder(i_net, p_net, frame_delay=8).decode(data, len(frames)): ...

`i_net` / `p_net` are the model mirrors of dcvc_b200.model (or the reference's own models: same API).
"""
from __future__ import annotations

import io
from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

from .stream import NalType, SPSHelper, read_header, read_ip_remaining, read_sps_remaining, write_ip, write_sps


@dataclass(frozen=True)
class Unit:
    """one coded unit of the sequence: an intra picture or a chunk of `count` P pictures starting at `first`"""
    first: int
    count: int
    is_intra: bool
    reset_feature_memory: int


def frame_schedule(frame_num: int, frame_delay: int, intra_period: int = -1, reset_interval: int = 32,
                   force_intra: bool = False) -> List[Unit]:
    """The reference's picture-type decisions (test_video.py:204-236), as a list.

    Picture 0 is intra; with intra_period > 1 every picture whose index is 1 modulo the period (except picture 1) is
    intra too (the period must be a multiple of the frame delay); intra_period == 1 or force_intra codes everything
    intra.  P pictures are coded `frame_delay` at a time (the tail chunk is shorter; the caller pads it by repeating
    the last picture, test_video.py:101-107).  A chunk starting at picture f resets the feature memory when
    (f + frame_delay) % reset_interval == 1."""
    if frame_delay < 1 or frame_num < 0:
        raise ValueError("bad frame_delay / frame_num")
    if intra_period > 1 and intra_period % frame_delay != 0:
        raise ValueError("intra_period must be a multiple of the frame delay")
    units, f = [], 0
    while f < frame_num:
        intra = force_intra or f == 0 or intra_period == 1
        if intra_period > 1 and f != 1 and f % intra_period == 1:
            intra = True
        if intra:
            units.append(Unit(f, 1, True, 0))
            f += 1
            continue
        count = min(frame_delay, frame_num - f)
        reset = 1 if (reset_interval > 0 and (f + frame_delay) % reset_interval == 1) else 0
        units.append(Unit(f, count, False, reset))
        f += count
    return units


Planes = Tuple["object", "object", "object"]      # (y [H,W], u [H/2,W/2], v [H/2,W/2]) uint8 tensors


class UnitTimer:
    """The reference driver's timing of one coded unit (test_video.py:219-221 + 263-267, 284-285 + 321-325): device
    synchronise, start event, the unit's work (model call + container framing on the host), end event, synchronise."""

    def __init__(self, device=None):
        import torch
        self.torch, self.device = torch, device
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def start(self):
        self.torch.cuda.synchronize(self.device)
        self.e0.record()

    def stop(self) -> float:
        self.e1.record()
        self.torch.cuda.synchronize(self.device)
        return self.e0.elapsed_time(self.e1)


def _device_to_model(planes: Sequence[Planes], frame_delay: int):
    """8-bit planes of 1 .. frame_delay pictures -> fp16 [1, 3*n, H, W] channels_last on the device of the planes"""
    import torch

    from .frame_io import yuv420_to_frame
    y0 = planes[0][0]
    H, W = y0.shape
    x = torch.empty((1, 3 * len(planes), H, W), dtype=torch.float16, device=y0.device).contiguous(memory_format=torch.channels_last)
    for i, (y, u, v) in enumerate(planes):
        yuv420_to_frame(y, u, v, out=x, channel=3 * i)
    return x


def _device_from_model(x_hat, height: int, width: int) -> Planes:
    from .frame_io import frame_to_yuv420
    return frame_to_yuv420(x_hat, height, width)


class SequenceEncoder:
    def __init__(self, i_net, p_net, height: int, width: int, qp_i: int, qp_p: int, frame_delay: int = 8,
                 intra_period: int = -1, reset_interval: int = 32, force_intra: bool = False,
                 to_model: Optional[Callable] = None):
        self.i_net, self.p_net = i_net, p_net
        self.height, self.width = height, width
        self.qp_i, self.qp_p = qp_i, qp_p
        self.frame_delay, self.intra_period, self.reset_interval = frame_delay, intra_period, reset_interval
        self.force_intra = force_intra
        self.to_model = to_model or _device_to_model
        self.padding_r, self.padding_b = i_net.get_padding_size(height, width, 16)     # test_video.py:189
        self.bits: List[int] = []          # per picture, like the reference's `bits` list (chunk bits on its first picture)
        self.timer = None                  # UnitTimer: the reference's per-unit timing protocol (test_video.py:219-267)
        self.unit_ms: List[float] = []

    def encode(self, frames: Iterable[Planes]) -> bytes:
        frames = list(frames)
        out = io.BytesIO()
        sps_helper = SPSHelper()
        self.bits = []
        for unit in frame_schedule(len(frames), self.frame_delay, self.intra_period, self.reset_interval, self.force_intra):
            group = list(frames[unit.first:unit.first + unit.count])
            if not unit.is_intra:
                group += [group[-1]] * (self.frame_delay - len(group))      # tail chunk: repeat the last picture
            x = self.to_model(group, self.frame_delay)
            if self.timer:
                self.timer.start()
            if unit.is_intra:
                qp = self.qp_i
                encoded = self.i_net.compress(x, qp, self.padding_b, self.padding_r)
                if not self.force_intra:
                    self.p_net.clear_dpb()
                    self.p_net.add_ref_feature_from_frame(encoded["x_hat"])
            else:
                qp = self.qp_p
                encoded = self.p_net.compress(x, qp, unit.reset_feature_memory, self.padding_b, self.padding_r)
            sps = {"sps_id": -1, "height": self.height, "width": self.width}
            sps_id, is_new = sps_helper.get_sps_id(sps)
            sps["sps_id"] = sps_id
            n = write_sps(out, sps) if is_new else 0
            n += write_ip(out, unit.is_intra, sps_id, qp, encoded["ec_parallel"], unit.reset_feature_memory,
                          encoded["bit_stream"])
            self.bits += [n * 8] + [0] * (unit.count - 1)
            if self.timer:
                self.unit_ms.append(self.timer.stop())
        return out.getvalue()


class SequenceDecoder:
    def __init__(self, i_net, p_net, frame_delay: int = 8, force_intra: bool = False,
                 from_model: Optional[Callable] = None):
        self.i_net, self.p_net = i_net, p_net
        self.frame_delay, self.force_intra = frame_delay, force_intra
        self.from_model = from_model or _device_from_model
        self.timer = None
        self.unit_ms: List[float] = []

    def decode(self, data: bytes, frame_num: int):
        """yields (y, u, v) uint8 planes of `frame_num` pictures in display order"""
        f = io.BytesIO(data)
        sps_helper = SPSHelper()
        done = 0
        while done < frame_num:
            if self.timer:
                self.timer.start()
            header = read_header(f)
            while header["nal_type"] == NalType.NAL_SPS:
                sps_helper.add_sps_by_id(read_sps_remaining(f, header["sps_id"]))
                header = read_header(f)
            sps = sps_helper.get_sps_by_id(header["sps_id"])
            if sps is None:
                raise ValueError("coded unit refers to an unknown sequence parameter set")
            qp, ec_part, reset, bit_stream = read_ip_remaining(f)
            if header["nal_type"] == NalType.NAL_I:
                decoded = self.i_net.decompress(bit_stream, sps, qp, ec_part)
                if not self.force_intra:
                    self.p_net.clear_dpb()
                    self.p_net.add_ref_feature_from_frame(decoded["x_hat"], False)
                count = 1
            elif header["nal_type"] == NalType.NAL_P:
                decoded = self.p_net.decompress(bit_stream, sps, qp, ec_part, reset)
                count = min(self.frame_delay, frame_num - done)
            else:
                raise ValueError("unexpected unit type in a DCVC-UF stream")
            x_hat = decoded["x_hat"]
            if self.timer:
                self.unit_ms.append(self.timer.stop())
            for i in range(count):
                yield self.from_model(x_hat[i] if isinstance(x_hat, (list, tuple)) else x_hat, sps["height"], sps["width"])
            done += count
