"""The DCVC-UF bitstream container: the byte framing the reference's driver puts around every coded picture / chunk
(reference: src/utils/stream_helper.py:37-192, used by test_video.py:202-266 on the encoder side and :281-310 on the
decoder side).  Same names and call shapes as the reference's helpers, own implementation; byte-for-byte parity is pinned
by tests/golden/stream_container.json (minted by running the reference's own functions, tests/golden/make_golden.py).

Layout (all big-endian):
  SPS unit   : [type=0 : 4 | sps_id : 4]  height(varint)  width(varint)
  I / P unit : [type=1|2 : 4 | sps_id : 4]  qp(8)  [ec_parallel : 7 | reset_feature_memory : 1]  length(varint)  payload
  varint     : 0xxxxxxx (7 bit) | 10xxxxxx + 1 byte (14 bit) | 11xxxxxx + 3 bytes (30 bit)
A third header form (type >= 3: frame count + packed sps ids, stream_helper.py:46-57) is parsed for completeness; the
reference never writes it.
"""
from __future__ import annotations

import enum


class NalType(enum.IntEnum):
    NAL_SPS = 0
    NAL_I = 1
    NAL_P = 2


# ------------------------------------------------------------------------------------------------ primitives
def write_uint_adaptive(f, a: int) -> int:
    """varint (stream_helper.py:106-130); returns the number of bytes written"""
    if a < 0 or a >= (1 << 30):
        raise ValueError("value does not fit the 30-bit adaptive integer")
    if a < (1 << 7):
        f.write(bytes((a,)))
        return 1
    if a < (1 << 14):
        f.write(bytes((0x80 | (a >> 8), a & 0xff)))
        return 2
    f.write(bytes((0xc0 | (a >> 24), (a >> 16) & 0xff, (a >> 8) & 0xff, a & 0xff)))
    return 4


def _take(f, n: int) -> bytes:
    b = f.read(n)
    if len(b) != n:
        raise EOFError("bitstream container ends inside a unit")
    return b


def read_uint_adaptive(f) -> int:
    """stream_helper.py:60-74"""
    first = _take(f, 1)[0]
    if first < 0x80:
        return first
    second = _take(f, 1)[0]
    if (first >> 6) == 0x02:
        return ((first & 0x3f) << 8) | second
    rest = _take(f, 2)
    return ((first & 0x3f) << 24) | (second << 16) | (rest[0] << 8) | rest[1]


# ------------------------------------------------------------------------------------------------ units
def write_sps(f, sps: dict) -> int:
    """stream_helper.py:147-157"""
    if not 0 <= sps["sps_id"] < 16:
        raise ValueError("sps_id must fit 4 bits")
    f.write(bytes(((int(NalType.NAL_SPS) << 4) | sps["sps_id"],)))
    return 1 + write_uint_adaptive(f, sps["height"]) + write_uint_adaptive(f, sps["width"])


def write_ip(f, is_i_frame: bool, sps_id: int, qp: int, ec_part: int, reset_feature_memory: int, bit_stream: bytes) -> int:
    """stream_helper.py:133-144"""
    if not (0 <= sps_id < 16 and 0 <= qp < 256 and 0 <= ec_part < 128 and reset_feature_memory in (0, 1)):
        raise ValueError("header field out of range")
    kind = NalType.NAL_I if is_i_frame else NalType.NAL_P
    f.write(bytes(((int(kind) << 4) | sps_id, qp, (ec_part << 1) | reset_feature_memory)))
    n = 3 + write_uint_adaptive(f, len(bit_stream))
    if len(bit_stream):
        f.write(bytes(bit_stream))
    return n + len(bit_stream)


def read_header(f) -> dict:
    """stream_helper.py:37-57"""
    flag = _take(f, 1)[0]
    kind = flag >> 4
    header = {"nal_type": NalType(kind)}
    if kind < 3:
        header["sps_id"] = flag & 0x0f
        return header
    frame_num = (flag & 0x0f) + 1
    ids = []
    for _ in range(0, frame_num, 2):
        b = _take(f, 1)[0]
        ids += [b >> 4, b & 0x0f]
    header["frame_num"] = frame_num
    header["sps_ids"] = ids[:frame_num]
    return header


def read_sps_remaining(f, sps_id: int) -> dict:
    """stream_helper.py:88-93"""
    height = read_uint_adaptive(f)
    width = read_uint_adaptive(f)
    return {"sps_id": sps_id, "height": height, "width": width}


def read_ip_remaining(f):
    """-> (qp, ec_part, reset_feature_memory, bit_stream)   (stream_helper.py:77-85)"""
    qp = _take(f, 1)[0]
    flag = _take(f, 1)[0]
    n = read_uint_adaptive(f)
    return qp, (flag >> 1) & 0x7f, flag & 0x01, _take(f, n)


class SPSHelper:
    """sequence-parameter-set bookkeeping of the driver (stream_helper.py:163-192): ids are handed out in order of
    first use of a (height, width) pair, at most 16 of them."""

    def __init__(self):
        self.spss = []

    def add_sps_by_id(self, sps: dict) -> None:
        for i, s in enumerate(self.spss):
            if s["sps_id"] == sps["sps_id"]:
                self.spss[i] = dict(sps)
                return
        self.spss.append(dict(sps))

    def get_sps_by_id(self, sps_id: int):
        for s in self.spss:
            if s["sps_id"] == sps_id:
                return s
        return None

    def get_sps_id(self, target_sps: dict):
        """-> (sps_id, is_new)"""
        highest = -1
        for s in self.spss:
            if s["height"] == target_sps["height"] and s["width"] == target_sps["width"]:
                return s["sps_id"], False
            highest = max(highest, s["sps_id"])
        if highest >= 15:
            raise ValueError("more than 16 picture sizes in one stream")
        new = dict(target_sps)
        new["sps_id"] = highest + 1
        self.spss.append(new)
        return new["sps_id"], True
