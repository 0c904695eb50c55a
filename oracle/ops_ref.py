"""Per-op oracles (TEST INFRASTRUCTURE ONLY — never imported by the product path).

Functional restatements of the reference's fused ops, following the `sm < 75` at:: compositions the
reference itself carries as the mathematical definition of each CUTLASS kernel
(e.g. src/layers/extensions/inference/cutlass/conv1x1_bias.cu:527-530,
cutlass/conv1x1_bias_wsilu_chunk_add.cu:364-377) and the PyTorch modules of src/layers/layers.py.
All functions take / return NCHW float32 tensors (device-agnostic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def wsilu(x):
    # src/layers/layers.py:106-111
    return torch.sigmoid(4.0 * x) * x


def chunk_add4(x):
    # src/layers/layers.py:114-125 (after the activation)
    return x[:, 0::4] + x[:, 1::4] + x[:, 2::4] + x[:, 3::4]


def conv1x1(x, w, b=None, act=False, chunk_add=False, res1=None, res2=None, q=None):
    """conv1x1_bias{,_wsilu,_wsilu_chunk_add,_shortcut,_shortcut2,_with_quant,_shortcut_with_quant}"""
    y = F.conv2d(x, w, b)
    if act:
        y = wsilu(y)
    if chunk_add:
        y = chunk_add4(y)
    if res1 is not None:
        y = y + res1
    if res2 is not None:
        y = y + res2
    if q is not None:
        y = y * q.view(1, -1, 1, 1)
    return y


def conv3x3_s2(x, w, b):
    # nn.Conv2d(C, C2, 3, stride=2, padding=1)  (src/models/image_model.py:62)
    return F.conv2d(x, w, b, stride=2, padding=1)


def conv2x2_s2(x, w, b):
    # pixel_unshuffle(2) + 1x1  (ResidualBlockWithStride2.down, src/layers/layers.py:176-188)
    return F.conv2d(F.pixel_unshuffle(x, 2), w, b)


def tconv2x2(x, w, b=None):
    # 1x1 + PixelShuffle(2)  (SubpelConv2x, src/layers/layers.py:92-103)
    return F.pixel_shuffle(F.conv2d(x, w, b), 2)


def dw3x3(x, w, b=None):
    return F.conv2d(x, w, b, padding=1, groups=x.shape[1])


def unshuffle8_pad(x, pad_b, pad_r):
    # replicate pad bottom/right then pixel_unshuffle(8)  (cat_and_pad.cu:7-51)
    xp = F.pad(x, (0, pad_r, 0, pad_b), mode="replicate")
    return F.pixel_unshuffle(xp, 8)


def shuffle8_clamp(x, clamp=True):
    y = F.pixel_shuffle(x, 8)
    return torch.clamp(y, -0.5, 0.5) if clamp else y


def round_half_away(x: torch.Tensor) -> torch.Tensor:
    # CUDA round()/roundf: ties away from zero (stream.cu:587-588)
    return torch.sign(x) * torch.floor(torch.abs(x) + 0.5)


# ----------------------------------------------------------------------------- scale index LUT
def scale_index_lut() -> np.ndarray:
    """fp16 bit pattern -> Gaussian table row, in the reference's half arithmetic
    (scale_to_index, elementwise/stream.cu:77-87; constants def_const.h:6-12)."""
    bits = np.arange(65536, dtype=np.uint16)
    s = bits.view(np.float16).astype(np.float32)
    lo = np.float32(np.float16(0.11))
    hi = np.float32(np.float16(16.0))
    s = np.where(s >= lo, s, lo)  # NaN / negatives -> minimum
    s = np.minimum(s, hi)
    log_min = np.float32(-2.2073)
    log_max = np.float32(2.7726)
    recip = np.float32(1.0) / ((log_max - log_min) / np.float32(127.0))
    l = np.log(s.astype(np.float32)).astype(np.float16)
    d = (l.astype(np.float32) - np.float32(np.float16(log_min))).astype(np.float16)
    m = (d.astype(np.float32) * np.float32(np.float16(recip))).astype(np.float16)
    idx = np.floor(m.astype(np.float32))
    return np.clip(idx, 0, 127).astype(np.uint8)


def active_group(step: int, H: int, W: int) -> np.ndarray:
    """[H, W] channel-group index that mask_<step> selects (common_model.py:174-195)."""
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    p = (hh % 2) * 2 + (ww % 2)
    return p ^ [0, 3, 2, 1][step]


def mask_4x(step: int, C: int, H: int, W: int) -> np.ndarray:
    """bool [C, H, W] — get_mask_4x restated (common_model.py:174-195, dmci_proxy.cpp:678-699)."""
    g = active_group(step, H, W)
    G = C // 4
    grp = (np.arange(C) // G)[:, None, None]
    return grp == g[None]


def entropy_enc_step_np(step, y, q_enc, scales, means, skip_thres, lut):
    """Encoder step in half arithmetic on NHWC float16 numpy arrays [H, W, C].
    Returns (y_hat [H,W,C] fp16 with zeros outside the active group, symbols int16 compacted,
    y_q int [H,W,C])  — process_with_mask + fold + build_index_enc + compaction
    (stream.cu:548-630, 931-949, 130-161, 261-282)."""
    H, W, Cc = y.shape
    G = Cc // 4
    m = np.transpose(mask_4x(step, Cc, H, W), (1, 2, 0))
    ys = y if q_enc is None else (y.astype(np.float32) * q_enc.astype(np.float32)[None, None]).astype(np.float16)
    means_hat = np.where(m, means, np.float16(0))
    s_hat = np.where(m, scales, np.float16(0))
    y_res = np.where(m, (ys.astype(np.float32) - means_hat.astype(np.float32)).astype(np.float16), np.float16(0))
    r = y_res.astype(np.float32)
    y_q = np.sign(r) * np.floor(np.abs(r) + 0.5)
    thres = np.float32(np.float16(skip_thres))
    cond = s_hat.astype(np.float32) > thres
    y_q = np.where(cond, y_q, 0.0)
    y_q = np.clip(y_q, -128, 127)
    y_hat = (y_q + means_hat.astype(np.float32)).astype(np.float16)
    # fold 4 -> 1 and compaction in NHWC order of the quarter tensor
    g = active_group(step, H, W)
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ch = g[..., None] * G + np.arange(G)[None, None]
    yq_w = np.take_along_axis(y_q, ch, axis=2).astype(np.int32)
    s_w = np.take_along_axis(scales, ch, axis=2)
    idx = lut[s_w.view(np.uint16)].astype(np.int32)
    keep = s_w.astype(np.float32) > thres
    sym = ((yq_w << 8) + idx).astype(np.int16)
    return y_hat, sym.reshape(-1)[keep.reshape(-1)], y_q.astype(np.int32)


def entropy_dec_index_np(step, scales, skip_thres, lut):
    H, W, Cc = scales.shape
    G = Cc // 4
    g = active_group(step, H, W)
    ch = g[..., None] * G + np.arange(G)[None, None]
    s_w = np.take_along_axis(scales, ch, axis=2)
    thres = np.float32(np.float16(skip_thres))
    keep = s_w.astype(np.float32) > thres
    idx = lut[s_w.view(np.uint16)]
    return idx.reshape(-1)[keep.reshape(-1)], keep


def entropy_dec_restore_np(step, scales, means, skip_thres, decoded):
    """Returns y_hat of this step [H,W,C] fp16 (zeros outside the active group)."""
    H, W, Cc = scales.shape
    G = Cc // 4
    g = active_group(step, H, W)
    ch = g[..., None] * G + np.arange(G)[None, None]
    s_w = np.take_along_axis(scales, ch, axis=2)
    m_w = np.take_along_axis(means, ch, axis=2)
    thres = np.float32(np.float16(skip_thres))
    keep = (s_w.astype(np.float32) > thres).reshape(-1)
    q = np.zeros(H * W * G, dtype=np.float32)
    q[keep] = decoded.astype(np.float32)
    yh_w = (q.reshape(H, W, G) + m_w.astype(np.float32)).astype(np.float16)
    out = np.zeros((H, W, Cc), dtype=np.float16)
    np.put_along_axis(out, ch, yh_w, axis=2)
    return out
