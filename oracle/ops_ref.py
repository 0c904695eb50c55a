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


def conv3x3_ps2(x, w, b):
    """SubpelConv2x with a 3x3 kernel (/root/reference/src/layers/layers.py:92-103, HT-L decoder.up,
    video_model_ht.py:41): conv2d(pad 1) + bias, then pixel_shuffle(2)"""
    return F.pixel_shuffle(F.conv2d(x, w, b, padding=1), 2)


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


# ---------------------------------------------------------------------------------------------- frame IO (§8 f2)
def yuv420_to_frame(y: np.ndarray, u: np.ndarray, v: np.ndarray) -> torch.Tensor:
    """get_src_frame for yuv420 sources (/root/reference/test_video.py:66-122): ycbcr420_to_444_np nearest-neighbour
    chroma (src/utils/transforms.py:69-80: scipy.ndimage.zoom(uv, (1, 2, 2), order=0) == 2x2 replication), then on the
    device x.half(); x / 255.0; x - 0.5 (fp32 op-math, one rounding to half per op).  Returns fp16 [1,3,H,W]."""
    uv = np.stack([u, v]).astype(np.float32)
    uv = uv.repeat(2, axis=1).repeat(2, axis=2)
    yuv = np.concatenate([y[None].astype(np.float32), uv], axis=0)
    x = torch.from_numpy(yuv).unsqueeze(0).half()
    x = (x.float() / 255.0).half()
    x = (x.float() - 0.5).half()
    return x


def frame_to_yuv420(x_hat: torch.Tensor, height: int, width: int):
    """the reference's save path (/root/reference/test_video.py:352-361) on x_hat[:, :, :height, :width]:
    yuv_444_to_420(x_hat + 0.5) (src/utils/transforms.py:83-90: avg_pool2d 2x2, fp32 accumulate), * 255, clamp,
    Y .round() (half to even) .byte(), UV .byte() (truncation).  Every op rounds to half once."""
    x = x_hat[:, :, :height, :width].half()
    x = (x.float() + 0.5).half()
    y = x[:, :1]
    uv = x[:, 1:]
    uv = F.avg_pool2d(uv.float(), kernel_size=2, stride=2).half()
    y = torch.clamp((y.float() * 255).half().float(), 0, 255).round().to(torch.uint8)
    uv = torch.clamp((uv.float() * 255).half().float(), 0, 255).to(torch.uint8)
    return y[0, 0].numpy(), uv[0, 0].numpy(), uv[0, 1].numpy()


# ------------------------------------------------------------------------- DCVC-family ops named by north_star (§8 f4)
def gdn(x, gamma, beta, inverse=False):
    """GDN.forward (/root/reference/DCVC-family/DCVC/src/layers/gdn.py:52-67) with the effective (reparametrised)
    gamma [C,C] and beta [C]: norm = conv2d(x^2, gamma, beta); out = x * rsqrt(norm) (inverse: x * sqrt(norm))."""
    Cc = x.shape[1]
    norm = F.conv2d(x ** 2, gamma.reshape(Cc, Cc, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


def torch_warp(feature, flow):
    """the reference's PyTorch fallback (/root/reference/DCVC-family/DCVC-FM/src/models/block_mc.py:34-58):
    grid_sample(bilinear, padding_mode='border', align_corners=True) on fp32"""
    B, _, H, W = flow.shape
    hor = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, -1, H, -1)
    ver = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([hor, ver], 1) + torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], 1)
    return F.grid_sample(feature, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


def warp_bilinear_half(im: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """block_mc_forward_kernel<__half> (/root/reference/DCVC-family/DCVC-FM/src/models/extensions/block_mc_kernel.cu:
    25-73) restated in numpy: im fp16 [C,H,W], flow fp16 [2,H,W]; fp32 positions / weights, weights rounded to half,
    four chained half FMAs (one rounding each; evaluated exactly in float64 before the rounding)."""
    Cc, H, W = im.shape
    f32 = np.float32
    xs = np.arange(W, dtype=f32)[None, :]
    ys = np.arange(H, dtype=f32)[:, None]
    x_pos = np.clip(flow[0].astype(f32) + xs, f32(0), f32(W - 1)).astype(f32)
    y_pos = np.clip(flow[1].astype(f32) + ys, f32(0), f32(H - 1)).astype(f32)
    x0 = np.floor(x_pos).astype(np.int64)
    y0 = np.floor(y_pos).astype(np.int64)
    x1 = np.minimum(x0 + 1, W - 1)
    y1 = np.minimum(y0 + 1, H - 1)
    w_r = (x_pos - x0.astype(f32)).astype(f32)
    w_l = (f32(1) - w_r).astype(f32)
    w_b = (y_pos - y0.astype(f32)).astype(f32)
    w_t = (f32(1) - w_b).astype(f32)
    wa = (w_l * w_t).astype(f32).astype(np.float16)
    wb = (w_l * w_b).astype(f32).astype(np.float16)
    wc = (w_r * w_t).astype(f32).astype(np.float16)
    wd = (w_r * w_b).astype(f32).astype(np.float16)

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float16)

    r = np.zeros((Cc, H, W), dtype=np.float16)
    r = fma(im[:, y0, x0], wa[None], r)
    r = fma(im[:, y1, x0], wb[None], r)
    r = fma(im[:, y0, x1], wc[None], r)
    r = fma(im[:, y1, x1], wd[None], r)
    return r
