"""Torch-facing wrappers of the op-level C ABI (one per kernel family of SURVEY.md §2.2).

Activations are fp16 CUDA tensors in NHWC layout, shape [H, W, C]; a channel slice `t[..., a:b]`
of a wider buffer is accepted as long as its pixel pitch is uniform (this is how the reference's
"cat" buffers are consumed without copies, cutlass/conv1x1_kernel.h:82,102-107).
PyTorch here only provides device memory and the stream; all math runs in libdcvc_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_GDN, ACT_IGDN, ACT_NONE, ACT_WSILU, GEMM_CONV2X2_S2, GEMM_CONV3X3_PS2, GEMM_CONV3X3_S2, GEMM_PW, GEMM_TCONV2X2,
                   DcbTailDesc, EntropyStep, GemmDesc, View)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def view_of(t: torch.Tensor | None) -> View:
    if t is None:
        return View(None, 0, 0, 0, 0)
    assert t.is_cuda and t.dtype == torch.float16 and t.dim() == 3, "expect fp16 CUDA [H,W,C]"
    H, W, Cc = t.shape
    assert t.stride(2) == 1 and (H == 1 or t.stride(0) == W * t.stride(1)), "non-uniform pixel pitch"
    return View(t.data_ptr(), Cc, t.stride(1), W, H)


def pack_weight(kind: int, w: torch.Tensor) -> torch.Tensor:
    """Re-lay a conv weight [Cout, Cin, kh, kw] into the packed K-major GEMM operand (on the GPU)."""
    lib = _lib.load()
    wh = w.detach().to("cpu", torch.float16).contiguous()
    cout, cin, kh, kw = wh.shape
    dst = torch.empty(wh.numel(), dtype=torch.float16)
    _lib.check(lib.dcvc_pack_weight(kind, wh.data_ptr(), cout, cin, kh, kw, dst.data_ptr()), "pack_weight")
    return dst.to(w.device if w.is_cuda else "cuda")


def gemm(kind, x, w_packed, N, out, bias=None, act=ACT_NONE, chunk_add=False, res1=None, res2=None,
         qscale=None):
    lib = _lib.load()
    d = GemmDesc()
    d.kind = kind
    d.inp = view_of(x)
    d.out = view_of(out)
    d.res1 = view_of(res1)
    d.res2 = view_of(res2)
    d.weight = w_packed.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.qscale = qscale.data_ptr() if qscale is not None else None
    d.N = N
    d.act = act
    d.chunk_add = 1 if chunk_add else 0
    _lib.check(lib.dcvc_op_gemm(C.byref(d), _stream()), "op_gemm")
    return out


def dcb_tail(t2, x, y, w3, b3, wf0, bf0, wf2, bf2, t1n=None, w0n=None, b0n=None, qscale=None, shortcut=False):
    """dc.3 -> ffn.0 -> ffn.2 (-> next block's dc.0) of a DepthConvBlock in one launch (include/dcvc_b200.h:
    dcvc_op_dcb_tail).  Weights are [N, K] fp16 CUDA tensors.  Returns False when the shape is not eligible."""
    lib = _lib.load()
    d = DcbTailDesc()
    d.t2, d.x, d.y, d.t1n = view_of(t2), view_of(x), view_of(y), view_of(t1n)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    d.w3, d.b3, d.wf0, d.bf0, d.wf2, d.bf2 = ptr(w3), ptr(b3), ptr(wf0), ptr(bf0), ptr(wf2), ptr(bf2)
    d.w0n, d.b0n, d.qscale = ptr(w0n), ptr(b0n), ptr(qscale)
    d.shortcut = 1 if shortcut else 0
    rc = lib.dcvc_op_dcb_tail(C.byref(d), _stream())
    if rc == 2:
        return False
    _lib.check(rc, "op_dcb_tail")
    return True


def dw3x3(x, w9c, out):
    lib = _lib.load()
    vi, vo = view_of(x), view_of(out)
    _lib.check(lib.dcvc_op_dw3x3(C.byref(vi), C.byref(vo), w9c.data_ptr(), _stream()), "op_dw3x3")
    return out


def unshuffle8_pad(x_nchw: torch.Tensor, out: torch.Tensor):
    """x: fp16 [1, Cs, H, W] (any strides); out: [H8, W8, Cs*64]"""
    lib = _lib.load()
    _, Cs, H, W = x_nchw.shape
    vo = view_of(out)
    _lib.check(lib.dcvc_op_unshuffle8_pad(x_nchw.data_ptr(), Cs, H, W, x_nchw.stride(1), x_nchw.stride(2),
                                          x_nchw.stride(3), C.byref(vo), _stream()), "op_unshuffle8_pad")
    return out


def shuffle8_clamp(x, out_hwc, clamp=True):
    lib = _lib.load()
    vi = view_of(x)
    Cs = x.shape[2] // 64
    _lib.check(lib.dcvc_op_shuffle8_clamp(C.byref(vi), out_hwc.data_ptr(), Cs, 1 if clamp else 0, _stream()),
               "op_shuffle8_clamp")
    return out_hwc


def pad_crop(x, out):
    lib = _lib.load()
    vi, vo = view_of(x), view_of(out)
    _lib.check(lib.dcvc_op_pad_crop(C.byref(vi), C.byref(vo), _stream()), "op_pad_crop")
    return out


def scale_channels(x, q, out):
    lib = _lib.load()
    vi, vo = view_of(x), view_of(out)
    _lib.check(lib.dcvc_op_scale_channels(C.byref(vi), q.data_ptr(), C.byref(vo), _stream()), "op_scale_channels")
    return out


def round_z(z, z_hat, z_i8):
    lib = _lib.load()
    _lib.check(lib.dcvc_op_round_z(z.data_ptr(), z_hat.data_ptr(), z_i8.data_ptr(), z.numel(), _stream()), "op_round_z")
    return z_hat, z_i8


def int8_to_half(x, out):
    lib = _lib.load()
    _lib.check(lib.dcvc_op_int8_to_half(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "op_int8_to_half")
    return out


def scale_index_lut() -> np.ndarray:
    lib = _lib.load()
    lut = np.zeros(65536, dtype=np.uint8)
    _lib.check(lib.dcvc_scale_index_lut(lut.ctypes.data_as(C.c_void_p)), "scale_index_lut")
    return lut


class EntropyBuffers:
    """Scratch of one entropy-parameter step for a [H, W, 4G] latent."""

    def __init__(self, H, W, G, device="cuda"):
        n = H * W
        self.H, self.W, self.G = H, W, G
        self.sym_raw = torch.empty(n * G, dtype=torch.int16, device=device)
        self.idx_raw = torch.empty(n * G, dtype=torch.uint8, device=device)
        self.counts = torch.empty(n, dtype=torch.int32, device=device)
        self.offsets = torch.empty(n + 1, dtype=torch.int32, device=device)
        self.total = torch.zeros(1, dtype=torch.int32, device=device)
        self.compact_i16 = torch.empty(n * G, dtype=torch.int16, device=device)
        self.compact_u8 = torch.empty(n * G, dtype=torch.uint8, device=device)


def _estep(b: EntropyBuffers, step, scales, means, acc, skip_thres, y=None, q_enc=None, decoded=None,
           compact=None) -> EntropyStep:
    a = EntropyStep()
    a.H, a.W, a.G, a.step = b.H, b.W, b.G, step
    a.y = y.data_ptr() if y is not None else None
    a.y_pitch = y.stride(1) if y is not None else 0
    a.q_enc = q_enc.data_ptr() if q_enc is not None else None
    a.scales = scales.data_ptr()
    a.means = means.data_ptr() if means is not None else None
    a.p_pitch = scales.stride(1)
    a.y_hat_acc = acc.data_ptr() if acc is not None else None
    a.acc_pitch = acc.stride(1) if acc is not None else 0
    a.skip_thres = float(skip_thres)
    a.sym_raw = b.sym_raw.data_ptr()
    a.idx_raw = b.idx_raw.data_ptr()
    a.counts = b.counts.data_ptr()
    a.offsets = b.offsets.data_ptr()
    a.total = b.total.data_ptr()
    a.compact = compact.data_ptr() if compact is not None else None
    a.decoded = decoded.data_ptr() if decoded is not None else None
    return a


def entropy_enc_step(b, step, y, q_enc, scales, means, acc, skip_thres):
    """Returns the compacted int16 symbols of this step (device tensor slice)."""
    lib = _lib.load()
    a = _estep(b, step, scales, means, acc, skip_thres, y=y, q_enc=q_enc, compact=b.compact_i16)
    _lib.check(lib.dcvc_op_entropy_enc_step(C.byref(a), _stream()), "op_entropy_enc_step")
    n = int(b.total.item())
    return b.compact_i16[:n]


def entropy_dec_index(b, step, scales, skip_thres):
    lib = _lib.load()
    a = _estep(b, step, scales, None, None, skip_thres, compact=b.compact_u8)
    _lib.check(lib.dcvc_op_entropy_dec_index(C.byref(a), _stream()), "op_entropy_dec_index")
    n = int(b.total.item())
    return b.compact_u8[:n]


def entropy_dec_restore(b, step, scales, means, acc, skip_thres, decoded_i8):
    lib = _lib.load()
    a = _estep(b, step, scales, means, acc, skip_thres, decoded=decoded_i8)
    _lib.check(lib.dcvc_op_entropy_dec_restore(C.byref(a), _stream()), "op_entropy_dec_restore")
    return acc


# ------------------------------------------------------------------ DCVC-family ops named by north_star (§8 f4)
def square(x, out):
    lib = _lib.load()
    vi, vo = view_of(x), view_of(out)
    _lib.check(lib.dcvc_op_square(C.byref(vi), C.byref(vo), _stream()), "op_square")
    return out


def gdn(x, gamma, beta, out, inverse=False, scratch=None):
    """Generalised divisive normalisation (DCVC-family/DCVC/src/layers/gdn.py:52-67) on an NHWC fp16 tensor:
    out = x * rsqrt(beta + gamma . x^2) (inverse: * sqrt): the elementwise square, then ONE pw_gemm whose epilogue
    multiplies the block input (residual operand) by rsqrt / sqrt of the accumulator.
    gamma: effective [C, C] (after the non-negative reparametrisation), beta: effective [C]."""
    Cc = x.shape[2]
    sq = scratch if scratch is not None else torch.empty_like(x)
    square(x, sq)
    wp = pack_weight(GEMM_PW, gamma.reshape(Cc, Cc, 1, 1))
    return gemm(GEMM_PW, sq, wp, Cc, out, bias=beta.half().to(x.device), act=ACT_IGDN if inverse else ACT_GDN, res1=x)


def warp_bilinear(im, flow, out):
    """bilinear backward warp with border clamp (block_mc_kernel.cu:25-73): im/out fp16 [H,W,C], flow fp16 [2,H,W]"""
    assert flow.dtype == torch.float16 and flow.is_cuda and flow.dim() == 3 and flow.shape[0] == 2
    lib = _lib.load()
    vi, vo = view_of(im), view_of(out)
    _lib.check(lib.dcvc_op_warp_bilinear(C.byref(vi), flow.data_ptr(), flow.stride(0), flow.stride(1), flow.stride(2),
                                         C.byref(vo), _stream()), "op_warp_bilinear")
    return out
