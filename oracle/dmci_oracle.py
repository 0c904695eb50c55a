"""CPU oracle for DCVC-UF-Intra (DMCI).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product path (dcvc_b200/, inference_extensions_cuda/) never does.

Two restatements, both functional over a plain state_dict (no nn.Module):

* `forward_one_frame`  — the reference's pure-PyTorch training forward,
  src/models/image_model.py:150-192 + src/models/common_model.py:123-132,231-282 (fp32, torch.round =
  ties-to-even, no clamps, no skip).  PINNED against tests/golden/dmci_forward_*.npz, which were
  produced by importing the reference modules themselves (tests/golden/make_golden.py).

* `compress` / `decompress` — the control flow of the reference's CUDA proxy,
  src/layers/extensions/inference/dmci_proxy.cpp:296-421 (compress), :423-602 (decompress),
  :804-882 (worker: symbol order, ec_parallel, z coding), with the per-element semantics of
  elementwise/stream.cu (half arithmetic, ties-away rounding, clamps, skip threshold), the weight
  folding of layers_proxy.cpp:160-206, and the reference's own rANS coder (oracle/_ref, built from
  /root/reference/src/cpp/py_rans by oracle/build_ref.py).  The reference's CUDA proxy cannot run
  here (no GPU in the authoring container, CUTLASS un-vendored), so this leg is PARITY UNPINNED
  at the NN-output level by anything but `forward_one_frame`'s goldens; the integer leg (symbols
  -> bytes) is pinned by the reference coder itself.

`emulate_fp16=True` rounds activations to fp16 at exactly the kernel boundaries of the CUDA
implementation (fp16 storage, fp32 accumulate), which is what the GPU parity tests compare with.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref
from .build_ref import import_ref_shim

QP_NUM = 64
CH_Y = 256
CH_Z = 128


def _pad_to(v, p):
    return (v + p - 1) // p * p


class DmciOracle:
    def __init__(self, state_dict, skip_thres: float = 0.0, emulate_fp16: bool = True, threads: int | None = None):
        self.emu = emulate_fp16
        self.skip_thres = float(skip_thres)
        if threads:
            torch.set_num_threads(threads)
        # the deployed model is .half(): weights are fp16 values (test_video.py:27-29)
        self.sd = {}
        for k, v in state_dict.items():
            t = v.detach().cpu()
            if t.is_floating_point():
                t = t.half().float() if emulate_fp16 else t.float()
            self.sd[k] = t
        self.lut = ops_ref.scale_index_lut()
        self._folded = {}
        self._cdf = None

    # ------------------------------------------------------------------ helpers
    def r(self, x):
        """fp16 storage rounding at a kernel boundary"""
        return x.half().float() if self.emu else x

    def _w(self, k):
        return self.sd[k]

    def _folded_bias(self, p):
        # dw-conv bias folded into the next 1x1 bias (layers_proxy.cpp:175-178)
        if p not in self._folded:
            w3 = self._w(p + "dc.3.weight")[:, :, 0, 0]
            b = w3 @ self._w(p + "dc.2.bias") + self._w(p + "dc.3.bias")
            self._folded[p] = self.r(b)
        return self._folded[p]

    def dcb(self, x, p, shortcut=False, q=None):
        """DepthConvBlock (src/layers/layers.py:152-159 == layers_proxy.cpp:71-101)"""
        if (p + "adaptor.weight") in self.sd:
            x = self.r(ops_ref.conv1x1(x, self._w(p + "adaptor.weight"), self._w(p + "adaptor.bias")))
        t = self.r(ops_ref.conv1x1(x, self._w(p + "dc.0.weight"), self._w(p + "dc.0.bias"), act=True))
        t = self.r(ops_ref.dw3x3(t, self._w(p + "dc.2.weight")))
        o = self.r(ops_ref.conv1x1(t, self._w(p + "dc.3.weight"), self._folded_bias(p), res1=x))
        t = self.r(ops_ref.conv1x1(o, self._w(p + "ffn.0.weight"), self._w(p + "ffn.0.bias"), act=True,
                                   chunk_add=True))
        out = ops_ref.conv1x1(t, self._w(p + "ffn.2.weight"), self._w(p + "ffn.2.bias"), res1=o,
                              res2=x if shortcut else None, q=q)
        return self.r(out)

    def rb_up(self, x, p):
        """ResidualBlockUpsample, shortcut=True (layers.py:162-173)"""
        t = self.r(ops_ref.tconv2x2(x, self._w(p + "up.conv.0.weight")))
        return self.dcb(t, p + "conv.", shortcut=True)

    def rb_down(self, x, p):
        """ResidualBlockWithStride2, shortcut=True (layers.py:176-188)"""
        t = self.r(ops_ref.conv2x2_s2(x, self._w(p + "down.weight"), self._w(p + "down.bias")))
        return self.dcb(t, p + "conv.", shortcut=True)

    # ------------------------------------------------------------------ sub-networks
    def encoder(self, x_unshuffled, qp):
        q = self.r(self._w("q_scale_enc")[qp])
        t = self.dcb(x_unshuffled, "enc.enc_1.", q=q)   # enc_1 then * q_enc (fused in the epilogue)
        for i in range(6):
            t = self.dcb(t, f"enc.enc_2.{i}.")
        return self.r(ops_ref.conv3x3_s2(t, self._w("enc.enc_2.6.weight"), self._w("enc.enc_2.6.bias")))

    def hyper_enc(self, y):
        t = self.dcb(y, "hyper_enc.conv.0.")
        t = self.rb_down(t, "hyper_enc.conv.1.")
        return self.rb_down(t, "hyper_enc.conv.2.")

    def hyper_dec(self, z_hat):
        t = self.rb_up(z_hat, "hyper_dec.conv.0.")
        t = self.rb_up(t, "hyper_dec.conv.1.")
        return self.dcb(t, "hyper_dec.conv.2.")

    def prior_fusion(self, x):
        t = self.dcb(x, "y_prior_fusion.conv.0.")
        t = self.dcb(t, "y_prior_fusion.conv.1.")
        t = self.dcb(t, "y_prior_fusion.conv.2.")
        return self.r(ops_ref.conv1x1(t, self._w("y_prior_fusion.conv.3.weight"), self._w("y_prior_fusion.conv.3.bias")))

    def spatial_prior(self, cat, k):
        t = self.dcb(cat, f"y_spatial_prior_adaptor_{k}.")
        for i in range(3):
            t = self.dcb(t, f"y_spatial_prior.conv.{i}.")
        return self.r(ops_ref.conv1x1(t, self._w("y_spatial_prior.conv.3.weight"), self._w("y_spatial_prior.conv.3.bias")))

    def decoder(self, y_hat, qp):
        t = self.rb_up(y_hat, "dec.dec_1.0.")
        q = self.r(self._w("q_scale_dec")[qp])
        for i in range(1, 13):
            t = self.dcb(t, f"dec.dec_1.{i}.", q=q if i == 12 else None)  # * q_dec fused in the last block
        return self.dcb(t, "dec.dec_2.")

    # ------------------------------------------------------------------ reference training forward
    @torch.inference_mode()
    def forward_one_frame(self, x, qp: int):
        """image_model.py:150-192 (recon + quantised latents; bit estimation omitted)."""
        assert not self.emu, "the training-forward restatement is an fp32 path"
        q_y_enc = self._w("q_scale_y_enc")[qp].view(1, -1, 1, 1)
        q_y_dec = self._w("q_scale_y_dec")[qp].view(1, -1, 1, 1)
        y = self.encoder(F.pixel_unshuffle(x, 8), qp)
        z = self.hyper_enc(y)
        z_hat = torch.round(z)
        params = self.prior_fusion(self.hyper_dec(z_hat))
        yH, yW = y.shape[2:]
        params = params[:, :, :yH, :yW]
        scales, means = params.chunk(2, 1)
        y = y * q_y_enc
        common = ops_ref.conv1x1(params, self._w("y_spatial_prior_reduction.weight"),
                                 self._w("y_spatial_prior_reduction.bias"))
        B, C, H, W = y.shape
        y_hat_so_far = None
        y_q_tot = torch.zeros_like(y)
        s_tot = torch.zeros_like(y)
        for k in range(4):
            mask = torch.from_numpy(ops_ref.mask_4x(k, C, H, W))[None]
            if k > 0:
                scales, means = self.spatial_prior(torch.cat((y_hat_so_far, common), 1), k).chunk(2, 1)
            means_hat = means * mask
            y_res = (y - means_hat) * mask
            y_q = torch.round(y_res)
            y_hat = y_q + means_hat
            y_hat_so_far = y_hat if k == 0 else y_hat_so_far + y_hat
            y_q_tot = y_q_tot + y_q
            s_tot = s_tot + scales * mask
        y_hat = y_hat_so_far * q_y_dec
        x_hat = F.pixel_shuffle(self.decoder(y_hat, qp), 8)
        return {"x_hat": x_hat, "y_q": y_q_tot, "z_hat": z_hat, "scales_hat": s_tot, "y_hat": y_hat}

    # ------------------------------------------------------------------ CDF tables + coder
    def cdf_tables(self):
        """(z_cdf, z_len, y_cdf, y_len) via the reference's quantiser (entropy_models.py:113-217)."""
        if self._cdf is None:
            self._cdf = build_cdf_tables_with_ref(self.sd)
        return self._cdf

    def _coder(self):
        ref = import_ref_shim()
        if ref is None:
            raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
        z_cdf, z_len, y_cdf, y_len = self.cdf_tables()
        enc, dec = ref.RansEncoder(), ref.RansDecoder()
        for c in (enc, dec):
            c.set_cdf(z_cdf, z_len, 0)
            c.set_cdf(y_cdf, y_len, 1)
        return enc, dec

    # ------------------------------------------------------------------ proxy restatement
    def _geometry(self, H, W):
        Hp, Wp = _pad_to(H, 16), _pad_to(W, 16)
        H16, W16 = Hp // 16, Wp // 16
        return Hp, Wp, H16, W16, _pad_to(H16, 4), _pad_to(W16, 4)

    def _hyper(self, z_hat, H16, W16):
        params = self.prior_fusion(self.hyper_dec(z_hat))[:, :, :H16, :W16]
        reduced = self.r(ops_ref.conv1x1(params, self._w("y_spatial_prior_reduction.weight"),
                                         self._w("y_spatial_prior_reduction.bias")))
        return params, reduced

    @staticmethod
    def _nhwc16(t):  # [1,C,H,W] float -> numpy fp16 [H,W,C]
        return t[0].permute(1, 2, 0).contiguous().half().numpy()

    @staticmethod
    def _canon(t):
        """copy into canonical NCHW strides: torch's CPU convolutions choose their summation order from
        the stride pattern (a permuted [1,C,1,1] view counts as channels_last), and the oracle must be
        bit-reproducible between its own encoder and decoder."""
        out = torch.empty(t.shape, dtype=t.dtype)
        out.copy_(t)
        return out

    @staticmethod
    def _nchw32(a):  # numpy fp16 [H,W,C] -> [1,C,H,W] float
        return DmciOracle._canon(torch.from_numpy(a.astype(np.float32)).permute(2, 0, 1).unsqueeze(0))

    @torch.inference_mode()
    def compress(self, x, qp: int, padding_b: int, padding_r: int):
        """x: float [1,3,H,W] holding fp16-representable values in [-0.5, 0.5].
        Returns dict(bit_stream bytes, x_hat [1,3,Hp,Wp], ec_parallel, symbols [4 int16 arrays], z_i8)."""
        assert self.emu, "the proxy restatement models the fp16 deployment"
        _, _, H, W = x.shape
        Hp, Wp, H16, W16, H16p, W16p = self._geometry(H, W)
        assert padding_b == Hp - H and padding_r == Wp - W
        xu = ops_ref.unshuffle8_pad(x, padding_b, padding_r)
        y = self.encoder(xu, qp)
        y_pad = F.pad(y, (0, W16p - W16, 0, H16p - H16), mode="replicate")
        z = self.hyper_enc(y_pad)
        z_hat = self._canon(torch.clamp(ops_ref.round_half_away(z), -64, 63))  # round_z (stream.cu:862-884)
        z_i8 = z_hat[0].permute(1, 2, 0).contiguous().numpy().astype(np.int8).reshape(-1)
        params, reduced = self._hyper(z_hat, H16, W16)

        y_np = self._nhwc16(y)
        q_enc = self.sd["q_scale_y_enc"][qp].half().numpy()
        red_np = self._nhwc16(reduced)
        acc = np.zeros((H16, W16, CH_Y), dtype=np.float16)
        symbols = []
        p_np = self._nhwc16(params)
        for k in range(4):
            if k > 0:
                cat = np.concatenate([acc, red_np], axis=2)
                p_np = self._nhwc16(self.spatial_prior(self._nchw32(cat), k))
            y_hat_k, sym_k, _ = ops_ref.entropy_enc_step_np(k, y_np, q_enc, p_np[..., :CH_Y], p_np[..., CH_Y:],
                                                            self.skip_thres, self.lut)
            acc = (acc.astype(np.float32) + y_hat_k.astype(np.float32)).astype(np.float16)
            symbols.append(sym_k)
        q_dec = self.sd["q_scale_y_dec"][qp].half().numpy()
        y_hat = (acc.astype(np.float32) * q_dec.astype(np.float32)[None, None]).astype(np.float16)
        x_hat = ops_ref.shuffle8_clamp(self.decoder(self._nchw32(y_hat), qp), True)

        total = sum(len(s) for s in symbols)
        ec_parallel = max(1, min(8, total // 32768))                      # dmc_common.cpp:31-35
        enc, _ = self._coder()
        enc.reset()
        enc.set_entropy_coder_parallel(ec_parallel)
        for k in (3, 2, 1, 0):                                            # dmci_proxy.cpp:839-841
            enc.encode_y(np.ascontiguousarray(symbols[k]))
        enc.encode_z(z_i8, qp * CH_Z, CH_Z)                               # dmci_proxy.cpp:843-844
        enc.flush()
        stream = bytes(np.asarray(enc.get_encoded_stream()).tobytes())
        return {"bit_stream": stream, "x_hat": x_hat, "ec_parallel": ec_parallel, "symbols": symbols,
                "z_i8": z_i8, "y_hat": y_hat}

    @torch.inference_mode()
    def decompress(self, bit_stream: bytes, qp: int, height: int, width: int, ec_parallel: int):
        assert self.emu
        Hp, Wp, H16, W16, H16p, W16p = self._geometry(height, width)
        zh, zw = (height + 63) // 64, (width + 63) // 64                   # dmci_proxy.cpp:432-433
        _, dec = self._coder()
        dec.set_entropy_coder_parallel(ec_parallel)
        dec.set_stream(np.frombuffer(bit_stream, dtype=np.uint8))
        n_z = CH_Z * zh * zw
        dec.decode_z(n_z, qp * CH_Z, CH_Z)
        z_i8 = dec.get_decoded(n_z)
        z_hat = self._canon(torch.from_numpy(z_i8.astype(np.float32)).view(zh, zw, CH_Z).permute(2, 0, 1).unsqueeze(0))
        params, reduced = self._hyper(z_hat, H16, W16)
        red_np = self._nhwc16(reduced)
        acc = np.zeros((H16, W16, CH_Y), dtype=np.float16)
        p_np = self._nhwc16(params)
        for k in range(4):
            if k > 0:
                cat = np.concatenate([acc, red_np], axis=2)
                p_np = self._nhwc16(self.spatial_prior(self._nchw32(cat), k))
            idx, _ = ops_ref.entropy_dec_index_np(k, p_np[..., :CH_Y], self.skip_thres, self.lut)
            dec.decode_y(np.ascontiguousarray(idx))
            decoded = dec.get_decoded(len(idx))
            y_hat_k = ops_ref.entropy_dec_restore_np(k, p_np[..., :CH_Y], p_np[..., CH_Y:], self.skip_thres, decoded)
            acc = (acc.astype(np.float32) + y_hat_k.astype(np.float32)).astype(np.float16)
        q_dec = self.sd["q_scale_y_dec"][qp].half().numpy()
        y_hat = (acc.astype(np.float32) * q_dec.astype(np.float32)[None, None]).astype(np.float16)
        x_hat = ops_ref.shuffle8_clamp(self.decoder(self._nchw32(y_hat), qp), True)
        return {"x_hat": x_hat, "y_hat": y_hat}


def build_cdf_tables_with_ref(sd):
    """CDF tables exactly as CompressionModel.update() builds them (entropy_models.py:113-217), but
    with the *reference's* pmf_to_quantized_cdf from oracle/_ref.  Used to pin dcvc_b200.entropy."""
    import math
    ref = import_ref_shim()
    if ref is None:
        raise RuntimeError("oracle/_ref missing")

    def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: int(pmf_length[i])], tail_mass[i]), dim=0)
            length = prob.size(0)
            prob1 = prob.clone()
            center = (length - 1) // 2
            prob1[0] = prob[center]
            for j in range(1, center + 1):
                prob1[2 * j - 1] = prob[center + j]
                prob1[2 * j] = prob[center - j]
            c = torch.IntTensor(ref.pmf_to_quantized_cdf(prob1.tolist()))
            cdf[i, : c.size(0)] = c
        return cdf

    M = 8
    # Gaussian
    table = torch.exp(torch.linspace(math.log(0.11), math.log(16.0), 128))
    zeros = torch.zeros_like(table)
    sym_range = zeros + M
    dist = torch.distributions.normal.Normal(0.0, table)
    for i in range(M, 1, -1):
        sym_range = torch.where(torch.squeeze(dist.cdf(zeros + i)) > 0.999, i, sym_range)
    sym_range = sym_range.int()
    pmf_length = 2 * sym_range + 1
    max_length = 2 * M + 1
    samples = (torch.arange(max_length) - sym_range[:, None]).float()
    dist = torch.distributions.normal.Normal(0.0, table[:, None])
    upper, lower = dist.cdf(samples + 0.5), dist.cdf(samples - 0.5)
    y_cdf = pmf_to_cdf(upper - lower, 2 * lower[:, :1], pmf_length, max_length).numpy()
    y_len = (pmf_length + 2).int().numpy()

    # factorised z
    h, b, a = sd["bit_estimator_z.h"].float(), sd["bit_estimator_z.b"].float(), sd["bit_estimator_z.a"].float()

    def fwd(x):
        for i in range(4):
            x = x * F.softplus(h[:, :, i:i + 1, None]) + b[:, :, i:i + 1, None]
            if i != 3:
                x = x + torch.tanh(x) * torch.tanh(a[:, :, i:i + 1, None])
        return torch.sigmoid(x)

    zeros = torch.zeros((h.shape[0], h.shape[1], 1, 1))
    sym_range = zeros + M
    for i in range(M, 1, -1):
        sym_range = torch.where(torch.logical_and(fwd(zeros - i) < 0.001, fwd(zeros + i) > 0.999), i, sym_range)
    sym_range = sym_range.int()
    pmf_length = (sym_range * 2 + 1).reshape(-1)
    samples = torch.arange(max_length)[None, None, None, :] - sym_range
    lower, upper = fwd(samples - 0.5), fwd(samples + 0.5)
    pmf = (upper - lower)[:, :, 0, :].reshape(-1, max_length)
    up_r = fwd(sym_range.float())
    tail = (lower[:, :, 0, :1] + (1.0 - up_r[:, :, 0, -1:])).reshape(-1, 1)
    z_cdf = pmf_to_cdf(pmf, tail, pmf_length, max_length).numpy()
    z_len = (pmf_length + 2).int().numpy()
    return z_cdf, z_len, y_cdf, y_len
