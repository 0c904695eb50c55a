"""CPU oracle for the DCVC-UF low-delay model (DMC of src/models/video_model_ld.py).  TEST INFRASTRUCTURE ONLY.

Two restatements, both functional over a plain state_dict:

* `compress` / `decompress` / `add_ref_feature_from_frame` — the control flow of the reference's CUDA proxy,
  /root/reference/src/layers/extensions/inference/dmc_ld_proxy.cpp:407-418 (add_ref_feature_from_frame), :420-486
  (compress), :488-593 (decompress), :639-658 (memory / context updates), with the half arithmetic of
  elementwise/stream.cu (means-only refinement with fixed scales over the two checkerboard masks, one symbol stream
  over the whole latent) and the reference's own rANS coder.  PARITY UNPINNED at the NN-output level (the CUDA proxy
  cannot run in the authoring container); self-consistency (decoder state == encoder state) and closeness to the
  pinned forward are tested in tests/test_cpu_parity.py.

* the reference's pure-PyTorch training forward (`forward_one_frame`,
/root/reference/src/models/video_model_ld.py:308-343, with `forward_prior_2x`,
/root/reference/src/models/common_model.py:212-229 and the 2-step checkerboard masks of
common_model.py:157-172), restated functionally over a plain state_dict and PINNED against
tests/golden/ld_forward_64x64.npz, which tests/golden/make_golden.py produced by importing the reference
modules themselves.  The product's counterpart of the CUDA proxy for this model (dmc_ld_proxy.cpp) is
dcvc_b200/csrc/codec_ld.cu; tests/test_ld_gpu.py compares it with this oracle.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref
from .dmci_oracle import _pad_to
from .hts_oracle import HtsOracle, _h

LD_CH_Y = 128
LD_CH_Z = 128


def mask_2x(step: int, C: int, H: int, W: int) -> np.ndarray:
    """get_mask_2x (common_model.py:157-172): the first half of the channels is coded on the (0,0)/(1,1)
    checkerboard phase in step 0 and on (0,1)/(1,0) in step 1, the second half the other way round."""
    yy, xx = np.mgrid[0:H, 0:W]
    m0 = ((yy + xx) % 2 == 0)   # micro mask ((1, 0), (0, 1))
    m1 = ~m0                    # micro mask ((0, 1), (1, 0))
    first, second = (m0, m1) if step == 0 else (m1, m0)
    out = np.zeros((C, H, W), dtype=np.float32)
    out[: C // 2] = first
    out[C // 2:] = second
    return out


class LdOracle(HtsOracle):
    """state: ref_feature (self.feature_p), memory, ctx — as DMC.apply_feature_adaptor / set_ref_feature"""

    # ---------------------------------------------------------------- sub-networks (video_model_ld.py)
    def feature_adaptor_i(self, f):
        return self.seq(f, "feature_adaptor_i.conv.", 4)

    def feature_adaptor_m(self, memory, feature):
        return self.seq(torch.cat((memory, feature), 1), "feature_adaptor_m.conv.", 4)

    def feature_extractor(self, memory):
        return self.seq(memory, "feature_extractor.conv.", 5)

    def ld_encoder(self, x_unshuffled, ctx, qp):
        """Encoder.internal_forward (:52-58): conv1 (2 blocks), conv2, * quant_step, 3x3/s2"""
        t = self.seq(torch.cat((x_unshuffled, ctx), 1), "encoder.conv1.", 2)
        q = self.r(self._w("q_encoder")[qp])
        t = self.dcb(t, "encoder.conv2.", q=q)
        return self.r(ops_ref.conv3x3_s2(t, self._w("encoder.down.weight"), self._w("encoder.down.bias")))

    def ld_hyper_enc(self, y):
        t = self.dcb(y, "hyper_encoder.conv.0.")
        t = self.rbd(t, "hyper_encoder.conv.1.", False)
        return self.rbd(t, "hyper_encoder.conv.2.", False)

    def ld_hyper_dec(self, z_hat):
        t = self.rbu(z_hat, "hyper_decoder.conv.0.", False)
        t = self.rbu(t, "hyper_decoder.conv.1.", False)
        return self.dcb(t, "hyper_decoder.conv.2.")

    def ld_prior_params(self, z_hat, memory, qp):
        """res_prior_param_decoder (:249-253): PriorFusion(hyper, temporal * q_feature)"""
        temporal = self.rbd(memory, "temporal_prior_encoder.conv.", False)
        hyper = self.ld_hyper_dec(z_hat)
        q = self.r(self._w("q_feature")[qp]).view(1, -1, 1, 1)
        t = self.seq(torch.cat((hyper, self.r(temporal * q)), 1), "y_prior_fusion.conv.", 3)
        return self.r(ops_ref.conv1x1(t, self._w("y_prior_fusion.conv.3.weight"), self._w("y_prior_fusion.conv.3.bias")))

    def ld_spatial_prior(self, y_hat_0, params):
        t = self.seq(torch.cat((y_hat_0, params), 1), "y_spatial_prior.conv.", 2)
        return self.r(ops_ref.conv1x1(t, self._w("y_spatial_prior.conv.2.weight"), self._w("y_spatial_prior.conv.2.bias")))

    def ld_decoder(self, y_hat, ctx, qp):
        """Decoder.internal_forward (:35-40): up, conv1 on cat(feature, ctx), conv2 (1x1), * quant_step"""
        up = self.r(ops_ref.tconv2x2(y_hat, self._w("decoder.up.conv.0.weight")))
        t = self.seq(torch.cat((up, ctx), 1), "decoder.conv1.", 3)
        q = self.r(self._w("q_decoder")[qp])
        return self.r(ops_ref.conv1x1(t, self._w("decoder.conv2.weight"), self._w("decoder.conv2.bias"), q=q))

    def ld_recon_head(self, feature, for_reset=False):
        t = self.seq(feature, "recon_head.conv.", 3)
        out = self.r(ops_ref.conv1x1(t, self._w("recon_head.head.weight"), self._w("recon_head.head.bias")))
        return out if for_reset else F.pixel_shuffle(out, 8)

    # ---------------------------------------------------------------- reference training forward (fp32)
    @torch.inference_mode()
    def forward_one_frame(self, x, qp: int, reset_feature_memory=False):
        """video_model_ld.py:308-343 (recon, latents and the carried state; bit estimation omitted)"""
        assert not self.emu, "the training-forward restatement is an fp32 path"
        if self.memory is None:                                   # apply_feature_adaptor (:222-227)
            self.memory = self.feature_adaptor_i(self.feature_p)
        else:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        self.ctx = self.feature_extractor(self.memory)
        y = self.ld_encoder(F.pixel_unshuffle(x, 8), self.ctx, qp)
        z = self.ld_hyper_enc(y)
        z_hat = torch.round(z)
        params = self.ld_prior_params(z_hat, self.memory, qp)
        # forward_prior_2x (common_model.py:212-229)
        quant_step, scales, means = params.chunk(3, 1)
        quant_step = torch.clamp_min(quant_step, 0.5)
        y = y * (1.0 / quant_step)
        B, C, H, W = y.shape
        m0 = torch.from_numpy(mask_2x(0, C, H, W))[None]
        m1 = torch.from_numpy(mask_2x(1, C, H, W))[None]

        def step(mean, mask):
            means_hat = mean * mask
            y_q = torch.round((y - means_hat) * mask)
            return y_q, y_q + means_hat

        y_q_0, y_hat_0 = step(means, m0)
        means_1 = self.ld_spatial_prior(y_hat_0, params)
        y_q_1, y_hat_1 = step(means_1, m1)
        y_hat = (y_hat_0 + y_hat_1) * quant_step
        feature = self.ld_decoder(y_hat, self.ctx, qp)
        x_hat = self.ld_recon_head(feature)
        # set_ref_feature (:268-273)
        self.feature_p = feature
        if reset_feature_memory:
            f = self.ld_recon_head(feature, for_reset=True)
            self.clear_dpb()
            self.feature_p = f
        return {"x_hat": x_hat, "y_q": y_q_0 + y_q_1, "z_hat": z_hat, "scales_hat": scales * m0 + scales * m1,
                "y_hat": y_hat, "feature": feature}

    # ---------------------------------------------------------------- proxy restatement (fp16 emulation)
    def add_ref_feature_from_frame(self, frame, apply_adaptor=True):
        """dmc_ld_proxy.cpp:407-418; frame: [1,3,Hp,Wp] (already padded reconstruction)"""
        self.feature_i = self._canon(F.pixel_unshuffle(frame, 8))
        if apply_adaptor:
            self.memory = self.feature_adaptor_i(self.feature_i)
            self.ctx = self.feature_extractor(self.memory)
        self.memory_has_value = apply_adaptor

    def _ld_params(self, z_hat, qp, H16, W16):
        temporal = self.rbd(self.memory, "temporal_prior_encoder.conv.", False)
        hyper = self.ld_hyper_dec(z_hat)[:, :, :H16, :W16]   # crop_hyper_params
        q = self.r(self._w("q_feature")[qp]).view(1, -1, 1, 1)
        t = self.seq(torch.cat((hyper, self.r(temporal * q)), 1), "y_prior_fusion.conv.", 3)
        return self.r(ops_ref.conv1x1(t, self._w("y_prior_fusion.conv.3.weight"), self._w("y_prior_fusion.conv.3.bias")))

    def _two_steps(self, common, y_np=None, decoded_dense=None):
        """process_with_mask_no_scale + process_with_mask_no_scale_add_and_multiply (encoder, :444-449) resp.
        restore_y + restore_y_and_add_multiply (decoder, :579-581): scales stay fixed, only the means are refined.
        Encoder: y_np = y / clamp_min(q_dec, .5) [H,W,C] fp16 -> (y_hat, y_q, scales); decoder: decoded_dense = y_q."""
        p_np = self._nhwc16(common)
        q_dec, scales, means0 = p_np[..., :LD_CH_Y], p_np[..., LD_CH_Y:2 * LD_CH_Y], p_np[..., 2 * LD_CH_Y:]
        H, W, C = scales.shape
        thres = np.float32(np.float16(self.skip_thres))
        acc = np.zeros((H, W, C), dtype=np.float16)
        y_q_all = np.zeros((H, W, C), dtype=np.int32)
        means = means0
        for k in range(2):
            if k == 1:
                means = self._nhwc16(self.ld_spatial_prior(self._nchw32(acc), common))
            m = np.transpose(mask_2x(k, C, H, W), (1, 2, 0)) > 0
            means_hat = np.where(m, means, np.float16(0))
            if y_np is not None:
                res = np.where(m, _h(y_np.astype(np.float32) - means_hat.astype(np.float32)), np.float16(0)).astype(np.float32)
                yq = np.sign(res) * np.floor(np.abs(res) + 0.5)
                yq = np.where(np.where(m, scales, np.float16(0)).astype(np.float32) > thres, yq, 0.0)
                yq = np.clip(yq, -128, 127)
                y_q_all += yq.astype(np.int32)
            else:
                yq = np.where(m, decoded_dense, 0).astype(np.float32)
            y_hat = np.where(m, _h(yq + means_hat.astype(np.float32)), np.float16(0))
            acc = _h(acc.astype(np.float32) + y_hat.astype(np.float32))
        qd = np.maximum(q_dec.astype(np.float32), np.float32(0.5))
        return _h(acc.astype(np.float32) * qd), y_q_all, scales

    @torch.inference_mode()
    def compress(self, x, qp: int, reset_feature_memory: bool, padding_b: int, padding_r: int):
        """x: [1,3,H,W] fp16-representable; memory / ctx were produced by add_ref_feature_from_frame or at the end of
        the previous call (dmc_ld_proxy.cpp:420-486)."""
        assert self.emu
        _, _, H, W = x.shape
        Hp, Wp = _pad_to(H, 16), _pad_to(W, 16)
        H16, W16 = Hp // 16, Wp // 16
        H16p, W16p = _pad_to(H16, 4), _pad_to(W16, 4)
        xu = ops_ref.unshuffle8_pad(x, padding_b, padding_r)
        y = self.ld_encoder(xu, self.ctx, qp)
        y_pad = F.pad(y, (0, W16p - W16, 0, H16p - H16), mode="replicate")
        z = self.ld_hyper_enc(y_pad)
        z_hat = self._canon(torch.clamp(ops_ref.round_half_away(z), -64, 63))
        z_i8 = z_hat[0].permute(1, 2, 0).contiguous().numpy().astype(np.int8).reshape(-1)
        common = self._ld_params(z_hat, qp, H16, W16)
        q_dec = self._nhwc16(common)[..., :LD_CH_Y]
        rcp = _h(np.float32(1.0) / np.maximum(q_dec.astype(np.float32), np.float32(0.5)))
        y_np = _h(self._nhwc16(y).astype(np.float32) * rcp.astype(np.float32))
        y_hat, y_q, scales = self._two_steps(common, y_np=y_np)
        keep = scales.astype(np.float32) > np.float32(np.float16(self.skip_thres))
        idx = self.lut[scales.view(np.uint16)].astype(np.int32)
        sym = ((y_q << 8) + idx).astype(np.int16).reshape(-1)[keep.reshape(-1)]
        ec_parallel = max(1, min(8, len(sym) // 32768))
        enc, _ = self._coder()
        enc.reset()
        enc.set_entropy_coder_parallel(ec_parallel)
        enc.encode_y(np.ascontiguousarray(sym))
        enc.encode_z(z_i8, qp * LD_CH_Z, LD_CH_Z)
        enc.flush()
        stream = bytes(np.asarray(enc.get_encoded_stream()).tobytes())
        # enc_1: decoder, then memory / context for the NEXT frame (:468-472, 639-649)
        self.feature_p = self.ld_decoder(self._nchw32(y_hat), self.ctx, qp)
        if reset_feature_memory:
            self.memory = self.feature_adaptor_i(self.ld_recon_head(self.feature_p, for_reset=True))
        else:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        self.ctx = self.feature_extractor(self.memory)
        return {"bit_stream": stream, "ec_parallel": ec_parallel, "symbols": sym, "z_i8": z_i8, "y_hat": y_hat}

    @torch.inference_mode()
    def decompress(self, bit_stream: bytes, qp: int, height: int, width: int, ec_parallel: int,
                   reset_feature_memory: bool):
        """dmc_ld_proxy.cpp:488-593: the memory update of the previous frame is applied lazily here (:512-516, 651-658)"""
        assert self.emu
        Hp, Wp = _pad_to(height, 16), _pad_to(width, 16)
        H16, W16 = Hp // 16, Wp // 16
        zh, zw = (height + 63) // 64, (width + 63) // 64
        if self.memory_has_value:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        else:
            self.memory = self.feature_adaptor_i(self.feature_i)
        _, dec = self._coder()
        dec.set_entropy_coder_parallel(ec_parallel)
        dec.set_stream(np.frombuffer(bit_stream, dtype=np.uint8))
        n_z = LD_CH_Z * zh * zw
        dec.decode_z(n_z, qp * LD_CH_Z, LD_CH_Z)
        z_i8 = dec.get_decoded(n_z)
        z_hat = self._canon(torch.from_numpy(z_i8.astype(np.float32)).view(zh, zw, LD_CH_Z).permute(2, 0, 1).unsqueeze(0))
        common = self._ld_params(z_hat, qp, H16, W16)
        scales = self._nhwc16(common)[..., LD_CH_Y:2 * LD_CH_Y]
        keep = scales.astype(np.float32) > np.float32(np.float16(self.skip_thres))
        idx = self.lut[scales.view(np.uint16)].reshape(-1)[keep.reshape(-1)]
        dec.decode_y(np.ascontiguousarray(idx))
        decoded = dec.get_decoded(len(idx))
        self.ctx = self.feature_extractor(self.memory)
        dense = np.zeros(scales.size, dtype=np.int32)
        dense[keep.reshape(-1)] = decoded
        y_hat, _, _ = self._two_steps(common, decoded_dense=dense.reshape(scales.shape))
        self.feature_p = self.ld_decoder(self._nchw32(y_hat), self.ctx, qp)
        head = self.ld_recon_head(self.feature_p, for_reset=True)
        self.feature_i = head   # the head output doubles as the reset reference (:585-586)
        self.memory_has_value = not reset_feature_memory
        return {"x_hat": ops_ref.shuffle8_clamp(head, True), "y_hat": y_hat}
