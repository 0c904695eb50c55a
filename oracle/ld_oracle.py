"""CPU oracle for the DCVC-UF low-delay model (DMC of src/models/video_model_ld.py).  TEST INFRASTRUCTURE ONLY.

Round-1 scope: the reference's pure-PyTorch training forward (`forward_one_frame`,
/root/reference/src/models/video_model_ld.py:308-343, with `forward_prior_2x`,
/root/reference/src/models/common_model.py:212-229 and the 2-step checkerboard masks of
common_model.py:157-172), restated functionally over a plain state_dict and PINNED against
tests/golden/ld_forward_64x64.npz, which tests/golden/make_golden.py produced by importing the reference
modules themselves.  The CUDA proxy for this model (dmc_ld_proxy.cpp) is not built yet;
this oracle and dcvc_b200.spec.ld_spec are the groundwork its parity tests will stand on.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref
from .hts_oracle import HtsOracle


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
