"""CPU oracle for the DCVC-UF HT-L chunk model (DMC(ModelStructure.HTL) of src/models/video_model_ht.py).
TEST INFRASTRUCTURE ONLY.

Two restatements:

* `compress` / `decompress` / `add_ref_feature_from_frame` — the control flow of the reference's CUDA proxy,
  /root/reference/src/layers/extensions/inference/dmc_htl_proxy.cpp:583-594, 596-717 (compress: four
  process_with_mask steps with scale updates, one symbol run per step, y / clamp_min(q_dec, .5) first, final
  add_and_multiply_with_clamp_min), :719-915 (decompress: z, then four index / decode / restore round trips),
  with the half arithmetic of elementwise/stream.cu (oracle/ops_ref.py) and the reference's own rANS coder.
  PARITY UNPINNED at the NN-output level; encoder/decoder state identity is tested in tests/test_cpu_parity.py.

* the reference's pure-PyTorch training forward (`forward_one_frame`,
/root/reference/src/models/video_model_ht.py:452-496 with the `else` (non-HTS) branches of :26-317 and
`forward_prior_4x(..., spatial_prior_has_scales=True)`, /root/reference/src/models/common_model.py:231-282), restated
functionally over a plain state_dict and PINNED against tests/golden/htl_forward_64x64.npz (minted by importing the
reference modules, tests/golden/make_golden.py).  The product's counterpart of the CUDA proxy for this model
(dmc_htl_proxy.cpp) is the experimental dcvc_b200/csrc/codec_htl.cu, compared with this oracle by
tests/test_htl_gpu.py on a device and by tests/test_host_dry_run.py under kernel emulation.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref
from .dmci_oracle import CH_Y, CH_Z, _pad_to
from .hts_oracle import G, HtsOracle, _h


class HtlOracle(HtsOracle):
    # ---------------------------------------------------------------- sub-networks (video_model_ht.py, non-HTS branches)
    def rbu_bias(self, x, p, shortcut):
        """ResidualBlockUpsample with force_bias (layers.py:92-103,162-173): 1x1 conv + bias, pixel_shuffle(2), block"""
        t = F.conv2d(x, self._w(p + "up.conv.0.weight"), self._w(p + "up.conv.0.bias"))
        return self.dcb(self.r(F.pixel_shuffle(t, 2)), p + "conv.", shortcut=shortcut)

    def feature_adaptor_i(self, f):
        return self.seq(f, "feature_adaptor_i.conv.", 3)

    def feature_adaptor_m(self, memory, feature):
        return self.seq(torch.cat((memory, feature), 1), "feature_adaptor_m.conv.", 10)

    def feature_extractor(self, memory):
        return self.seq(memory, "feature_extractor.conv.", 2)

    def v_encoder(self, x_unshuffled, ctx, qp):
        q = self.r(self._w("q_encoder")[qp])
        t = self.seq(torch.cat((x_unshuffled, ctx), 1), "encoder.conv1.", 7, q_last=q)
        return self.r(ops_ref.conv3x3_s2(t, self._w("encoder.down.weight"), self._w("encoder.down.bias")))

    def v_hyper_enc(self, y):
        t = self.dcb(y, "hyper_encoder.conv.0.")
        t = self.rbd(t, "hyper_encoder.conv.1.", True)
        return self.rbd(t, "hyper_encoder.conv.2.", True)

    def v_hyper_dec(self, z_hat):
        t = self.rbu_bias(z_hat, "hyper_decoder.conv.0.", True)
        t = self.rbu_bias(t, "hyper_decoder.conv.1.", True)
        return self.dcb(t, "hyper_decoder.conv.2.")

    def temporal_prior(self, memory, qp):
        q = self.r(self._w("q_feature")[qp])
        t = self.r(memory * q.view(1, -1, 1, 1))
        return self.rbd(t, "temporal_prior_encoder.conv.", True)

    def v_spatial_prior_sm(self, y_hat_so_far, reduced, k):
        """adaptor_k(cat(y_hat_so_far, common)) -> 3 blocks -> 1x1: (scales, means)"""
        t = self.dcb(torch.cat((y_hat_so_far, reduced), 1), f"y_spatial_prior_adaptor_{k}.")
        t = self.seq(t, "y_spatial_prior.conv.", 3)
        return self.r(ops_ref.conv1x1(t, self._w("y_spatial_prior.conv.3.weight"), self._w("y_spatial_prior.conv.3.bias")))

    def v_decoder(self, y_hat, ctx, qp):
        """Decoder (:41-60): 3x3 SubpelConv2x with bias, 11 blocks on cat(feature, ctx), * quant_step"""
        up = F.conv2d(y_hat, self._w("decoder.up.conv.0.weight"), self._w("decoder.up.conv.0.bias"), padding=1)
        up = self.r(F.pixel_shuffle(up, 2))
        q = self.r(self._w("q_decoder")[qp])
        return self.seq(torch.cat((up, ctx), 1), "decoder.conv1.", 11, q_last=q)

    def recon_head_one(self, feature, i, common=None):
        t = self.seq(feature, f"recon_head.conv.{i}.", 5)
        return self.r(ops_ref.conv1x1(t, self._w(f"recon_head.conv.{i}.5.weight"), self._w(f"recon_head.conv.{i}.5.bias"))), None

    def recon_head(self, feature):
        return [self.recon_head_one(feature, i)[0] for i in range(G)]

    # ---------------------------------------------------------------- reference training forward (fp32)
    @torch.inference_mode()
    def forward_one_frame(self, x, qp: int, reset_feature_memory=False):
        """video_model_ht.py:452-496 with spatial_prior_has_scales=True; state as the reference"""
        assert not self.emu
        if self.memory is None:
            self.memory = self.feature_adaptor_i(self.feature_p)
        else:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        self.ctx = self.feature_extractor(self.memory)
        y = self.v_encoder(F.pixel_unshuffle(x, 8), self.ctx, qp)
        z = self.v_hyper_enc(y)
        z_hat = torch.round(z)
        params = self.v_prior_fusion(self.v_hyper_dec(z_hat), self.temporal_prior(self.memory, qp))
        quant_step, scales, means = params.chunk(3, 1)
        quant_step = torch.clamp_min(quant_step, 0.5)
        y = y * (1.0 / quant_step)
        reduced = ops_ref.conv1x1(params, self._w("y_spatial_prior_reduction.weight"), self._w("y_spatial_prior_reduction.bias"))
        B, C, H, W = y.shape
        y_hat_so_far = None
        y_q_tot = torch.zeros_like(y)
        for k in range(4):
            mask = torch.from_numpy(ops_ref.mask_4x(k, C, H, W))[None]
            if k > 0:
                scales, means = self.v_spatial_prior_sm(y_hat_so_far, reduced, k).chunk(2, 1)
            means_hat = means * mask
            y_q = torch.round((y - means_hat) * mask)
            y_hat = y_q + means_hat
            y_hat_so_far = y_hat if k == 0 else y_hat_so_far + y_hat
            y_q_tot = y_q_tot + y_q
        y_hat = y_hat_so_far * quant_step
        feature = self.v_decoder(y_hat, self.ctx, qp)
        x_hat = [F.pixel_shuffle(h, 8) for h in self.recon_head(feature)]
        self.feature_p = feature
        if reset_feature_memory:          # set_ref_feature (:406-411): recon_head(feature, for_reset=True) = last head
            head, _ = self.recon_head_one(feature, G - 1)
            self.memory = None
            self.ctx = None
            self.feature_p = head
        return {"x_hat": x_hat, "y_q": y_q_tot, "z_hat": z_hat, "y": y, "feature": feature}

    # ---------------------------------------------------------------- proxy restatement (fp16 emulation)
    # add_ref_feature_from_frame: HtsOracle's (dmc_htl_proxy.cpp:583-594 == dmc_hts_proxy.cpp:492-502)

    def _sp_np(self, acc, red_np, k):
        cat = np.concatenate([acc, red_np], axis=2)
        t = self._nchw32(cat)
        return self._nhwc16(self.v_spatial_prior_sm(t[:, :CH_Y], t[:, CH_Y:], k))

    @torch.inference_mode()
    def compress(self, x, qp: int, reset_feature_memory: bool, padding_b: int, padding_r: int):
        """x: [1,24,H,W] fp16-representable (dmc_htl_proxy.cpp:596-717)"""
        assert self.emu
        _, _, H, W = x.shape
        Hp, Wp = _pad_to(H, 16), _pad_to(W, 16)
        H16, W16 = Hp // 16, Wp // 16
        H16p, W16p = _pad_to(H16, 4), _pad_to(W16, 4)
        xu = ops_ref.unshuffle8_pad(x, padding_b, padding_r)
        y = self.v_encoder(xu, self.ctx, qp)
        y_pad = F.pad(y, (0, W16p - W16, 0, H16p - H16), mode="replicate")
        z = self.v_hyper_enc(y_pad)
        z_hat = self._canon(torch.clamp(ops_ref.round_half_away(z), -64, 63))
        z_i8 = z_hat[0].permute(1, 2, 0).contiguous().numpy().astype(np.int8).reshape(-1)
        common = self._params(z_hat, qp, H16, W16)
        p_np = self._nhwc16(common)
        q_dec, scales, means = p_np[..., :CH_Y], p_np[..., CH_Y:2 * CH_Y], p_np[..., 2 * CH_Y:]
        rcp = _h(np.float32(1.0) / np.maximum(q_dec.astype(np.float32), np.float32(0.5)))
        y_np = _h(self._nhwc16(y).astype(np.float32) * rcp.astype(np.float32))   # divide_with_clamp_min_inplace
        red_np = self._nhwc16(self.r(ops_ref.conv1x1(common, self._w("y_spatial_prior_reduction.weight"),
                                                      self._w("y_spatial_prior_reduction.bias"))))
        acc = np.zeros((H16, W16, CH_Y), dtype=np.float16)
        symbols = []
        for k in range(4):
            if k > 0:
                sm = self._sp_np(acc, red_np, k)
                scales, means = sm[..., :CH_Y], sm[..., CH_Y:]
            y_hat_k, sym_k, _ = ops_ref.entropy_enc_step_np(k, y_np, None, scales, means, self.skip_thres, self.lut)
            acc = _h(acc.astype(np.float32) + y_hat_k.astype(np.float32))
            symbols.append(sym_k)
        y_hat = _h(acc.astype(np.float32) * np.maximum(q_dec.astype(np.float32), np.float32(0.5)))
        total = sum(len(v) for v in symbols)
        ec_parallel = max(1, min(8, total // 32768))
        enc, _ = self._coder()
        enc.reset()
        enc.set_entropy_coder_parallel(ec_parallel)
        for k in (3, 2, 1, 0):
            enc.encode_y(np.ascontiguousarray(symbols[k]))
        enc.encode_z(z_i8, qp * CH_Z, CH_Z)
        enc.flush()
        stream = bytes(np.asarray(enc.get_encoded_stream()).tobytes())
        # enc_1: decoder, then memory / context for the NEXT chunk
        self.feature_p = self.v_decoder(self._nchw32(y_hat), self.ctx, qp)
        if reset_feature_memory:
            head, _ = self.recon_head_one(self.feature_p, G - 1)
            self.memory = self.feature_adaptor_i(head)
        else:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        self.ctx = self.feature_extractor(self.memory)
        return {"bit_stream": stream, "ec_parallel": ec_parallel, "symbols": symbols, "z_i8": z_i8, "y_hat": y_hat}

    @torch.inference_mode()
    def decompress(self, bit_stream: bytes, qp: int, height: int, width: int, ec_parallel: int,
                   reset_feature_memory: bool):
        """dmc_htl_proxy.cpp:719-915"""
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
        n_z = CH_Z * zh * zw
        dec.decode_z(n_z, qp * CH_Z, CH_Z)
        z_i8 = dec.get_decoded(n_z)
        z_hat = self._canon(torch.from_numpy(z_i8.astype(np.float32)).view(zh, zw, CH_Z).permute(2, 0, 1).unsqueeze(0))
        common = self._params(z_hat, qp, H16, W16)
        p_np = self._nhwc16(common)
        q_dec, scales, means = p_np[..., :CH_Y], p_np[..., CH_Y:2 * CH_Y], p_np[..., 2 * CH_Y:]
        red_np = self._nhwc16(self.r(ops_ref.conv1x1(common, self._w("y_spatial_prior_reduction.weight"),
                                                      self._w("y_spatial_prior_reduction.bias"))))
        self.ctx = self.feature_extractor(self.memory)
        acc = np.zeros((H16, W16, CH_Y), dtype=np.float16)
        for k in range(4):
            if k > 0:
                sm = self._sp_np(acc, red_np, k)
                scales, means = sm[..., :CH_Y], sm[..., CH_Y:]
            idx, _ = ops_ref.entropy_dec_index_np(k, scales, self.skip_thres, self.lut)
            dec.decode_y(np.ascontiguousarray(idx))
            decoded = dec.get_decoded(len(idx))
            y_hat_k = ops_ref.entropy_dec_restore_np(k, scales, means, self.skip_thres, decoded)
            acc = _h(acc.astype(np.float32) + y_hat_k.astype(np.float32))
        y_hat = _h(acc.astype(np.float32) * np.maximum(q_dec.astype(np.float32), np.float32(0.5)))
        self.feature_p = self.v_decoder(self._nchw32(y_hat), self.ctx, qp)
        heads = self.recon_head(self.feature_p)
        self.feature_i = heads[G - 1]
        self.memory_has_value = not reset_feature_memory
        return {"x_hat": [ops_ref.shuffle8_clamp(h, True) for h in heads], "y_hat": y_hat}
