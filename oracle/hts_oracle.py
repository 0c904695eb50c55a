"""CPU oracle for the DCVC-UF HT-S chunk codec (8 frames per chunk).  TEST INFRASTRUCTURE ONLY.

* `forward_one_frame` restates the reference's training forward, src/models/video_model_ht.py:452-496
  (+ :357-362 apply_feature_adaptor, :406-411 set_ref_feature, common_model.py:231-282 with
  spatial_prior_has_scales=False) in fp32; PINNED by tests/golden/hts_forward_*.npz minted from
  the reference modules.
* `add_ref_feature_from_frame` / `compress` / `decompress` restate the CUDA proxy control flow,
  src/layers/extensions/inference/dmc_hts_proxy.cpp:492-502 (reference feature), :504-585 (compress),
  :587-710 (decompress), :764-851 (worker, adaptor scheduling), with the half-precision per-element
  semantics of elementwise/stream.cu:422-452 (divide_with_clamp), :548-683 (process_with_mask
  variants), :685-754 (restore_y variants) and the reference's own rANS coder (oracle/_ref).
  The CUDA proxy itself cannot run here: PARITY UNPINNED at that level (see dmci_oracle.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref
from .dmci_oracle import CH_Y, CH_Z, DmciOracle, _pad_to

G = 8  # frames per chunk


def _h(a):
    return a.astype(np.float16)


class HtsOracle(DmciOracle):
    def __init__(self, state_dict, skip_thres=0.0, emulate_fp16=True, threads=None):
        super().__init__(state_dict, skip_thres, emulate_fp16, threads)
        self.clear_dpb()

    def clear_dpb(self):
        self.memory = None        # [1,512,H8,W8]
        self.ctx = None
        self.feature_p = None     # decoder output of the previous chunk
        self.feature_i = None     # [1,192,H8,W8] reference feature of an intra frame / reset head
        self.memory_has_value = False

    # ---------------------------------------------------------------- blocks with explicit shortcut flag
    def rbu(self, x, p, shortcut):
        t = self.r(ops_ref.tconv2x2(x, self._w(p + "up.conv.0.weight")))
        return self.dcb(t, p + "conv.", shortcut=shortcut)

    def rbd(self, x, p, shortcut):
        t = self.r(ops_ref.conv2x2_s2(x, self._w(p + "down.weight"), self._w(p + "down.bias")))
        return self.dcb(t, p + "conv.", shortcut=shortcut)

    def seq(self, x, prefix, n, q_last=None):
        for i in range(n):
            x = self.dcb(x, f"{prefix}{i}.", q=q_last if i == n - 1 else None)
        return x

    # ---------------------------------------------------------------- sub-networks (video_model_ht.py)
    def feature_adaptor_i(self, f):
        return self.seq(f, "feature_adaptor_i.conv.", 4)

    def feature_adaptor_m(self, memory, feature):
        return self.seq(torch.cat((memory, feature), 1), "feature_adaptor_m.conv.", 6)

    def feature_extractor(self, memory):
        return self.seq(memory, "feature_extractor.conv.", 5)

    def v_encoder(self, x_unshuffled, ctx, qp):
        q = self.r(self._w("q_encoder")[qp])
        t = self.seq(torch.cat((x_unshuffled, ctx), 1), "encoder.conv1.", 6, q_last=q)
        return self.r(ops_ref.conv3x3_s2(t, self._w("encoder.down.weight"), self._w("encoder.down.bias")))

    def v_hyper_enc(self, y):
        t = self.dcb(y, "hyper_encoder.conv.0.")
        t = self.rbd(t, "hyper_encoder.conv.1.", False)
        return self.rbd(t, "hyper_encoder.conv.2.", False)

    def v_hyper_dec(self, z_hat):
        t = self.rbu(z_hat, "hyper_decoder.conv.0.", False)
        t = self.rbu(t, "hyper_decoder.conv.1.", False)
        return self.dcb(t, "hyper_decoder.conv.2.")

    def temporal_prior(self, memory, qp):
        q = self.r(self._w("q_feature")[qp])
        t = self.r(memory * q.view(1, -1, 1, 1))
        return self.rbd(t, "temporal_prior_encoder.conv.", False)

    def v_prior_fusion(self, hyper, temporal):
        t = self.seq(torch.cat((hyper, temporal), 1), "y_prior_fusion.conv.", 3)
        return self.r(ops_ref.conv1x1(t, self._w("y_prior_fusion.conv.3.weight"), self._w("y_prior_fusion.conv.3.bias")))

    def v_spatial_prior(self, y_hat_so_far, reduced, k):
        t = self.dcb(torch.cat((y_hat_so_far, reduced), 1), f"y_spatial_prior_adaptor_{k}.")
        t = self.seq(t, "y_spatial_prior.conv.", 3)
        return self.r(ops_ref.conv1x1(t, self._w("y_spatial_prior.conv.3.weight"), self._w("y_spatial_prior.conv.3.bias")))

    def v_decoder(self, y_hat, ctx, qp):
        up = self.r(ops_ref.tconv2x2(y_hat, self._w("decoder.up.conv.0.weight")))
        q = self.r(self._w("q_decoder")[qp])
        return self.seq(torch.cat((up, ctx), 1), "decoder.conv1.", 7, q_last=q)

    def recon_head_one(self, feature, i, common=None):
        if common is None:
            common = self.dcb(feature, f"recon_head.conv1.{i // 2}.0.")
        t = self.seq(common, f"recon_head.conv2.{i}.", 3)
        return self.r(ops_ref.conv1x1(t, self._w(f"recon_head.conv2.{i}.3.weight"), self._w(f"recon_head.conv2.{i}.3.bias"))), common

    def recon_head(self, feature):
        outs, common = [], None
        for i in range(G):
            head, common = self.recon_head_one(feature, i, None if i % 2 == 0 else common)
            outs.append(head)
        return outs

    # ---------------------------------------------------------------- reference training forward (fp32)
    @torch.inference_mode()
    def forward_one_frame(self, x, qp: int, reset_feature_memory=False):
        """video_model_ht.py:452-496; state (ref_feature -> self.feature_p, memory, ctx) as the reference."""
        assert not self.emu
        # apply_feature_adaptor (:357-362)
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
                means = self.v_spatial_prior(y_hat_so_far, reduced, k)
            means_hat = means * mask
            y_q = torch.round((y - means_hat) * mask)
            y_hat = y_q + means_hat
            y_hat_so_far = y_hat if k == 0 else y_hat_so_far + y_hat
            y_q_tot = y_q_tot + y_q
        y_hat = y_hat_so_far * quant_step
        feature = self.v_decoder(y_hat, self.ctx, qp)
        x_hat = [F.pixel_shuffle(h, 8) for h in self.recon_head(feature)]
        # set_ref_feature (:406-411)
        self.feature_p = feature
        if reset_feature_memory:
            head, _ = self.recon_head_one(feature, G - 1)
            self.memory = None
            self.ctx = None
            self.feature_p = head
        return {"x_hat": x_hat, "y_q": y_q_tot, "z_hat": z_hat, "y": y, "feature": feature}

    # ---------------------------------------------------------------- proxy restatement (fp16 emulation)
    def add_ref_feature_from_frame(self, frame, apply_adaptor=True):
        """dmc_hts_proxy.cpp:492-502; frame: [1,3,Hp,Wp] (already padded reconstruction)."""
        self.feature_i = self._canon(F.pixel_unshuffle(frame, 8))
        if apply_adaptor:
            self.memory = self.feature_adaptor_i(self.feature_i)
            self.ctx = self.feature_extractor(self.memory)
        self.memory_has_value = apply_adaptor

    def _params(self, z_hat, qp, H16, W16):
        temporal = self.temporal_prior(self.memory, qp)
        hyper = self.v_hyper_dec(z_hat)[:, :, :H16, :W16]
        common = self.v_prior_fusion(hyper, temporal)
        return common

    def _means_only_steps(self, common, y_np=None, decoded_dense=None):
        """the 4 means-refinement steps with fixed scales (dmc_hts_proxy.cpp:531-557 enc, :676-702 dec).
        Encoder: y_np = y already divided by clamp_min(q_dec, .5) -> returns (acc, y_q int [H,W,C]);
        decoder: decoded_dense = y_q [H,W,C] -> returns acc."""
        p_np = self._nhwc16(common)
        q_dec, scales, means0 = p_np[..., :CH_Y], p_np[..., CH_Y:2 * CH_Y], p_np[..., 2 * CH_Y:]
        reduced = self.r(ops_ref.conv1x1(common, self._w("y_spatial_prior_reduction.weight"),
                                         self._w("y_spatial_prior_reduction.bias")))
        H, W, C = scales.shape
        thres = np.float32(np.float16(self.skip_thres))
        acc = np.zeros((H, W, C), dtype=np.float16)
        y_q_all = np.zeros((H, W, C), dtype=np.int32)
        means = means0
        for k in range(4):
            if k > 0:
                means = self._nhwc16(self.v_spatial_prior(self._nchw32(acc), reduced, k))
            m = np.transpose(ops_ref.mask_4x(k, C, H, W), (1, 2, 0))
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
        y_hat = _h(acc.astype(np.float32) * qd)
        return y_hat, y_q_all, scales

    def _update_memory(self, reset_feature_memory_prev):
        pass

    @torch.inference_mode()
    def compress(self, x, qp: int, reset_feature_memory: bool, padding_b: int, padding_r: int):
        """x: [1,24,H,W] fp16-representable.  Encoder state machine: memory/ctx were produced at the end of
        the previous call (dmc_hts_proxy.cpp:575-576, 832-842)."""
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
        q_dec = p_np[..., :CH_Y]
        # y / clamp_min(q_dec, .5) in half: y * hrcp(max(q, .5))  (stream.cu:422-443)
        rcp = _h(np.float32(1.0) / np.maximum(q_dec.astype(np.float32), np.float32(0.5)))
        y_np = _h(self._nhwc16(y).astype(np.float32) * rcp.astype(np.float32))
        y_hat, y_q, scales = self._means_only_steps(common, y_np=y_np)
        # one symbol stream over the whole latent, NHWC order (dmc_hts_proxy.cpp:559-562)
        keep = scales.astype(np.float32) > np.float32(np.float16(self.skip_thres))
        idx = self.lut[scales.view(np.uint16)].astype(np.int32)
        sym = ((y_q << 8) + idx).astype(np.int16).reshape(-1)[keep.reshape(-1)]
        ec_parallel = max(1, min(8, len(sym) // 32768))
        enc, _ = self._coder()
        enc.reset()
        enc.set_entropy_coder_parallel(ec_parallel)
        enc.encode_y(np.ascontiguousarray(sym))
        enc.encode_z(z_i8, qp * CH_Z, CH_Z)
        enc.flush()
        stream = bytes(np.asarray(enc.get_encoded_stream()).tobytes())
        # enc_1: decoder, then the memory/context for the NEXT chunk
        self.feature_p = self.v_decoder(self._nchw32(y_hat), self.ctx, qp)
        if reset_feature_memory:
            head, _ = self.recon_head_one(self.feature_p, G - 1)
            self.memory = self.feature_adaptor_i(head)
        else:
            self.memory = self.feature_adaptor_m(self.memory, self.feature_p)
        self.ctx = self.feature_extractor(self.memory)
        return {"bit_stream": stream, "ec_parallel": ec_parallel, "symbols": sym, "z_i8": z_i8, "y_hat": y_hat}

    @torch.inference_mode()
    def decompress(self, bit_stream: bytes, qp: int, height: int, width: int, ec_parallel: int,
                   reset_feature_memory: bool):
        """Decoder state machine: the memory update of the previous chunk is applied lazily here
        (dmc_hts_proxy.cpp:613-620, 844-851)."""
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
        scales = p_np[..., CH_Y:2 * CH_Y]
        keep = scales.astype(np.float32) > np.float32(np.float16(self.skip_thres))
        idx = self.lut[scales.view(np.uint16)].reshape(-1)[keep.reshape(-1)]
        dec.decode_y(np.ascontiguousarray(idx))
        decoded = dec.get_decoded(len(idx))
        self.ctx = self.feature_extractor(self.memory)
        dense = np.zeros(scales.size, dtype=np.int32)
        dense[keep.reshape(-1)] = decoded
        y_hat, _, _ = self._means_only_steps(common, decoded_dense=dense.reshape(scales.shape))
        self.feature_p = self.v_decoder(self._nchw32(y_hat), self.ctx, qp)
        heads = self.recon_head(self.feature_p)
        self.feature_i = heads[G - 1]
        self.memory_has_value = not reset_feature_memory
        return {"x_hat": [ops_ref.shuffle8_clamp(h, True) for h in heads], "y_hat": y_hat}
