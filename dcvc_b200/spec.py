"""Parameter specification of the DCVC-UF models (state_dict key -> shape) and a seeded synthetic
checkpoint generator.

The reference ships no checkpoints (checkpoints/.gitkeep) and there is no network, so every parity
test and the benchmark use *seeded synthetic weights*: Xavier-like random convolutions, damped
residual branches so activations stay O(1) through 13+ blocks, and hand-set final layers of the
entropy-parameter networks so that the predicted scales span the 0.11..16 Gaussian table and the
quantised latents stay within a few levels (random Xavier alone gives ~17 bpp).

Key names and shapes follow the reference modules exactly (src/models/image_model.py:21-147,
src/layers/layers.py:128-188, src/models/entropy_models.py:78-90); tests/test_spec.py checks this
against the reference tree when it is available and against a committed fixture otherwise.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

QP_NUM = 64
G_CH_SRC = 3 * 8 * 8   # image_model.py:15
G_CH_ENC_DEC = 384     # image_model.py:16
G_CH_Y = 256           # image_model.py:17
G_CH_Z = 128           # image_model.py:18


def _conv(spec, prefix, cout, cin, k=1, bias=True):
    spec[prefix + "weight"] = (cout, cin, k, k)
    if bias:
        spec[prefix + "bias"] = (cout,)


def depth_conv_block(spec, prefix, in_ch, out_ch, dcb2=False, force_adaptor=False):
    """DepthConvBlock (src/layers/layers.py:128-159)"""
    if in_ch != out_ch or force_adaptor:
        _conv(spec, prefix + "adaptor.", out_ch, in_ch)
    r = 2 if dcb2 else 1
    inner = out_ch // r
    _conv(spec, prefix + "dc.0.", inner, out_ch)
    spec[prefix + "dc.2.weight"] = (inner, 1, 3, 3)
    spec[prefix + "dc.2.bias"] = (inner,)
    _conv(spec, prefix + "dc.3.", out_ch, inner)
    _conv(spec, prefix + "ffn.0.", out_ch * 4 // r, out_ch)
    _conv(spec, prefix + "ffn.2.", out_ch, out_ch // r)


def residual_block_upsample(spec, prefix, in_ch, out_ch, dcb2=False, force_bias=False):
    """ResidualBlockUpsample (layers.py:162-173): SubpelConv2x(1x1, bias only with force_bias) + DepthConvBlock"""
    _conv(spec, prefix + "up.conv.0.", out_ch * 4, in_ch, bias=force_bias)
    depth_conv_block(spec, prefix + "conv.", out_ch, out_ch, dcb2=dcb2)


def residual_block_stride2(spec, prefix, in_ch, out_ch, dcb2=False):
    """ResidualBlockWithStride2 (layers.py:176-188): pixel_unshuffle(2) + 1x1 + DepthConvBlock"""
    _conv(spec, prefix + "down.", out_ch, in_ch * 4)
    depth_conv_block(spec, prefix + "conv.", out_ch, out_ch, dcb2=dcb2)


def dmci_spec() -> "OrderedDict[str, tuple]":
    """state_dict layout of DMCI (src/models/image_model.py:126-148), nn.Module registration order
    is irrelevant to consumers (set_param looks keys up by name, dmci_proxy.cpp:604-652)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["bit_estimator_z.h"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.b"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.a"] = (QP_NUM, G_CH_Z, 3)
    s["q_scale_enc"] = (QP_NUM, G_CH_ENC_DEC)
    s["q_scale_dec"] = (QP_NUM, G_CH_ENC_DEC)
    s["q_scale_y_enc"] = (QP_NUM, G_CH_Y)
    s["q_scale_y_dec"] = (QP_NUM, G_CH_Y)
    depth_conv_block(s, "enc.enc_1.", G_CH_SRC, G_CH_ENC_DEC)
    for i in range(6):
        depth_conv_block(s, f"enc.enc_2.{i}.", G_CH_ENC_DEC, G_CH_ENC_DEC)
    _conv(s, "enc.enc_2.6.", G_CH_Y, G_CH_ENC_DEC, k=3)
    depth_conv_block(s, "hyper_enc.conv.0.", G_CH_Y, G_CH_Z)
    residual_block_stride2(s, "hyper_enc.conv.1.", G_CH_Z, G_CH_Z)
    residual_block_stride2(s, "hyper_enc.conv.2.", G_CH_Z, G_CH_Z)
    residual_block_upsample(s, "hyper_dec.conv.0.", G_CH_Z, G_CH_Z)
    residual_block_upsample(s, "hyper_dec.conv.1.", G_CH_Z, G_CH_Z)
    depth_conv_block(s, "hyper_dec.conv.2.", G_CH_Z, G_CH_Y)
    depth_conv_block(s, "y_prior_fusion.conv.0.", G_CH_Y, G_CH_Y * 2)
    depth_conv_block(s, "y_prior_fusion.conv.1.", G_CH_Y * 2, G_CH_Y * 2)
    depth_conv_block(s, "y_prior_fusion.conv.2.", G_CH_Y * 2, G_CH_Y * 2)
    _conv(s, "y_prior_fusion.conv.3.", G_CH_Y * 2, G_CH_Y * 2)
    _conv(s, "y_spatial_prior_reduction.", G_CH_Y, G_CH_Y * 2)
    for i in (1, 2, 3):
        depth_conv_block(s, f"y_spatial_prior_adaptor_{i}.", G_CH_Y * 2, G_CH_Y * 2, force_adaptor=True)
    for i in range(3):
        depth_conv_block(s, f"y_spatial_prior.conv.{i}.", G_CH_Y * 2, G_CH_Y * 2)
    _conv(s, "y_spatial_prior.conv.3.", G_CH_Y * 2, G_CH_Y * 2)
    residual_block_upsample(s, "dec.dec_1.0.", G_CH_Y, G_CH_ENC_DEC)
    for i in range(1, 13):
        depth_conv_block(s, f"dec.dec_1.{i}.", G_CH_ENC_DEC, G_CH_ENC_DEC)
    depth_conv_block(s, "dec.dec_2.", G_CH_ENC_DEC, G_CH_SRC)
    return s


G_FRAME_DELAY = 8                      # video_model_ht.py:16
G_CH_SRC_D = G_CH_SRC * G_FRAME_DELAY  # 1536
G_CH_D = 512                           # video_model_ht.py:21
G_CH_M = 512                           # video_model_ht.py:22
G_CH_RECON = 256                       # video_model_ht.py:23


def hts_spec() -> "OrderedDict[str, tuple]":
    """state_dict layout of DMC(ModelStructure.HTS) (src/models/video_model_ht.py:26-345)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["bit_estimator_z.h"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.b"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.a"] = (QP_NUM, G_CH_Z, 3)
    s["q_encoder"] = (QP_NUM, G_CH_D)
    s["q_decoder"] = (QP_NUM, G_CH_D)
    s["q_feature"] = (QP_NUM, G_CH_D)
    # FeatureAdaptorI / FeatureAdaptorM / FeatureExtractor (video_model_ht.py:95-164)
    depth_conv_block(s, "feature_adaptor_i.conv.0.", G_CH_SRC, G_CH_M, dcb2=True)
    for i in range(1, 4):
        depth_conv_block(s, f"feature_adaptor_i.conv.{i}.", G_CH_M, G_CH_M, dcb2=True)
    depth_conv_block(s, "feature_adaptor_m.conv.0.", G_CH_M + G_CH_D, G_CH_M, dcb2=True)
    for i in range(1, 6):
        depth_conv_block(s, f"feature_adaptor_m.conv.{i}.", G_CH_M, G_CH_M, dcb2=True)
    for i in range(5):
        depth_conv_block(s, f"feature_extractor.conv.{i}.", G_CH_D, G_CH_D, dcb2=True)
    # Encoder (:63-92)
    depth_conv_block(s, "encoder.conv1.0.", G_CH_SRC_D + G_CH_D, G_CH_D, dcb2=True)
    for i in range(1, 6):
        depth_conv_block(s, f"encoder.conv1.{i}.", G_CH_D, G_CH_D, dcb2=True)
    _conv(s, "encoder.down.", G_CH_Y, G_CH_D, k=3)
    # HyperEncoder / HyperDecoder (:167-198)
    depth_conv_block(s, "hyper_encoder.conv.0.", G_CH_Y, G_CH_Y)
    residual_block_stride2(s, "hyper_encoder.conv.1.", G_CH_Y, G_CH_Y)
    residual_block_stride2(s, "hyper_encoder.conv.2.", G_CH_Y, G_CH_Z)
    residual_block_upsample(s, "hyper_decoder.conv.0.", G_CH_Z, G_CH_Y)
    residual_block_upsample(s, "hyper_decoder.conv.1.", G_CH_Y, G_CH_Y)
    depth_conv_block(s, "hyper_decoder.conv.2.", G_CH_Y, G_CH_Y)
    # TemporalPriorEncoder (:307-317), PriorFusion (:201-212)
    residual_block_stride2(s, "temporal_prior_encoder.conv.", G_CH_D, G_CH_Y * 2)
    for i in range(3):
        depth_conv_block(s, f"y_prior_fusion.conv.{i}.", G_CH_Y * 3, G_CH_Y * 3)
    _conv(s, "y_prior_fusion.conv.3.", G_CH_Y * 3, G_CH_Y * 3)
    _conv(s, "y_spatial_prior_reduction.", G_CH_Y, G_CH_Y * 3)
    for i in (1, 2, 3):
        depth_conv_block(s, f"y_spatial_prior_adaptor_{i}.", G_CH_Y * 2, G_CH_Y * 2, force_adaptor=True)
    for i in range(3):
        depth_conv_block(s, f"y_spatial_prior.conv.{i}.", G_CH_Y * 2, G_CH_Y * 2)
    _conv(s, "y_spatial_prior.conv.3.", G_CH_Y, G_CH_Y * 2)
    # Decoder (:26-60)
    _conv(s, "decoder.up.conv.0.", G_CH_D * 4, G_CH_Y, bias=False)
    depth_conv_block(s, "decoder.conv1.0.", G_CH_D * 2, G_CH_D, dcb2=True)
    for i in range(1, 7):
        depth_conv_block(s, f"decoder.conv1.{i}.", G_CH_D, G_CH_D, dcb2=True)
    # ReconHead (:215-275)
    for i in range(G_FRAME_DELAY // 2):
        depth_conv_block(s, f"recon_head.conv1.{i}.0.", G_CH_D, G_CH_D)
    for i in range(G_FRAME_DELAY):
        depth_conv_block(s, f"recon_head.conv2.{i}.0.", G_CH_D, G_CH_RECON)
        depth_conv_block(s, f"recon_head.conv2.{i}.1.", G_CH_RECON, G_CH_RECON)
        depth_conv_block(s, f"recon_head.conv2.{i}.2.", G_CH_RECON, G_CH_RECON)
        _conv(s, f"recon_head.conv2.{i}.3.", G_CH_SRC, G_CH_RECON)
    return s


def htl_spec() -> "OrderedDict[str, tuple]":
    """state_dict layout of DMC(ModelStructure.HTL) (src/models/video_model_ht.py: the `else` branches of :26-317).
    Used by the experimental HT-L codec (csrc/codec_htl.cu, opt-in), the oracle (oracle/htl_oracle.py) and its goldens."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["bit_estimator_z.h"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.b"] = (QP_NUM, G_CH_Z, 4)
    s["bit_estimator_z.a"] = (QP_NUM, G_CH_Z, 3)
    s["q_encoder"] = (QP_NUM, G_CH_D)
    s["q_decoder"] = (QP_NUM, G_CH_D)
    s["q_feature"] = (QP_NUM, G_CH_D)
    depth_conv_block(s, "feature_adaptor_i.conv.0.", G_CH_SRC, G_CH_M)
    for i in range(1, 3):
        depth_conv_block(s, f"feature_adaptor_i.conv.{i}.", G_CH_M, G_CH_M)
    depth_conv_block(s, "feature_adaptor_m.conv.0.", G_CH_M + G_CH_D, G_CH_M)
    for i in range(1, 10):
        depth_conv_block(s, f"feature_adaptor_m.conv.{i}.", G_CH_M, G_CH_M)
    for i in range(2):
        depth_conv_block(s, f"feature_extractor.conv.{i}.", G_CH_D, G_CH_D)
    depth_conv_block(s, "encoder.conv1.0.", G_CH_SRC_D + G_CH_D, G_CH_D)
    for i in range(1, 7):
        depth_conv_block(s, f"encoder.conv1.{i}.", G_CH_D, G_CH_D)
    _conv(s, "encoder.down.", G_CH_Y, G_CH_D, k=3)
    depth_conv_block(s, "hyper_encoder.conv.0.", G_CH_Y, G_CH_Y)
    residual_block_stride2(s, "hyper_encoder.conv.1.", G_CH_Y, G_CH_Y)
    residual_block_stride2(s, "hyper_encoder.conv.2.", G_CH_Y, G_CH_Z)
    residual_block_upsample(s, "hyper_decoder.conv.0.", G_CH_Z, G_CH_Y, force_bias=True)
    residual_block_upsample(s, "hyper_decoder.conv.1.", G_CH_Y, G_CH_Y, force_bias=True)
    depth_conv_block(s, "hyper_decoder.conv.2.", G_CH_Y, G_CH_Y)
    residual_block_stride2(s, "temporal_prior_encoder.conv.", G_CH_D, G_CH_Y * 2)
    for i in range(3):
        depth_conv_block(s, f"y_prior_fusion.conv.{i}.", G_CH_Y * 3, G_CH_Y * 3)
    _conv(s, "y_prior_fusion.conv.3.", G_CH_Y * 3, G_CH_Y * 3)
    _conv(s, "y_spatial_prior_reduction.", G_CH_Y, G_CH_Y * 3)
    for i in (1, 2, 3):
        depth_conv_block(s, f"y_spatial_prior_adaptor_{i}.", G_CH_Y * 2, G_CH_Y * 2, force_adaptor=True)
    for i in range(3):
        depth_conv_block(s, f"y_spatial_prior.conv.{i}.", G_CH_Y * 2, G_CH_Y * 2)
    _conv(s, "y_spatial_prior.conv.3.", G_CH_Y * 2, G_CH_Y * 2)   # (scales, means)
    _conv(s, "decoder.up.conv.0.", G_CH_D * 4, G_CH_Y, k=3)      # 3x3 SubpelConv2x with bias
    depth_conv_block(s, "decoder.conv1.0.", G_CH_D * 2, G_CH_D)
    for i in range(1, 11):
        depth_conv_block(s, f"decoder.conv1.{i}.", G_CH_D, G_CH_D)
    for i in range(G_FRAME_DELAY):
        depth_conv_block(s, f"recon_head.conv.{i}.0.", G_CH_D, G_CH_RECON)
        for j in range(1, 5):
            depth_conv_block(s, f"recon_head.conv.{i}.{j}.", G_CH_RECON, G_CH_RECON)
        _conv(s, f"recon_head.conv.{i}.5.", G_CH_SRC, G_CH_RECON)
    return s


# DCVC-UF low-delay model constants (src/models/video_model_ld.py:16-21)
LD_CH_SRC_D = 3 * 8 * 8
LD_CH_Y = 128
LD_CH_Z = 128
LD_CH_D = 256
LD_CH_M = 256


def ld_spec() -> "OrderedDict[str, tuple]":
    """state_dict layout of the low-delay DMC (src/models/video_model_ld.py:24-211).  Used by the oracle
    (oracle/ld_oracle.py), its goldens and the LD codec (csrc/codec_ld.cu)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["bit_estimator_z.h"] = (QP_NUM, LD_CH_Z, 4)
    s["bit_estimator_z.b"] = (QP_NUM, LD_CH_Z, 4)
    s["bit_estimator_z.a"] = (QP_NUM, LD_CH_Z, 3)
    s["q_encoder"] = (QP_NUM, LD_CH_D)
    s["q_decoder"] = (QP_NUM, LD_CH_D)
    s["q_feature"] = (QP_NUM, LD_CH_Y * 2)
    # FeatureAdaptorI / M, FeatureExtractor (:61-107)
    depth_conv_block(s, "feature_adaptor_i.conv.0.", LD_CH_SRC_D, LD_CH_M, dcb2=True)
    for i in range(1, 4):
        depth_conv_block(s, f"feature_adaptor_i.conv.{i}.", LD_CH_M, LD_CH_M, dcb2=True)
    depth_conv_block(s, "feature_adaptor_m.conv.0.", LD_CH_M + LD_CH_D, LD_CH_M, dcb2=True)
    for i in range(1, 4):
        depth_conv_block(s, f"feature_adaptor_m.conv.{i}.", LD_CH_M, LD_CH_M, dcb2=True)
    for i in range(5):
        depth_conv_block(s, f"feature_extractor.conv.{i}.", LD_CH_M, LD_CH_M, dcb2=True)
    # Encoder (:43-58)
    depth_conv_block(s, "encoder.conv1.0.", LD_CH_SRC_D + LD_CH_M, LD_CH_D, dcb2=True)
    depth_conv_block(s, "encoder.conv1.1.", LD_CH_D, LD_CH_D, dcb2=True)
    depth_conv_block(s, "encoder.conv2.", LD_CH_D, LD_CH_D, dcb2=True)
    _conv(s, "encoder.down.", LD_CH_Y, LD_CH_D, k=3)
    # HyperEncoder / HyperDecoder (:110-135)
    depth_conv_block(s, "hyper_encoder.conv.0.", LD_CH_Y, LD_CH_Z, dcb2=True)
    residual_block_stride2(s, "hyper_encoder.conv.1.", LD_CH_Z, LD_CH_Z, dcb2=True)
    residual_block_stride2(s, "hyper_encoder.conv.2.", LD_CH_Z, LD_CH_Z, dcb2=True)
    residual_block_upsample(s, "hyper_decoder.conv.0.", LD_CH_Z, LD_CH_Z, dcb2=True)
    residual_block_upsample(s, "hyper_decoder.conv.1.", LD_CH_Z, LD_CH_Z, dcb2=True)
    depth_conv_block(s, "hyper_decoder.conv.2.", LD_CH_Z, LD_CH_Y, dcb2=True)
    # TemporalPriorEncoder (:182-188), PriorFusion (:138-149), SpatialPrior (:169-179)
    residual_block_stride2(s, "temporal_prior_encoder.conv.", LD_CH_M, LD_CH_Y * 2, dcb2=True)
    for i in range(3):
        depth_conv_block(s, f"y_prior_fusion.conv.{i}.", LD_CH_Y * 3, LD_CH_Y * 3, dcb2=True)
    _conv(s, "y_prior_fusion.conv.3.", LD_CH_Y * 3, LD_CH_Y * 3)
    depth_conv_block(s, "y_spatial_prior.conv.0.", LD_CH_Y * 4, LD_CH_Y * 2, dcb2=True)
    depth_conv_block(s, "y_spatial_prior.conv.1.", LD_CH_Y * 2, LD_CH_Y * 2, dcb2=True)
    _conv(s, "y_spatial_prior.conv.2.", LD_CH_Y, LD_CH_Y * 2)
    # Decoder (:24-40), ReconHead (:152-166)
    _conv(s, "decoder.up.conv.0.", LD_CH_D * 4, LD_CH_Y, bias=False)
    depth_conv_block(s, "decoder.conv1.0.", LD_CH_D + LD_CH_M, LD_CH_D, dcb2=True)
    depth_conv_block(s, "decoder.conv1.1.", LD_CH_D, LD_CH_D, dcb2=True)
    depth_conv_block(s, "decoder.conv1.2.", LD_CH_D, LD_CH_D, dcb2=True)
    _conv(s, "decoder.conv2.", LD_CH_D, LD_CH_D)
    for i in range(3):
        depth_conv_block(s, f"recon_head.conv.{i}.", LD_CH_D, LD_CH_D, dcb2=True)
    _conv(s, "recon_head.head.", LD_CH_SRC_D, LD_CH_D)
    return s


# hand-calibrated gains (checked with the reference modules on random inputs): keep z within a few
# levels, the predicted (scale, mean) noise around the hand-set biases small, x_hat inside [-0.5, 0.5]
_WEIGHT_GAIN = {
    "hyper_enc.conv.0.adaptor.weight": 0.25,
    "y_prior_fusion.conv.3.weight": 0.06,
    "y_spatial_prior.conv.3.weight": 0.06,
    "dec.dec_2.adaptor.weight": 0.08,
    "dec.dec_2.dc.3.weight": 0.3,
    "dec.dec_2.ffn.2.weight": 0.3,
    # HT-S
    "decoder.conv1.0.adaptor.weight": 0.35,          # keep the memory/context recurrence contractive
    "feature_adaptor_m.conv.0.adaptor.weight": 0.35,
    "feature_adaptor_i.conv.0.adaptor.weight": 0.5,
    # LD
    "y_spatial_prior.conv.2.weight": 0.06,
    "recon_head.head.weight": 0.5,
}
for _i in range(G_FRAME_DELAY):
    _WEIGHT_GAIN[f"recon_head.conv2.{_i}.3.weight"] = 0.5
    _WEIGHT_GAIN[f"recon_head.conv.{_i}.5.weight"] = 0.5   # HT-L


def synth_state_dict(spec, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic checkpoint (float32, CPU) for a spec from this module."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    qp = torch.arange(QP_NUM, dtype=torch.float32) / (QP_NUM - 1)
    for name, shape in spec.items():
        if name.startswith("bit_estimator_z."):
            leaf = name.rsplit(".", 1)[1]
            mean = {"h": 0.5, "b": 0.0, "a": 0.0}[leaf]
            t = torch.randn(shape, generator=g) * 0.2 + mean
        elif name.startswith("q_scale") or name in ("q_encoder", "q_decoder", "q_feature"):
            jitter = 1.0 + 0.05 * torch.randn(shape, generator=g)
            if name in ("q_scale_enc", "q_encoder"):
                t = (0.6 + 0.8 * qp)[:, None] * jitter
            elif name == "q_feature":
                t = (0.8 + 0.4 * qp)[:, None] * jitter
            elif name in ("q_scale_dec", "q_decoder"):
                t = (1.0 / (0.6 + 0.8 * qp))[:, None] * jitter
            elif name == "q_scale_y_enc":
                t = (0.4 + 1.2 * qp)[:, None] * jitter
            else:
                t = (1.0 / (0.4 + 1.2 * qp))[:, None] * jitter
        elif name.endswith(".weight"):
            cout, cin, kh, kw = shape
            if cin == 1 and kh == 3:  # depthwise: keep the variance of the input
                std = 1.0 / 3.0
            else:
                std = math.sqrt(2.0 / ((cin + cout) * kh * kw))
            t = torch.randn(shape, generator=g) * std
            if ".dc.3." in name or ".ffn.2." in name:
                t = t * 0.3  # damp the residual branches
            if name in ("enc.enc_2.6.weight", "encoder.down.weight"):
                t = t * 1.5  # latent spread: a few quantisation levels
            t = t * _WEIGHT_GAIN.get(name, 1.0)
        elif name.endswith(".bias"):
            t = torch.randn(shape, generator=g) * 0.02
            if name == "y_prior_fusion.conv.3.bias" and shape[0] in (3 * G_CH_Y, 3 * LD_CH_Y):   # HT / LD: (q_dec, scales, means)
                third = shape[0] // 3
                perm = torch.randperm(third, generator=g)
                t[:third] = torch.rand(third, generator=g) * 1.0 + 0.6
                t[third:2 * third] = torch.exp(torch.linspace(math.log(0.02), math.log(1.5), third))[perm]
                t[2 * third:] = torch.randn(third, generator=g) * 0.3
            elif name == "y_spatial_prior.conv.3.bias" and shape[0] == G_CH_Y:    # HT-S: means only
                t = torch.randn(shape, generator=g) * 0.3
            elif name == "y_spatial_prior.conv.2.bias" and shape[0] == LD_CH_Y:   # LD: means only
                t = torch.randn(shape, generator=g) * 0.3
            elif name in ("y_prior_fusion.conv.3.bias", "y_spatial_prior.conv.3.bias"):
                half = shape[0] // 2
                perm = torch.randperm(half, generator=g)
                t[:half] = torch.exp(torch.linspace(math.log(0.02), math.log(1.5), half))[perm]  # scales
                t[half:] = torch.randn(half, generator=g) * 0.3         # means
        else:
            raise KeyError(f"synth_state_dict: unhandled parameter {name}")
        out[name] = t.float().contiguous()
    return out
