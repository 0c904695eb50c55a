"""Host-side entropy-model tables (CDFs) — the integer contract between the kernels and the rANS
coder.  Mirrors the reference's `src/models/entropy_models.py` (GaussianEncoder.update :184-217,
BitEstimator.update :113-149, EntropyCoder.pmf_to_cdf / reorder_prob :45-75) with the same names and
argument meaning, so `CompressionModel.update(skip_thres)` callers are served unchanged.  The only
native dependency is libdcvc_b200.so (pmf_to_quantized_cdf), not MLCodec_extensions_cpp.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

MAX_ENTROPY_CODING_VALUE = 8  # entropy_models.py:12


def pmf_to_quantized_cdf(pmf) -> torch.Tensor:
    """ryg-style 16-bit quantisation (src/cpp/py_rans/py_rans.cpp:36-94) via the C ABI."""
    lib = _lib.load()
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    _lib.check(lib.dcvc_pmf_to_quantized_cdf(p.ctypes.data_as(C.c_void_p), p.size,
                                             out.ctypes.data_as(C.c_void_p)), "pmf_to_quantized_cdf")
    return torch.from_numpy(out.astype(np.int32))


def reorder_prob(prob: torch.Tensor) -> torch.Tensor:
    """symmetric pmf [-r..r] + tail  ->  0, +1, -1, +2, -2, ... + tail (entropy_models.py:45-57)"""
    length = prob.size(0)
    out = prob.clone()
    center = (length - 1) // 2
    out[0] = prob[center]
    for i in range(1, center + 1):
        out[2 * i - 1] = prob[center + i]
        out[2 * i] = prob[center - i]
    return out


def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length) -> torch.Tensor:
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i, p in enumerate(pmf):
        prob = torch.cat((p[: int(pmf_length[i])], tail_mass[i]), dim=0)
        c = pmf_to_quantized_cdf(reorder_prob(prob).tolist())
        cdf[i, : c.size(0)] = c
    return cdf


def gaussian_scale_table(scale_min=0.11, scale_max=16.0, levels=128) -> torch.Tensor:
    return torch.exp(torch.linspace(math.log(scale_min), math.log(scale_max), levels))


def gaussian_cdf_tables():
    """(quantized_cdf int32 [128, 19], cdf_length int32 [128]) — GaussianEncoder.update."""
    table = gaussian_scale_table()
    zeros = torch.zeros_like(table)
    sym_range = zeros + MAX_ENTROPY_CODING_VALUE
    dist = torch.distributions.normal.Normal(0.0, table)
    for i in range(MAX_ENTROPY_CODING_VALUE, 1, -1):
        probs = torch.squeeze(dist.cdf(zeros + i))
        sym_range = torch.where(probs > 0.999, i, sym_range)
    sym_range = sym_range.int()
    pmf_length = 2 * sym_range + 1
    max_length = 2 * MAX_ENTROPY_CODING_VALUE + 1
    samples = (torch.arange(max_length) - sym_range[:, None]).float()
    dist = torch.distributions.normal.Normal(0.0, table[:, None])
    upper = dist.cdf(samples + 0.5)
    lower = dist.cdf(samples - 0.5)
    pmf = upper - lower
    tail_mass = 2 * lower[:, :1]
    cdf = pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
    return cdf.numpy(), (pmf_length + 2).reshape(-1).int().numpy()


def _bit_estimator_prob(x, h, b, a):
    # src/layers/layers.py:13-19 (accumulated probability of the factorised z model)
    for i in range(4):
        x = x * F.softplus(h[:, :, i:i + 1, None]) + b[:, :, i:i + 1, None]
        if i != 3:
            x = x + torch.tanh(x) * torch.tanh(a[:, :, i:i + 1, None])
    return torch.sigmoid(x)


@torch.inference_mode()
def bit_estimator_cdf_tables(h: torch.Tensor, b: torch.Tensor, a: torch.Tensor):
    """(quantized_cdf int32 [qp*ch, 19], cdf_length int32 [qp*ch]) — BitEstimator.update.
    h, b: [qp, ch, 4]; a: [qp, ch, 3] (float32)."""
    h, b, a = h.float().cpu(), b.float().cpu(), a.float().cpu()
    qp_num, channel = h.shape[0], h.shape[1]
    zeros = torch.zeros((qp_num, channel, 1, 1))

    def fwd(x):
        return _bit_estimator_prob(x, h, b, a)

    sym_range = zeros + MAX_ENTROPY_CODING_VALUE
    for i in range(MAX_ENTROPY_CODING_VALUE, 1, -1):
        neg = fwd(zeros - i)
        pos = fwd(zeros + i)
        sym_range = torch.where(torch.logical_and(neg < 0.001, pos > 0.999), i, sym_range)
    sym_range = sym_range.int()
    pmf_length = sym_range * 2 + 1
    max_length = MAX_ENTROPY_CODING_VALUE * 2 + 1
    samples = torch.arange(max_length)[None, None, None, :] - sym_range
    lower = fwd(samples - 0.5)
    upper = fwd(samples + 0.5)
    pmf = (upper - lower)[:, :, 0, :]
    upper_r = fwd(sym_range.float())
    tail_mass = lower[:, :, 0, :1] + (1.0 - upper_r[:, :, 0, -1:])
    pmf = pmf.reshape([-1, max_length])
    tail_mass = tail_mass.reshape([-1, 1])
    pmf_length = pmf_length.reshape([-1])
    cdf = pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
    return cdf.numpy(), (pmf_length + 2).reshape(-1).int().numpy()
