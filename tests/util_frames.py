"""Synthetic test frames (no datasets offline): band-limited noise, fp16-representable, in [-0.5, 0.5]
(SURVEY.md §8d recipe)."""
import numpy as np
import torch


def synth_frame(h, w, seed, channels=3):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.random((1, channels, h + 4, w + 4)).astype(np.float32))
    x = torch.nn.functional.avg_pool2d(x, 5, 1)
    x = (x - x.mean()) / x.std() * 0.18
    return x.clamp(-0.5, 0.5).half().float()


def psnr(a, b):
    mse = torch.mean((a.float() - b.float()) ** 2).item()
    return 10 * np.log10(1.0 / max(mse, 1e-12))
