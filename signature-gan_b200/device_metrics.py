"""Ink statistics of image batches on the device — B200 counterpart of the reference's
`calculate_stroke_density` / `calculate_foreground_ratio` (src/utils/metrics.py:118-174, used by
evaluate_vanilla_gan_signatures.py:306-332). Same signatures and returned dicts; the per-pixel work (rescale decision,
threshold, per-image mean) is ONE pass of libsiggan `sg_ink_stats` over the images where they already are (HBM), and
only 12 bytes per image travel to the host for the numpy summary the reference computes there too. No CPU fallback.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

import _siggan_lib as L


def _ink_fraction(images: torch.Tensor, threshold: float) -> np.ndarray:
    """float32 per-image fraction of pixels below `threshold`, after the reference's `(images + 1) / 2` rescale when the
    batch minimum is negative (metrics.py:129-137 / 158-165)."""
    if images.device.type != "cuda":
        raise RuntimeError("siggan_b200 device_metrics runs on CUDA only (no CPU path); images are on " + str(images.device))
    if images.dim() != 4 or images.shape[1] != 1:
        raise NotImplementedError("siggan_b200 implements the grayscale (N, 1, H, W) configuration")
    x = images.contiguous().float()
    n, pixels = x.shape[0], x.shape[2] * x.shape[3]
    buf = torch.empty(3, n, dtype=torch.int32, device=x.device)        # count_raw | count_rescaled | minimum (float bits)
    mins = buf[2].view(torch.float32)
    L.check(L.load_library().sg_ink_stats(L.ptr(x), n, pixels, float(threshold), L.ptr(buf[0]), L.ptr(buf[1]), L.ptr(mins),
                                          L.current_stream(x.device)), "sg_ink_stats")
    host = buf.cpu()
    rescale = bool(host[2].view(torch.float32).min() < 0)
    counts = host[1 if rescale else 0].numpy()
    return counts.astype(np.float32) / np.float32(pixels)              # == stroke.view(N, -1).mean(dim=1) in float32


def calculate_stroke_density(images: torch.Tensor, threshold: float = 0.5) -> Dict[str, float]:
    d = _ink_fraction(images, threshold)
    return {"mean": float(np.mean(d)), "std": float(np.std(d)), "min": float(np.min(d)), "max": float(np.max(d))}


def calculate_foreground_ratio(images: torch.Tensor, threshold: float = 0.5) -> Dict[str, float]:
    r = _ink_fraction(images, threshold)
    return {"mean": float(np.mean(r)), "std": float(np.std(r)),
            "percentiles": {"25": float(np.percentile(r, 25)), "50": float(np.percentile(r, 50)),
                            "75": float(np.percentile(r, 75))}}
