"""CPU oracle for the training-time augmentation of the input pipeline (TEST INFRASTRUCTURE — never imported by the
product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).

Reference: src/data_loader_signatures.py:154-219 `get_train_transforms` — on an already resized 8-bit grayscale PIL image
    RandomRotation(degrees=5, fill=255)  -> RandomAffine(degrees=0, scale=(0.9, 1.1), fill=255)
    [-> RandomHorizontalFlip]  -> ToTensor  -> Normalize(0.5, 0.5)
The arithmetic lives in third-party dependencies that ARE installed in this image (requirements.txt: torchvision,
Pillow; here torchvision 0.26 / Pillow 12.2): torchvision.transforms.functional.rotate / affine call
PIL.Image.rotate / Image.transform(AFFINE, NEAREST), whose C code (libImaging/Geometry.c) resamples by
nearest neighbour in 16.16 fixed point (`affine_fixed`) for a general matrix and with a running double-precision
coordinate (`ImagingScaleAffine`) for a pure scale. Restated below in numpy integer / float64 arithmetic.

Pinning: tests/test_augment_oracle.py compares this file bit for bit with torchvision + Pillow themselves on random
images, angles and scales (both are importable wherever the tests run), and with tests/golden/augment_64.pt, produced
by the reference's own `get_train_transforms` pipeline (tests/golden/make_golden_augment.py).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

FILL = 255  # data_loader_signatures.py:184,194 — white background


def _fix(v: float) -> int:
    """Geometry.c FIX(v) = FLOOR(v * 65536 + 0.5), FLOOR truncating for v >= 0 and flooring below."""
    t = v * 65536.0 + 0.5
    return int(math.floor(t)) if t < 0.0 else int(t)


def rotation_matrix(angle_deg: float, size: int) -> Tuple[float, ...]:
    """PIL.Image.rotate's inverse matrix for expand=0, center=None (Image.py `rotate`)."""
    angle = angle_deg % 360.0
    cx = cy = size / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return tuple(m)


def rotation_fixed(angle_deg: float, size: int) -> Tuple[int, int, int, int, int, int]:
    """16.16 fixed-point coefficients (a0, a1, a2, a3, a4, a5) of `affine_fixed`; a2 / a5 carry the half-pixel offset.
    Angles PIL short-cuts (0 -> copy, 180 -> transpose, 90/270 on a square image -> transpose) map to the identical
    exact permutation here."""
    ang = angle_deg % 360.0
    one = 65536
    if ang == 0:
        return (one, 0, one // 2, 0, one, one // 2)
    if ang in (90.0, 180.0, 270.0):
        # exact transposes: source index = permutation of the output index; written as fixed-point coefficients that
        # reproduce it for every pixel of a size x size image
        if ang == 180.0:
            return (-one, 0, size * one - one // 2, 0, -one, size * one - one // 2)
        if ang == 90.0:    # ROTATE_90 (counter-clockwise): out[y][x] = in[x][size-1-y]
            return (0, -one, size * one - one // 2, one, 0, one // 2)
        return (0, one, one // 2, -one, 0, size * one - one // 2)   # 270
    m = rotation_matrix(angle_deg, size)
    return (_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def rotate_nearest(img: np.ndarray, fixed: Sequence[int], fill: int = FILL) -> np.ndarray:
    """`affine_fixed`: xx = a2 + y*a1 + x*a0, yy = a5 + y*a4 + x*a3 (int32), source pixel (yy >> 16, xx >> 16)."""
    a0, a1, a2, a3, a4, a5 = (int(v) for v in fixed)
    h, w = img.shape
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + y * a1 + x * a0) >> 16
    yin = (a5 + y * a4 + x * a3) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.full_like(img, fill)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def scale_params(scale: float, size: int) -> Tuple[float, float, float, float]:
    """torchvision F.affine(angle=0, translate=(0,0), scale, shear=(0,0)) on a PIL image: center = (w/2, h/2),
    `_get_inverse_affine_matrix`; returns what `ImagingScaleAffine` starts from: (a0, xo, a4, yo) with
    xo = a[2] + a[0]*0.5, yo = a[5] + a[4]*0.5."""
    cx = cy = size * 0.5
    rot = math.radians(0.0)
    sx = sy = math.radians(0.0)
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [v / scale for v in m]
    m[2] += m[0] * (-cx) + m[1] * (-cy)
    m[5] += m[3] * (-cx) + m[4] * (-cy)
    m[2] += cx
    m[5] += cy
    assert m[1] == 0 and m[3] == 0      # pure scale -> the ImagingScaleAffine path
    return (m[0], m[2] + m[0] * 0.5, m[4], m[5] + m[4] * 0.5)


def _coord_table(a: float, o: float, n_out: int, n_in: int) -> np.ndarray:
    """COORD(v) = v < 0 ? -1 : (int)v of a coordinate accumulated by repeated `o += a` in double; -1 = outside."""
    tab = np.empty(n_out, dtype=np.int64)
    for i in range(n_out):
        v = -1 if o < 0.0 else int(o)
        tab[i] = v if 0 <= v < n_in else -1
        o += a
    return tab


def scale_nearest(img: np.ndarray, params: Sequence[float], fill: int = FILL) -> np.ndarray:
    a0, xo, a4, yo = params
    h, w = img.shape
    xt, yt = _coord_table(a0, xo, w, w), _coord_table(a4, yo, h, h)
    out = np.full_like(img, fill)
    ys, xs = np.nonzero((yt >= 0)[:, None] & (xt >= 0)[None, :])
    out[ys, xs] = img[yt[ys], xt[xs]]
    return out


def to_normalised(img_u8: np.ndarray) -> np.ndarray:
    """ToTensor (uint8 -> float32 / 255) then Normalize(0.5, 0.5) (sub, div), both in float32."""
    f = img_u8.astype(np.float32) / np.float32(255.0)
    return (f - np.float32(0.5)) / np.float32(0.5)


def augment(img_u8: np.ndarray, angle: float, scale: float, flip: bool = False) -> np.ndarray:
    """One image through the training pipeline with the sampled (angle, scale, flip). Returns float32 (S, S)."""
    size = img_u8.shape[0]
    r = rotate_nearest(img_u8, rotation_fixed(angle, size))
    s = scale_nearest(r, scale_params(scale, size))     # RandomAffine runs even for scale == 1.0 (identity lookup)
    if flip:
        s = s[:, ::-1]
    return to_normalised(np.ascontiguousarray(s))


def parameter_tables(angles: Sequence[float], scales: Sequence[float], size: int):
    """Per-image tables the CUDA kernel consumes: int32 [B][6] rotation coefficients, float64 [B][4] scale start /
    step — the same numbers this oracle resamples with."""
    rot = np.array([rotation_fixed(float(a), size) for a in angles], dtype=np.int32).reshape(-1, 6)
    sc = np.array([scale_params(float(s), size) for s in scales], dtype=np.float64).reshape(-1, 4)
    return rot, sc


# --------------------------------------------------------------------------------------------------------------
# Ink statistics (src/utils/metrics.py:118-174 calculate_stroke_density / calculate_foreground_ratio), restated in numpy
# --------------------------------------------------------------------------------------------------------------
def ink_counts(images: np.ndarray, threshold: float = 0.5):
    """images (N, 1, H, W) float32. Returns (count_raw, count_rescaled, minimum) per image, the three columns
    sg_ink_stats produces: #(x < t), #((x + 1) / 2 < t) with float32 roundings, min(x)."""
    x = images.astype(np.float32).reshape(images.shape[0], -1)
    t = np.float32(threshold)
    res = (x + np.float32(1.0)) / np.float32(2.0)
    return (x < t).sum(axis=1).astype(np.int32), (res < t).sum(axis=1).astype(np.int32), x.min(axis=1)


def ink_fraction(images: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    raw, res, mn = ink_counts(images, threshold)
    counts = res if mn.min() < 0 else raw                       # `if images.min() < 0: images = (images + 1) / 2`
    return counts.astype(np.float32) / np.float32(images.shape[2] * images.shape[3])


def stroke_density(images: np.ndarray, threshold: float = 0.5):
    d = ink_fraction(images, threshold)
    return {"mean": float(np.mean(d)), "std": float(np.std(d)), "min": float(np.min(d)), "max": float(np.max(d))}


def foreground_ratio(images: np.ndarray, threshold: float = 0.5):
    r = ink_fraction(images, threshold)
    return {"mean": float(np.mean(r)), "std": float(np.std(r)),
            "percentiles": {k: float(np.percentile(r, int(k))) for k in ("25", "50", "75")}}
