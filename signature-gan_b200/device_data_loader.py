"""Device-resident training-data loader — B200 counterpart of the reference's `create_data_loader`
(src/data_loader_signatures.py:249-321: `SignatureDataset` + `get_train_transforms` in a 4-worker DataLoader).

The reference decodes and augments every image on the CPU, per sample, with PIL: ~10^3-10^4 images/s, against a
training step that consumes > 3x10^5 images/s. Here the resized 8-bit grayscale images live in
HBM once ((N, S, S) uint8: 4 KB per 64x64 image), every batch is drawn by a device permutation, and the
augmentation of `get_train_transforms` (:154-219) — RandomRotation(±deg, fill=255), RandomAffine(scale, fill=255),
optional horizontal flip, ToTensor, Normalize(0.5, 0.5) — runs in ONE kernel (libsiggan `sg_augment_batch`) that is
bit-identical to torchvision + Pillow for the same sampled parameters. Per-image (angle, scale, flip) are sampled on
the host (numpy Generator), turned into Pillow's fixed-point tables by `sg_augment_params`, and uploaded (56 bytes
per image). No CPU or torch fallback: batches come out of the CUDA kernel or the loader raises.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Iterator, Optional, Sequence, Tuple, Union

import numpy as np
import torch

import _siggan_lib as L

SUPPORTED_EXTENSIONS = (".png", ".jpg", ".jpeg", ".bmp", ".tiff")        # data_loader_signatures.py:38


class DeviceSignatureLoader:
    """Iterates (batch, 1, S, S) float32 CUDA tensors in [-1, 1] — what `create_data_loader(...)` yields.

    images: (N, S, S) or (N, 1, S, S) uint8 (already resized, white background = 255), any device; S in {64, 128}.
    Argument names follow `create_data_loader` / `get_train_transforms`.
    """

    def __init__(self, images: torch.Tensor, batch_size: int = 64, augment: bool = True, rotation_degrees: float = 5.0,
                 scale_range: Tuple[float, float] = (0.9, 1.1), horizontal_flip: bool = False, shuffle: bool = True,
                 drop_last: bool = True, seed: int = 0, device: Union[str, torch.device] = "cuda") -> None:
        if images.dtype != torch.uint8:
            raise TypeError("DeviceSignatureLoader keeps the dataset as uint8 (the reference's 8-bit grayscale images)")
        if images.dim() == 4:
            images = images[:, 0]
        if images.dim() != 3 or images.shape[1] != images.shape[2] or images.shape[1] not in (64, 128):
            raise ValueError(f"images must be (N, S, S) with S in {{64, 128}}, got {tuple(images.shape)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("siggan_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.lib = L.load_library()
        self.pool = images.to(self.device).contiguous()
        self.image_size = int(images.shape[1])
        self.batch_size, self.augment = int(batch_size), bool(augment)
        self.rotation_degrees, self.scale_range = float(rotation_degrees), (float(scale_range[0]), float(scale_range[1]))
        self.horizontal_flip, self.shuffle, self.drop_last = bool(horizontal_flip), bool(shuffle), bool(drop_last)
        self.rng = np.random.default_rng(seed)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)

    # -- reference DataLoader surface ---------------------------------------------------------------
    def __len__(self) -> int:
        n = self.pool.shape[0]
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    @property
    def dataset(self) -> torch.Tensor:
        return self.pool

    def __iter__(self) -> Iterator[torch.Tensor]:
        n = self.pool.shape[0]
        order = (torch.randperm(n, device=self.device, generator=self._gen) if self.shuffle
                 else torch.arange(n, device=self.device)).to(torch.int32)
        nb = len(self)
        if nb == 0:
            return
        sizes = [min(self.batch_size, n - b * self.batch_size) for b in range(nb)]
        # the host side of batch b+1 (parameter sampling + sg_augment_params, which releases the GIL) runs on a helper
        # thread while the consumer trains on batch b
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(self._host_tables, sizes[0])
            for b in range(nb):
                tables = pending.result()
                if b + 1 < nb:
                    pending = pool.submit(self._host_tables, sizes[b + 1])
                idx = order[b * self.batch_size:b * self.batch_size + sizes[b]]
                yield self._launch(idx, *tables)

    # -- one batch ------------------------------------------------------------------------------------
    def sample_params(self, n: int) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
        """(angles in degrees, scales, flips) as `get_train_transforms` samples them: uniform(-deg, deg),
        uniform(scale_range), Bernoulli(0.5) (RandomRotation.get_params, RandomAffine.get_params, RandomHorizontalFlip)."""
        if not self.augment:
            return np.zeros(n), np.ones(n), None
        d = self.rotation_degrees
        angles = self.rng.uniform(-d, d, n) if d > 0 else np.zeros(n)
        lo, hi = self.scale_range
        scales = self.rng.uniform(lo, hi, n) if (lo, hi) != (1.0, 1.0) else np.ones(n)
        flips = (self.rng.random(n) < 0.5).astype(np.uint8) if self.horizontal_flip else None
        return angles, scales, flips

    def batch(self, index: torch.Tensor, angles: Optional[Sequence[float]] = None,
              scales: Optional[Sequence[float]] = None, flips: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Augmented batch of the pool images `index` (int32 CUDA tensor). angles / scales / flips override the
        sampled parameters (parity tests inject the values the reference's transforms drew)."""
        n = int(index.numel())
        return self._launch(index, *self._host_tables(n, angles, scales, flips))

    def _host_tables(self, n: int, angles=None, scales=None, flips=None) -> Tuple[torch.Tensor, int, int, bool]:
        """Pinned staging block [rot_fixed int32 n*6 | scale_affine float64 n*4 | flip uint8 n] for one batch."""
        if angles is None and scales is None and flips is None:
            angles, scales, flips = self.sample_params(n)
        angles = np.ascontiguousarray(np.zeros(n) if angles is None else angles, dtype=np.float64)
        scales = np.ascontiguousarray(np.ones(n) if scales is None else scales, dtype=np.float64)
        if angles.shape != (n,) or scales.shape != (n,):
            raise ValueError("angles / scales must have one entry per image")
        rot_bytes, sc_bytes = n * 24, n * 32
        host = torch.empty(rot_bytes + sc_bytes + n, dtype=torch.uint8).pin_memory()
        base = host.data_ptr()
        L.check(self.lib.sg_augment_params(angles.ctypes.data, scales.ctypes.data, n, self.image_size, base,
                                           base + rot_bytes), "sg_augment_params")
        if flips is not None:
            host[rot_bytes + sc_bytes:] = torch.as_tensor(np.ascontiguousarray(flips, dtype=np.uint8))
        return host, rot_bytes, sc_bytes, flips is not None

    def _launch(self, index: torch.Tensor, host: torch.Tensor, rot_bytes: int, sc_bytes: int, has_flip: bool) -> torch.Tensor:
        n, S = int(index.numel()), self.image_size
        dev = host.to(self.device, non_blocking=True)     # one H2D copy: 56 (57) bytes per image
        out = torch.empty(n, 1, S, S, dtype=torch.float32, device=self.device)
        index = index.to(device=self.device, dtype=torch.int32).contiguous()
        d0 = dev.data_ptr()
        L.check(self.lib.sg_augment_batch(L.ptr(self.pool), L.ptr(index), d0, d0 + rot_bytes,
                                          d0 + rot_bytes + sc_bytes if has_flip else None, n, S, L.ptr(out),
                                          L.current_stream(self.device)), "sg_augment_batch")
        # `dev` and `host` may be released now: both allocators are stream-ordered (the caching host allocator tracks the
        # pending non-blocking copy, the device block is only reused by work enqueued after the kernel on this stream)
        return out

    # -- construction from a directory, like create_data_loader(data_dir, ...) ---------------------------
    @staticmethod
    def load_directory_uint8(data_dir: Union[str, Path], image_size: int = 64,
                             extensions: Tuple[str, ...] = SUPPORTED_EXTENSIONS) -> torch.Tensor:
        """(N, S, S) uint8 pool from the image files directly inside `data_dir` (same discovery rule as
        SignatureDataset._collect_image_paths, data_loader_signatures.py:88-103: listed extensions in lower and upper
        case, sorted). Each file goes once through what SignatureDataset.__getitem__ + the first transform do per
        access: Image.open(...).convert('L') and transforms.Resize((S, S)) (= PIL bilinear resize)."""
        from PIL import Image
        root = Path(data_dir)
        if not root.exists():
            raise ValueError(f"Directory does not exist: {data_dir}")
        paths = set()
        for ext in extensions:
            paths.update(root.glob(f"*{ext}"))
            paths.update(root.glob(f"*{ext.upper()}"))
        paths = sorted(paths)
        if not paths:
            raise ValueError(f"No images found in {data_dir}")
        imgs = [np.asarray(Image.open(p).convert("L").resize((image_size, image_size), Image.BILINEAR)) for p in paths]
        return torch.from_numpy(np.stack(imgs))

    @classmethod
    def from_directory(cls, data_dir: Union[str, Path], image_size: int = 64, **kwargs) -> "DeviceSignatureLoader":
        """Decode + grayscale + resize every image ONCE on the host, then keep the uint8 pool in HBM."""
        return cls(cls.load_directory_uint8(data_dir, image_size), **kwargs)
