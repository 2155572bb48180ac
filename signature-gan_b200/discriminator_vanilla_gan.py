"""Discriminator of the signature Vanilla GAN — B200 drop-in for the reference's src/discriminator_vanilla_gan.py.

Same classes / constructor signatures / attributes / sub-module names (so `state_dict()` keys, shapes and
dtypes match, reference disc…:18-81, 84-283, 285-344, 347-370); `forward` is executed by libsiggan.so
(include/siggan.h: sg_d_forward / sg_d_backward). The torch sub-modules only own the fp32 master
parameters, which are views into one flat CUDA buffer. No CPU or eager fallback.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

import _siggan_lib as L


class DownsampleBlock(nn.Module):
    """Parameter holder for Conv2d(k4,s2,p1)+bias, LeakyReLU, Dropout2d (reference disc…:18-81)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 4, stride: int = 2, padding: int = 1,
                 use_batch_norm: bool = False, use_spectral_norm: bool = False, dropout: float = 0.25,
                 leaky_slope: float = 0.2) -> None:
        super().__init__()
        if use_batch_norm:
            raise NotImplementedError("siggan_b200 covers the blocks the reference Discriminator builds "
                                      "(use_batch_norm is never set by it, disc…:131-194)")
        conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=True)
        if use_spectral_norm:
            # same wrapper as the reference (disc…:61-62): registers weight_orig / weight_u / weight_v with torch's
            # own initialisation, state-dict versioning and load hooks; its forward pre-hook never runs here (the
            # holder is not called) — Discriminator.forward does the power iteration through sg_spectral_norm_weight.
            conv = spectral_norm(conv)
        mods: List[nn.Module] = [conv, nn.LeakyReLU(leaky_slope, inplace=True)]
        if dropout > 0:
            mods.append(nn.Dropout2d(dropout))
        self.block = nn.Sequential(*mods)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("DownsampleBlock is a parameter holder in siggan_b200; run it through Discriminator.forward")


class _SpectralWeights:
    """Effective weights of one forward of the spectral-norm variant: flat copy of the parameters with every
    weight divided by its sigma, plus the (u, v, sigma) that forward used (torch SpectralNorm.compute_weight)."""

    def __init__(self, eff: torch.Tensor, layers: list, sigma: torch.Tensor, scratch: torch.Tensor) -> None:
        self.eff, self.layers, self.sigma, self.scratch = eff, layers, sigma, scratch

    def backward(self, sctx: "L.Context", gflat: torch.Tensor, dev: torch.device) -> None:
        """dL/dweight -> dL/dweight_orig, in place in the flat gradient buffer."""
        for k, (off, rows, cols, u, v) in enumerate(self.layers):
            L.check(sctx.lib.sg_spectral_norm_backward(
                L.ptr(self.eff) + 4 * off, L.ptr(u), L.ptr(v), L.ptr(self.sigma) + 4 * k, rows, cols,
                L.ptr(gflat) + 4 * off, L.ptr(self.scratch), L.current_stream(dev)), "sg_spectral_norm_backward")


class _DiscriminatorFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = sg_d_forward, backward = sg_d_backward."""

    @staticmethod
    def forward(ctx, disc: "Discriminator", save: bool, masks: Optional[torch.Tensor], x: torch.Tensor,
                *params: torch.Tensor) -> torch.Tensor:
        sctx, fp = disc._ctx, disc._flat
        B, dev = x.shape[0], x.device
        prob = torch.empty(B, 1, dtype=torch.float32, device=dev)
        ws = None
        if save:
            ws = torch.empty(int(sctx.lib.sg_d_workspace_bytes(sctx.handle, B)), dtype=torch.uint8, device=dev)
        # spectral-norm variant: the layers run on weight_orig / sigma; what this call's power iteration left in
        # (u, v, sigma) is kept for its own backward (a later forward of the same module iterates again)
        sn = disc._spectral_weights(dev) if disc.use_spectral_norm else None
        weights = sn.eff if sn is not None else fp.flat
        L.check(sctx.lib.sg_d_forward(sctx.handle, L.ptr(weights), L.ptr(x), B, L.ptr(masks), L.ptr(ws), L.ptr(prob),
                                      None, L.current_stream(dev)), "sg_d_forward")
        ctx.disc, ctx.ws, ctx.masks, ctx.x, ctx.B, ctx.sn = disc, ws, masks, x, B, (sn if save else None)
        ctx.x_needs_grad = x.requires_grad
        ctx.w_needs_grad = any(p.requires_grad for p in params)
        return prob

    @staticmethod
    def backward(ctx, grad_prob: torch.Tensor):
        disc: Discriminator = ctx.disc
        sctx, fp = disc._ctx, disc._flat
        if ctx.ws is None:
            raise RuntimeError("Discriminator backward without saved activations")
        grad_prob = grad_prob.contiguous().float()
        dev = grad_prob.device
        gflat = fp.fresh_grad() if ctx.w_needs_grad else None
        dx = torch.empty_like(ctx.x) if ctx.x_needs_grad else None
        sn = ctx.sn
        weights = sn.eff if sn is not None else fp.flat
        L.check(sctx.lib.sg_d_backward(sctx.handle, L.ptr(weights), L.ptr(ctx.x), L.ptr(ctx.ws), L.ptr(ctx.masks),
                                       L.ptr(grad_prob), ctx.B, L.ptr(gflat), L.ptr(dx), L.current_stream(dev)),
                "sg_d_backward")
        if sn is not None and gflat is not None:
            sn.backward(sctx, gflat, dev)
        ctx.ws = ctx.sn = None
        grads = fp.grad_views(gflat) if gflat is not None else [None] * len(fp.params)
        return (None, None, None, dx, *grads)


class Discriminator(nn.Module):
    """image (B, 1, S, S) -> P(real) (B, 1); S in {64, 128} (reference disc…:84-283).

    4/5 x DownsampleBlock (1->64->128->256->512[->512]) -> Flatten (NCHW order) -> Linear(8192, 1) -> Sigmoid.
    Dropout2d is active only in training mode: one Bernoulli(1-p) draw per (sample, channel), scaled by 1/(1-p).
    """

    def __init__(self, input_size: int = 64, input_channels: int = 1, use_spectral_norm: bool = False,
                 dropout: float = 0.25, leaky_slope: float = 0.2, base_features: int = 64) -> None:
        super().__init__()
        if input_size not in [64, 128]:
            raise ValueError(f"input_size must be 64 or 128, got {input_size}")
        # base_features (an extension, last in the signature; the reference's ladder starts at 64, disc…:131-194): 128
        # selects the "2x hidden width" variant of the width / resolution sweep (bf16 mode, no spectral norm)
        if base_features not in (64, 128):
            raise ValueError(f"base_features must be 64 or 128, got {base_features}")
        if base_features == 128 and use_spectral_norm:
            raise NotImplementedError("the 2x-width Discriminator is built without spectral norm")
        self.base_features = base_features
        self._width_mult = base_features // 64
        self.input_size = input_size
        self.input_channels = input_channels
        self.use_spectral_norm = use_spectral_norm
        self.dropout = dropout
        self.leaky_slope = leaky_slope
        ladder = [input_channels] + [c * self._width_mult
                                     for c in ([64, 128, 256, 512] if input_size == 64 else [64, 128, 256, 512, 512])]
        self.conv_blocks = nn.Sequential(*[
            DownsampleBlock(a, b, use_spectral_norm=use_spectral_norm, dropout=dropout, leaky_slope=leaky_slope)
            for a, b in zip(ladder[:-1], ladder[1:])])
        self.flatten = nn.Flatten()
        fc = nn.Linear(ladder[-1] * 4 * 4, 1)
        if use_spectral_norm:
            fc = spectral_norm(fc)                  # disc…:199-202
        self.classifier = nn.Sequential(fc, nn.Sigmoid())
        self.apply(self._init_weights)
        self._flat = L.FlatParams(self, L.SG_NET_D)
        self._ctx: Optional[L.Context] = None
        self._precision = L.precision_from_env()
        #: parity hook: a list of (B, C_i) keep-scale tensors (values {0, 1/(1-p)}) used instead of the RNG
        self.mask_override: Optional[List[torch.Tensor]] = None

    def _init_weights(self, module: nn.Module) -> None:
        """DCGAN initialisation (reference disc…:212-239)."""
        if isinstance(module, (nn.Conv2d, nn.Linear)):
            nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.BatchNorm2d):
            nn.init.normal_(module.weight, mean=1.0, std=0.02)
            nn.init.zeros_(module.bias)

    # -- plumbing ---------------------------------------------------------------------------------
    def _prepare(self, device: torch.device) -> None:
        if self.input_channels != 1:
            raise NotImplementedError("siggan_b200 implements the grayscale (input_channels=1) configuration")
        self._ctx = L.Context.get(device, self.input_size, getattr(self, "_latent_hint", 100), self._precision,
                                  self.leaky_slope, width_mult=self._width_mult)
        self._flat.sync(self._ctx)

    def _sn_modules(self) -> List[nn.Module]:
        return [blk.block[0] for blk in self.conv_blocks] + [self.classifier[0]]

    def _spectral_weights(self, device: torch.device, classifier: bool = True) -> _SpectralWeights:
        """One SpectralNorm.compute_weight per wrapped layer, as the reference's forward pre-hooks do on every call
        (disc…:61-62, 201-202): a power iteration that updates weight_u / weight_v in place when the module is in
        training mode, none in eval mode; then weight = weight_orig / sigma."""
        sctx, fp = self._ctx, self._flat
        mods = self._sn_modules()
        eff = fp.flat.clone()                       # biases travel unchanged
        sigma = torch.empty(len(mods), dtype=torch.float32, device=device)
        scratch = torch.empty(1024, dtype=torch.float32, device=device)
        by_name = {n: k for k, n in enumerate(fp._names)}
        layers = []
        names = [f"conv_blocks.{i}.block.0" for i in range(len(self.conv_blocks))] + ["classifier.0"]
        if not classifier:                          # forward_features never calls the classifier (disc…:262-274)
            names, mods = names[:-1], mods[:-1]
        for k, (name, m) in enumerate(zip(names, mods)):
            off, n, shape = fp.layout[by_name[name + ".weight_orig"]]
            rows, cols = shape[0], n // shape[0]
            hook = next(h for h in m._forward_pre_hooks.values() if hasattr(h, "n_power_iterations"))
            u, v = m.weight_u, m.weight_v
            if u.device != device or v.device != device or not u.is_contiguous() or not v.is_contiguous():
                raise RuntimeError("spectral-norm buffers must live on the module's CUDA device")
            iters = hook.n_power_iterations if self.training else 0
            L.check(sctx.lib.sg_spectral_norm_weight(
                L.ptr(fp.flat) + 4 * off, L.ptr(u), L.ptr(v), rows, cols, iters, float(hook.eps),
                L.ptr(eff) + 4 * off, L.ptr(sigma) + 4 * k, L.ptr(scratch), L.current_stream(device)),
                "sg_spectral_norm_weight")
            # torch clones u / v after iterating so that a backward of this forward is not disturbed by the next one
            layers.append((off, rows, cols, u.clone() if iters else u, v.clone() if iters else v))
        return _SpectralWeights(eff, layers, sigma, scratch)

    def set_precision(self, precision: str) -> "Discriminator":
        self._precision = {"bf16": L.SG_PREC_BF16, "fp32": L.SG_PREC_FP32}[precision]
        return self

    def _masks(self, B: int, device: torch.device) -> Optional[torch.Tensor]:
        if not self.training or self.dropout <= 0:
            return None
        sctx = self._ctx
        n = int(sctx.lib.sg_d_mask_count(sctx.handle, B))
        if self.mask_override is not None:
            flat = torch.cat([m.to(device=device, dtype=torch.float32).reshape(-1) for m in self.mask_override])
            if flat.numel() != n:
                raise RuntimeError(f"mask_override has {flat.numel()} elements, expected {n}")
            return flat.contiguous()
        masks = torch.empty(n, dtype=torch.float32, device=device)
        L.check(sctx.lib.sg_dropout_masks(sctx.handle, L.DROPOUT.seed(), L.DROPOUT.advance(n), B,
                                          float(self.dropout), L.ptr(masks), L.current_stream(device)), "sg_dropout_masks")
        return masks

    def _check_input(self, x: torch.Tensor) -> torch.device:
        dev = self.classifier[0].bias.device
        if dev.type != "cuda" or x.device.type != "cuda":
            raise RuntimeError(f"siggan_b200 Discriminator runs on CUDA only (no CPU path); module on {dev}, input on {x.device}")
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != self.input_size or x.shape[3] != self.input_size:
            raise RuntimeError(f"Discriminator expects (batch, 1, {self.input_size}, {self.input_size}), got {tuple(x.shape)}")
        return dev

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = self._check_input(x)
        self._prepare(dev)
        x = x.contiguous().float()
        params = self._flat.params
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _DiscriminatorFn.apply(self, save, self._masks(x.shape[0], dev), x, *params)

    @torch.no_grad()
    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """(B, 512*4*4) features before the classifier, flattened in the reference's NCHW order (disc…:262-274)."""
        dev = self._check_input(x)
        self._prepare(dev)
        x = x.contiguous().float()
        sctx, fp = self._ctx, self._flat
        B = x.shape[0]
        feat = torch.empty(B, int(sctx.lib.sg_d_feature_count(sctx.handle)), dtype=torch.float32, device=dev)
        weights = self._spectral_weights(dev, classifier=False).eff if self.use_spectral_norm else fp.flat
        L.check(sctx.lib.sg_d_forward(sctx.handle, L.ptr(weights), L.ptr(x), B, L.ptr(self._masks(B, dev)), None, None,
                                      L.ptr(feat), L.current_stream(dev)), "sg_d_forward")
        return feat

    # -- reference API ----------------------------------------------------------------------------
    def get_num_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_input_shape(self) -> Tuple[int, int, int]:
        return (self.input_channels, self.input_size, self.input_size)


class MinibatchDiscrimination(nn.Module):
    """Minibatch-discrimination layer (reference disc…:285-344). Defined for API parity; the reference never
    instantiates it, so it is off the accelerated path and evaluated with plain tensor ops."""

    def __init__(self, in_features: int, out_features: int, kernel_dims: int = 5) -> None:
        super().__init__()
        self.in_features, self.out_features, self.kernel_dims = in_features, out_features, kernel_dims
        self.T = nn.Parameter(torch.randn(in_features, out_features, kernel_dims) * 0.02)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = torch.einsum("bi,iok->bok", x, self.T)
        l1 = (m.unsqueeze(0) - m.unsqueeze(1)).abs().sum(dim=3)
        return torch.cat([x, torch.exp(-l1).sum(dim=1)], dim=1)


def create_discriminator(input_size: int = 64, input_channels: int = 1, use_spectral_norm: bool = False,
                         dropout: float = 0.25) -> Discriminator:
    return Discriminator(input_size=input_size, input_channels=input_channels, use_spectral_norm=use_spectral_norm,
                         dropout=dropout)
