"""Generator of the signature Vanilla GAN — B200 drop-in for the reference's src/generator_vanilla_gan.py.

Same class names, constructor signatures, attributes, sub-module names (hence identical `state_dict()`
keys, shapes and dtypes) and methods as the reference (gen…:17-66, 69-237, 240-260), but `forward`
does not run the torch sub-modules: they only hold the fp32 master parameters, which live as views
into one flat CUDA buffer, and the arithmetic is done by libsiggan.so (hand-written sm_100a kernels,
include/siggan.h: sg_g_forward / sg_g_backward). There is no CPU or eager fallback: calling a module
that is not on a CUDA device raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

import _siggan_lib as L


class UpsampleBlock(nn.Module):
    """Parameter holder for ConvTranspose2d(k4,s2,p1) [+ BatchNorm2d] + ReLU (reference gen…:17-66)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 4, stride: int = 2, padding: int = 1,
                 output_padding: int = 0, use_batch_norm: bool = True) -> None:
        super().__init__()
        mods = [nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                   padding=padding, output_padding=output_padding, bias=not use_batch_norm)]
        if use_batch_norm:
            mods.append(nn.BatchNorm2d(out_channels))
        mods.append(nn.ReLU(inplace=True))
        self.block = nn.Sequential(*mods)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("UpsampleBlock is a parameter holder in siggan_b200; run it through Generator.forward")


class _GeneratorFn(torch.autograd.Function):
    """One autograd node for the whole network: forward = sg_g_forward, backward = sg_g_backward."""

    @staticmethod
    def forward(ctx, gen: "Generator", save: bool, z: torch.Tensor, *params: torch.Tensor) -> torch.Tensor:
        sctx, fp = gen._ctx, gen._flat
        B = z.shape[0]
        dev = z.device
        out = torch.empty(B, 1, gen.output_size, gen.output_size, dtype=torch.float32, device=dev)
        ws = None
        if save:
            ws = torch.empty(int(sctx.lib.sg_g_workspace_bytes(sctx.handle, B)), dtype=torch.uint8, device=dev)
        train = 1 if gen.training else 0
        L.check(sctx.lib.sg_g_forward(sctx.handle, L.ptr(fp.flat), L.ptr(fp.stats), L.ptr(z), B, train, L.ptr(ws),
                                      L.ptr(out), None, L.current_stream(dev)), "sg_g_forward")
        ctx.gen, ctx.ws, ctx.B, ctx.train = gen, ws, B, train
        ctx.z_needs_grad = z.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        gen: Generator = ctx.gen
        sctx, fp = gen._ctx, gen._flat
        if ctx.ws is None:
            raise RuntimeError("Generator backward without saved activations")
        grad_out = grad_out.contiguous().float()
        gflat = fp.fresh_grad()
        dz = torch.empty(ctx.B, gen.latent_dim, dtype=torch.float32, device=grad_out.device) if ctx.z_needs_grad else None
        L.check(sctx.lib.sg_g_backward(sctx.handle, L.ptr(fp.flat), L.ptr(ctx.ws), L.ptr(grad_out), ctx.B, ctx.train,
                                       L.ptr(gflat), L.ptr(dz), L.current_stream(grad_out.device)), "sg_g_backward")
        ctx.ws = None
        return (None, None, dz, *fp.grad_views(gflat))


class Generator(nn.Module):
    """z (B, latent_dim) -> image (B, 1, S, S) in [-1, 1]; S in {64, 128} (reference gen…:69-237).

    fc: Linear(latent, C0*16) + BatchNorm1d + ReLU -> view (B, C0, 4, 4) -> 4/5 x UpsampleBlock ->
    Conv2d(32, 1, 3, padding=1) + Tanh. Training mode uses batch statistics and updates the running
    statistics (momentum 0.1, unbiased variance), eval mode uses the running statistics.
    """

    def __init__(self, latent_dim: int = 100, output_size: int = 64, output_channels: int = 1,
                 base_features: int = 256) -> None:
        super().__init__()
        if output_size not in [64, 128]:
            raise ValueError(f"output_size must be 64 or 128, got {output_size}")
        self.latent_dim = latent_dim
        self.output_size = output_size
        self.output_channels = output_channels
        self.base_features = base_features
        self.init_size = 4
        self.init_channels = base_features if output_size == 64 else base_features * 2
        feat = self.init_channels * self.init_size * self.init_size
        self.fc = nn.Sequential(nn.Linear(latent_dim, feat), nn.BatchNorm1d(feat), nn.ReLU(inplace=True))
        # base_features is inert in the reference (any value but 256 fails in its first ConvTranspose2d, whose widths are
        # literals, gen…:131-149). Here 512 selects the "2x hidden width" variant of the width / resolution sweep: the same
        # blocks with every channel count doubled (bf16 mode only); other values keep the reference's ladder and fail as
        # the reference does, at the first use.
        self._width_mult = 2 if base_features == 512 else 1
        ladder = [c * self._width_mult for c in ([256, 128, 64, 32, 32] if output_size == 64 else [512, 256, 128, 64, 32, 32])]
        self.upsample_blocks = nn.Sequential(*[UpsampleBlock(a, b) for a, b in zip(ladder[:-1], ladder[1:])])
        self.final_conv = nn.Sequential(nn.Conv2d(ladder[-1], output_channels, kernel_size=3, stride=1, padding=1,
                                                  bias=True), nn.Tanh())
        self.apply(self._init_weights)
        self._flat = L.FlatParams(self, L.SG_NET_G)
        self._ctx: Optional[L.Context] = None
        self._precision = L.precision_from_env()
        self._act_slope = 0.0      # ReLU; ablation_generator.ConfigurableGenerator sets the LeakyReLU slope

    def _init_weights(self, module: nn.Module) -> None:
        """DCGAN initialisation (reference gen…:168-187): N(0, 0.02) weights, zero biases, BN gamma ~ N(1, 0.02)."""
        if isinstance(module, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, (nn.BatchNorm1d, nn.BatchNorm2d)):
            nn.init.normal_(module.weight, mean=1.0, std=0.02)
            nn.init.zeros_(module.bias)

    # -- plumbing ---------------------------------------------------------------------------------
    def _bn_modules(self):
        return [self.fc[1]] + [blk.block[1] for blk in self.upsample_blocks]

    def _prepare(self, device: torch.device) -> None:
        if self.output_channels != 1:
            raise NotImplementedError("siggan_b200 implements the grayscale (output_channels=1) configuration the "
                                      "reference trains and serves")
        if self.init_channels != (256 if self.output_size == 64 else 512) * self._width_mult:
            raise RuntimeError("base_features other than 256 is not a working configuration of the reference either "
                               "(its first ConvTranspose2d is hard-wired to 256/512 input channels)")
        bns = self._bn_modules()
        self._ctx = L.Context.get(device, self.output_size, self.latent_dim, self._precision, 0.2, bns[0].eps,
                                  bns[0].momentum, self._act_slope, self._width_mult)
        self._flat.sync(self._ctx, bns)

    def set_precision(self, precision: str) -> "Generator":
        """'bf16' (tensor-core path) or 'fp32' (validation mode)."""
        self._precision = {"bf16": L.SG_PREC_BF16, "fp32": L.SG_PREC_FP32}[precision]
        return self

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        if z.dim() != 2 or z.shape[1] != self.latent_dim:
            raise RuntimeError(f"Generator expects latent vectors of shape (batch, {self.latent_dim}), got {tuple(z.shape)}")
        dev = self.fc[0].weight.device
        if dev.type != "cuda" or z.device.type != "cuda":
            raise RuntimeError("siggan_b200 Generator runs on CUDA only (no CPU path); module on "
                               f"{dev}, input on {z.device}")
        if self.training and z.shape[0] < 2:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
        self._prepare(dev)
        z = z.contiguous().float()
        params = self._flat.params
        save = torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in params))
        out = _GeneratorFn.apply(self, save, z, *params)
        if self.training:
            torch._foreach_add_([bn.num_batches_tracked for bn in self._bn_modules()], 1)
        return out

    @torch.no_grad()
    def sample_uint8(self, z: torch.Tensor) -> torch.Tensor:
        """Sampling egress: returns the uint8 images ((x+1)*127.5 clipped, utils/inference.py:129) computed in the
        final kernel's epilogue, shape (B, 1, S, S)."""
        dev = self.fc[0].weight.device
        self._prepare(dev)
        z = z.contiguous().float()
        B = z.shape[0]
        out = torch.empty(B, 1, self.output_size, self.output_size, dtype=torch.uint8, device=dev)
        sctx, fp = self._ctx, self._flat
        L.check(sctx.lib.sg_g_forward(sctx.handle, L.ptr(fp.flat), L.ptr(fp.stats), L.ptr(z), B,
                                      1 if self.training else 0, None, None, L.ptr(out), L.current_stream(dev)),
                "sg_g_forward")
        return out

    @torch.no_grad()
    def sample_uint8_to_host(self, n_samples: int, batch: int = 16384, latents: Optional[torch.Tensor] = None,
                             generator: Optional[torch.Generator] = None,
                             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Bulk sampling egress (the batched loop of utils/inference.py:106-134 + its uint8 conversion :129): returns a
        pinned host tensor (n_samples, 1, S, S) uint8. Chunks of `batch` latents go through the generator with the uint8
        conversion fused into the last kernel; the device->host copy of chunk i runs on a second stream while chunk
        i+1 is computed (two device buffers), so the PCIe transfer is hidden behind the generator instead of added to
        it. `latents` (n_samples, latent_dim; host or device) are used if given, else drawn on the device. `out`: an
        optional pinned uint8 host tensor to fill (page-locking hundreds of MB costs more than generating them)."""
        dev = self.fc[0].weight.device
        self._prepare(dev)
        S = self.output_size
        if out is None:
            out = torch.empty(n_samples, 1, S, S, dtype=torch.uint8, pin_memory=True)
        elif out.shape != (n_samples, 1, S, S) or out.dtype != torch.uint8 or out.device.type != "cpu":
            raise ValueError("out must be a uint8 host tensor of shape (n_samples, 1, S, S)")
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        bufs = [torch.empty(min(batch, n_samples), 1, S, S, dtype=torch.uint8, device=dev) for _ in range(2)]
        copied = [None, None]   # event: the copy out of bufs[k] has finished
        sctx, fp = self._ctx, self._flat
        for i, b0 in enumerate(range(0, n_samples, batch)):
            n = min(batch, n_samples - b0)
            if latents is not None:
                z = latents[b0:b0 + n].to(dev, non_blocking=True).contiguous().float()
            else:
                z = torch.randn(n, self.latent_dim, device=dev, generator=generator)
            k = i % 2
            if copied[k] is not None:
                compute.wait_event(copied[k])
            L.check(sctx.lib.sg_g_forward(sctx.handle, L.ptr(fp.flat), L.ptr(fp.stats), L.ptr(z), n,
                                          1 if self.training else 0, None, None, L.ptr(bufs[k]), L.current_stream(dev)),
                    "sg_g_forward")
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                out[b0:b0 + n].copy_(bufs[k][:n], non_blocking=True)
                copied[k] = torch.cuda.Event()
                copied[k].record(copy)
            z.record_stream(compute)
        copy.synchronize()
        compute.synchronize()
        return out

    # -- reference API ----------------------------------------------------------------------------
    def generate_latent(self, n_samples: int, device: Optional[torch.device] = None) -> torch.Tensor:
        if device is None:
            device = next(self.parameters()).device
        return torch.randn(n_samples, self.latent_dim, device=device)

    def get_num_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_output_shape(self) -> Tuple[int, int, int]:
        return (self.output_channels, self.output_size, self.output_size)


def create_generator(latent_dim: int = 100, output_size: int = 64, output_channels: int = 1) -> Generator:
    return Generator(latent_dim=latent_dim, output_size=output_size, output_channels=output_channels)
