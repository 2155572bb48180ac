"""ConfigurableGenerator — B200 drop-in for the generator the reference's ablation script defines for itself
(src/ablation_vanilla_gan_signatures.py:159-213 `UpsampleBlockConfigurable`, :216-328 `ConfigurableGenerator`): the
Vanilla-GAN generator with ReLU *or* LeakyReLU(0.2) after every BatchNorm (fc stage :265-268, upsample blocks :204-207).

Same constructor signature, attributes, sub-module names and `state_dict()` keys as the reference class (they equal the
plain Generator's: the activation modules hold no parameters). The arithmetic runs in libsiggan.so through the same
sg_g_forward / sg_g_backward entry points; the activation slope is part of the library context (`sg_config.g_act_slope`):
the GEMM epilogues, the BatchNorm-apply kernel, the final Conv3x3 kernels and every backward gate take it. The fastest
ReLU-only kernels (fused eval tail, `mma.sync` final-conv kernels, `convt4`'s fused eval epilogue) step aside for their
general siblings when the slope is not 0. `AblationGANTrainer` (ablation…:339-467) drives this module, the Discriminator
(spectral norm on or off) and stock `torch.optim.Adam` optimizers through the ordinary module / autograd path.
"""
from __future__ import annotations

import torch.nn as nn

from generator_vanilla_gan import Generator


class UpsampleBlockConfigurable(nn.Module):
    """Parameter holder for ConvTranspose2d(k4,s2,p1) [+ BatchNorm2d] + ReLU / LeakyReLU (ablation…:159-213)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 4, stride: int = 2, padding: int = 1,
                 output_padding: int = 0, use_batch_norm: bool = True, activation: str = "relu",
                 leaky_slope: float = 0.2) -> None:
        super().__init__()
        mods = [nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                                   output_padding=output_padding, bias=not use_batch_norm)]
        if use_batch_norm:
            mods.append(nn.BatchNorm2d(out_channels))
        mods.append(nn.LeakyReLU(leaky_slope, inplace=True) if activation == "leaky_relu" else nn.ReLU(inplace=True))
        self.block = nn.Sequential(*mods)

    def forward(self, x):
        raise RuntimeError("UpsampleBlockConfigurable is a parameter holder in siggan_b200; run it through "
                           "ConfigurableGenerator.forward")


class ConfigurableGenerator(Generator):
    """z (B, latent_dim) -> image (B, 1, S, S); activation 'relu' or 'leaky_relu' (ablation…:216-328). Anything but
    'leaky_relu' means ReLU, as in the reference (:204-207)."""

    def __init__(self, latent_dim: int = 100, output_size: int = 64, output_channels: int = 1, base_features: int = 256,
                 activation: str = "relu", leaky_slope: float = 0.2) -> None:
        super().__init__(latent_dim=latent_dim, output_size=output_size, output_channels=output_channels,
                         base_features=base_features)
        self.activation = activation
        leaky = activation == "leaky_relu"
        if leaky and not 0.0 <= float(leaky_slope) < 1.0:
            raise ValueError(f"leaky_slope must be in [0, 1), got {leaky_slope}")
        self._act_slope = float(leaky_slope) if leaky else 0.0
        # the holders mirror the reference's module tree (repr / children), parameters are the ones Generator created
        ladder = [256, 128, 64, 32, 32] if output_size == 64 else [512, 256, 128, 64, 32, 32]
        blocks = []
        for old, (a, b) in zip(self.upsample_blocks, zip(ladder[:-1], ladder[1:])):
            blk = UpsampleBlockConfigurable(a, b, activation=activation, leaky_slope=leaky_slope)
            blk.block[0], blk.block[1] = old.block[0], old.block[1]
            blocks.append(blk)
        self.upsample_blocks = nn.Sequential(*blocks)
        if leaky:
            self.fc[2] = nn.LeakyReLU(leaky_slope, inplace=True)
