"""VanillaGAN container — B200 drop-in for the reference's src/vanilla_gan_model.py.

Public surface (constructor, attributes `.generator .discriminator .criterion .g_optimizer .d_optimizer
.current_epoch .global_step .d_losses .g_losses`, methods, returned dict keys, checkpoint formats) mirrors
reference vanilla…:28-660. Differences are all under the hood:
  * `criterion` is an `nn.BCELoss` subclass whose CUDA path is one fused kernel each way (sg_bce_*),
  * the optimizers are `torch.optim.Adam` subclasses that update the flat parameter buffer with one kernel
    (sg_adam_step) and keep torch's `state_dict()` layout,
  * `train_discriminator_step` / `train_generator_step` / `train_step` enqueue the whole D / G step through
    sg_train_step (no autograd graph, one host sync per step for the returned floats).
The unchanged reference trainer (train…:281-376) drives `.generator/.discriminator/.criterion/.*_optimizer`
directly; that path goes through the autograd Functions of the two modules and is equally supported.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch
import torch.nn as nn
import torch.optim as optim

import _siggan_lib as L
from discriminator_vanilla_gan import Discriminator, create_discriminator  # noqa: F401
from generator_vanilla_gan import Generator, create_generator  # noqa: F401


# ------------------------------------------------------------------------------------------------
# nn.BCELoss on probabilities with fused CUDA kernels (reference vanilla…:107)
# ------------------------------------------------------------------------------------------------
class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        lib = L.load_library()
        loss = torch.empty((), dtype=torch.float32, device=prob.device)
        L.check(lib.sg_bce_forward(L.ptr(prob), L.ptr(target), prob.numel(), L.ptr(loss),
                                   L.current_stream(prob.device)), "sg_bce_forward")
        ctx.save_for_backward(prob, target)
        return loss

    @staticmethod
    def backward(ctx, grad_loss: torch.Tensor):
        prob, target = ctx.saved_tensors
        lib = L.load_library()
        dprob = torch.empty_like(prob)
        g = grad_loss.contiguous().float()
        L.check(lib.sg_bce_backward(L.ptr(prob), L.ptr(target), prob.numel(), L.ptr(g), L.ptr(dprob),
                                    L.current_stream(prob.device)), "sg_bce_backward")
        return dprob, None


class BCELoss(nn.BCELoss):
    """Mean binary cross-entropy on probabilities: log terms clamped at -100, gradient (p-y)/max(p(1-p),1e-12)/n."""

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if input.device.type != "cuda":
            raise RuntimeError("siggan_b200 BCELoss runs on CUDA only (no CPU path)")
        if self.weight is not None or self.reduction != "mean":
            raise NotImplementedError("only the reference's configuration nn.BCELoss() (mean, unweighted) is implemented")
        if input.shape != target.shape:
            raise ValueError(f"Using a target size ({target.shape}) that is different to the input size ({input.shape})")
        return _BCEFn.apply(input.contiguous().float(), target.contiguous().float())


# ------------------------------------------------------------------------------------------------
# torch.optim.Adam surface over the flat parameter buffer (reference vanilla…:110-120)
# ------------------------------------------------------------------------------------------------
class FusedAdam(optim.Adam):
    """`torch.optim.Adam` whose `step()` is one kernel over the owning module's flat parameter buffer.

    `state_dict()` / `load_state_dict()` / `param_groups` / `zero_grad()` are torch's own, so checkpoints written
    by the reference trainer (train…:402-444) load here and vice versa. Only the reference's configuration is
    accelerated (single param group, no weight decay, no amsgrad, not maximize)."""

    def __init__(self, module: nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        super().__init__(module.parameters(), lr=lr, betas=betas, eps=eps)
        self._module = module
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._step_t = torch.tensor(0.0)
        self._steps = 0

    # flat moment buffers, exposed to torch's state machinery as per-parameter views
    def _ensure_state(self) -> None:
        fp: L.FlatParams = self._module._flat
        flat = fp.flat
        if flat is None:
            raise RuntimeError("FusedAdam: the module has no flat CUDA parameters yet (move it to CUDA and run it once)")
        stale = self._m is None or self._m.device != flat.device or self._m.numel() != flat.numel()
        if not stale:
            st = self.state.get(fp.params[0])
            stale = st is None or "exp_avg" not in st or st["exp_avg"].data_ptr() != self._m.data_ptr() + 4 * fp.layout[0][0]
        if not stale:
            return
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        steps = 0
        for p, (off, n, shape) in zip(fp.params, fp.layout):
            st = self.state.get(p)
            if st is not None and "exp_avg" in st:   # adopt moments restored by load_state_dict
                m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps = int(float(st["step"]))
        self._m, self._v, self._steps = m, v, steps
        self._step_t = torch.tensor(float(steps))
        for p, (off, n, shape) in zip(fp.params, fp.layout):
            self.state[p] = {"step": self._step_t, "exp_avg": m[off:off + n].view(shape),
                             "exp_avg_sq": v[off:off + n].view(shape)}

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._m = None   # re-adopt the restored moments at the next step

    def _hyper(self):
        g = self.param_groups[0]
        if len(self.param_groups) != 1 or g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("FusedAdam accelerates the reference's Adam configuration only")
        return float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])

    def advance(self) -> int:
        """Book-keeping for callers that ran the update kernel themselves (fused train step)."""
        self._steps += 1
        self._step_t.fill_(float(self._steps))
        return self._steps

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        fp: L.FlatParams = self._module._flat
        if fp.flat is None or fp.flat.device.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA only (no CPU path)")
        grads = [p.grad for p in fp.params]
        if all(g is None for g in grads):
            return loss
        if any(g is None for g in grads):
            raise NotImplementedError("FusedAdam needs a gradient for every parameter of the network")
        self._ensure_state()
        gflat = fp.flat_grad_if_contiguous()
        if gflat is None:   # gradients produced elsewhere (e.g. clipped copies): gather them once
            gflat = fp.grad_staging()
            torch._foreach_copy_(fp.grad_views(gflat), [g.float() for g in grads])
        lr, b1, b2, eps = self._hyper()
        step = self.advance()
        lib = L.load_library()
        L.check(lib.sg_adam_step(L.ptr(fp.flat), L.ptr(gflat), L.ptr(self._m), L.ptr(self._v), fp.flat.numel(), lr, b1,
                                 b2, eps, step, L.current_stream(fp.flat.device)), "sg_adam_step")
        return loss


# ------------------------------------------------------------------------------------------------
# VanillaGAN
# ------------------------------------------------------------------------------------------------
_METRIC_KEYS_D = ["d_loss", "d_loss_real", "d_loss_fake", "d_real_acc", "d_fake_acc", "d_real_mean", "d_fake_mean"]
_METRIC_KEYS_G = ["g_loss", "g_fake_mean"]


class VanillaGAN(nn.Module):
    """Generator + Discriminator + BCE + two Adams (reference vanilla…:28-130)."""

    def __init__(self, latent_dim: int = 100, image_size: int = 64, image_channels: int = 1, g_lr: float = 2e-4,
                 d_lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999, label_smoothing: float = 0.9,
                 use_spectral_norm: bool = False, device: Optional[str] = None, width_mult: int = 1) -> None:
        super().__init__()
        if width_mult not in (1, 2):      # extension (last in the signature): 2 = the "2x hidden width" sweep variant
            raise ValueError(f"width_mult must be 1 or 2, got {width_mult}")
        self.width_mult = width_mult
        self.latent_dim, self.image_size, self.image_channels = latent_dim, image_size, image_channels
        self.g_lr, self.d_lr, self.beta1, self.beta2 = g_lr, d_lr, beta1, beta2
        self.label_smoothing = label_smoothing
        self.use_spectral_norm = use_spectral_norm
        if device is None:
            self._device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        else:
            self._device = torch.device(device)
        self.generator = Generator(latent_dim=latent_dim, output_size=image_size, output_channels=image_channels,
                                   base_features=256 * width_mult)
        self.discriminator = Discriminator(input_size=image_size, input_channels=image_channels,
                                           use_spectral_norm=use_spectral_norm, base_features=64 * width_mult)
        self.discriminator._latent_hint = latent_dim
        self.criterion = BCELoss()
        self.g_optimizer = FusedAdam(self.generator, lr=g_lr, betas=(beta1, beta2))
        self.d_optimizer = FusedAdam(self.discriminator, lr=d_lr, betas=(beta1, beta2))
        self.current_epoch = 0
        self.global_step = 0
        self.d_losses: list = []
        self.g_losses: list = []
        self._metrics: Optional[torch.Tensor] = None
        self._d_loss_hist: Optional[torch.Tensor] = None
        #: parity hook for the fused step: dict(real=[...], fake=[...]) of (B, C_i) keep-scale tensors
        self.mask_override: Optional[Dict[str, list]] = None
        self.to(self._device)

    @property
    def device(self) -> torch.device:
        return self._device

    # The Dropout2d mask stream is process-wide (shared with the Discriminator module path, _siggan_lib.DROPOUT).
    @property
    def _dropout_offset(self) -> int:
        return L.DROPOUT.offset

    @_dropout_offset.setter
    def _dropout_offset(self, value: int) -> None:
        L.DROPOUT.offset = int(value)

    def to(self, device: Union[str, torch.device]) -> "VanillaGAN":  # reference vanilla…:136-150
        if isinstance(device, str):
            device = torch.device(device)
        self._device = device
        super().to(device)
        return self

    def _get_labels(self, batch_size: int, real: bool = True, smooth: bool = True) -> torch.Tensor:
        if real:
            value = self.label_smoothing if smooth else 1.0
            return torch.full((batch_size, 1), value, device=self._device)
        return torch.zeros(batch_size, 1, device=self._device)

    # -- fused step plumbing ----------------------------------------------------------------------
    def _fused_ready(self) -> L.Context:
        if self._device.type != "cuda":
            raise RuntimeError("siggan_b200 VanillaGAN trains on CUDA only (no CPU path)")
        g, d = self.generator, self.discriminator
        g._prepare(g.fc[0].weight.device)
        d._prepare(d.classifier[0].bias.device)
        self.g_optimizer._ensure_state()
        self.d_optimizer._ensure_state()
        if self._metrics is None or self._metrics.device != g._flat.flat.device:
            self._metrics = torch.zeros(12, dtype=torch.float32, device=g._flat.flat.device)
        return g._ctx

    def _state(self, masks: Optional[Dict[str, torch.Tensor]] = None) -> L.SgTrainState:
        g, d = self.generator, self.discriminator
        glr, b1, b2, eps = self.g_optimizer._hyper()
        dlr = self.d_optimizer._hyper()[0]
        st = L.SgTrainState()
        st.g_params, st.g_running_stats = L.ptr(g._flat.flat), L.ptr(g._flat.stats)
        st.g_exp_avg, st.g_exp_avg_sq = L.ptr(self.g_optimizer._m), L.ptr(self.g_optimizer._v)
        st.d_params = L.ptr(d._flat.flat)
        st.d_exp_avg, st.d_exp_avg_sq = L.ptr(self.d_optimizer._m), L.ptr(self.d_optimizer._v)
        st.g_step, st.d_step = self.g_optimizer._steps, self.d_optimizer._steps
        st.g_lr, st.d_lr, st.beta1, st.beta2, st.eps = glr, dlr, b1, b2, eps
        st.label_smoothing = float(self.label_smoothing)
        st.dropout_p = float(d.dropout)
        st.seed = L.DROPOUT.seed()
        st.offset = L.DROPOUT.offset
        if masks is not None:
            st.masks_real, st.masks_fake = L.ptr(masks["real"]), L.ptr(masks["fake"])
        st.world_size = 1
        return st

    def _flat_masks(self, B: int) -> Optional[Dict[str, torch.Tensor]]:
        if self.mask_override is None:
            return None
        dev = self.generator._flat.flat.device
        return {k: torch.cat([m.to(device=dev, dtype=torch.float32).reshape(-1) for m in self.mask_override[k]]).contiguous()
                for k in ("real", "fake")}

    def _allreduce(self, flat_grad: torch.Tensor) -> None:
        """Data-parallel: average the flat gradient bucket across ranks (NCCL over NVLink) before the update."""
        from data_parallel import average_gradients_
        average_gradients_(flat_grad)

    # -- spectral-norm variant: the same two steps through the module path -----------------------------------
    # sg_train_step runs the Discriminator on its raw parameters; with use_spectral_norm the layers must run on
    # weight_orig / sigma with a power iteration per forward (disc…:61-62, 201-202), which the Discriminator module does
    # (sg_spectral_norm_weight / _backward around sg_d_forward / sg_d_backward). The steps below are the reference's own
    # statements (vanilla…:203-252, 273-306) on those modules; metrics stay on the device.
    def _grad_allreduce(self, module: nn.Module) -> None:
        import data_parallel as dp
        if dp.world()[1] > 1:
            flat = module._flat.flat_grad_if_contiguous()
            for g in ([flat] if flat is not None else [p.grad for p in module.parameters()]):
                dp.average_gradients_(g)

    def _layered_discriminator_step(self, real_images: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
        D, G = self.discriminator, self.generator
        D.train()
        G.eval()
        self._fused_ready()
        dev = G._flat.flat.device
        real = real_images.to(dev, non_blocking=True).contiguous().float()
        B = real.shape[0]
        noise = torch.randn(B, self.latent_dim, device=dev) if noise is None else noise.to(dev).contiguous().float()
        self.d_optimizer.zero_grad()
        masks = self.mask_override
        D.mask_override = masks["real"] if masks is not None else None
        real_preds = D(real)
        d_loss_real = self.criterion(real_preds, self._get_labels(B, real=True, smooth=True))
        with torch.no_grad():
            fake = G(noise)
        D.mask_override = masks["fake"] if masks is not None else None
        fake_preds = D(fake)
        D.mask_override = None
        d_loss_fake = self.criterion(fake_preds, self._get_labels(B, real=False))
        d_loss = d_loss_real + d_loss_fake
        d_loss.backward()
        self._grad_allreduce(D)
        self.d_optimizer.step()
        with torch.no_grad():
            self._metrics[:7] = torch.stack([d_loss, d_loss_real, d_loss_fake, (real_preds > 0.5).float().mean(),
                                             (fake_preds < 0.5).float().mean(), real_preds.mean(), fake_preds.mean()])
        return self._metrics

    def _layered_generator_step(self, batch_size: int, noise: Optional[torch.Tensor]) -> torch.Tensor:
        D, G = self.discriminator, self.generator
        G.train()
        D.eval()
        self._fused_ready()
        dev = G._flat.flat.device
        noise = torch.randn(batch_size, self.latent_dim, device=dev) if noise is None else noise.to(dev).contiguous().float()
        self.g_optimizer.zero_grad()
        fake_preds = D(G(noise))
        g_loss = self.criterion(fake_preds, self._get_labels(noise.shape[0], real=True, smooth=False))
        g_loss.backward()
        self._grad_allreduce(G)
        self.g_optimizer.step()
        with torch.no_grad():
            self._metrics[7:9] = torch.stack([g_loss, fake_preds.mean()])
        return self._metrics

    def discriminator_step_async(self, real_images: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueue one D step (reference vanilla…:180-252) without synchronising; metrics stay on the device."""
        if self.use_spectral_norm:
            return self._layered_discriminator_step(real_images, noise)
        self.discriminator.train()
        self.generator.eval()
        sctx = self._fused_ready()
        dev = self.generator._flat.flat.device
        real = real_images.to(dev, non_blocking=True).contiguous().float()
        B = real.shape[0]
        if noise is None:
            noise = torch.randn(B, self.latent_dim, device=dev)
        noise = noise.to(dev).contiguous().float()
        self.d_optimizer.zero_grad()
        masks = self._flat_masks(B)
        st = self._state(masks)
        dgrads = self.discriminator._flat.grad_staging()
        stream = L.current_stream(dev)
        args = (sctx.handle, C.byref(st), L.ptr(real), L.ptr(noise), None, B, L.ptr(dgrads), None, L.ptr(self._metrics))
        import data_parallel as dp
        if dp.world()[1] > 1:
            # data-parallel: the all-reduce of the bucket's tail (classifier + last conv block, 76 % of the bytes,
            # final first) runs on the communication stream while the rest of the D backward executes
            tail = int(sctx.lib.sg_d_grad_tail_offset(sctx.handle))
            own = sctx.comm_world() > 1      # library-owned NCCL communicator (data_parallel.init_library_comm)
            L.check(sctx.lib.sg_train_step(*args, 11, stream), "sg_train_step(D backward, last block)")
            w_tail = sctx.allreduce_grads(L.SG_NET_D, dgrads, tail, -1, overlap=True) if own else dp.all_reduce_start(dgrads[tail:])
            L.check(sctx.lib.sg_train_step(*args, 12, stream), "sg_train_step(D backward, remaining blocks)")
            w_head = sctx.allreduce_grads(L.SG_NET_D, dgrads, 0, tail, overlap=True) if own else dp.all_reduce_start(dgrads[:tail])
            if own:
                sctx.allreduce_join()
            else:
                dp.all_reduce_finish([w_tail, w_head], dgrads)
        else:
            L.check(sctx.lib.sg_train_step(*args, 1, stream), "sg_train_step(D backward)")
        L.check(sctx.lib.sg_train_step(*args, 2, stream), "sg_train_step(D update)")
        self.d_optimizer.advance()
        self.discriminator._flat.expose(dgrads)
        if st.dropout_p > 0:
            L.DROPOUT.advance(int(sctx.lib.sg_d_mask_count(sctx.handle, 2 * B)))
        return self._metrics

    def generator_step_async(self, batch_size: int, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueue one G step (reference vanilla…:254-306) without synchronising."""
        if self.use_spectral_norm:
            return self._layered_generator_step(batch_size, noise)
        self.generator.train()
        self.discriminator.eval()
        sctx = self._fused_ready()
        dev = self.generator._flat.flat.device
        if noise is None:
            noise = torch.randn(batch_size, self.latent_dim, device=dev)
        noise = noise.to(dev).contiguous().float()
        B = noise.shape[0]
        self.g_optimizer.zero_grad()
        st = self._state()
        ggrads = self.generator._flat.grad_staging()
        stream = L.current_stream(dev)
        args = (sctx.handle, C.byref(st), None, None, L.ptr(noise), B, None, L.ptr(ggrads), L.ptr(self._metrics))
        import data_parallel as dp
        if dp.world()[1] > 1:
            # the upsample blocks' and the final conv's gradients (the bucket from block 0's weight on) are averaged while
            # the fc stage of the backward runs; its own gradients follow
            tail = int(sctx.lib.sg_g_grad_tail_offset(sctx.handle))
            own = sctx.comm_world() > 1
            L.check(sctx.lib.sg_train_step(*args, 31, stream), "sg_train_step(G backward, upsample blocks)")
            w_tail = sctx.allreduce_grads(L.SG_NET_G, ggrads, tail, -1, overlap=True) if own else dp.all_reduce_start(ggrads[tail:])
            L.check(sctx.lib.sg_train_step(*args, 32, stream), "sg_train_step(G backward, fc stage)")
            w_head = sctx.allreduce_grads(L.SG_NET_G, ggrads, 0, tail, overlap=True) if own else dp.all_reduce_start(ggrads[:tail])
            if own:
                sctx.allreduce_join()
            else:
                dp.all_reduce_finish([w_tail, w_head], ggrads)
        else:
            L.check(sctx.lib.sg_train_step(*args, 3, stream), "sg_train_step(G backward)")
        L.check(sctx.lib.sg_train_step(*args, 4, stream), "sg_train_step(G update)")
        self.g_optimizer.advance()
        self.generator._flat.expose(ggrads)
        torch._foreach_add_([bn.num_batches_tracked for bn in self.generator._bn_modules()], 1)
        return self._metrics

    # -- reference API ----------------------------------------------------------------------------
    def train_discriminator_step(self, real_images: torch.Tensor, noise: Optional[torch.Tensor] = None) -> Dict[str, float]:
        m = self.discriminator_step_async(real_images, noise).tolist()
        out = dict(zip(_METRIC_KEYS_D, m[:7]))
        self.d_losses.append(out["d_loss"])
        self.global_step += 1
        return out

    def train_generator_step(self, batch_size: int, noise: Optional[torch.Tensor] = None) -> Dict[str, float]:
        m = self.generator_step_async(batch_size, noise).tolist()
        out = dict(zip(_METRIC_KEYS_G, m[7:9]))
        self.g_losses.append(out["g_loss"])
        return out

    def train_step_async(self, real_images: torch.Tensor, n_critic: int = 1) -> torch.Tensor:
        """D step(s) + G step enqueued back to back; returns the 12-float device metrics tensor (no host sync). With
        n_critic > 1 the loss of every critic step is kept in `self._d_loss_hist` (device, n_critic floats)."""
        if n_critic > 1:
            dev = self.generator.fc[0].weight.device
            if self._d_loss_hist is None or self._d_loss_hist.numel() != n_critic or self._d_loss_hist.device != dev:
                self._d_loss_hist = torch.zeros(n_critic, dtype=torch.float32, device=dev)
        if n_critic == 1 and not self.use_spectral_norm and self.mask_override is None and self._whole_step_in_library():
            return self._step_in_library(real_images)
        for i in range(n_critic):
            m = self.discriminator_step_async(real_images)
            if n_critic > 1:
                self._d_loss_hist[i:i + 1].copy_(m[0:1])
        return self.generator_step_async(real_images.size(0))

    def _whole_step_in_library(self) -> bool:
        """A single process, or data-parallel ranks whose context owns its NCCL communicator (init_library_comm); ranks
        that all-reduce through torch.distributed drive the phases themselves."""
        import data_parallel as dp
        if dp.world()[1] == 1:
            return True
        g = self.generator
        return getattr(g, "_ctx", None) is not None and g._ctx.comm_world() > 1

    def _step_in_library(self, real_images: torch.Tensor) -> torch.Tensor:
        """One D step + G step as ONE library call (sg_train_step phase 0). Inside the call nobody else writes the
        parameters, so the D update emits the Discriminator's weight packs itself and the G step re-packs nothing; with a
        communicator the bucket all-reduces are issued by libsiggan on its communication stream between the phases."""
        # module modes: the library's phases do not read them; leave the ones the reference's G step leaves behind
        # (vanilla…:274-275: G.train(), D.eval()) without walking the module trees when they are already set
        if not self.generator.training:
            self.generator.train()
        if self.discriminator.training:
            self.discriminator.eval()
        sctx = self._fused_ready()
        dev = self.generator._flat.flat.device
        real = real_images.to(dev, non_blocking=True).contiguous().float()
        B = real.shape[0]
        # two draws, as the separate D / G steps make them (the same RNG stream whichever path runs)
        noise = [torch.randn(B, self.latent_dim, device=dev), torch.randn(B, self.latent_dim, device=dev)]
        self.d_optimizer.zero_grad()
        self.g_optimizer.zero_grad()
        st = self._state()
        dgrads, ggrads = self.discriminator._flat.grad_staging(), self.generator._flat.grad_staging()
        L.check(sctx.lib.sg_train_step(sctx.handle, C.byref(st), L.ptr(real), L.ptr(noise[0]), L.ptr(noise[1]), B,
                                       L.ptr(dgrads), L.ptr(ggrads), L.ptr(self._metrics), 0, L.current_stream(dev)),
                "sg_train_step(whole step)")
        self.d_optimizer.advance()
        self.g_optimizer.advance()
        self.discriminator._flat.expose(dgrads)
        self.generator._flat.expose(ggrads)
        if st.dropout_p > 0:
            L.DROPOUT.advance(int(sctx.lib.sg_d_mask_count(sctx.handle, 2 * B)))
        torch._foreach_add_([bn.num_batches_tracked for bn in self.generator._bn_modules()], 1)
        return self._metrics

    def train_step(self, real_images: torch.Tensor, n_critic: int = 1) -> Dict[str, float]:
        """reference vanilla…:308-336; one device->host read for all returned floats. Every critic step's own loss
        goes into `d_losses` (the reference appends per train_discriminator_step call, vanilla…:241)."""
        metrics = self.train_step_async(real_images, n_critic)
        if n_critic > 1:
            both = torch.cat([metrics, self._d_loss_hist]).tolist()
            m, per_step = both[:12], both[12:]
        else:
            m = metrics.tolist()
            per_step = [m[0]]
        out = dict(zip(_METRIC_KEYS_D, m[:7]))
        out.update(zip(_METRIC_KEYS_G, m[7:9]))
        self.d_losses.extend(per_step)
        self.global_step += n_critic
        self.g_losses.append(out["g_loss"])
        return out

    @torch.no_grad()
    def generate(self, n_samples: int, device: Optional[Union[str, torch.device]] = None,
                 noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.generator.eval()
        if device is None:
            device = self._device
        elif isinstance(device, str):
            device = torch.device(device)
        noise = torch.randn(n_samples, self.latent_dim, device=device) if noise is None else noise.to(device)
        return self.generator(noise)

    @torch.no_grad()
    def generate_interpolation(self, n_steps: int = 10, z_start: Optional[torch.Tensor] = None,
                               z_end: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Linear interpolation between two latents (reference vanilla…:373-409)."""
        self.generator.eval()
        if z_start is None:
            z_start = torch.randn(1, self.latent_dim, device=self._device)
        if z_end is None:
            z_end = torch.randn(1, self.latent_dim, device=self._device)
        alphas = torch.linspace(0, 1, n_steps, device=self._device).view(-1, 1)
        return self.generator(z_start * (1 - alphas) + z_end * alphas)

    def get_config(self) -> Dict[str, Any]:
        gp, dp = self.generator.get_num_params(), self.discriminator.get_num_params()
        return {"latent_dim": self.latent_dim, "image_size": self.image_size, "image_channels": self.image_channels,
                "g_lr": self.g_lr, "d_lr": self.d_lr, "beta1": self.beta1, "beta2": self.beta2,
                "label_smoothing": self.label_smoothing, "use_spectral_norm": self.use_spectral_norm,
                "current_epoch": self.current_epoch, "global_step": self.global_step, "g_params": gp, "d_params": dp,
                "total_params": gp + dp, **({"width_mult": self.width_mult} if self.width_mult != 1 else {})}

    def save(self, path: Union[str, Path], save_optimizer: bool = True, save_history: bool = True) -> None:
        """Checkpoint in the reference's format (vanilla…:433-474): `<path>.pt` + `<path>_config.json`."""
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        ckpt = {"config": self.get_config(), "generator_state_dict": self.generator.state_dict(),
                "discriminator_state_dict": self.discriminator.state_dict(), "current_epoch": self.current_epoch,
                "global_step": self.global_step, "saved_at": datetime.now().isoformat()}
        if save_optimizer:
            if self.generator._flat.flat is not None and self.generator._flat.flat.device.type == "cuda":
                self.g_optimizer._ensure_state()
                self.d_optimizer._ensure_state()
            ckpt["g_optimizer_state_dict"] = self.g_optimizer.state_dict()
            ckpt["d_optimizer_state_dict"] = self.d_optimizer.state_dict()
        if save_history:
            ckpt["d_losses"], ckpt["g_losses"] = self.d_losses, self.g_losses
        ckpt["dropout_stream"] = L.DROPOUT.state_dict()   # extra key (ignored by the reference's loader): resume continues the masks
        torch.save(ckpt, f"{path}.pt")
        with open(f"{path}_config.json", "w") as f:
            json.dump(self.get_config(), f, indent=2)
        print(f"Model saved to {path}.pt")

    def load(self, path: Union[str, Path], load_optimizer: bool = True, load_history: bool = True,
             map_location: Optional[str] = None) -> None:
        path = Path(path)
        if not path.suffix:
            path = Path(f"{path}.pt")
        if map_location is None:
            map_location = str(self._device)
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        self.generator.load_state_dict(ckpt["generator_state_dict"])
        self.discriminator.load_state_dict(ckpt["discriminator_state_dict"])
        self.current_epoch = ckpt.get("current_epoch", 0)
        self.global_step = ckpt.get("global_step", 0)
        if load_optimizer and "g_optimizer_state_dict" in ckpt:
            self.g_optimizer.load_state_dict(ckpt["g_optimizer_state_dict"])
            self.d_optimizer.load_state_dict(ckpt["d_optimizer_state_dict"])
        if "dropout_stream" in ckpt:
            L.DROPOUT.load_state_dict(ckpt["dropout_stream"])
        if load_history and "d_losses" in ckpt:
            self.d_losses = ckpt.get("d_losses", [])
            self.g_losses = ckpt.get("g_losses", [])
        print(f"Model loaded from {path}")
        print(f"  Epoch: {self.current_epoch}, Global Step: {self.global_step}")

    @classmethod
    def from_checkpoint(cls, path: Union[str, Path], device: Optional[str] = None) -> "VanillaGAN":
        path = Path(path)
        if not path.suffix:
            path = Path(f"{path}.pt")
        cfg = torch.load(path, map_location="cpu", weights_only=False)["config"]
        model = cls(latent_dim=cfg["latent_dim"], image_size=cfg["image_size"], image_channels=cfg["image_channels"],
                    g_lr=cfg["g_lr"], d_lr=cfg["d_lr"], beta1=cfg["beta1"], beta2=cfg["beta2"],
                    label_smoothing=cfg["label_smoothing"], use_spectral_norm=cfg["use_spectral_norm"], device=device,
                    width_mult=cfg.get("width_mult", 1))
        model.load(path, load_optimizer=True, load_history=True)
        return model

    def set_learning_rates(self, g_lr: float, d_lr: float) -> None:
        for group in self.g_optimizer.param_groups:
            group["lr"] = g_lr
        for group in self.d_optimizer.param_groups:
            group["lr"] = d_lr
        self.g_lr, self.d_lr = g_lr, d_lr

    def get_recent_losses(self, n: int = 100) -> Dict[str, float]:
        d_recent = self.d_losses[-n:] if self.d_losses else [0]
        g_recent = self.g_losses[-n:] if self.g_losses else [0]
        return {"avg_d_loss": sum(d_recent) / len(d_recent), "avg_g_loss": sum(g_recent) / len(g_recent)}

    def summary(self) -> str:
        c = self.get_config()
        bar, thin = "=" * 60, "-" * 60
        lines = [bar, "VanillaGAN Model Summary", bar, f"Device: {self._device}",
                 f"Latent Dimension: {c['latent_dim']}", f"Image Size: {c['image_size']}x{c['image_size']}",
                 f"Image Channels: {c['image_channels']}", thin, "Generator:", f"  Parameters: {c['g_params']:,}",
                 f"  Learning Rate: {c['g_lr']}", thin, "Discriminator:", f"  Parameters: {c['d_params']:,}",
                 f"  Learning Rate: {c['d_lr']}", f"  Spectral Norm: {c['use_spectral_norm']}", thin,
                 f"Total Parameters: {c['total_params']:,}", f"Label Smoothing: {c['label_smoothing']}",
                 f"Adam Betas: ({c['beta1']}, {c['beta2']})", thin, "Training State:",
                 f"  Current Epoch: {c['current_epoch']}", f"  Global Step: {c['global_step']}", bar]
        return "\n".join(lines)


def create_vanilla_gan(latent_dim: int = 100, image_size: int = 64, use_spectral_norm: bool = False,
                       device: Optional[str] = None) -> VanillaGAN:
    return VanillaGAN(latent_dim=latent_dim, image_size=image_size, image_channels=1,
                      use_spectral_norm=use_spectral_norm, device=device)
