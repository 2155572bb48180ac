"""CPU oracle for the signature-GAN hot path (TEST INFRASTRUCTURE — never imported by the product).

A functional fp32 restatement, on CPU torch tensors, of what the reference's Generator /
Discriminator / BCELoss / Adam compute during a D step and a G step, with every backward pass
written out by hand (no autograd), so that both the forward formulas and the gradient formulas of
the CUDA path have an independent checker. Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file.

Pinning: the reference ships no golden vectors or value-asserting tests, so the
oracle is pinned against outputs of the reference itself, produced in the build container by
`tests/golden/make_golden.py` (which imports /root/reference/src unmodified) and committed under
`tests/golden/`. `tests/test_oracle_golden.py` checks the oracle against them.

Parameters travel as a dict keyed exactly like the reference's `state_dict()`:
  G: fc.0.{weight,bias} fc.1.{weight,bias,running_mean,running_var,num_batches_tracked}
     upsample_blocks.{i}.block.0.weight  upsample_blocks.{i}.block.1.{...}  final_conv.0.{weight,bias}
  D: conv_blocks.{i}.block.0.{weight,bias}  classifier.0.{weight,bias}
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5       # torch.nn.BatchNorm default, used by gen…:58,126
BN_MOMENTUM = 0.1   # idem
LEAKY_SLOPE = 0.2   # disc…:46,71
DROPOUT_P = 0.25    # disc…:45,75


def g_channels(image_size: int, width: int = 1) -> List[int]:
    """Channel ladder of the upsample blocks (gen…:131-149). width = 2: the "2x hidden width" variant of BASELINE
    configs[4] — not a configuration of the reference (its base_features is inert): the same blocks
    (gen…:33-42) assembled with doubled channel counts."""
    if image_size == 64:
        return [c * width for c in (256, 128, 64, 32, 32)]
    if image_size == 128:
        return [c * width for c in (512, 256, 128, 64, 32, 32)]
    raise ValueError(f"output_size must be 64 or 128, got {image_size}")


def d_channels(image_size: int, in_ch: int = 1, width: int = 1) -> List[int]:
    """Channel ladder of the downsample blocks (disc…:131-194); width as in g_channels (disc…:36-47 blocks)."""
    if image_size == 64:
        return [in_ch] + [c * width for c in (64, 128, 256, 512)]
    if image_size == 128:
        return [in_ch] + [c * width for c in (64, 128, 256, 512, 512)]
    raise ValueError(f"input_size must be 64 or 128, got {image_size}")


# --------------------------------------------------------------------------------------------
# BatchNorm (training: batch statistics, biased variance to normalise, unbiased for running_var)
# --------------------------------------------------------------------------------------------
def _bn_forward(y: Tensor, sd: Dict[str, Tensor], prefix: str, train: bool, new_stats: Dict[str, Tensor]):
    dims = [0] if y.dim() == 2 else [0, 2, 3]
    shape = [1, -1] if y.dim() == 2 else [1, -1, 1, 1]
    gamma, beta = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if train:
        n = y.numel() // y.shape[1]
        mean = y.mean(dim=dims)
        var = y.var(dim=dims, unbiased=False)
        new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_mean"] + BN_MOMENTUM * mean
        new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_var"] + BN_MOMENTUM * var * (
            n / max(n - 1, 1))
        new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    rstd = torch.rsqrt(var + BN_EPS)
    xhat = (y - mean.view(shape)) * rstd.view(shape)
    return xhat * gamma.view(shape) + beta.view(shape), xhat, rstd


def _bn_backward(dout: Tensor, xhat: Tensor, rstd: Tensor, gamma: Tensor, train: bool):
    """Returns (dy, dgamma, dbeta). Training mode differentiates through the batch statistics."""
    dims = [0] if dout.dim() == 2 else [0, 2, 3]
    shape = [1, -1] if dout.dim() == 2 else [1, -1, 1, 1]
    dbeta = dout.sum(dim=dims)
    dgamma = (dout * xhat).sum(dim=dims)
    scale = (gamma * rstd).view(shape)
    if train:
        n = dout.numel() // dout.shape[1]
        dy = scale * (dout - dbeta.view(shape) / n - xhat * dgamma.view(shape) / n)
    else:
        dy = scale * dout
    return dy, dgamma, dbeta


# --------------------------------------------------------------------------------------------
# Generator (gen…:189-209): fc -> view -> 4-5 x (ConvT k4 s2 p1, BN2d, ReLU) -> Conv3x3 -> Tanh
# --------------------------------------------------------------------------------------------
def _g_act(a: Tensor, act_slope: float) -> Tensor:
    """ReLU (gen…:60,127) or, for the ablation's ConfigurableGenerator, LeakyReLU(act_slope) (ablation…:204-207, 265-268)."""
    return torch.relu(a) if act_slope == 0.0 else F.leaky_relu(a, act_slope)


def _g_act_grad(a: Tensor, act_slope: float) -> Tensor:
    """Derivative of the activation from its OUTPUT (the reference's activations are in-place; the sign is preserved)."""
    pos = a > 0
    return pos.to(a.dtype) if act_slope == 0.0 else torch.where(pos, torch.ones_like(a), torch.full_like(a, act_slope))


def g_forward(sd: Dict[str, Tensor], z: Tensor, image_size: int = 64, train: bool = False, act_slope: float = 0.0):
    """Returns (image, cache, new_stats). `cache` holds every intermediate (keys documented inline). act_slope > 0:
    the ablation script's ConfigurableGenerator with activation="leaky_relu" (ablation…:216-328; same layers)."""
    ch = g_channels(image_size, sd["fc.0.weight"].shape[0] // (16 * g_channels(image_size)[0]))
    cache: Dict[str, Tensor] = {"z": z}
    new_stats: Dict[str, Tensor] = {}
    y = F.linear(z, sd["fc.0.weight"], sd["fc.0.bias"])                       # gen…:125
    cache["fc.y"] = y
    a, xhat, rstd = _bn_forward(y, sd, "fc.1", train, new_stats)              # gen…:126
    a = _g_act(a, act_slope)                                                  # gen…:127 / ablation…:265-268
    cache["fc.xhat"], cache["fc.rstd"], cache["fc.a"] = xhat, rstd, a
    x = a.view(-1, ch[0], 4, 4)                                               # gen…:201
    for i in range(len(ch) - 1):
        p = f"upsample_blocks.{i}.block"
        cache[f"up{i}.in"] = x
        y = F.conv_transpose2d(x, sd[p + ".0.weight"], None, stride=2, padding=1)   # gen…:46-54
        cache[f"up{i}.y"] = y
        a, xhat, rstd = _bn_forward(y, sd, p + ".1", train, new_stats)        # gen…:58
        x = _g_act(a, act_slope)                                              # gen…:60 / ablation…:204-207
        cache[f"up{i}.xhat"], cache[f"up{i}.rstd"], cache[f"up{i}.a"] = xhat, rstd, x
    pre = F.conv2d(x, sd["final_conv.0.weight"], sd["final_conv.0.bias"], stride=1, padding=1)  # gen…:154-161
    out = torch.tanh(pre)                                                     # gen…:162
    cache["final.in"], cache["out"] = x, out
    return out, cache, new_stats


def g_backward(sd: Dict[str, Tensor], cache: Dict[str, Tensor], dout: Tensor, image_size: int = 64,
               train: bool = True, taps: Optional[Dict[str, Tensor]] = None, act_slope: float = 0.0) -> Dict[str, Tensor]:
    """Gradients of every Generator parameter given d(loss)/d(image). `taps` (optional dict) receives the upstream
    gradient of every stage: `up{i}.dbn` / `fc.dbn` = gradient w.r.t. the stage's BatchNorm output with ReLU' applied
    (what the per-layer parity tests feed to one stage of the CUDA backward)."""
    ch = g_channels(image_size, sd["fc.0.weight"].shape[0] // (16 * g_channels(image_size)[0]))
    g: Dict[str, Tensor] = {}
    dpre = dout * (1.0 - cache["out"] ** 2)                                   # tanh'
    xin = cache["final.in"]
    g["final_conv.0.bias"] = dpre.sum(dim=[0, 2, 3])
    g["final_conv.0.weight"] = torch.nn.grad.conv2d_weight(xin, sd["final_conv.0.weight"].shape, dpre, stride=1,
                                                           padding=1)
    da = F.conv_transpose2d(dpre, sd["final_conv.0.weight"], None, stride=1, padding=1)
    for i in reversed(range(len(ch) - 1)):
        p = f"upsample_blocks.{i}.block"
        dbn = da * _g_act_grad(cache[f"up{i}.a"], act_slope)                  # ReLU' / LeakyReLU'
        if taps is not None:
            taps[f"up{i}.dbn"] = dbn
        dy, dgam, dbet = _bn_backward(dbn, cache[f"up{i}.xhat"], cache[f"up{i}.rstd"], sd[p + ".1.weight"], train)
        g[p + ".1.weight"], g[p + ".1.bias"] = dgam, dbet
        w = sd[p + ".0.weight"]                                               # (Cin, Cout, 4, 4)
        x = cache[f"up{i}.in"]
        # dW[ci,co,ky,kx] = sum x[n,ci,iy,ix] * dy[n,co,2iy-1+ky,2ix-1+kx]
        g[p + ".0.weight"] = torch.nn.grad.conv2d_weight(dy, w.shape, x, stride=2, padding=1)
        da = F.conv2d(dy, w, None, stride=2, padding=1)                       # data gradient of ConvT
    dfc = da.reshape(da.shape[0], -1) * _g_act_grad(cache["fc.a"], act_slope)
    if taps is not None:
        taps["fc.dbn"] = dfc
    dy, dgam, dbet = _bn_backward(dfc, cache["fc.xhat"], cache["fc.rstd"], sd["fc.1.weight"], train)
    g["fc.1.weight"], g["fc.1.bias"] = dgam, dbet
    g["fc.0.weight"] = dy.t() @ cache["z"]
    g["fc.0.bias"] = dy.sum(dim=0)
    g["__dz"] = dy @ sd["fc.0.weight"]
    return g


# --------------------------------------------------------------------------------------------
# Discriminator (disc…:241-260): 4-5 x (Conv k4 s2 p1 + bias, LeakyReLU, Dropout2d) -> Linear -> Sigmoid
# --------------------------------------------------------------------------------------------
def d_forward(sd: Dict[str, Tensor], x: Tensor, image_size: int = 64,
              masks: Optional[List[Tensor]] = None):
    """`masks[i]` is the Dropout2d keep-scale of block i, shape (B, C_i) with values {0, 1/(1-p)}; None = eval."""
    ch = d_channels(image_size, x.shape[1])
    cache: Dict[str, Tensor] = {}
    a = x
    for i in range(len(ch) - 1):
        p = f"conv_blocks.{i}.block.0"
        cache[f"c{i}.in"] = a
        y = F.conv2d(a, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1)     # disc…:51-58
        a = torch.where(y > 0, y, y * LEAKY_SLOPE)                                   # disc…:71
        if masks is not None:
            a = a * masks[i].view(masks[i].shape[0], -1, 1, 1)                       # disc…:75 (per (n,c) mask)
        cache[f"c{i}.a"] = a
    feat = a.reshape(a.shape[0], -1)                                                 # disc…:197 NCHW flatten
    logit = F.linear(feat, sd["classifier.0.weight"], sd["classifier.0.bias"])       # disc…:200
    prob = torch.sigmoid(logit)                                                      # disc…:206
    cache["feat"], cache["logit"], cache["prob"] = feat, logit, prob
    return prob, cache


def d_backward(sd: Dict[str, Tensor], cache: Dict[str, Tensor], dprob: Tensor, image_size: int = 64,
               masks: Optional[List[Tensor]] = None, need_dx: bool = False,
               taps: Optional[Dict[str, Tensor]] = None):
    """`taps` (optional dict) receives `c{i}.dy`, the gradient w.r.t. block i's convolution output (after LeakyReLU'
    and the dropout mask) — the upstream gradient the per-layer parity tests feed to one block of the CUDA backward."""
    ch = d_channels(image_size, cache["c0.in"].shape[1])
    g: Dict[str, Tensor] = {}
    p = cache["prob"]
    dlogit = dprob * p * (1.0 - p)                                                   # sigmoid'
    g["classifier.0.weight"] = dlogit.t() @ cache["feat"]
    g["classifier.0.bias"] = dlogit.sum(dim=0)
    da = (dlogit @ sd["classifier.0.weight"]).view_as(cache[f"c{len(ch) - 2}.a"])
    for i in reversed(range(len(ch) - 1)):
        name = f"conv_blocks.{i}.block.0"
        a = cache[f"c{i}.a"]
        if masks is not None:
            da = da * masks[i].view(masks[i].shape[0], -1, 1, 1)
        # LeakyReLU(inplace) backward keys on the sign of its output; dropped channels carry zero gradient anyway.
        dy = da * torch.where(a > 0, torch.ones_like(a), torch.full_like(a, LEAKY_SLOPE))
        if taps is not None:
            taps[f"c{i}.dy"] = dy
        w = sd[name + ".weight"]
        xin = cache[f"c{i}.in"]
        g[name + ".weight"] = torch.nn.grad.conv2d_weight(xin, w.shape, dy, stride=2, padding=1)
        g[name + ".bias"] = dy.sum(dim=[0, 2, 3])
        if i > 0 or need_dx:
            da = F.conv_transpose2d(dy, w, None, stride=2, padding=1)
    if need_dx:
        g["__dx"] = da
    return g


# --------------------------------------------------------------------------------------------
# Spectral-norm variant of the Discriminator (disc…:61-62 Conv2d, :201-202 Linear wrapped in
# torch.nn.utils.spectral_norm; built by ablation_vanilla_gan_signatures.py:367-371). The arithmetic lives in
# torch (SpectralNorm.compute_weight, torch/nn/utils/spectral_norm.py); restated here.
# --------------------------------------------------------------------------------------------
SN_EPS = 1e-12  # spectral_norm default eps


def sn_layer_names(image_size: int) -> List[str]:
    return [f"conv_blocks.{i}.block.0" for i in range(len(d_channels(image_size)) - 1)] + ["classifier.0"]


def spectral_weight(w_orig: Tensor, u: Tensor, v: Tensor, power_iterations: int, eps: float = SN_EPS):
    """One compute_weight call. Returns (weight, u', v', sigma); u', v' are what the buffers hold afterwards.
    W = weight_orig as (Cout, -1); v <- normalize(W^T u); u <- normalize(W v); sigma = u.(W v); weight = W_orig/sigma.
    normalize(x) = x / max(||x||_2, eps)."""
    W = w_orig.reshape(w_orig.shape[0], -1)
    for _ in range(power_iterations):
        t = W.t() @ u
        v = t / torch.clamp(t.norm(), min=eps)
        s = W @ v
        u = s / torch.clamp(s.norm(), min=eps)
    sigma = torch.dot(u, W @ v)
    return w_orig / sigma, u, v, sigma


def spectral_grad(g_eff: Tensor, w_eff: Tensor, u: Tensor, v: Tensor, sigma: Tensor) -> Tensor:
    """dL/dweight_orig from dL/dweight with u, v constant: (G - <G, weight> u v^T) / sigma."""
    inner = (g_eff * w_eff).sum()
    return (g_eff - inner * torch.outer(u, v).view_as(g_eff)) / sigma


def sn_effective(sd: Dict[str, Tensor], image_size: int, train: bool):
    """sd keyed like the reference's SN state_dict ({p}.weight_orig / weight_u / weight_v / bias). Returns the plain
    state dict the layers run on, the per-layer (u', v', sigma), and the buffers after the call."""
    eff: Dict[str, Tensor] = {}
    aux: Dict[str, Tuple[Tensor, Tensor, Tensor]] = {}
    new_buffers: Dict[str, Tensor] = {}
    for p in sn_layer_names(image_size):
        w, u, v, sigma = spectral_weight(sd[p + ".weight_orig"], sd[p + ".weight_u"], sd[p + ".weight_v"],
                                         1 if train else 0)
        eff[p + ".weight"], eff[p + ".bias"] = w, sd[p + ".bias"]
        aux[p] = (u, v, sigma)
        new_buffers[p + ".weight_u"], new_buffers[p + ".weight_v"] = u, v
    return eff, aux, new_buffers


def d_forward_sn(sd: Dict[str, Tensor], x: Tensor, image_size: int = 64, masks: Optional[List[Tensor]] = None,
                 train: bool = False):
    """Forward of Discriminator(use_spectral_norm=True). `train` = module.training (power iteration on/off); masks as
    in d_forward. Returns (prob, cache, new_buffers)."""
    eff, aux, new_buffers = sn_effective(sd, image_size, train)
    prob, cache = d_forward(eff, x, image_size, masks)
    cache["__eff"], cache["__aux"] = eff, aux
    return prob, cache, new_buffers


def d_backward_sn(cache: Dict[str, Tensor], dprob: Tensor, image_size: int = 64,
                  masks: Optional[List[Tensor]] = None, need_dx: bool = False) -> Dict[str, Tensor]:
    eff, aux = cache["__eff"], cache["__aux"]
    g = d_backward(eff, cache, dprob, image_size, masks, need_dx)
    out: Dict[str, Tensor] = {}
    for k, val in g.items():
        if k.endswith(".weight"):
            p = k[:-len(".weight")]
            u, v, sigma = aux[p]
            out[p + ".weight_orig"] = spectral_grad(val, eff[k], u, v, sigma)
        else:
            out[k] = val
    return out


def sn_trainable_names(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k in sd if k.endswith(".bias") or k.endswith(".weight_orig")]


def d_step_sn(g_sd, d_sd, d_opt: "AdamState", real: Tensor, noise: Tensor, image_size: int = 64,
              masks_real=None, masks_fake=None, label_smoothing: float = 0.9, lr: float = 2e-4,
              b1: float = 0.5, b2: float = 0.999, apply_update: bool = True):
    """vanilla…:180-252 with VanillaGAN(use_spectral_norm=True): D.train -> BOTH forwards run a power iteration (the
    second starts from the buffers the first left), each backward uses its own forward's (u, v, sigma)."""
    p_real, c_real, buf = d_forward_sn(d_sd, real, image_size, masks_real, train=True)
    d_sd.update(buf)
    fake, _, _ = g_forward(g_sd, noise, image_size, train=False)
    p_fake, c_fake, buf = d_forward_sn(d_sd, fake, image_size, masks_fake, train=True)
    d_sd.update(buf)
    t_real = torch.full_like(p_real, label_smoothing)
    loss_real, loss_fake = bce(p_real, t_real), bce(p_fake, torch.zeros_like(p_fake))
    g_real = d_backward_sn(c_real, bce_grad(p_real, t_real), image_size, masks_real)
    g_fake = d_backward_sn(c_fake, bce_grad(p_fake, torch.zeros_like(p_fake)), image_size, masks_fake)
    grads = {k: g_real[k] + g_fake[k] for k in g_real}
    if apply_update:
        d_opt.apply(d_sd, grads, lr, b1, b2)
    metrics = {
        "d_loss": float(loss_real + loss_fake), "d_loss_real": float(loss_real), "d_loss_fake": float(loss_fake),
        "d_real_acc": float((p_real > 0.5).float().mean()), "d_fake_acc": float((p_fake < 0.5).float().mean()),
        "d_real_mean": float(p_real.mean()), "d_fake_mean": float(p_fake.mean()),
    }
    return metrics, grads


def g_step_sn(g_sd, d_sd, g_opt: "AdamState", noise: Tensor, image_size: int = 64, lr: float = 2e-4,
              b1: float = 0.5, b2: float = 0.999, apply_update: bool = True):
    """vanilla…:254-306 with the SN discriminator in eval mode: no power iteration, sigma from the stored u, v."""
    fake, gc, new_stats = g_forward(g_sd, noise, image_size, train=True)
    p, dc, _ = d_forward_sn(d_sd, fake, image_size, None, train=False)
    ones = torch.ones_like(p)
    loss = bce(p, ones)
    dg = d_backward(dc["__eff"], dc, bce_grad(p, ones), image_size, None, need_dx=True)
    grads = g_backward(g_sd, gc, dg["__dx"], image_size, train=True)
    grads.pop("__dz")
    for k, v in new_stats.items():
        g_sd[k] = v
    if apply_update:
        g_opt.apply(g_sd, grads, lr, b1, b2)
    return {"g_loss": float(loss), "g_fake_mean": float(p.mean())}, grads


def make_sn_state_dict(image_size: int = 64, seed: int = 0) -> Dict[str, Tensor]:
    """SN-variant state dict in the reference's key order: kaiming-scale weight_orig (the reference's DCGAN init
    does not reach weight_orig, disc…:212-239 initialises the derived `weight`), unit-norm u / v."""
    _, d = make_state_dicts(image_size, 100, seed)
    out: Dict[str, Tensor] = {}
    for k, p in enumerate(sn_layer_names(image_size)):
        w = d[p + ".weight"] * 2.5
        out[p + ".bias"] = d[p + ".bias"]
        out[p + ".weight_orig"] = w
        u = hash_normal((w.shape[0],), seed * 1000 + 500 + k)
        v = hash_normal((w[0].numel(),), seed * 1000 + 520 + k)
        out[p + ".weight_u"] = u / u.norm()
        out[p + ".weight_v"] = v / v.norm()
    return out


# --------------------------------------------------------------------------------------------
# nn.BCELoss on probabilities (vanilla…:107): log clamped at -100, mean reduction
# --------------------------------------------------------------------------------------------
def bce(p: Tensor, y: Tensor) -> Tensor:
    logp = torch.clamp(torch.log(p), min=-100.0)
    log1p = torch.clamp(torch.log(1.0 - p), min=-100.0)
    return (-(y * logp + (1.0 - y) * log1p)).mean()


def bce_grad(p: Tensor, y: Tensor) -> Tensor:
    """d(mean BCE)/dp as ATen computes it: (p - y) / max(p (1-p), 1e-12) / numel."""
    return (p - y) / torch.clamp(p * (1.0 - p), min=1e-12) / p.numel()


# --------------------------------------------------------------------------------------------
# torch.optim.Adam (vanilla…:110-120): wd = 0, amsgrad = False, eps = 1e-8
# --------------------------------------------------------------------------------------------
def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float, b2: float,
                eps: float = 1e-8) -> None:
    """In-place single-tensor Adam, `step` is the 1-based count after this update."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


class AdamState:
    def __init__(self, params: Dict[str, Tensor], names: List[str]):
        self.names = names
        self.m = {k: torch.zeros_like(params[k]) for k in names}
        self.v = {k: torch.zeros_like(params[k]) for k in names}
        self.step = 0

    def apply(self, params: Dict[str, Tensor], grads: Dict[str, Tensor], lr: float, b1: float, b2: float):
        self.step += 1
        for k in self.names:
            adam_update(params[k], grads[k], self.m[k], self.v[k], self.step, lr, b1, b2)


def trainable_names(sd: Dict[str, Tensor]) -> List[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


# --------------------------------------------------------------------------------------------
# The two step algorithms of the production trainer (train…:281-376 == vanilla…:180-306)
# --------------------------------------------------------------------------------------------
def d_step(g_sd, d_sd, d_opt: AdamState, real: Tensor, noise: Tensor, image_size: int = 64,
           masks_real=None, masks_fake=None, label_smoothing: float = 0.9, lr: float = 2e-4,
           b1: float = 0.5, b2: float = 0.999, apply_update: bool = True):
    """D.train / G.eval: D(real) vs 0.9, D(G(noise)) vs 0, summed loss, Adam on D."""
    B = real.shape[0]
    p_real, c_real = d_forward(d_sd, real, image_size, masks_real)
    loss_real = bce(p_real, torch.full_like(p_real, label_smoothing))
    fake, _, _ = g_forward(g_sd, noise, image_size, train=False)       # torch.no_grad(), running-stat BN
    p_fake, c_fake = d_forward(d_sd, fake, image_size, masks_fake)
    loss_fake = bce(p_fake, torch.zeros_like(p_fake))
    g_real = d_backward(d_sd, c_real, bce_grad(p_real, torch.full_like(p_real, label_smoothing)), image_size,
                        masks_real)
    g_fake = d_backward(d_sd, c_fake, bce_grad(p_fake, torch.zeros_like(p_fake)), image_size, masks_fake)
    grads = {k: g_real[k] + g_fake[k] for k in g_real}
    if apply_update:
        d_opt.apply(d_sd, grads, lr, b1, b2)
    metrics = {
        "d_loss": float(loss_real + loss_fake), "d_loss_real": float(loss_real), "d_loss_fake": float(loss_fake),
        "d_real_acc": float((p_real > 0.5).float().mean()), "d_fake_acc": float((p_fake < 0.5).float().mean()),
        "d_real_mean": float(p_real.mean()), "d_fake_mean": float(p_fake.mean()),
    }
    return metrics, grads, {"fake": fake, "p_real": p_real, "p_fake": p_fake, "c_real": c_real, "c_fake": c_fake}


def g_step(g_sd, d_sd, g_opt: AdamState, noise: Tensor, image_size: int = 64, lr: float = 2e-4,
           b1: float = 0.5, b2: float = 0.999, apply_update: bool = True):
    """G.train / D.eval: BCE(D(G(noise)), 1), backward through D into G, Adam on G, BN running stats updated."""
    fake, gc, new_stats = g_forward(g_sd, noise, image_size, train=True)
    p, dc = d_forward(d_sd, fake, image_size, None)                     # D.eval(): no dropout
    ones = torch.ones_like(p)
    loss = bce(p, ones)
    dg = d_backward(d_sd, dc, bce_grad(p, ones), image_size, None, need_dx=True)
    grads = g_backward(g_sd, gc, dg["__dx"], image_size, train=True)
    grads.pop("__dz")
    for k, v in new_stats.items():
        g_sd[k] = v
    if apply_update:
        g_opt.apply(g_sd, grads, lr, b1, b2)
    d_side_grads = {k: v for k, v in dg.items() if not k.startswith("__")}
    return ({"g_loss": float(loss), "g_fake_mean": float(p.mean())}, grads,
            {"fake": fake, "p": p, "gc": gc, "dc": dc, "d_grads": d_side_grads, "dx": dg["__dx"]})


# --------------------------------------------------------------------------------------------
# Deterministic inputs shared by the golden generator, the tests and bench.py
# --------------------------------------------------------------------------------------------
def _hash_uniform(n: int, seed: int) -> Tensor:
    """Counter-based uniform(0,1) floats that do not depend on torch's RNG streams (splitmix64)."""
    import numpy as np
    base = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + base
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = ((x >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
    return torch.from_numpy(u)


def hash_normal(shape, seed: int, mean: float = 0.0, std: float = 1.0) -> Tensor:
    """Box-Muller on the counter-based uniforms; fp32, reproducible anywhere."""
    n = 1
    for s in shape:
        n *= s
    half = (n + 1) // 2
    u1 = _hash_uniform(half, 2 * seed + 1)
    u2 = _hash_uniform(half, 2 * seed + 2)
    r = torch.sqrt(-2.0 * torch.log(u1))
    zz = torch.cat([r * torch.cos(2 * math.pi * u2), r * torch.sin(2 * math.pi * u2)])[:n]
    return (zz * std + mean).to(torch.float32).reshape(*shape)


def hash_uniform(shape, seed: int) -> Tensor:
    n = 1
    for s in shape:
        n *= s
    return _hash_uniform(n, seed).to(torch.float32).reshape(*shape)


def make_state_dicts(image_size: int = 64, latent_dim: int = 100, seed: int = 0, in_ch: int = 1, width: int = 1):
    """DCGAN-style init (gen…:168-187, disc…:212-239) from the counter-based generator: N(0,.02) weights,
    zero bias, BN gamma ~ N(1,.02). Running stats are perturbed away from (0,1) so eval-mode BN is exercised."""
    gch, dch = g_channels(image_size, width), d_channels(image_size, in_ch, width)
    s = seed * 1000
    g: Dict[str, Tensor] = {}
    f0 = gch[0] * 16

    def bn(prefix, c, k):
        g[prefix + ".weight"] = hash_normal((c,), s + k, 1.0, 0.02)
        g[prefix + ".bias"] = hash_normal((c,), s + k + 1, 0.0, 0.02)
        g[prefix + ".running_mean"] = hash_normal((c,), s + k + 2, 0.0, 0.05)
        g[prefix + ".running_var"] = 0.05 + 0.1 * hash_uniform((c,), s + k + 3)
        g[prefix + ".num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)

    g["fc.0.weight"] = hash_normal((f0, latent_dim), s + 1, 0.0, 0.02)
    g["fc.0.bias"] = hash_normal((f0,), s + 2, 0.0, 0.02)
    bn("fc.1", f0, 10)
    for i in range(len(gch) - 1):
        g[f"upsample_blocks.{i}.block.0.weight"] = hash_normal((gch[i], gch[i + 1], 4, 4), s + 20 + i, 0.0, 0.02)
        bn(f"upsample_blocks.{i}.block.1", gch[i + 1], 30 + 10 * i)
    g["final_conv.0.weight"] = hash_normal((in_ch, gch[-1], 3, 3), s + 90, 0.0, 0.02)
    g["final_conv.0.bias"] = hash_normal((in_ch,), s + 91, 0.0, 0.02)
    d: Dict[str, Tensor] = {}
    for i in range(len(dch) - 1):
        d[f"conv_blocks.{i}.block.0.weight"] = hash_normal((dch[i + 1], dch[i], 4, 4), s + 100 + i, 0.0, 0.02)
        d[f"conv_blocks.{i}.block.0.bias"] = hash_normal((dch[i + 1],), s + 110 + i, 0.0, 0.02)
    d["classifier.0.weight"] = hash_normal((1, dch[-1] * 16), s + 120, 0.0, 0.02)
    d["classifier.0.bias"] = hash_normal((1,), s + 121, 0.0, 0.02)
    return g, d


def make_dropout_masks(batch: int, image_size: int, seed: int, p: float = DROPOUT_P, width: int = 1) -> List[Tensor]:
    dch = d_channels(image_size, 1, width)
    return [(hash_uniform((batch, c), seed * 100 + i) >= p).float() / (1.0 - p) for i, c in enumerate(dch[1:])]


def synthetic_signatures(n: int, size: int = 64, seed: int = 1234) -> Tensor:
    """Synthetic signature-like strokes: white background (+1), dark ink (-1), (n,1,S,S) fp32."""
    u = hash_uniform((n, 3, 16), seed)           # per image, per stroke: 16 curve parameters
    t = torch.linspace(0, 1, 192).view(1, 1, -1)
    ink = torch.zeros(n, size * size)

    def U(k, lo, hi):
        return (lo + (hi - lo) * u[:, :, k]).unsqueeze(-1)

    x = U(0, .4, .6) + U(2, .5, .8) * (t - .5)
    y = U(1, .35, .65) + U(3, -.1, .1) * (t - .5)
    for k in range(1, 4):
        x = x + U(3 + k, -.06, .06) / k * torch.sin(2 * math.pi * (k * t + U(9 + k, 0, 1)))
        y = y + U(6 + k, -.18, .18) / k * torch.sin(2 * math.pi * (k * t + U(12 + k, 0, 1)))
    xi = (x.clamp(0, 1) * (size - 1)).round().long().reshape(n, -1)
    yi = (y.clamp(0, 1) * (size - 1)).round().long().reshape(n, -1)
    ink.scatter_(1, yi * size + xi, 1.0)
    ink = ink.view(n, 1, size, size)
    ink = F.max_pool2d(ink, 3, 1, 1)
    ink = F.avg_pool2d(ink, 3, 1, 1)
    return (1.0 - 2.0 * ink).clamp(-1, 1)


def metric_batches(size: int = 64, n: int = 24) -> Dict[str, Tensor]:
    """Seeded image batches for the ink-statistics checks (shared by tests/golden/make_golden_metrics.py and the tests):
    generator-range [-1, 1] images, their [0, 1] counterpart (no rescale branch), and one with exact zeros / values at
    the rescale boundary."""
    sig = synthetic_signatures(n, size, seed=31)
    noise = hash_uniform((n, 1, size, size), 77) * 2.0 - 1.0
    soft = (0.6 * sig + 0.4 * noise).clamp(-1, 1)
    return {"signed": soft, "unit": (soft + 1.0) / 2.0,
            "edge": torch.where(soft.abs() < 0.02, torch.zeros_like(soft), soft)}
