"""Golden fixture for the ablation script's LeakyReLU generator.

Runs ONLY in the build container (needs /root/reference). Imports the reference's UNMODIFIED
src/ablation_vanilla_gan_signatures.py (matplotlib, absent from this image, is satisfied by tests/stubs/matplotlib: the
module only plots in its report functions), builds its `ConfigurableGenerator(activation="leaky_relu")` (ablation…:216-328)
and its Discriminator, loads deterministic weights and records

  * the generator's training-mode and eval-mode outputs, the BatchNorm running statistics after the training-mode pass,
  * every parameter gradient of one G-loss backward through the Discriminator in train mode WITHOUT dropout noise
    (dropout p = 0 so that no mask has to be captured) — `AblationGANTrainer`'s G update (ablation…:441-448) with its
    smoothed label 0.9. Gradients are stored as compact probes (L2 norm, mean, 256 sampled elements per tensor).

    python tests/golden/make_golden_ablation.py        # rewrites tests/golden/ablation_leaky_{64,128}.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("SIGGAN_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
sys.path.insert(0, REF_SRC)

import siggan_oracle as O  # noqa: E402

N_PROBE = 256


def probe(t: torch.Tensor, seed: int = 7):
    t = t.detach().to(torch.float32).reshape(-1)
    n = t.numel()
    idx = (O.hash_uniform((min(N_PROBE, n),), seed + n % 9973) * n).long().clamp_(0, n - 1)
    return {"numel": n, "norm": float(t.double().norm()), "mean": float(t.double().mean()), "idx": idx,
            "vals": t[idx].clone()}


def main():
    import ablation_vanilla_gan_signatures as A      # the reference script, unmodified
    from discriminator_vanilla_gan import Discriminator
    assert os.path.realpath(A.__file__).startswith(os.path.realpath(REF_SRC))
    for size, B in ((64, 8), (128, 8)):
        g_sd, d_sd = O.make_state_dicts(size, 100, seed=11)
        G = A.ConfigurableGenerator(latent_dim=100, output_size=size, activation="leaky_relu")
        D = Discriminator(input_size=size, dropout=0.0)
        G.load_state_dict(g_sd)
        D.load_state_dict(d_sd)
        z = O.hash_normal((B, 100), 77)
        out = {"size": size, "B": B, "seed": 11, "z_seed": 77, "leaky_slope": 0.2}
        G.eval()
        with torch.no_grad():
            out["eval.image"] = G(z).clone()
        G.train()
        D.train()
        img = G(z)
        out["train.image"] = img.detach().clone()
        out["train.stats"] = {k: v.detach().clone() for k, v in G.state_dict().items() if "running" in k or "tracked" in k}
        pred = D(img)
        loss = torch.nn.BCELoss()(pred, torch.full((B, 1), 0.9))
        loss.backward()
        out["g_loss"] = float(loss.detach())
        out["grads"] = {k: probe(p.grad) for k, p in G.named_parameters()}
        torch.save(out, os.path.join(HERE, f"ablation_leaky_{size}.pt"))
        print(f"wrote ablation_leaky_{size}.pt: g_loss {out['g_loss']:.6f}, "
              f"{sum(v['numel'] for v in out['grads'].values())} gradient elements probed")


if __name__ == "__main__":
    main()
