"""Golden fixture for the spectral-norm Discriminator variant (reference disc…:61-62, 201-202).

Runs ONLY in the build container (needs /root/reference). Drives the reference's unmodified
`Discriminator(use_spectral_norm=True)` on CPU fp32: two training-mode forwards (power iteration, captured
Dropout2d masks), the backward of BCE(prob, 0.9) through the first one, then an eval-mode forward.

    python tests/golden/make_golden_sn.py         # rewrites tests/golden/sn_64.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O, capture_dropout, probe  # noqa: E402  (also puts the reference on sys.path)


def main(size: int = 64, B: int = 4):
    from discriminator_vanilla_gan import Discriminator
    sd = O.make_sn_state_dict(size, seed=3)
    D = Discriminator(size, use_spectral_norm=True)
    D.load_state_dict(sd)
    x = O.synthetic_signatures(B, size, seed=7)
    out = {"size": size, "B": B, "keys": list(D.state_dict().keys())}
    D.train()
    torch.manual_seed(321)
    rec, hooks = capture_dropout(D)
    p1 = D(x)
    loss = torch.nn.BCELoss()(p1, torch.full_like(p1, 0.9))
    loss.backward()
    out["train1.masks"] = [m.clone() for m in rec]
    out["train1.prob"] = p1.detach().clone()
    out["train1.loss"] = float(loss)
    for k, p in D.named_parameters():
        out[f"train1.grad.{k}"] = probe(p.grad)
    for k, v in D.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            out[f"train1.buf.{k}"] = v.clone()
    for name in O.sn_layer_names(size):
        mod = D.get_submodule(name)
        out[f"train1.weight.{name}"] = probe(mod.weight)          # weight_orig / sigma of that forward
    rec.clear()
    with torch.no_grad():
        p2 = D(x)
    out["train2.masks"] = [m.clone() for m in rec]
    out["train2.prob"] = p2.clone()
    for h in hooks:
        h.remove()
    for k, v in D.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            out[f"train2.buf.{k}"] = v.clone()
    D.eval()
    with torch.no_grad():
        out["eval.prob"] = D(x).clone()
        out["eval.feat"] = probe(D.forward_features(x))
    torch.save(out, os.path.join(HERE, f"sn_{size}.pt"))
    print(f"sn_{size}.pt: {len(out)} entries; prob={p1.detach().flatten().tolist()} loss={float(loss):.6f}")




def make_sn_steps(size: int = 64, B: int = 4, steps: int = 2):
    """The reference's VanillaGAN(use_spectral_norm=True) through train_discriminator_step / train_generator_step
    (vanilla…:180-306) with injected noise and captured dropout masks -> tests/golden/sn_steps_64.pt."""
    from vanilla_gan_model import VanillaGAN
    g_sd, _ = O.make_state_dicts(size, 100, seed=6)
    d_sd = O.make_sn_state_dict(size, seed=6)
    gan = VanillaGAN(latent_dim=100, image_size=size, use_spectral_norm=True, device="cpu")
    gan.generator.load_state_dict(g_sd)
    gan.discriminator.load_state_dict(d_sd)
    out = {"size": size, "B": B, "steps": steps, "metrics": [], "masks": []}
    torch.manual_seed(55)
    for s in range(steps):
        real = O.synthetic_signatures(B, size, seed=400 + s)
        nd, ng = O.hash_normal((B, 100), 500 + s), O.hash_normal((B, 100), 600 + s)
        rec, hooks = capture_dropout(gan.discriminator)
        md = gan.train_discriminator_step(real, noise=nd)
        for h in hooks:
            h.remove()
        nblk = len(gan.discriminator.conv_blocks)
        out["masks"].append({"real": [m.clone() for m in rec[:nblk]], "fake": [m.clone() for m in rec[nblk:]]})
        for k, p in gan.discriminator.named_parameters():
            out[f"s{s}.d_grad.{k}"] = probe(p.grad)
            out[f"s{s}.d_param.{k}"] = probe(p)
        for k, v in gan.discriminator.state_dict().items():
            if k.endswith(("weight_u", "weight_v")):
                out[f"s{s}.d_buf.{k}"] = v.clone()
        mg = gan.train_generator_step(B, noise=ng)
        for k, p in gan.generator.named_parameters():
            out[f"s{s}.g_grad.{k}"] = probe(p.grad)
            out[f"s{s}.g_param.{k}"] = probe(p)
        md.update(mg)
        out["metrics"].append(md)
    torch.save(out, os.path.join(HERE, f"sn_steps_{size}.pt"))
    print(f"sn_steps_{size}.pt: {len(out)} entries; metrics[0]={out['metrics'][0]}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
    make_sn_steps()
