"""Generate the golden fixtures that pin oracle/siggan_oracle.py to the reference.

Runs ONLY in the build container (needs /root/reference). It imports the reference's
generator_vanilla_gan / discriminator_vanilla_gan / vanilla_gan_model modules unmodified, loads
deterministic weights (oracle.make_state_dicts), drives the reference's own forward passes and
`train_discriminator_step` / `train_generator_step` on CPU fp32, and stores compact probes of every
tensor on the path (shape, L2 norm, mean, 96 sampled elements) plus all scalar metrics.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("SIGGAN_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF_SRC)

import siggan_oracle as O  # noqa: E402

N_PROBE = 96


def probe(t: torch.Tensor, seed: int = 7):
    t = t.detach().to(torch.float32).reshape(-1)
    n = t.numel()
    idx = (O.hash_uniform((min(N_PROBE, n),), seed + n % 9973) * n).long().clamp_(0, n - 1)
    return {"numel": n, "norm": float(t.double().norm()), "mean": float(t.double().mean()), "idx": idx,
            "vals": t[idx].clone()}


def capture_dropout(disc):
    """Record the Dropout2d keep-scale per (n, c) of every block, in call order."""
    rec, hooks = [], []

    def hook(mod, inp, out):
        keep = (out.detach().abs().amax(dim=(2, 3)) > 0).float() / (1.0 - mod.p)
        rec.append(keep)

    for m in disc.modules():
        if isinstance(m, torch.nn.Dropout2d):
            hooks.append(m.register_forward_hook(hook))
    return rec, hooks


def layer_outputs(seq_modules):
    acts, hooks = {}, []
    for name, mod in seq_modules:
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: acts.__setitem__(name, o.detach().clone())))
    return acts, hooks


def make_forward(size: int, B: int = 4):
    from generator_vanilla_gan import Generator
    from discriminator_vanilla_gan import Discriminator
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=1)
    G, D = Generator(100, size), Discriminator(size)
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    z = O.hash_normal((B, 100), 11)
    real = O.synthetic_signatures(B, size, seed=5)
    out = {"size": size, "B": B}
    nb = len(O.g_channels(size)) - 1
    # ---- G eval / train with per-layer activations
    for mode in ("eval", "train"):
        G.load_state_dict(g_sd)
        G.train(mode == "train")
        names = [("fc", G.fc)] + [(f"up{i}", G.upsample_blocks[i]) for i in range(nb)]
        acts, hooks = layer_outputs(names)
        with torch.no_grad():
            img = G(z)
        for h in hooks:
            h.remove()
        out[f"g_{mode}.out"] = probe(img)
        for k, v in acts.items():
            out[f"g_{mode}.{k}"] = probe(v)
        if mode == "train":
            for k, v in G.state_dict().items():
                if "running" in k or "num_batches" in k:
                    out[f"g_train.stats.{k}"] = v.clone() if v.numel() == 1 else probe(v)
    # ---- D eval and D train (captured dropout masks)
    D.eval()
    names = [(f"c{i}", D.conv_blocks[i]) for i in range(len(D.conv_blocks))]
    acts, hooks = layer_outputs(names)
    with torch.no_grad():
        p = D(real)
        feats = D.forward_features(real)
    for h in hooks:
        h.remove()
    out["d_eval.prob"] = p.clone()
    out["d_eval.feat"] = probe(feats)
    for k, v in acts.items():
        out[f"d_eval.{k}"] = probe(v)
    D.train()
    torch.manual_seed(123)
    rec, hooks = capture_dropout(D)
    with torch.no_grad():
        p = D(real)
    for h in hooks:
        h.remove()
    out["d_train.masks"] = [m.clone() for m in rec]
    out["d_train.prob"] = p.clone()
    # ---- one autograd backward of each net, for the hand-written backward formulas
    G.load_state_dict(g_sd)
    G.train()
    D.eval()
    img = G(z)
    pr = D(img)
    loss = torch.nn.BCELoss()(pr, torch.ones_like(pr))
    loss.backward()
    out["bwd.loss"] = float(loss)
    for k, v in G.named_parameters():
        out[f"bwd.g_grad.{k}"] = probe(v.grad)
    for k, v in D.named_parameters():
        out[f"bwd.d_grad.{k}"] = probe(v.grad)
    torch.save(out, os.path.join(HERE, f"forward_{size}.pt"))
    print(f"forward_{size}.pt: {len(out)} entries")


def make_steps(size: int, B: int = 4, steps: int = 3):
    from vanilla_gan_model import VanillaGAN
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=2)
    gan = VanillaGAN(latent_dim=100, image_size=size, device="cpu")
    gan.generator.load_state_dict(g_sd)
    gan.discriminator.load_state_dict(d_sd)
    out = {"size": size, "B": B, "steps": steps, "metrics": [], "masks": []}
    torch.manual_seed(99)
    for s in range(steps):
        real = O.synthetic_signatures(B, size, seed=100 + s)
        nd = O.hash_normal((B, 100), 200 + s)
        ng = O.hash_normal((B, 100), 300 + s)
        rec, hooks = capture_dropout(gan.discriminator)
        md = gan.train_discriminator_step(real, noise=nd)          # vanilla…:180-252
        for h in hooks:
            h.remove()
        nblk = len(gan.discriminator.conv_blocks)
        out["masks"].append({"real": [m.clone() for m in rec[:nblk]], "fake": [m.clone() for m in rec[nblk:]]})
        for k, p in gan.discriminator.named_parameters():
            out[f"s{s}.d_grad.{k}"] = probe(p.grad)
            out[f"s{s}.d_param.{k}"] = probe(p)
        mg = gan.train_generator_step(B, noise=ng)                 # vanilla…:254-306
        for k, p in gan.generator.named_parameters():
            out[f"s{s}.g_grad.{k}"] = probe(p.grad)
            out[f"s{s}.g_param.{k}"] = probe(p)
        for k, v in gan.generator.state_dict().items():
            if "running" in k:
                out[f"s{s}.g_stats.{k}"] = probe(v)
        md.update(mg)
        out["metrics"].append(md)
    # Adam state layout (torch.optim.Adam.state_dict) after `steps` updates
    st = gan.d_optimizer.state_dict()
    out["d_adam.keys"] = sorted(st["state"][0].keys())
    out["d_adam.step"] = float(st["state"][0]["step"])
    out["d_adam.exp_avg.0"] = probe(st["state"][0]["exp_avg"])
    out["d_adam.exp_avg_sq.0"] = probe(st["state"][0]["exp_avg_sq"])
    out["d_adam.param_group_keys"] = sorted(k for k in st["param_groups"][0].keys())
    torch.save(out, os.path.join(HERE, f"steps_{size}.pt"))
    print(f"steps_{size}.pt: {len(out)} entries; metrics[0]={out['metrics'][0]}")


def make_contract():
    """State-dict keys / shapes / dtypes and module attributes that the drop-in boundary must reproduce."""
    from generator_vanilla_gan import Generator
    from discriminator_vanilla_gan import Discriminator
    out = {}
    for size in (64, 128):
        G, D = Generator(100, size), Discriminator(size)
        out[f"g{size}"] = [(k, tuple(v.shape), str(v.dtype)) for k, v in G.state_dict().items()]
        out[f"d{size}"] = [(k, tuple(v.shape), str(v.dtype)) for k, v in D.state_dict().items()]
        out[f"g{size}.nparams"], out[f"d{size}.nparams"] = G.get_num_params(), D.get_num_params()
    Dsn = Discriminator(64, use_spectral_norm=True)
    out["d64sn"] = [(k, tuple(v.shape), str(v.dtype)) for k, v in Dsn.state_dict().items()]
    torch.save(out, os.path.join(HERE, "contract.pt"))
    print("contract.pt written")


if __name__ == "__main__":
    torch.set_num_threads(8)
    make_contract()
    for size in (64, 128):
        make_forward(size)
    make_steps(64)
    make_steps(128, B=2, steps=2)
