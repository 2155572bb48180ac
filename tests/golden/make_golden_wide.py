"""Golden fixture for the "2x hidden width" variant of BASELINE.json configs[4].

The reference has no such configuration (its `base_features` argument is inert: any value but 256 fails in the first
ConvTranspose2d), so the fixture comes from a model ASSEMBLED FROM THE REFERENCE'S OWN BLOCKS: its
channel-parameterised `UpsampleBlock` (src/generator_vanilla_gan.py:17-66) and `DownsampleBlock`
(src/discriminator_vanilla_gan.py:18-81), imported unmodified from /root/reference/src, stacked exactly as
`Generator.__init__` / `Discriminator.__init__` stack them (gen…:124-163, disc…:131-207) with every channel count doubled.
Records, at 64x64 and B = 16: eval / training-mode generator images, running statistics, the discriminator's probabilities
(eval and with captured Dropout2d masks), and probes of every parameter gradient of the G loss (labels 1) and the D loss
(real vs 0.9, fake vs 0).

    python tests/golden/make_golden_wide.py        # rewrites tests/golden/wide2_64.pt
"""
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("SIGGAN_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF_SRC)

import siggan_oracle as O  # noqa: E402

N_PROBE = 256


def probe(t, seed=7):
    t = t.detach().to(torch.float32).reshape(-1)
    n = t.numel()
    idx = (O.hash_uniform((min(N_PROBE, n),), seed + n % 9973) * n).long().clamp_(0, n - 1)
    return {"numel": n, "norm": float(t.double().norm()), "mean": float(t.double().mean()), "idx": idx, "vals": t[idx].clone()}


def build(size, width):
    from generator_vanilla_gan import UpsampleBlock            # the reference's blocks, unmodified
    from discriminator_vanilla_gan import DownsampleBlock
    assert os.path.realpath(sys.modules["generator_vanilla_gan"].__file__).startswith(os.path.realpath(REF_SRC))
    gch, dch = O.g_channels(size, width), O.d_channels(size, 1, width)

    class G(nn.Module):                                         # gen…:124-163, 189-209 with the doubled ladder
        def __init__(self):
            super().__init__()
            f0 = gch[0] * 16
            self.fc = nn.Sequential(nn.Linear(100, f0), nn.BatchNorm1d(f0), nn.ReLU(inplace=True))
            self.upsample_blocks = nn.Sequential(*[UpsampleBlock(a, b) for a, b in zip(gch[:-1], gch[1:])])
            self.final_conv = nn.Sequential(nn.Conv2d(gch[-1], 1, kernel_size=3, stride=1, padding=1, bias=True), nn.Tanh())

        def forward(self, z):
            x = self.fc(z).view(-1, gch[0], 4, 4)
            return self.final_conv(self.upsample_blocks(x))

    class D(nn.Module):                                         # disc…:131-207, 241-260 with the doubled ladder
        def __init__(self):
            super().__init__()
            self.conv_blocks = nn.Sequential(*[DownsampleBlock(a, b) for a, b in zip(dch[:-1], dch[1:])])
            self.classifier = nn.Sequential(nn.Linear(dch[-1] * 16, 1), nn.Sigmoid())

        def forward(self, x):
            return self.classifier(self.conv_blocks(x).flatten(1))

    return G(), D()


def main():
    size, width, B, seed = 64, 2, 16, 21
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=seed, width=width)
    G, D = build(size, width)
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    z = O.hash_normal((B, 100), 55)
    real = O.synthetic_signatures(B, size, seed=8)
    out = {"size": size, "width": width, "B": B, "seed": seed, "z_seed": 55, "real_seed": 8}
    G.eval(); D.eval()
    with torch.no_grad():
        img = G(z)
        out["eval.image"] = img.clone()
        out["eval.prob_fake"] = D(img).clone()
        out["eval.prob_real"] = D(real).clone()
    # ---- G loss (vanilla…:273-306: G.train, D.eval, labels 1)
    G.train(); D.eval()
    img = G(z)
    out["train.image"] = img.detach().clone()
    out["train.stats"] = {k: v.detach().clone() for k, v in G.state_dict().items() if "running" in k or "tracked" in k}
    loss = nn.BCELoss()(D(img), torch.ones(B, 1))
    loss.backward()
    out["g_loss"] = float(loss.detach())
    out["g_grads"] = {k: probe(p.grad) for k, p in G.named_parameters()}
    # ---- D loss (vanilla…:203-236) with the Dropout2d masks captured per block
    D.train(); D.zero_grad()
    rec = []
    hooks = [m.register_forward_hook(lambda mod, i, o: rec.append((o.detach().abs().amax(dim=(2, 3)) > 0).float() / (1 - mod.p)))
             for m in D.modules() if isinstance(m, nn.Dropout2d)]
    p_real = D(real)
    masks_real = list(rec); rec.clear()
    p_fake = D(img.detach())
    masks_fake = list(rec)
    for h in hooks:
        h.remove()
    d_loss = nn.BCELoss()(p_real, torch.full((B, 1), 0.9)) + nn.BCELoss()(p_fake, torch.zeros(B, 1))
    d_loss.backward()
    out["d_loss"] = float(d_loss.detach())
    out["masks_real"], out["masks_fake"] = masks_real, masks_fake
    out["train.prob_real"], out["train.prob_fake"] = p_real.detach().clone(), p_fake.detach().clone()
    out["d_grads"] = {k: probe(p.grad) for k, p in D.named_parameters()}
    torch.save(out, os.path.join(HERE, f"wide{width}_{size}.pt"))
    print(f"wrote wide{width}_{size}.pt: g_loss {out['g_loss']:.6f} d_loss {out['d_loss']:.6f}; "
          f"G {sum(p.numel() for p in G.parameters())} / D {sum(p.numel() for p in D.parameters())} parameters")


if __name__ == "__main__":
    main()
