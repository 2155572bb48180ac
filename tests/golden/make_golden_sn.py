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


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
