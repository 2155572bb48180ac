"""Golden fixture for the input-pipeline augmentation (reference src/data_loader_signatures.py:154-219).

Runs ONLY in the build container (needs /root/reference). Pushes 8-bit synthetic signatures through the reference's
unmodified `get_train_transforms(...)` Compose (PIL + torchvision), recording the parameters its RandomRotation /
RandomAffine / RandomHorizontalFlip drew, and stores inputs, parameters and the float32 outputs.

    python tests/golden/make_golden_augment.py        # rewrites tests/golden/augment_{64,128}.pt
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O  # noqa: E402  (also puts the reference's src/ on sys.path)


def signatures_u8(n, size, seed):
    x = O.synthetic_signatures(n, size, seed=seed)[:, 0]
    return ((x + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)


def main(size, n, flip):
    from torchvision import transforms
    import data_loader_signatures as R
    rec = {"angle": [], "scale": [], "flip": []}
    rot_get, aff_get = transforms.RandomRotation.get_params, transforms.RandomAffine.get_params

    def rot_params(degrees):
        a = rot_get(degrees)
        rec["angle"].append(a)
        return a

    def aff_params(*args, **kw):
        out = aff_get(*args, **kw)
        assert out[0] == 0.0 and tuple(out[1]) == (0, 0) and tuple(out[3]) == (0.0, 0.0)
        rec["scale"].append(out[2])
        return out

    transforms.RandomRotation.get_params = staticmethod(rot_params)
    transforms.RandomAffine.get_params = staticmethod(aff_params)
    try:
        tf = R.get_train_transforms(image_size=size, horizontal_flip=flip)          # defaults: ±5°, scale (0.9, 1.1)
        imgs = signatures_u8(n, size, seed=21)
        torch.manual_seed(77)
        outs = []
        for i in range(n):
            pil = Image.fromarray(imgs[i].numpy())
            if flip:                                   # RandomHorizontalFlip draws torch.rand(1) < p inside forward
                state = torch.get_rng_state()
                probe = transforms.Compose(tf.transforms[:3])(pil)      # resize, rotation, affine (consumes the same draws)
                flipped = bool(torch.rand(1) < 0.5)
                torch.set_rng_state(state)
                rec["angle"].pop(), rec["scale"].pop()
                rec["flip"].append(int(flipped))
                del probe
            outs.append(tf(pil))
    finally:
        transforms.RandomRotation.get_params, transforms.RandomAffine.get_params = rot_get, aff_get
    out = {"size": size, "images": imgs, "angles": torch.tensor(rec["angle"], dtype=torch.float64),
           "scales": torch.tensor(rec["scale"], dtype=torch.float64),
           "flips": torch.tensor(rec["flip"], dtype=torch.uint8) if flip else None, "out": torch.stack(outs)}
    name = f"augment_{size}{'_flip' if flip else ''}.pt"
    torch.save(out, os.path.join(HERE, name))
    print(name, tuple(out["out"].shape), "angles", rec["angle"][:3], "scales", rec["scale"][:3], "flips", rec["flip"][:6])


if __name__ == "__main__":
    main(64, 8, False)
    main(64, 4, True)
    main(128, 2, False)
