"""Generate the reference loss curves that the 1k-step loss-band test compares against.

Runs ONLY in the build container (needs /root/reference). Drives the UNMODIFIED reference
`vanilla_gan_model.VanillaGAN.train_step` on CPU fp32 for STEPS steps of batch B on the deterministic synthetic
signature pool, once per seed, and stores the per-step metrics (d_loss, g_loss, d_real_mean, d_fake_mean).

    python tests/golden/make_loss_band.py            # rewrites tests/golden/loss_band_64.pt  (~15 min of CPU)

The initial weights come from oracle.make_state_dicts(seed) so that the CUDA run starts from the same point; noise
and dropout masks are NOT shared (CPU and CUDA RNG streams differ): the test compares smoothed curves to the
envelope over the reference seeds, not step-by-step values.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("SIGGAN_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF_SRC)

import siggan_oracle as O  # noqa: E402

SIZE, B, STEPS, POOL = 64, 32, 1000, 2048
SEEDS = (11, 12, 13)
KEYS = ("d_loss", "g_loss", "d_real_mean", "d_fake_mean")


def run(seed: int):
    from vanilla_gan_model import VanillaGAN
    torch.manual_seed(seed)
    gan = VanillaGAN(latent_dim=100, image_size=SIZE, device="cpu")
    g_sd, d_sd = O.make_state_dicts(SIZE, 100, seed=seed)
    gan.generator.load_state_dict(g_sd)
    gan.discriminator.load_state_dict(d_sd)
    pool = O.synthetic_signatures(POOL, SIZE, seed=1234)
    perm = torch.randperm(POOL, generator=torch.Generator().manual_seed(seed))
    out = {k: [] for k in KEYS}
    t0 = time.time()
    for i in range(STEPS):
        idx = perm[(i * B) % POOL:(i * B) % POOL + B]
        m = gan.train_step(pool[idx])
        for k in KEYS:
            out[k].append(float(m[k]))
        if i % 100 == 0:
            print(f"seed {seed} step {i} d={m['d_loss']:.4f} g={m['g_loss']:.4f} ({time.time() - t0:.0f}s)", flush=True)
    return {k: torch.tensor(v) for k, v in out.items()}


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("LOSS_BAND_THREADS", "4")))
    res = {"size": SIZE, "batch": B, "steps": STEPS, "pool": POOL, "seeds": list(SEEDS), "curves": {}}
    for s in SEEDS:
        res["curves"][s] = run(s)
    torch.save(res, os.path.join(HERE, f"loss_band_{SIZE}.pt"))
    print("saved")
