"""Second-half means of the 1000-step curves for each reference seed (the quantity tests/test_loss_band.py bounds)."""
import os, sys, torch
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for p in ("signature-gan_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import siggan_oracle as O
from _util import make_gan
ref = torch.load(os.path.join(ROOT, "tests/golden/loss_band_64.pt"))
size, B, steps, pool_n, seeds = ref["size"], ref["batch"], ref["steps"], ref["pool"], ref["seeds"]
KEYS = ("d_loss", "g_loss", "d_real_mean", "d_fake_mean")
for s in seeds:
    print("ref seed", s, {k: round(float(sum(ref["curves"][s][k][steps // 2:])) / (steps - steps // 2), 3) for k in KEYS})
pool = O.synthetic_signatures(pool_n, size, seed=1234).cuda()
for seed in list(seeds) + [s + 100 for s in seeds]:
    gan, _, _ = make_gan(size, seed if seed in seeds else seeds[0], "bf16")
    torch.manual_seed(seed)
    perm = torch.randperm(pool_n, generator=torch.Generator().manual_seed(seed))
    cur = {k: [] for k in KEYS}
    for i in range(steps):
        idx = perm[(i * B) % pool_n:(i * B) % pool_n + B].cuda()
        m = gan.train_step(pool[idx])
        for k in KEYS:
            cur[k].append(m[k])
    print("ours seed", seed, {k: round(float(sum(cur[k][steps // 2:])) / (steps - steps // 2), 3) for k in KEYS})
