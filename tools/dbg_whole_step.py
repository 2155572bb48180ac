import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "signature-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import siggan_oracle as O
import _siggan_lib as L
from _util import make_gan
B, size = 32, 64
reals = [O.synthetic_signatures(B, size, seed=70 + i).cuda() for i in range(5)]
names = ["g_flat", "d_flat", "g_stats", "g_m", "g_v", "d_m", "d_v", "metrics"]
snaps = {}
for whole in (True, False):
    gan, _, _ = make_gan(size, 9, "bf16")
    torch.manual_seed(123)
    L.DROPOUT.offset = 0
    for i, x in enumerate(reals):
        if whole:
            m = gan.train_step_async(x)
        else:
            gan.discriminator_step_async(x)
            m = gan.generator_step_async(x.size(0))
        torch.cuda.synchronize()
        snaps[(whole, i)] = [t.clone() for t in (gan.generator._flat.flat, gan.discriminator._flat.flat, gan.generator._flat.stats,
                             gan.g_optimizer._m, gan.g_optimizer._v, gan.d_optimizer._m, gan.d_optimizer._v, m)]
for i in range(5):
    print("step", i, {n: (float((a - b).abs().max()), int((a != b).sum())) for n, a, b in zip(names, snaps[(True, i)], snaps[(False, i)])})
