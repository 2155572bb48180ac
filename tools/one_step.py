"""One or two fused training steps at the benchmark batch (driver for targeted ncu captures: ncu -k regex:<kernel> ...)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "signature-gan_b200"))
from vanilla_gan_model import VanillaGAN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gan = VanillaGAN(latent_dim=100, image_size=64, device="cuda")
real = torch.rand(B, 1, 64, 64, device="cuda") * 2 - 1
for _ in range(2):
    m = gan.train_step(real)
torch.cuda.synchronize()
print("ok", m["d_loss"], m["g_loss"])
