import sys, os, torch
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "signature-gan_b200"))
from vanilla_gan_model import VanillaGAN
gan = VanillaGAN(latent_dim=100, image_size=64, device="cuda")
G = gan.generator; G.eval()
z = torch.randn(16384, 100, device="cuda")
with torch.no_grad():
    for _ in range(3): G(z)
    torch.cuda.synchronize()
    ctx = G._ctx
    ctx.profile(True)
    for _ in range(3): G(z)
    recs = ctx.profile_records()
    ctx.profile(False)
agg = {}
for name, ms, fl, by in recs:
    a = agg.setdefault(name, [0.0, 0.0, 0.0]); a[0] += ms / 3; a[1] = fl; a[2] = by
tot = sum(a[0] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:16s} {a[0]:.4f} ms  {a[1]/a[0]/1e9 if a[0] else 0:8.1f} TFLOP/s  {a[2]/a[0]/1e6 if a[0] else 0:8.1f} GB/s")
print("total", tot)
