"""Per-tensor gradient error report (CUDA vs oracle) for both precisions; diagnostic, not a test."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "signature-gan_b200"), os.path.join(ROOT, "tests")]
import torch
import siggan_oracle as O
from _util import make_gan, rel_err

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
for precision in ("fp32", "bf16"):
    gan, g_sd, d_sd = make_gan(size, 1, precision)
    G, D = gan.generator, gan.discriminator
    z = O.hash_normal((B, 100), 11)
    G.train(); D.eval()
    fake = G(z.cuda()); pred = D(fake)
    loss = gan.criterion(pred, torch.ones(B, 1, device="cuda")); loss.backward()
    img, gc, _ = O.g_forward(g_sd, z, size, train=True)
    pr, dc = O.d_forward(d_sd, img, size, None)
    ones = torch.ones_like(pr)
    dg = O.d_backward(d_sd, dc, O.bce_grad(pr, ones), size, None, need_dx=True)
    gg = O.g_backward(g_sd, gc, dg["__dx"], size, train=True)
    print(f"== {precision} size={size} B={B} loss {float(loss):.6f} vs {float(O.bce(pr, ones)):.6f}")
    for k, p in list(D.named_parameters()):
        print(f"  D {k:40s} rel={rel_err(p.grad, dg[k]):.3e} |ref|={dg[k].norm():.3e}")
    for k, p in list(G.named_parameters()):
        print(f"  G {k:40s} rel={rel_err(p.grad, gg[k]):.3e} |ref|={gg[k].norm():.3e}")
