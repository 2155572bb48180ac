"""The "2x hidden width" variant of BASELINE.json configs[4] (every channel count doubled; bf16 tensor-core mode) against
the oracle at width = 2 and against tests/golden/wide2_64.pt, which a model assembled from the reference's own
UpsampleBlock / DownsampleBlock produced (tests/golden/make_golden_wide.py)."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import check_grads, make_gan, rel_err, to64, tol

pytestmark = pytest.mark.gpu


def test_wide_modules_against_assembled_reference_fixture(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "wide2_64.pt"), weights_only=False)
    size, width, B = gold["size"], gold["width"], gold["B"]
    gan, g_sd, d_sd = make_gan(size, gold["seed"], "bf16", width=width)
    G, D = gan.generator, gan.discriminator
    assert G.get_num_params() == 3663425 and D.get_num_params() == 11030401
    z = O.hash_normal((B, 100), gold["z_seed"]).cuda()
    real = O.synthetic_signatures(B, size, seed=gold["real_seed"]).cuda()
    t = tol("bf16")
    G.eval(); D.eval()
    with torch.no_grad():
        img = G(z)
        assert rel_err(img, gold["eval.image"]) <= t
        assert rel_err(D(img), gold["eval.prob_fake"]) <= t and rel_err(D(real), gold["eval.prob_real"]) <= t
        u8 = G.sample_uint8(z)
        want = ((gold["eval.image"] + 1.0) * 127.5).clip(0, 255).to(torch.uint8)
        assert (u8.cpu().int() - want.int()).abs().max() <= 2
    # G loss through the module / autograd path (vanilla…:273-306)
    G.train(); D.eval()
    G.zero_grad()
    img = G(z)
    assert rel_err(img, gold["train.image"]) <= t
    sd = G.state_dict()
    for k, ref in gold["train.stats"].items():
        if "tracked" in k:
            assert int(sd[k]) == int(ref), k
        else:
            assert rel_err(sd[k], ref) <= t, k
    pred = D(img)
    loss = torch.nn.functional.binary_cross_entropy(pred, torch.ones_like(pred))
    assert abs(float(loss.detach()) - gold["g_loss"]) <= 2e-2
    loss.backward()
    ref = {}
    zc = z.cpu()
    for name, sg_, sd_, zz in (("f32", g_sd, d_sd, zc), ("f64", to64(g_sd), to64(d_sd), zc.double())):
        im, gc, _ = O.g_forward(sg_, zz, size, train=True)
        pr, dc = O.d_forward(sd_, im, size, None)
        dg = O.d_backward(sd_, dc, O.bce_grad(pr, torch.ones_like(pr)), size, None, need_dx=True)
        ref[name] = O.g_backward(sg_, gc, dg["__dx"], size, train=True)
    check_grads("bf16", "G", {k: p.grad for k, p in G.named_parameters()}, ref["f32"], ref["f64"],
                skip=("fc.0.bias", "final_conv.0.bias"))
    # D loss with the reference's captured Dropout2d masks (vanilla…:203-236)
    D.train(); D.zero_grad()
    D.mask_override = [m.cuda() for m in gold["masks_real"]]
    p_real = D(real)
    D.mask_override = [m.cuda() for m in gold["masks_fake"]]
    p_fake = D(img.detach())
    D.mask_override = None
    assert rel_err(p_real, gold["train.prob_real"]) <= t and rel_err(p_fake, gold["train.prob_fake"]) <= t
    bce = torch.nn.functional.binary_cross_entropy
    d_loss = bce(p_real, torch.full_like(p_real, 0.9)) + bce(p_fake, torch.zeros_like(p_fake))
    assert abs(float(d_loss.detach()) - gold["d_loss"]) <= 2e-2
    d_loss.backward()
    for k, p in D.named_parameters():
        pr = gold["d_grads"][k]
        got = p.grad.detach().double().cpu().reshape(-1)[pr["idx"]]
        e = (got - pr["vals"].double()).norm().item() / max(pr["vals"].double().norm().item(), 1e-12)
        assert e <= tol("bf16", "grad_d") * (2 if k.endswith("bias") else 1), (k, e)


def test_wide_fused_steps_against_oracle():
    """sg_train_step at width 2 (D step + G step with injected noise and masks) against the oracle: metrics and parameters."""
    size, width, B = 64, 2, 16
    gan, g_sd, d_sd = make_gan(size, 23, "bf16", width=width)
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    for s in range(2):
        real = O.synthetic_signatures(B, size, seed=300 + s)
        nd, ng = O.hash_normal((B, 100), 400 + s), O.hash_normal((B, 100), 500 + s)
        mk_r = O.make_dropout_masks(B, size, 600 + s, width=width)
        mk_f = O.make_dropout_masks(B, size, 700 + s, width=width)
        gan.mask_override = {"real": mk_r, "fake": mk_f}
        md = gan.train_discriminator_step(real.cuda(), noise=nd.cuda())
        mg = gan.train_generator_step(B, noise=ng.cuda())
        od, _, _ = O.d_step(g_sd, d_sd, d_opt, real, nd, size, mk_r, mk_f)
        og, _, _ = O.g_step(g_sd, d_sd, g_opt, ng, size)
        for k, v in {**od, **og}.items():
            got = {**md, **mg}[k]
            assert abs(got - v) <= 2e-2 * max(1.0, abs(v)) + (0.13 if k.endswith("acc") else 0), (s, k, got, v)
    gsd, dsd = gan.generator.state_dict(), gan.discriminator.state_dict()
    for k in O.trainable_names(d_sd):
        assert rel_err(dsd[k], d_sd[k]) <= tol("bf16", "param"), k
    for k in O.trainable_names(g_sd):
        if k != "fc.0.bias":      # its gradient is mathematically zero (bias in front of BatchNorm): Adam steps on rounding noise
            assert rel_err(gsd[k], g_sd[k]) <= tol("bf16", "param"), k


@pytest.mark.parametrize("size", [64, 128])
def test_wide_whole_step_runs_and_fp32_mode_refuses(size):
    gan, _, _ = make_gan(size, 4, "bf16", width=2)
    real = O.synthetic_signatures(32 if size == 64 else 16, size, seed=2).cuda()
    for _ in range(4):       # eager, capture, replay
        m = gan.train_step(real)
    assert all(v == v and abs(v) < 50 for v in m.values())
    assert all(bool(torch.isfinite(p).all()) for p in list(gan.generator.parameters()) + list(gan.discriminator.parameters()))
    from generator_vanilla_gan import Generator
    G = Generator(output_size=size, base_features=512).to("cuda").set_precision("fp32")
    with pytest.raises(RuntimeError, match="bf16"):
        G(torch.randn(4, 100, device="cuda"))
