"""GPU parity of the spectral-norm Discriminator variant (reference disc…:61-62, 201-202; built by
ablation_vanilla_gan_signatures.py:367-371 and trained there with a stock torch.optim.Adam): the drop-in
`Discriminator(use_spectral_norm=True)` -> sg_spectral_norm_weight / sg_d_forward / sg_d_backward /
sg_spectral_norm_backward against the CPU oracle and the reference-generated fixture tests/golden/sn_64.pt."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import rel_err, to64, tol

pytestmark = pytest.mark.gpu


def _make(size, seed, precision):
    from discriminator_vanilla_gan import Discriminator
    sd = O.make_sn_state_dict(size, seed=seed)
    D = Discriminator(size, use_spectral_norm=True).set_precision(precision)
    D.load_state_dict(sd)
    return D.cuda(), sd


def _buffers(D):
    return {k: v.detach().cpu().clone() for k, v in D.state_dict().items() if k.endswith(("weight_u", "weight_v"))}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("size,B", [(64, 32), (128, 8)])
def test_spectral_norm_forward_backward_vs_oracle(precision, size, B):
    D, sd = _make(size, 4, precision)
    x = O.synthetic_signatures(B, size, seed=9)
    masks = O.make_dropout_masks(B, size, seed=6)
    D.train()
    D.mask_override = masks
    p = D(x.cuda())
    target = torch.full_like(p, 0.9)
    loss = torch.nn.functional.binary_cross_entropy(p, target)
    loss.backward()
    # oracle in float64: power iteration, effective weights, forward, backward through sigma
    sd64 = to64(sd)
    ref_p, cache, buf = O.d_forward_sn(sd64, x.double(), size, [m.double() for m in masks], train=True)
    assert rel_err(p, ref_p) <= tol(precision), rel_err(p, ref_p)
    got_buf = _buffers(D)
    for k, v in buf.items():   # u, v are computed in fp32 from fp32 weights in both precisions
        assert rel_err(got_buf[k], v) <= 1e-5, (k, rel_err(got_buf[k], v))
    ref_g = O.d_backward_sn(cache, O.bce_grad(ref_p, torch.full_like(ref_p, 0.9)), size, [m.double() for m in masks])
    named = dict(D.named_parameters())
    assert sorted(named) == sorted(ref_g)
    for k, g in ref_g.items():
        e = rel_err(named[k].grad, g)
        limit = 5e-4 if precision == "fp32" else 6e-2 * (2 if k.endswith("bias") else 1)
        assert e <= limit, f"grad {k}: {e:.3e}"
    # a second training forward iterates again from the stored buffers; eval mode must not touch them
    with torch.no_grad():
        p2 = D(x.cuda())
    sd2 = dict(sd64)
    sd2.update(buf)
    ref_p2, _, buf2 = O.d_forward_sn(sd2, x.double(), size, [m.double() for m in masks], train=True)
    assert rel_err(p2, ref_p2) <= tol(precision)
    D.eval()
    before = _buffers(D)
    with torch.no_grad():
        pe = D(x.cuda())
        feat = D.forward_features(x.cuda())
    after = _buffers(D)
    assert all(torch.equal(before[k], after[k]) for k in before)
    sd3 = dict(sd2)
    sd3.update(buf2)
    ref_pe, ce, _ = O.d_forward_sn(sd3, x.double(), size, None, train=False)
    assert rel_err(pe, ref_pe) <= tol(precision)
    assert rel_err(feat, ce["feat"]) <= tol(precision)


def test_spectral_norm_against_golden_fixture(golden_dir):
    """fp32 validation mode straight against the reference's own outputs (no oracle in between)."""
    gold = torch.load(os.path.join(golden_dir, "sn_64.pt"), weights_only=False)
    size, B = gold["size"], gold["B"]
    D, sd = _make(size, 3, "fp32")
    assert list(D.state_dict().keys()) == gold["keys"]
    x = O.synthetic_signatures(B, size, seed=7).cuda()
    D.train()
    D.mask_override = gold["train1.masks"]
    p1 = D(x)
    loss = torch.nn.functional.binary_cross_entropy(p1, torch.full_like(p1, 0.9))
    loss.backward()
    assert torch.allclose(p1.detach().cpu(), gold["train1.prob"], atol=2e-6)
    assert abs(float(loss) - gold["train1.loss"]) < 1e-5
    for k, v in _buffers(D).items():
        assert torch.allclose(v, gold[f"train1.buf.{k}"], rtol=1e-4, atol=1e-6), k
    for k, p in D.named_parameters():
        pr = gold[f"train1.grad.{k}"]
        g = p.grad.detach().cpu().reshape(-1)
        scale = pr["norm"] / pr["numel"] ** 0.5
        err = (g[pr["idx"]] - pr["vals"]).double().norm().item()
        assert err <= 5e-3 * max(pr["vals"].double().norm().item(), scale * len(pr["idx"]) ** 0.5) + 1e-7, k
    D.mask_override = gold["train2.masks"]
    with torch.no_grad():
        p2 = D(x)
    assert torch.allclose(p2.cpu(), gold["train2.prob"], atol=2e-6)
    D.eval()
    with torch.no_grad():
        pe = D(x)
    assert torch.allclose(pe.cpu(), gold["eval.prob"], atol=2e-6)


def test_spectral_norm_trains_with_stock_adam():
    """ablation…:382-386: the SN discriminator is optimised by a plain torch.optim.Adam over D.parameters()."""
    D, _ = _make(64, 5, "bf16")
    opt = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    x = O.synthetic_signatures(64, 64, seed=2).cuda()
    noise = torch.rand(64, 1, 64, 64, device="cuda") * 2 - 1
    D.train()
    losses = []
    for _ in range(20):
        opt.zero_grad()
        pr, pf = D(x), D(noise)        # two forwards -> two power iterations, each backward uses its own (u, v, sigma)
        loss = torch.nn.functional.binary_cross_entropy(pr, torch.full_like(pr, 0.9)) + \
            torch.nn.functional.binary_cross_entropy(pf, torch.zeros_like(pf))
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in D.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vanilla_gan_with_spectral_norm_trains_like_the_reference(golden_dir, precision):
    """VanillaGAN(use_spectral_norm=True) (vanilla…:70-103): train_discriminator_step / train_generator_step must run the
    SN discriminator (power iteration per forward, weight_orig / sigma), not the raw-parameter fused step — checked
    against the oracle's d_step_sn / g_step_sn and the metrics of the unmodified reference (sn_steps_64.pt)."""
    from vanilla_gan_model import VanillaGAN
    gold = torch.load(os.path.join(golden_dir, "sn_steps_64.pt"), weights_only=False)
    size, B = gold["size"], gold["B"]
    g_sd, _ = O.make_state_dicts(size, 100, seed=6)
    d_sd = O.make_sn_state_dict(size, seed=6)
    gan = VanillaGAN(latent_dim=100, image_size=size, use_spectral_norm=True, device="cuda")
    gan.generator.set_precision(precision)
    gan.discriminator.set_precision(precision)
    gan.generator.load_state_dict(g_sd)
    gan.discriminator.load_state_dict(d_sd)
    assert gan.get_config()["use_spectral_norm"] is True
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.sn_trainable_names(d_sd))
    for s in range(gold["steps"]):
        real = O.synthetic_signatures(B, size, seed=400 + s)
        nd, ng = O.hash_normal((B, 100), 500 + s), O.hash_normal((B, 100), 600 + s)
        mk = gold["masks"][s]
        gan.mask_override = mk
        md = gan.train_discriminator_step(real.cuda(), noise=nd.cuda())
        mg = gan.train_generator_step(B, noise=ng.cuda())
        od, _ = O.d_step_sn(g_sd, d_sd, d_opt, real, nd, size, mk["real"], mk["fake"])
        og, _ = O.g_step_sn(g_sd, d_sd, g_opt, ng, size)
        mtol = 2e-4 if precision == "fp32" else 2e-2
        for k, v in {**od, **og}.items():
            got = {**md, **mg}[k]
            assert abs(got - v) <= mtol * max(1.0, abs(v)) + (0.26 if k.endswith("acc") else 0), (s, k, got, v)
            assert abs(got - gold["metrics"][s][k]) <= mtol * max(1.0, abs(v)) + (0.26 if k.endswith("acc") else 0), (s, k)
        got_d = gan.discriminator.state_dict()
        for k, v in d_sd.items():
            kind = 1e-5 if k.endswith(("weight_u", "weight_v")) and precision == "fp32" else tol(precision, "param")
            assert rel_err(got_d[k], v) <= max(kind, 1e-5), (s, k, rel_err(got_d[k], v))
        got_g = gan.generator.state_dict()
        for k in O.trainable_names(g_sd):
            if k != "fc.0.bias":
                assert rel_err(got_g[k], g_sd[k]) <= tol(precision, "param"), (s, k, rel_err(got_g[k], g_sd[k]))
    # train_step (D step + G step) goes the same way and keeps the reference's keys
    out = gan.train_step(O.synthetic_signatures(B, size, seed=3).cuda())
    assert set(out) == {"d_loss", "d_loss_real", "d_loss_fake", "d_real_acc", "d_fake_acc", "d_real_mean", "d_fake_mean",
                        "g_loss", "g_fake_mean"}
    assert all(v == v for v in out.values())


def test_ablation_trainer_ordering_with_external_generator():
    """AblationGANTrainer.train_epoch (ablation…:397-467): both networks stay in train mode, ONE generator pass per
    batch, the D step sees fake.detach(), the G loss re-runs D (dropout on, another power iteration) on the same fakes
    against the smoothed label 0.9. The generator is the script's own torch module (ConfigurableGenerator,
    ablation…:216-328); only D is ours — its image gradient must reach the external module through autograd."""
    size, B = 64, 8
    D, sd = _make(size, 7, "fp32")
    D.train()
    torch.manual_seed(0)
    Gx = torch.nn.Sequential(torch.nn.Linear(100, size * size), torch.nn.Tanh()).cuda()   # stand-in external generator
    d_opt = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    g_opt = torch.optim.Adam(Gx.parameters(), lr=2e-4, betas=(0.5, 0.999))
    real = O.synthetic_signatures(B, size, seed=12)
    z = O.hash_normal((B, 100), 13).cuda()
    m1, m2, m3 = (O.make_dropout_masks(B, size, seed=40 + i) for i in range(3))
    bce = torch.nn.functional.binary_cross_entropy
    fake = Gx(z).view(B, 1, size, size)
    fake.retain_grad()
    d_opt.zero_grad()
    D.mask_override = m1
    pr = D(real.cuda())
    D.mask_override = m2
    pf = D(fake.detach())
    d_loss = bce(pr, torch.full_like(pr, 0.9)) + bce(pf, torch.zeros_like(pf))
    d_loss.backward()
    d_opt.step()
    g_opt.zero_grad()
    D.mask_override = m3
    pg = D(fake)
    g_loss = bce(pg, torch.full_like(pg, 0.9))
    g_loss.backward()
    before = [p.detach().clone() for p in Gx.parameters()]
    g_opt.step()
    # ---- oracle (float64) of the same sequence
    sd64 = to64(sd)
    opt = O.AdamState(sd64, O.sn_trainable_names(sd64))
    fake64 = fake.detach().cpu().double()
    d64 = [[m.double() for m in ms] for ms in (m1, m2, m3)]
    p1, c1, buf = O.d_forward_sn(sd64, real.double(), size, d64[0], train=True)
    sd64.update(buf)
    p2, c2, buf = O.d_forward_sn(sd64, fake64, size, d64[1], train=True)
    sd64.update(buf)
    ga = O.d_backward_sn(c1, O.bce_grad(p1, torch.full_like(p1, 0.9)), size, d64[0])
    gb = O.d_backward_sn(c2, O.bce_grad(p2, torch.zeros_like(p2)), size, d64[1])
    ref_loss = float(O.bce(p1, torch.full_like(p1, 0.9)) + O.bce(p2, torch.zeros_like(p2)))
    assert abs(float(d_loss) - ref_loss) < 1e-5
    opt.apply(sd64, {k: ga[k] + gb[k] for k in ga}, 2e-4, 0.5, 0.999)
    p3, c3, buf = O.d_forward_sn(sd64, fake64, size, d64[2], train=True)
    g3 = O.d_backward_sn(c3, O.bce_grad(p3, torch.full_like(p3, 0.9)), size, d64[2], need_dx=True)
    assert abs(float(g_loss) - float(O.bce(p3, torch.full_like(p3, 0.9)))) < 1e-5
    assert rel_err(fake.grad, g3["__dx"]) <= 5e-4, rel_err(fake.grad, g3["__dx"])
    assert all(not torch.equal(a, b) for a, b in zip(before, Gx.parameters()))      # the external module did learn
