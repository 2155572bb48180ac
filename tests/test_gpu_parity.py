"""GPU parity tests: the CUDA path (through the drop-in modules -> ctypes -> libsiggan.so) against the CPU oracle on
the same seeded inputs, against the committed golden fixtures, and size-independent properties at full batch."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import check_grads, make_gan, rel_err, to64, tol

pytestmark = pytest.mark.gpu

PRECISIONS = ["bf16", "fp32"]


def _check(name, got, ref, limit):
    e = rel_err(got, ref)
    assert e <= limit, f"{name}: relative error {e:.3e} > {limit:.1e}"
    return e


# ------------------------------------------------------------------------------------------------
# forward passes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("size,B", [(64, 32), (128, 8)])
def test_generator_forward(precision, size, B):
    gan, g_sd, _ = make_gan(size, 1, precision)
    G = gan.generator
    z = O.hash_normal((B, 100), 11)
    G.eval()
    with torch.no_grad():
        img = G(z.cuda())
    ref, _, _ = O.g_forward(g_sd, z, size, train=False)
    assert img.shape == (B, 1, size, size) and img.dtype == torch.float32
    _check("G eval", img, ref, tol(precision))
    # eval forward with autograd enabled takes the un-fused path and must agree too
    img2 = G(z.cuda())
    _check("G eval (grad mode)", img2, ref, tol(precision))
    G.train()
    with torch.no_grad():
        img = G(z.cuda())
    ref, _, stats = O.g_forward(g_sd, z, size, train=True)
    _check("G train", img, ref, tol(precision))
    sd = G.state_dict()
    for k, v in stats.items():
        if "num_batches" in k:
            assert int(sd[k]) == int(v), k
        else:
            _check(k, sd[k], v, tol(precision))


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("size,B", [(64, 32), (128, 8)])
def test_discriminator_forward(precision, size, B):
    gan, _, d_sd = make_gan(size, 1, precision)
    D = gan.discriminator
    x = O.synthetic_signatures(B, size, seed=5)
    D.eval()
    with torch.no_grad():
        p = D(x.cuda())
        feat = D.forward_features(x.cuda())
    ref, cache = O.d_forward(d_sd, x, size, None)
    assert p.shape == (B, 1)
    _check("D eval prob", p, ref, tol(precision))
    _check("D features (NCHW order)", feat, cache["feat"], tol(precision))
    masks = O.make_dropout_masks(B, size, seed=3)
    D.train()
    D.mask_override = masks
    with torch.no_grad():
        p = D(x.cuda())
    ref, _ = O.d_forward(d_sd, x, size, masks)
    _check("D train prob", p, ref, tol(precision))
    D.mask_override = None
    with torch.no_grad():
        p1 = D(x.cuda())
    assert torch.isfinite(p1).all() and (p1 - p).abs().max() > 0     # library RNG masks differ from the injected ones


@pytest.mark.parametrize("size", [64, 128])
def test_forward_against_golden_fixtures(golden_dir, size):
    """CUDA fp32 validation mode straight against the reference-generated fixtures (no oracle in between)."""
    gold = torch.load(os.path.join(golden_dir, f"forward_{size}.pt"), weights_only=False)
    B = gold["B"]
    gan, _, _ = make_gan(size, 1, "fp32")
    z = O.hash_normal((B, 100), 11).cuda()
    real = O.synthetic_signatures(B, size, seed=5).cuda()
    for mode in ("eval", "train"):
        gan.generator.train(mode == "train")
        with torch.no_grad():
            img = gan.generator(z).cpu().reshape(-1)
        pr = gold[f"g_{mode}.out"]
        assert (img[pr["idx"]] - pr["vals"]).abs().max() < 2e-5
        assert abs(float(img.double().norm()) - pr["norm"]) < 1e-4 * pr["norm"]
    gan.discriminator.eval()
    with torch.no_grad():
        p = gan.discriminator(real).cpu()
    assert torch.allclose(p, gold["d_eval.prob"], atol=2e-6)
    gan.discriminator.train()
    gan.discriminator.mask_override = gold["d_train.masks"]
    with torch.no_grad():
        p = gan.discriminator(real).cpu()
    assert torch.allclose(p, gold["d_train.prob"], atol=2e-6)


# ------------------------------------------------------------------------------------------------
# backward through autograd (the path the unchanged reference trainer uses, train…:339-376)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("size,B", [(64, 64), (128, 16)])
def test_autograd_g_step_gradients(precision, size, B):
    gan, g_sd, d_sd = make_gan(size, 1, precision)
    G, D = gan.generator, gan.discriminator
    z = O.hash_normal((B, 100), 11)
    G.train()
    D.eval()
    fake = G(z.cuda())
    pred = D(fake)
    loss = gan.criterion(pred, torch.ones(B, 1, device="cuda"))
    loss.backward()
    refs = []
    for gs, ds, zz in ((g_sd, d_sd, z), (to64(g_sd), to64(d_sd), z.double())):
        img, gc, _ = O.g_forward(gs, zz, size, train=True)
        pr, dc = O.d_forward(ds, img, size, None)
        ones = torch.ones_like(pr)
        dg = O.d_backward(ds, dc, O.bce_grad(pr, ones), size, None, need_dx=True)
        gg = O.g_backward(gs, gc, dg["__dx"], size, train=True)
        refs.append((float(O.bce(pr, ones)), dg, gg))
    assert abs(loss.item() - refs[1][0]) < tol(precision) * 2
    # fc.0.bias sits in front of a BatchNorm: its gradient is mathematically zero, both sides hold rounding noise
    assert G.fc[0].bias.grad.abs().max() <= 1e-2 * G.fc[0].weight.grad.abs().max()
    check_grads(precision, "G", {k: p.grad for k, p in G.named_parameters()}, refs[0][2], refs[1][2])
    # autograd also fills D's (unused) weight gradients in the G step, like the reference
    check_grads(precision, "D", {k: p.grad for k, p in D.named_parameters()}, refs[0][1], refs[1][1])


@pytest.mark.parametrize("precision", PRECISIONS)
def test_autograd_d_step_like_reference_trainer(precision):
    """GANTrainer._train_discriminator (train…:281-337) written against the module API: two D passes, summed loss,
    backward, optimizer.step()."""
    size, B = 64, 32
    gan, g_sd, d_sd = make_gan(size, 2, precision)
    G, D = gan.generator, gan.discriminator
    real = O.synthetic_signatures(B, size, seed=100)
    noise = O.hash_normal((B, 100), 200)
    mk_r, mk_f = O.make_dropout_masks(B, size, 31), O.make_dropout_masks(B, size, 32)
    D.train()
    G.eval()
    gan.d_optimizer.zero_grad()
    D.mask_override = mk_r
    real_preds = D(real.cuda())
    d_loss_real = gan.criterion(real_preds, torch.full((B, 1), 0.9, device="cuda"))
    with torch.no_grad():
        fake = G(noise.cuda())
    D.mask_override = mk_f
    fake_preds = D(fake)
    d_loss_fake = gan.criterion(fake_preds, torch.zeros(B, 1, device="cuda"))
    d_loss = d_loss_real + d_loss_fake
    d_loss.backward()
    total_norm = torch.nn.utils.clip_grad_norm_(D.parameters(), 1e9)   # train…:262-279 must keep working
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    md, grads, _ = O.d_step(g_sd, d_sd, d_opt, real, noise, size, mk_r, mk_f, apply_update=False)
    _, grads64, _ = O.d_step(to64(g_sd), to64(d_sd), None, real.double(), noise.double(), size,
                             [m.double() for m in mk_r], [m.double() for m in mk_f], apply_update=False)
    assert abs(d_loss.item() - md["d_loss"]) < tol(precision) * 2
    assert abs(real_preds.mean().item() - md["d_real_mean"]) < tol(precision)
    ref_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    assert abs(total_norm.item() - ref_norm) <= 4e-2 * ref_norm
    check_grads(precision, "D", {k: p.grad for k, p in D.named_parameters()}, grads, grads64)
    gan.d_optimizer.step()
    d_opt.apply(d_sd, grads, 2e-4, 0.5, 0.999)
    for k, p in D.named_parameters():
        _check(f"D param {k}", p, d_sd[k], tol(precision, "param"))
    st = gan.d_optimizer.state_dict()
    assert sorted(st["state"][0].keys()) == ["exp_avg", "exp_avg_sq", "step"] and float(st["state"][0]["step"]) == 1.0
    _check("exp_avg", st["state"][0]["exp_avg"], d_opt.m["conv_blocks.0.block.0.weight"], 4e-2)


# ------------------------------------------------------------------------------------------------
# fused training steps (VanillaGAN.train_*_step, vanilla…:180-336)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("size,B,steps", [(64, 64, 3), (128, 16, 2)])
def test_fused_training_steps(precision, size, B, steps):
    gan, g_sd, d_sd = make_gan(size, 2, precision)
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    for s in range(steps):
        real = O.synthetic_signatures(B, size, seed=100 + s)
        nd, ng = O.hash_normal((B, 100), 200 + s), O.hash_normal((B, 100), 300 + s)
        mk_r, mk_f = O.make_dropout_masks(B, size, 40 + s), O.make_dropout_masks(B, size, 50 + s)
        gan.mask_override = {"real": mk_r, "fake": mk_f}
        md = gan.train_discriminator_step(real.cuda(), noise=nd.cuda())
        d_grads = {k: p.grad.clone() for k, p in gan.discriminator.named_parameters()}
        mg = gan.train_generator_step(B, noise=ng.cuda())
        g64, d64 = to64(g_sd), to64(d_sd)     # fp64 oracle from the same pre-step state, for the gradient checks
        _, ogr64, _ = O.d_step(g64, d64, None, real.double(), nd.double(), size, [m.double() for m in mk_r],
                               [m.double() for m in mk_f], apply_update=False)
        od, ogr, _ = O.d_step(g_sd, d_sd, d_opt, real, nd, size, mk_r, mk_f)
        d64 = to64(d_sd)
        _, ggr64, _ = O.g_step(g64, d64, None, ng.double(), size, apply_update=False)
        og, ggr, _ = O.g_step(g_sd, d_sd, g_opt, ng, size)
        # after the first update the two runs no longer start from identical parameters (Adam turns sub-tolerance
        # gradient differences into lr-sized parameter differences), so later steps get proportionally more slack
        drift = 1 + 4 * s
        mtol = tol(precision) * 2 * drift
        for k, v in od.items():
            slack = 1.0 / B + 1e-6 if k.endswith("acc") else mtol * max(1.0, abs(v))
            assert abs(md[k] - v) <= slack, (s, k, md[k], v)
        for k, v in og.items():
            assert abs(mg[k] - v) <= mtol * max(1.0, abs(v)), (s, k, mg[k], v)
        if s == 0:   # later steps start from (slightly) different parameters; parameters are compared instead
            check_grads(precision, "D", d_grads, ogr, ogr64)
            check_grads(precision, "G", {k: p.grad for k, p in gan.generator.named_parameters()}, ggr, ggr64)
        for k, p in gan.discriminator.named_parameters():
            _check(f"s{s} D param {k}", p, d_sd[k], tol(precision, "param") * (1 + s))  # noqa
        for k, p in gan.generator.named_parameters():
            if k != "fc.0.bias":
                _check(f"s{s} G param {k}", p, g_sd[k], tol(precision, "param") * (1 + s))
        sd = gan.generator.state_dict()
        for k in g_sd:
            if "running" in k:
                _check(f"s{s} {k}", sd[k], g_sd[k], tol(precision) * (1 if s == 0 else 20 * s))
            elif "num_batches" in k:
                assert int(sd[k]) == int(g_sd[k])
    assert gan.global_step == steps and len(gan.d_losses) == steps and len(gan.g_losses) == steps


def test_fused_steps_against_golden_fixtures(golden_dir):
    """fp32 validation mode against the reference's own train_*_step outputs (tests/golden/steps_64.pt)."""
    gold = torch.load(os.path.join(golden_dir, "steps_64.pt"), weights_only=False)
    B = gold["B"]
    gan, _, _ = make_gan(64, 2, "fp32")
    for s in range(gold["steps"]):
        real = O.synthetic_signatures(B, 64, seed=100 + s)
        gan.mask_override = gold["masks"][s]
        m = gan.train_discriminator_step(real.cuda(), noise=O.hash_normal((B, 100), 200 + s).cuda())
        m.update(gan.train_generator_step(B, noise=O.hash_normal((B, 100), 300 + s).cuda()))
        for k, v in gold["metrics"][s].items():
            assert abs(m[k] - v) <= 5e-4 * max(1.0, abs(v)), (s, k, m[k], v)
        for k, p in gan.discriminator.named_parameters():
            pr = gold[f"s{s}.d_param.{k}"]
            got = p.detach().cpu().reshape(-1)[pr["idx"]]
            assert (got - pr["vals"]).norm() <= 2e-3 * max(float(pr["vals"].norm()), 1e-6), (s, k)


# ------------------------------------------------------------------------------------------------
# size-independent properties at the benchmark batch, edge cases, egress, checkpoints
# ------------------------------------------------------------------------------------------------
def test_full_batch_properties():
    B, size = 4096, 64
    gan, _, _ = make_gan(size, 3, "bf16")
    G, D = gan.generator, gan.discriminator
    g = torch.Generator("cuda").manual_seed(0)
    z = torch.randn(B, 100, device="cuda", generator=g)
    G.eval()
    D.eval()
    with torch.no_grad():
        img = G(z)
        assert img.shape == (B, 1, size, size) and torch.isfinite(img).all() and img.abs().max() <= 1
        # eval-mode samples are independent of their batch mates: any slice equals the slice's own forward
        sub = G(z[1000:1128])
        assert torch.equal(sub, img[1000:1128])
        p = D(img)
        p2 = D(torch.cat([img[2048:], img[:2048]]))
        assert torch.equal(p2, torch.cat([p[2048:], p[:2048]]))
        u8 = G.sample_uint8(z[:256])
        ref8 = ((img[:256] + 1) * 127.5).clamp(0, 255).to(torch.uint8)
        assert torch.equal(u8, ref8)
    # two identically seeded runs of the fused step are bit-identical (deterministic reductions)
    outs = []
    for _ in range(2):
        gan2, _, _ = make_gan(size, 3, "bf16")
        gan2._dropout_offset = 0        # the Dropout2d mask stream is process-wide: both runs start it at the same point
        torch.manual_seed(7)
        real = torch.rand(B, 1, size, size, device="cuda", generator=torch.Generator("cuda").manual_seed(1)) * 2 - 1
        m = [gan2.train_step(real) for _ in range(2)][-1]
        outs.append((m, gan2.generator._flat.flat.clone(), gan2.discriminator._flat.flat.clone()))
        assert all(torch.isfinite(torch.tensor(v)) for v in m.values())
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_benchmark_batch_step_against_oracle():
    """The benchmarked configuration itself (64x64, B = 4096 per GPU, bf16: BASELINE.json configs[2]) against the CPU
    oracle: one D step and one G step of the fused path, gradients and metrics. Split-K factors, CTA-pair weight
    gradients and tile schedules are functions of the batch, so the small-batch parity tests do not cover them.
    (~1 minute of CPU time for the oracle at this size.)"""
    B, size = 4096, 64
    gan, g_sd, d_sd = make_gan(size, 4, "bf16")
    real = O.synthetic_signatures(B, size, seed=21)
    nd, ng = O.hash_normal((B, 100), 31), O.hash_normal((B, 100), 32)
    mk_r, mk_f = O.make_dropout_masks(B, size, 41), O.make_dropout_masks(B, size, 42)
    gan.mask_override = {"real": mk_r, "fake": mk_f}
    md = gan.train_discriminator_step(real.cuda(), noise=nd.cuda())
    d_got = {k: p.grad.detach().float().cpu().clone() for k, p in gan.discriminator.named_parameters()}
    mg = gan.train_generator_step(B, noise=ng.cuda())
    g_got = {k: p.grad.detach().float().cpu().clone() for k, p in gan.generator.named_parameters()}
    torch.cuda.synchronize()
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    od, d_ref, _ = O.d_step(g_sd, d_sd, d_opt, real, nd, size, mk_r, mk_f)
    og, g_ref, _ = O.g_step(g_sd, d_sd, g_opt, ng, size)
    for k, v in {**od, **og}.items():
        got = {**md, **mg}[k]
        assert abs(got - v) <= 1e-2 * max(1.0, abs(v)) + (0.02 if k.endswith("acc") else 0), (k, got, v)
    errs = {}
    for k, ref in d_ref.items():      # chained tolerances of tests/_util.py (errors accumulate down the backward chain)
        errs[k] = rel_err(d_got[k], ref)
        assert errs[k] <= tol("bf16", "grad_d") * (2 if k.endswith("bias") else 1), f"D grad {k}: {errs[k]:.3e}"
    num = den = 0.0
    for k, ref in g_ref.items():
        if k == "fc.0.bias":          # mathematically zero (bias in front of BatchNorm)
            continue
        errs[k] = rel_err(g_got[k], ref)
        assert errs[k] <= tol("bf16", "grad_g"), f"G grad {k}: {errs[k]:.3e}"
        num += ((g_got[k].double() - ref.double()) ** 2).sum().item()
        den += (ref.double() ** 2).sum().item()
    assert (num / den) ** 0.5 <= tol("bf16", "grad_g_all"), f"G gradient vector: {(num / den) ** 0.5:.3e}"
    print("\nB=4096 step vs oracle:", {k: f"{e:.1e}" for k, e in errs.items()}, "G vector", f"{(num / den) ** 0.5:.1e}")


def test_edge_cases():
    gan, g_sd, d_sd = make_gan(64, 1, "bf16")
    G, D = gan.generator, gan.discriminator
    G.eval()
    D.eval()
    for B in (1, 3, 5, 130):     # ragged batches: partial GEMM tiles, masked rows
        z = O.hash_normal((B, 100), 60 + B)
        with torch.no_grad():
            img = G(z.cuda())
            p = D(img)
        ref, _, _ = O.g_forward(g_sd, z, 64, train=False)
        pr, _ = O.d_forward(d_sd, ref, 64, None)
        assert rel_err(img, ref) <= 1e-2 and rel_err(p, pr) <= 1e-2
    G.train()
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        G(torch.randn(1, 100, device="cuda"))
    with pytest.raises(RuntimeError):
        G(torch.randn(1, 100, 1, 1, device="cuda"))          # the 4-D latent the reference's Linear rejects too
    with pytest.raises(RuntimeError):
        D(torch.randn(2, 1, 32, 32, device="cuda"))
    # BCE saturation semantics: log clamp at -100 and gradient denominator eps
    p = torch.tensor([[0.0], [1.0], [0.3]], device="cuda", requires_grad=True)
    t = torch.tensor([[1.0], [1.0], [1.0]], device="cuda")
    loss = gan.criterion(p, t)
    loss.backward()
    rp = torch.tensor([[0.0], [1.0], [0.3]], requires_grad=True)
    rl = torch.nn.BCELoss()(rp, torch.ones(3, 1))
    rl.backward()
    assert abs(loss.item() - rl.item()) < 1e-4 and torch.allclose(p.grad.cpu(), rp.grad, rtol=1e-5)


def test_checkpoint_round_trip(tmp_path):
    gan, _, _ = make_gan(64, 5, "bf16")
    real = O.synthetic_signatures(16, 64, seed=1).cuda()
    gan.train_step(real)
    gan.save(tmp_path / "ckpt")
    ck = torch.load(tmp_path / "ckpt.pt", weights_only=False)
    assert {"config", "generator_state_dict", "discriminator_state_dict", "current_epoch", "global_step", "saved_at",
            "g_optimizer_state_dict", "d_optimizer_state_dict", "d_losses", "g_losses"} <= set(ck)
    assert float(ck["g_optimizer_state_dict"]["state"][0]["step"]) == 1.0
    from vanilla_gan_model import VanillaGAN
    gan2 = VanillaGAN.from_checkpoint(tmp_path / "ckpt", device="cuda")
    for (k, a), (_, b) in zip(gan.state_dict().items(), gan2.state_dict().items()):
        assert torch.equal(a, b), k
    # the Dropout2d mask stream (process-wide, _siggan_lib.DROPOUT) is checkpointed: loading put it back where the save left it
    off = gan2._dropout_offset
    assert off == ck["dropout_stream"]["offset"] and off > 0
    torch.manual_seed(3)
    m1 = gan.train_step(real)
    torch.manual_seed(3)
    gan2._dropout_offset = off                       # what a resumed process would start from
    m2 = gan2.train_step(real)
    assert m1 == m2, (m1, m2)       # resumed run continues bit-identically (Adam moments + step restored)
    # a plain torch.optim.Adam state dict (what the reference trainer writes) loads into the fused optimizer
    ref_opt = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in gan.discriminator.parameters()],
                               lr=2e-4, betas=(0.5, 0.999))
    for p in ref_opt.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    ref_opt.step()
    gan2.d_optimizer.load_state_dict(ref_opt.state_dict())
    gan2.d_optimizer._ensure_state()
    assert gan2.d_optimizer._steps == 1 and torch.allclose(gan2.d_optimizer._m, torch.full_like(gan2.d_optimizer._m, 0.5))


def test_split_d_backward_phases_equal_the_fused_phase():
    """Phases 11 + 12 of sg_train_step (used to overlap the D gradient all-reduce with backward) == phase 1, bitwise."""
    import ctypes as C
    import _siggan_lib as L
    B, size = 48, 64
    real = O.synthetic_signatures(B, size, seed=3).cuda()
    noise = O.hash_normal((B, 100), 5).cuda()
    grads = []
    for split in (False, True):
        gan, _, _ = make_gan(size, 6, "bf16")
        torch.manual_seed(11)
        gan.discriminator.train()
        gan.generator.eval()
        sctx = gan._fused_ready()
        st = gan._state(None)
        dg = gan.discriminator._flat.grad_staging()
        dg.fill_(float("nan"))
        stream = L.current_stream(real.device)
        args = (sctx.handle, C.byref(st), L.ptr(real), L.ptr(noise), None, B, L.ptr(dg), None, L.ptr(gan._metrics))
        if split:
            L.check(sctx.lib.sg_train_step(*args, 11, stream), "phase 11")
            tail = int(sctx.lib.sg_d_grad_tail_offset(sctx.handle))
            torch.cuda.synchronize()
            assert torch.isfinite(dg[tail:]).all() and torch.isnan(dg[:tail]).all()   # only the tail is final
            L.check(sctx.lib.sg_train_step(*args, 12, stream), "phase 12")
        else:
            L.check(sctx.lib.sg_train_step(*args, 1, stream), "phase 1")
        torch.cuda.synchronize()
        grads.append(dg.clone())
    assert torch.isfinite(grads[0]).all() and torch.equal(grads[0], grads[1])


@pytest.mark.parametrize("size", [64, 128])
def test_split_g_backward_phases_equal_the_fused_phase(size):
    """Phases 31 + 32 of sg_train_step (the G bucket's upsample-block part is all-reduced while the fc stage runs) ==
    phase 3, bitwise; after phase 31 exactly the bucket from sg_g_grad_tail_offset() on is final."""
    import ctypes as C
    import _siggan_lib as L
    B = 24
    noise = O.hash_normal((B, 100), 5).cuda()
    grads, stats = [], []
    for split in (False, True):
        gan, _, _ = make_gan(size, 6, "bf16")
        gan.generator.train()
        gan.discriminator.eval()
        sctx = gan._fused_ready()
        st = gan._state(None)
        gg = gan.generator._flat.grad_staging()
        gg.fill_(float("nan"))
        stream = L.current_stream(noise.device)
        args = (sctx.handle, C.byref(st), None, None, L.ptr(noise), B, None, L.ptr(gg), L.ptr(gan._metrics))
        for rep in range(3):      # the third call replays the captured graphs of the phases
            if split:
                gg.fill_(float("nan"))
                L.check(sctx.lib.sg_train_step(*args, 31, stream), "phase 31")
                tail = int(sctx.lib.sg_g_grad_tail_offset(sctx.handle))
                torch.cuda.synchronize()
                assert torch.isfinite(gg[tail:]).all() and torch.isnan(gg[:tail]).all()
                L.check(sctx.lib.sg_train_step(*args, 32, stream), "phase 32")
            else:
                L.check(sctx.lib.sg_train_step(*args, 3, stream), "phase 3")
        torch.cuda.synchronize()
        grads.append(gg.clone())
        stats.append(gan.generator._flat.stats.clone())
    assert torch.isfinite(grads[0]).all() and torch.equal(grads[0], grads[1]) and torch.equal(stats[0], stats[1])


def test_whole_step_call_equals_the_four_phases():
    """sg_train_step(phase 0) — the D update emits the Discriminator's packs itself (fused Adam + pack kernel) and the G step
    re-packs nothing — leaves the same parameters, moments and BatchNorm statistics, bit for bit, as the four phases called
    one by one (each of which packs from the fp32 masters); 5 steps, so the captured graphs of both forms replay."""
    import _siggan_lib as L
    B, size = 32, 64
    reals = [O.synthetic_signatures(B, size, seed=70 + i).cuda() for i in range(5)]
    outs = []
    for whole in (True, False):
        gan, _, _ = make_gan(size, 9, "bf16")
        torch.manual_seed(123)
        L.DROPOUT.offset = 0
        for x in reals:
            if whole:
                m = gan.train_step_async(x)                         # one call
            else:
                gan.discriminator_step_async(x)                     # phases 1, 2
                m = gan.generator_step_async(x.size(0))             # phases 3, 4
        torch.cuda.synchronize()
        outs.append([t.clone() for t in (gan.generator._flat.flat, gan.discriminator._flat.flat, gan.generator._flat.stats,
                                         gan.g_optimizer._m, gan.g_optimizer._v, gan.d_optimizer._m, gan.d_optimizer._v, m)])
        assert gan.g_optimizer._steps == 5 and gan.d_optimizer._steps == 5
    for a, b in zip(*outs):
        assert torch.isfinite(a).all() and torch.equal(a, b)


def test_library_communicator_single_rank():
    """sg_comm_* / sg_allreduce_grads with a one-rank NCCL communicator (the box the GPU suite runs on has one GPU; the
    multi-rank path is exercised by bench.py --gpus N, which asserts bit-identical replicas): the id handshake works,
    the mean over one rank leaves the bucket unchanged in both the in-order and the overlapped form, ranges are checked."""
    import ctypes as C
    import _siggan_lib as L
    gan, _, _ = make_gan(64, 3, "bf16")
    sctx = gan._fused_ready()
    lib = sctx.lib
    assert lib.sg_comm_nccl_version() >= 21800
    assert lib.sg_comm_world_size(sctx.handle) == 0
    ident = C.create_string_buffer(128)
    L.check(lib.sg_comm_unique_id(ident, 128), "unique id")
    assert lib.sg_comm_unique_id(ident, 64) != 0 and b"128" in lib.sg_last_error()
    g = torch.randn(sctx.param_count(L.SG_NET_D), device="cuda")
    st = L.current_stream(g.device)
    assert lib.sg_allreduce_grads(sctx.handle, L.SG_NET_D, L.ptr(g), 0, -1, 0, st) != 0      # no communicator yet
    assert b"communicator" in lib.sg_last_error()
    L.check(lib.sg_comm_init(sctx.handle, ident.raw, 128, 0, 1), "comm init")
    try:
        assert lib.sg_comm_world_size(sctx.handle) == 1
        ref = g.clone()
        tail = int(lib.sg_d_grad_tail_offset(sctx.handle))
        L.check(lib.sg_allreduce_grads(sctx.handle, L.SG_NET_D, L.ptr(g), tail, -1, 1, st), "overlapped all-reduce")
        L.check(lib.sg_allreduce_grads(sctx.handle, L.SG_NET_D, L.ptr(g), 0, tail, 0, st), "in-order all-reduce")
        L.check(lib.sg_allreduce_join(sctx.handle, st), "join")
        torch.cuda.synchronize()
        assert torch.equal(g, ref)
        assert lib.sg_allreduce_grads(sctx.handle, L.SG_NET_D, L.ptr(g), 10, g.numel(), 0, st) != 0
        assert b"exceeds" in lib.sg_last_error()
    finally:
        L.check(lib.sg_comm_destroy(sctx.handle), "comm destroy")
    assert lib.sg_comm_world_size(sctx.handle) == 0


def test_bulk_sampler_matches_chunked_sampling():
    """sample_uint8_to_host (double-buffered egress) == sample_uint8 chunk by chunk, including a ragged last chunk."""
    gan, _, _ = make_gan(64, 8, "bf16")
    G = gan.generator
    G.eval()
    z = O.hash_normal((700, 100), 91)
    host = G.sample_uint8_to_host(700, batch=256, latents=z.pin_memory())
    assert host.shape == (700, 1, 64, 64) and host.dtype == torch.uint8 and host.is_pinned()
    ref = torch.cat([G.sample_uint8(z[i:i + 256].cuda()).cpu() for i in range(0, 700, 256)])
    assert torch.equal(host, ref)
    drawn = G.sample_uint8_to_host(300, batch=128, generator=torch.Generator("cuda").manual_seed(5))
    assert drawn.shape == (300, 1, 64, 64) and 0 < drawn.float().std()


@pytest.mark.parametrize("size,B", [(64, 300), (128, 37)])
def test_fused_eval_tail_matches_layered_path(size, B):
    """The one-kernel eval tail (sg_convt4_final.cu: last ConvT block + BatchNorm + ReLU + Conv3x3 + tanh, taken under
    no_grad) against the layered path the same module takes with autograd enabled, on a batch that gives every CTA
    several whole images plus a ragged remainder (the 3x3 row halo is carried from tile to tile inside an image and
    must not leak across images), with non-trivial running statistics; then the uint8-only egress (no fp32 image
    written) against the conversion of the fp32 image (utils/inference.py:129)."""
    gan, g_sd, _ = make_gan(size, 21, "bf16")
    G = gan.generator
    with torch.no_grad():   # running statistics away from their (0, 1) initial values
        for k, v in G.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(0.05 * O.hash_normal(tuple(v.shape), 300 + len(k)).to(v.device))
            elif k.endswith("running_var"):
                v.copy_(0.5 + O.hash_normal(tuple(v.shape), 400 + len(k)).abs().to(v.device))
    G.eval()
    z = O.hash_normal((B, 100), 77).cuda()
    with torch.no_grad():
        fused = G(z)
        u8 = G.sample_uint8(z)
    layered = G(z).detach()
    assert torch.isfinite(fused).all() and fused.abs().max() <= 1
    _check("fused tail vs layered path", fused, layered, 4e-3)   # both bf16 paths; they differ in rounding points only
    ref, _, _ = O.g_forward({k: v.cpu() for k, v in G.state_dict().items()}, z.cpu(), size, train=False)
    _check("fused tail vs oracle", fused, ref, tol("bf16"))
    assert torch.equal(u8, ((fused + 1) * 127.5).clamp(0, 255).to(torch.uint8))
    # first and last rows / columns are where the zero padding of the 3x3 stencil lives
    for sl in ((..., 0, slice(None)), (..., size - 1, slice(None)), (..., slice(None), 0), (..., slice(None), size - 1)):
        _check("border", fused[sl], ref[sl], 2e-2)
