"""Pins oracle/siggan_oracle.py to the reference through the committed golden fixtures
(tests/golden/*.pt, produced by tests/golden/make_golden.py from /root/reference/src)."""
import os

import pytest
import torch

import siggan_oracle as O

RTOL = 2e-4   # fp32 CPU forward
GTOL = 5e-3   # gradients: hand-written backward vs autograd differ by summation order, amplified by the
              # B=4 BatchNorm chain (rstd up to ~300)


def check_probe(name, t, pr, rtol=RTOL, atol=1e-6):
    t = t.detach().to(torch.float32).reshape(-1)
    assert t.numel() == pr["numel"], name
    scale = pr["norm"] / max(pr["numel"], 1) ** 0.5
    got = t[pr["idx"]]
    # norm-relative over the sampled elements (long fp32 reductions cancel, max-abs is too brittle)
    err = (got - pr["vals"]).double().norm().item()
    ref = pr["vals"].double().norm().item()
    assert err <= rtol * max(ref, scale * len(got) ** 0.5) + atol * len(got) ** 0.5, \
        f"{name}: sampled err {err} vs ref {ref} (rms {scale})"
    assert abs(float(t.double().norm()) - pr["norm"]) <= rtol * pr["norm"] + atol, f"{name}: norm"


@pytest.mark.parametrize("size", [64, 128])
def test_forward_and_backward_match_reference(golden_dir, size):
    gold = torch.load(os.path.join(golden_dir, f"forward_{size}.pt"), weights_only=False)
    B = gold["B"]
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=1)
    z = O.hash_normal((B, 100), 11)
    real = O.synthetic_signatures(B, size, seed=5)
    nb = len(O.g_channels(size)) - 1
    for mode in ("eval", "train"):
        img, cache, stats = O.g_forward(g_sd, z, size, train=(mode == "train"))
        check_probe(f"g_{mode}.out", img, gold[f"g_{mode}.out"])
        check_probe(f"g_{mode}.fc", cache["fc.a"], gold[f"g_{mode}.fc"])
        for i in range(nb):
            check_probe(f"g_{mode}.up{i}", cache[f"up{i}.a"], gold[f"g_{mode}.up{i}"])
        if mode == "train":
            for k, v in stats.items():
                ref = gold[f"g_train.stats.{k}"]
                if isinstance(ref, dict):
                    check_probe(k, v, ref)
                else:
                    assert int(v) == int(ref), k
    p, cache = O.d_forward(d_sd, real, size, None)
    assert torch.allclose(p, gold["d_eval.prob"], rtol=RTOL, atol=1e-6)
    check_probe("d_eval.feat", cache["feat"], gold["d_eval.feat"])
    for i in range(len(O.d_channels(size)) - 1):
        check_probe(f"d_eval.c{i}", cache[f"c{i}.a"], gold[f"d_eval.c{i}"])
    p, _ = O.d_forward(d_sd, real, size, gold["d_train.masks"])
    assert torch.allclose(p, gold["d_train.prob"], rtol=RTOL, atol=1e-6)
    # hand-written backward vs the reference's autograd
    img, gc, _ = O.g_forward(g_sd, z, size, train=True)
    pr, dc = O.d_forward(d_sd, img, size, None)
    ones = torch.ones_like(pr)
    assert abs(float(O.bce(pr, ones)) - gold["bwd.loss"]) < 1e-5
    dg = O.d_backward(d_sd, dc, O.bce_grad(pr, ones), size, None, need_dx=True)
    gg = O.g_backward(g_sd, gc, dg["__dx"], size, train=True)
    for k in O.trainable_names(g_sd):
        # fc.0.bias feeds a BatchNorm: its gradient is mathematically zero (pure rounding noise in both).
        check_probe(f"g_grad.{k}", gg[k], gold[f"bwd.g_grad.{k}"], rtol=GTOL, atol=2e-8 if k != "fc.0.bias" else 1e-6)
    for k in O.trainable_names(d_sd):
        check_probe(f"d_grad.{k}", dg[k], gold[f"bwd.d_grad.{k}"], rtol=GTOL, atol=2e-8)


@pytest.mark.parametrize("size", [64, 128])
def test_training_steps_match_reference(golden_dir, size):
    gold = torch.load(os.path.join(golden_dir, f"steps_{size}.pt"), weights_only=False)
    B = gold["B"]
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=2)
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.trainable_names(d_sd))
    for s in range(gold["steps"]):
        real = O.synthetic_signatures(B, size, seed=100 + s)
        nd, ng = O.hash_normal((B, 100), 200 + s), O.hash_normal((B, 100), 300 + s)
        mk = gold["masks"][s]
        md, dgr, _ = O.d_step(g_sd, d_sd, d_opt, real, nd, size, mk["real"], mk["fake"])
        for k in O.trainable_names(d_sd):
            check_probe(f"s{s}.d_grad.{k}", dgr[k], gold[f"s{s}.d_grad.{k}"], rtol=GTOL, atol=2e-8)
            check_probe(f"s{s}.d_param.{k}", d_sd[k], gold[f"s{s}.d_param.{k}"])
        mg, ggr, _ = O.g_step(g_sd, d_sd, g_opt, ng, size)
        for k in O.trainable_names(g_sd):
            if k != "fc.0.bias":
                check_probe(f"s{s}.g_grad.{k}", ggr[k], gold[f"s{s}.g_grad.{k}"], rtol=GTOL, atol=2e-8)
                check_probe(f"s{s}.g_param.{k}", g_sd[k], gold[f"s{s}.g_param.{k}"])
        for k in g_sd:
            if "running" in k:
                check_probe(f"s{s}.g_stats.{k}", g_sd[k], gold[f"s{s}.g_stats.{k}"])
        md.update(mg)
        for k, v in gold["metrics"][s].items():
            assert abs(md[k] - v) <= 2e-4 * max(1.0, abs(v)), (s, k, md[k], v)
    assert gold["d_adam.keys"] == ["exp_avg", "exp_avg_sq", "step"]
    assert gold["d_adam.step"] == gold["steps"]
    check_probe("d_adam.exp_avg", d_opt.m["conv_blocks.0.block.0.weight"], gold["d_adam.exp_avg.0"], atol=2e-8)
    check_probe("d_adam.exp_avg_sq", d_opt.v["conv_blocks.0.block.0.weight"], gold["d_adam.exp_avg_sq.0"], atol=1e-12)


def test_closed_form_facts():
    """Facts SURVEY.md §8c lists, restated numerically on the oracle."""
    # ConvTranspose index rule oy = 2*iy - 1 + ky
    x = torch.zeros(1, 1, 4, 4); x[0, 0, 1, 2] = 1.0
    w = torch.zeros(1, 1, 4, 4); w[0, 0, 3, 0] = 1.0
    y = torch.nn.functional.conv_transpose2d(x, w, stride=2, padding=1)
    assert y[0, 0, 2 * 1 - 1 + 3, 2 * 2 - 1 + 0] == 1.0 and y.sum() == 1.0
    # BCE clamp and gradient epsilon
    p = torch.tensor([[0.0], [1.0]]); t = torch.tensor([[1.0], [1.0]])
    assert float(O.bce(p, t)) == 50.0
    assert float(O.bce_grad(p, t)[0]) == pytest.approx(-1.0 / 1e-12 / 2)
    # dropout masks take values {0, 4/3}
    m = O.make_dropout_masks(8, 64, seed=3)
    vals = torch.cat([v.reshape(-1) for v in m]).unique().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and vals[1] == pytest.approx(4 / 3)
    img = O.synthetic_signatures(4, 64)
    assert img.shape == (4, 1, 64, 64) and img.max() <= 1 and img.min() >= -1 and (img < 0).any()


def test_spectral_norm_variant_matches_reference(golden_dir):
    """Discriminator(use_spectral_norm=True) (disc…:61-62, 201-202): power iteration, weight_orig / sigma, the
    gradient through sigma, eval mode without iteration — against tests/golden/sn_64.pt (make_golden_sn.py)."""
    gold = torch.load(os.path.join(golden_dir, "sn_64.pt"), weights_only=False)
    size, B = gold["size"], gold["B"]
    sd = O.make_sn_state_dict(size, seed=3)
    assert list(sd.keys()) == gold["keys"]
    x = O.synthetic_signatures(B, size, seed=7)
    masks = gold["train1.masks"]
    p1, cache, buf1 = O.d_forward_sn(sd, x, size, masks, train=True)
    assert torch.allclose(p1, gold["train1.prob"], rtol=RTOL, atol=1e-6)
    target = torch.full_like(p1, 0.9)
    assert abs(float(O.bce(p1, target)) - gold["train1.loss"]) < 1e-5
    for k, v in buf1.items():
        assert torch.allclose(v, gold[f"train1.buf.{k}"], rtol=RTOL, atol=1e-6), k
    for name in O.sn_layer_names(size):
        check_probe(f"weight.{name}", cache["__eff"][name + ".weight"], gold[f"train1.weight.{name}"])
    grads = O.d_backward_sn(cache, O.bce_grad(p1, target), size, masks)
    for k in gold["keys"]:
        if k.endswith("weight_orig") or k.endswith("bias"):
            check_probe(f"grad.{k}", grads[k], gold[f"train1.grad.{k}"], rtol=GTOL)
    sd2 = dict(sd)
    sd2.update(buf1)
    p2, _, buf2 = O.d_forward_sn(sd2, x, size, gold["train2.masks"], train=True)
    assert torch.allclose(p2, gold["train2.prob"], rtol=RTOL, atol=1e-6)
    for k, v in buf2.items():
        assert torch.allclose(v, gold[f"train2.buf.{k}"], rtol=RTOL, atol=1e-6), k
    sd3 = dict(sd2)
    sd3.update(buf2)
    pe, ce, buf3 = O.d_forward_sn(sd3, x, size, None, train=False)
    assert torch.allclose(pe, gold["eval.prob"], rtol=RTOL, atol=1e-6)
    check_probe("eval.feat", ce["feat"], gold["eval.feat"])
    for k, v in buf3.items():
        assert torch.equal(v, sd3[k]), f"{k}: eval mode must not iterate"


def test_spectral_norm_training_steps_match_reference(golden_dir):
    """VanillaGAN(use_spectral_norm=True).train_discriminator_step / train_generator_step (vanilla…:180-306) of the
    unmodified reference vs the oracle's d_step_sn / g_step_sn (tests/golden/sn_steps_64.pt)."""
    gold = torch.load(os.path.join(golden_dir, "sn_steps_64.pt"), weights_only=False)
    size, B = gold["size"], gold["B"]
    g_sd, _ = O.make_state_dicts(size, 100, seed=6)
    d_sd = O.make_sn_state_dict(size, seed=6)
    g_opt = O.AdamState(g_sd, O.trainable_names(g_sd))
    d_opt = O.AdamState(d_sd, O.sn_trainable_names(d_sd))
    for s in range(gold["steps"]):
        real = O.synthetic_signatures(B, size, seed=400 + s)
        nd, ng = O.hash_normal((B, 100), 500 + s), O.hash_normal((B, 100), 600 + s)
        mk = gold["masks"][s]
        md, dgr = O.d_step_sn(g_sd, d_sd, d_opt, real, nd, size, mk["real"], mk["fake"])
        for k in O.sn_trainable_names(d_sd):
            check_probe(f"s{s}.d_grad.{k}", dgr[k], gold[f"s{s}.d_grad.{k}"], rtol=GTOL, atol=2e-8)
            check_probe(f"s{s}.d_param.{k}", d_sd[k], gold[f"s{s}.d_param.{k}"])
        for k in d_sd:
            if k.endswith(("weight_u", "weight_v")):
                assert torch.allclose(d_sd[k], gold[f"s{s}.d_buf.{k}"], rtol=RTOL, atol=1e-6), (s, k)
        mg, ggr = O.g_step_sn(g_sd, d_sd, g_opt, ng, size)
        for k in O.trainable_names(g_sd):
            if k != "fc.0.bias":
                check_probe(f"s{s}.g_grad.{k}", ggr[k], gold[f"s{s}.g_grad.{k}"], rtol=GTOL, atol=2e-8)
                check_probe(f"s{s}.g_param.{k}", g_sd[k], gold[f"s{s}.g_param.{k}"])
        md.update(mg)
        for k, v in gold["metrics"][s].items():
            assert abs(md[k] - v) <= 2e-4 * max(1.0, abs(v)), (s, k, md[k], v)


@pytest.mark.parametrize("size", [64, 128])
def test_leaky_relu_generator_matches_reference_ablation_class(golden_dir, size):
    """g_forward / g_backward with act_slope = 0.2 == the ablation script's own ConfigurableGenerator(activation=
    "leaky_relu") (ablation…:216-328), driven unmodified by tests/golden/make_golden_ablation.py."""
    gold = torch.load(os.path.join(golden_dir, f"ablation_leaky_{size}.pt"), weights_only=False)
    B, slope = gold["B"], gold["leaky_slope"]
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=gold["seed"])
    z = O.hash_normal((B, 100), gold["z_seed"])
    img, _, _ = O.g_forward(g_sd, z, size, train=False, act_slope=slope)
    assert torch.allclose(img, gold["eval.image"], rtol=RTOL, atol=2e-6)
    img, gc, stats = O.g_forward(g_sd, z, size, train=True, act_slope=slope)
    assert torch.allclose(img, gold["train.image"], rtol=RTOL, atol=2e-6)
    relu_img, _, _ = O.g_forward(g_sd, z, size, train=True)
    assert not torch.allclose(relu_img, gold["train.image"], atol=1e-3)      # the slope matters
    for k, ref in gold["train.stats"].items():
        assert torch.allclose(stats[k].to(ref.dtype), ref, rtol=RTOL, atol=1e-6), k
    pr, dc = O.d_forward(d_sd, img, size, None)
    target = torch.full_like(pr, 0.9)                                        # ablation…:421,445: the smoothed label
    assert abs(float(O.bce(pr, target)) - gold["g_loss"]) < 1e-5
    dg = O.d_backward(d_sd, dc, O.bce_grad(pr, target), size, None, need_dx=True)
    gg = O.g_backward(g_sd, gc, dg["__dx"], size, train=True, act_slope=slope)
    for k in O.trainable_names(g_sd):
        check_probe(f"g_grad.{k}", gg[k], gold["grads"][k], rtol=GTOL, atol=2e-8 if k != "fc.0.bias" else 1e-6)


def test_double_width_matches_assembled_reference_blocks(golden_dir):
    """The oracle at width = 2 (BASELINE configs[4]'s "2x hidden width") == a model assembled from the
    reference's own channel-parameterised UpsampleBlock / DownsampleBlock with doubled channel counts
    (tests/golden/make_golden_wide.py): forward in both modes, running statistics, G-loss and D-loss gradients."""
    gold = torch.load(os.path.join(golden_dir, "wide2_64.pt"), weights_only=False)
    size, width, B = gold["size"], gold["width"], gold["B"]
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=gold["seed"], width=width)
    assert g_sd["upsample_blocks.3.block.0.weight"].shape == (64, 64, 4, 4) and d_sd["classifier.0.weight"].shape == (1, 16384)
    z = O.hash_normal((B, 100), gold["z_seed"])
    real = O.synthetic_signatures(B, size, seed=gold["real_seed"])
    img, _, _ = O.g_forward(g_sd, z, size, train=False)
    assert torch.allclose(img, gold["eval.image"], rtol=RTOL, atol=2e-6)
    assert torch.allclose(O.d_forward(d_sd, img, size, None)[0], gold["eval.prob_fake"], rtol=RTOL, atol=1e-6)
    assert torch.allclose(O.d_forward(d_sd, real, size, None)[0], gold["eval.prob_real"], rtol=RTOL, atol=1e-6)
    img, gc, stats = O.g_forward(g_sd, z, size, train=True)
    assert torch.allclose(img, gold["train.image"], rtol=RTOL, atol=2e-6)
    for k, ref in gold["train.stats"].items():
        assert torch.allclose(stats[k].to(ref.dtype), ref, rtol=RTOL, atol=1e-6), k
    pr, dc = O.d_forward(d_sd, img, size, None)
    ones = torch.ones_like(pr)
    assert abs(float(O.bce(pr, ones)) - gold["g_loss"]) < 1e-5
    dg = O.d_backward(d_sd, dc, O.bce_grad(pr, ones), size, None, need_dx=True)
    gg = O.g_backward(g_sd, gc, dg["__dx"], size, train=True)
    for k in O.trainable_names(g_sd):
        check_probe(f"g_grad.{k}", gg[k], gold["g_grads"][k], rtol=GTOL, atol=2e-8 if k != "fc.0.bias" else 1e-6)
    # D loss: real vs 0.9 + fake vs 0 with the reference's captured masks
    p_r, c_r = O.d_forward(d_sd, real, size, gold["masks_real"])
    p_f, c_f = O.d_forward(d_sd, img.detach(), size, gold["masks_fake"])
    assert torch.allclose(p_r, gold["train.prob_real"], rtol=RTOL, atol=1e-6)
    assert torch.allclose(p_f, gold["train.prob_fake"], rtol=RTOL, atol=1e-6)
    loss = float(O.bce(p_r, torch.full_like(p_r, 0.9)) + O.bce(p_f, torch.zeros_like(p_f)))
    assert abs(loss - gold["d_loss"]) < 1e-5
    g_r = O.d_backward(d_sd, c_r, O.bce_grad(p_r, torch.full_like(p_r, 0.9)), size, gold["masks_real"])
    g_f = O.d_backward(d_sd, c_f, O.bce_grad(p_f, torch.zeros_like(p_f)), size, gold["masks_fake"])
    for k in O.trainable_names(d_sd):
        check_probe(f"d_grad.{k}", g_r[k] + g_f[k], gold["d_grads"][k], rtol=GTOL, atol=2e-8)
