"""ConfigurableGenerator (the ablation script's generator with ReLU / LeakyReLU, ablation…:159-328) on the GPU against the
oracle (act_slope) and against tests/golden/ablation_leaky_{64,128}.pt, which the reference's own class produced."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import check_grads, rel_err, to64, tol

pytestmark = pytest.mark.gpu


def _make(size, seed, precision, activation="leaky_relu", dropout=0.25):
    from ablation_generator import ConfigurableGenerator
    from discriminator_vanilla_gan import Discriminator
    g_sd, d_sd = O.make_state_dicts(size, 100, seed=seed)
    G = ConfigurableGenerator(latent_dim=100, output_size=size, activation=activation).to("cuda")
    D = Discriminator(input_size=size, dropout=dropout).to("cuda")
    G.set_precision(precision)
    D.set_precision(precision)
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    return G, D, g_sd, d_sd


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("size", [64, 128])
def test_leaky_generator_against_reference_fixture(golden_dir, precision, size):
    gold = torch.load(os.path.join(golden_dir, f"ablation_leaky_{size}.pt"), weights_only=False)
    B = gold["B"]
    G, D, g_sd, d_sd = _make(size, gold["seed"], precision, dropout=0.0)
    z = O.hash_normal((B, 100), gold["z_seed"]).cuda()
    G.eval()
    with torch.no_grad():
        img = G(z)                               # fused-epilogue path (LeakyReLU in the GEMM epilogues)
        u8 = G.sample_uint8(z)
    assert rel_err(img, gold["eval.image"]) <= tol(precision), rel_err(img, gold["eval.image"])
    want = ((gold["eval.image"] + 1.0) * 127.5).clip(0, 255).to(torch.uint8)
    assert (u8.cpu().int() - want.int()).abs().max() <= (2 if precision == "bf16" else 1)
    img2 = G(z)                                  # grad mode: the layered path (bn_apply with the slope)
    assert rel_err(img2, gold["eval.image"]) <= tol(precision)
    # ---- the ablation trainer's G update (ablation…:441-448): both modules in train mode, smoothed label
    G.train()
    D.train()
    G.zero_grad()
    img = G(z)
    assert rel_err(img, gold["train.image"]) <= tol(precision)
    sd = G.state_dict()
    for k, ref in gold["train.stats"].items():
        if "tracked" in k:
            assert int(sd[k]) == int(ref), k
        else:
            assert rel_err(sd[k], ref) <= tol(precision), k
    pred = D(img)
    loss = torch.nn.functional.binary_cross_entropy(pred, torch.full_like(pred, 0.9))
    assert abs(float(loss) - gold["g_loss"]) <= (2e-2 if precision == "bf16" else 1e-5)
    loss.backward()
    # gradients: the same bounds as the plain Generator's chained test (_util.check_grads: fp32 within 10x of the fp32 CPU
    # oracle's own distance to float64; bf16 per-tensor angle + whole-vector norm), against the oracle pinned to the
    # reference's class by test_oracle_golden.py, and the reference's own sampled gradient values
    zc = z.cpu()
    ref = {}
    for name, sd_g, sd_d, zz in (("f32", g_sd, d_sd, zc), ("f64", to64(g_sd), to64(d_sd), zc.double())):
        im, gc, _ = O.g_forward(sd_g, zz, size, train=True, act_slope=0.2)
        pr, dc = O.d_forward(sd_d, im, size, None)
        dg = O.d_backward(sd_d, dc, O.bce_grad(pr, torch.full_like(pr, 0.9)), size, None, need_dx=True)
        ref[name] = O.g_backward(sd_g, gc, dg["__dx"], size, train=True, act_slope=0.2)
    got = {k: p.grad for k, p in G.named_parameters()}
    # final_conv.0.bias: ONE sum of cancelling terms (at 128x128 / B = 4 its bf16 value is 50 % off and still 1e-4 of the
    # gradient vector's norm): bounded through the whole-vector norm only
    check_grads(precision, "G", got, ref["f32"], ref["f64"], skip=("fc.0.bias", "final_conv.0.bias"))
    for k, g in got.items():
        pr = gold["grads"][k]
        if k == "fc.0.bias" or pr["numel"] < 64:      # mathematically zero / single sums of cancelling terms
            continue
        sampled = g.detach().double().cpu().reshape(-1)[pr["idx"]]
        e_probe = (sampled - pr["vals"].double()).norm().item() / max(pr["vals"].double().norm().item(), 1e-12)
        assert e_probe <= (0.25 if precision == "bf16" else 5e-2), (k, e_probe)


def test_relu_configuration_equals_the_plain_generator():
    """activation="relu" (and anything but "leaky_relu", ablation…:204-207) is the plain Generator, bit for bit."""
    from generator_vanilla_gan import Generator
    G, _, g_sd, _ = _make(64, 5, "bf16", activation="relu")
    P = Generator(latent_dim=100, output_size=64).to("cuda")
    P.load_state_dict(g_sd)
    z = O.hash_normal((24, 100), 3).cuda()
    for mode in ("eval", "train"):
        getattr(G, mode)()
        getattr(P, mode)()
        with torch.no_grad():
            assert torch.equal(G(z), P(z))
    assert list(G.state_dict().keys()) == list(P.state_dict().keys())
    assert isinstance(G.fc[2], torch.nn.ReLU)


def test_leaky_generator_module_contract():
    from ablation_generator import ConfigurableGenerator
    G = ConfigurableGenerator(latent_dim=50, output_size=64, activation="leaky_relu")
    assert G.activation == "leaky_relu" and isinstance(G.fc[2], torch.nn.LeakyReLU)
    assert isinstance(G.upsample_blocks[3].block[2], torch.nn.LeakyReLU) and G.upsample_blocks[3].block[2].negative_slope == 0.2
    with pytest.raises(ValueError):
        ConfigurableGenerator(output_size=32)
    G = G.to("cuda")
    G.train()
    out = G(torch.randn(6, 50, device="cuda"))
    assert out.shape == (6, 1, 64, 64) and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0
    out.mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in G.parameters())
    # a stock optimizer steps it, as AblationGANTrainer does (ablation…:376-380)
    opt = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    before = G.fc[0].weight.detach().clone()
    opt.step()
    assert not torch.equal(before, G.fc[0].weight)
