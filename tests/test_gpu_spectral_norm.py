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
