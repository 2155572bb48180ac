"""1000-step loss-band test (SURVEY.md §8d "parity gates"): the CUDA path, trained with VanillaGAN.train_step from the
reference's initial weights on the same synthetic pool and batch order, must keep the smoothed d_loss / g_loss /
mean D outputs inside the envelope of three seeds of the UNMODIFIED reference on CPU (tests/golden/loss_band_64.pt,
written by tests/golden/make_loss_band.py), widened by the stated band; no NaN anywhere. Noise and dropout streams
differ between CPU and CUDA, so curves are compared as bands, not step by step.

The stated band: SURVEY.md suggested +-0.15 (losses) / +-0.1 (means) around the envelope, but the reference does not
meet that against itself — leaving one reference seed out, its EMA(50) leaves the envelope of the other two by up to
0.31 (d_loss), 1.40 (g_loss), 0.12 (d_real_mean / d_fake_mean): GAN training is chaotic at batch 32. The band used
here is therefore max(suggested band, 1.25 x the reference's own worst leave-one-out excursion), computed from the
fixture, plus a check on the time average over the second half of the run, which must lie within the range of the
reference seeds' averages widened by TWICE their spread (at least 0.05) on either side. One spread was too tight a
window to draw from three samples: 12 runs of this implementation (tools/loss_band_probe.py, six seeds x two builds,
profiles/r01_loss_band_probe.log) put the second-half g_loss mean at 1.81 .. 2.56 (mean 2.18) against 2.01 .. 2.22
(mean 2.15) for the three reference seeds — same centre, run-to-run scatter of about +-0.2 — so the +-1-spread window
[1.81, 2.42] rejected about one run in six whatever the kernels were."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import make_gan

pytestmark = pytest.mark.gpu

KEYS = ("d_loss", "g_loss", "d_real_mean", "d_fake_mean")
BAND = {"d_loss": 0.15, "g_loss": 0.15, "d_real_mean": 0.1, "d_fake_mean": 0.1}


def ema(x: torch.Tensor, span: int = 50) -> torch.Tensor:
    a = 2.0 / (span + 1)
    out = torch.empty_like(x)
    acc = float(x[0])
    for i, v in enumerate(x.tolist()):
        acc = a * v + (1 - a) * acc
        out[i] = acc
    return out


@pytest.mark.parametrize("precision", ["bf16"])
def test_thousand_step_loss_band(golden_dir, precision):
    ref = torch.load(os.path.join(golden_dir, "loss_band_64.pt"))
    size, B, steps, pool_n = ref["size"], ref["batch"], ref["steps"], ref["pool"]
    seeds = ref["seeds"]
    ref_ema = {k: torch.stack([ema(torch.tensor(ref["curves"][s][k], dtype=torch.float64)) for s in seeds]) for k in KEYS}
    lo = {k: ref_ema[k].min(0).values for k in KEYS}
    hi = {k: ref_ema[k].max(0).values for k in KEYS}
    band, avg_lo, avg_hi = {}, {}, {}
    for k in KEYS:
        loo = 0.0
        for i in range(len(seeds)):
            others = torch.stack([ref_ema[k][j] for j in range(len(seeds)) if j != i])
            out = torch.maximum(others.min(0).values - ref_ema[k][i], ref_ema[k][i] - others.max(0).values)
            loo = max(loo, float(out.clamp_min(0)[50:].max()))
        band[k] = max(BAND[k], 1.25 * loo)
        means = [sum(ref["curves"][s][k][steps // 2:]) / (steps - steps // 2) for s in seeds]
        spread = max(max(means) - min(means), 0.05)
        avg_lo[k], avg_hi[k] = min(means) - 2 * spread, max(means) + 2 * spread
    pool = O.synthetic_signatures(pool_n, size, seed=1234).cuda()
    worst = {k: 0.0 for k in KEYS}
    for seed in seeds[:2]:
        gan, _, _ = make_gan(size, seed, precision)
        torch.manual_seed(seed)
        perm = torch.randperm(pool_n, generator=torch.Generator().manual_seed(seed))
        cur = {k: [] for k in KEYS}
        for i in range(steps):
            idx = perm[(i * B) % pool_n:(i * B) % pool_n + B].cuda()
            m = gan.train_step(pool[idx])
            for k in KEYS:
                cur[k].append(m[k])
        for k in KEYS:
            c = torch.tensor(cur[k], dtype=torch.float64)
            assert torch.isfinite(c).all(), f"seed {seed}: {k} is not finite"
            e = ema(c)
            # the first steps of an EMA are dominated by single noisy values: judge from step 50 on
            out = torch.maximum(lo[k] - e, e - hi[k]).clamp_min(0)[50:]
            worst[k] = max(worst[k], float(out.max()))
            avg = float(c[steps // 2:].mean())
            assert avg_lo[k] <= avg <= avg_hi[k], f"seed {seed}: mean {k} over the second half {avg:.3f} outside " \
                                                  f"[{avg_lo[k]:.3f}, {avg_hi[k]:.3f}]"
    for k in KEYS:
        assert worst[k] <= band[k], f"{k}: EMA leaves the reference envelope by {worst[k]:.3f} (> {band[k]:.3f})"
