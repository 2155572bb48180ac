"""AsyncCheckpointWriter: checkpoints in the reference trainer's format (train…:402-444) written from a
stream-ordered snapshot while training continues; they must equal what a synchronous `state_dict()` capture at the same
point holds, and load through the reference's own `load_checkpoint` statements (train…:446-486) and a stock Adam."""
import os

import pytest
import torch

import siggan_oracle as O
from _util import make_gan

pytestmark = pytest.mark.gpu


def _clone_sd(sd):
    return {k: v.detach().cpu().clone() for k, v in sd.items()}


def test_async_checkpoints_equal_synchronous_captures(tmp_path):
    from checkpoint_writer import AsyncCheckpointWriter
    from vanilla_gan_model import VanillaGAN
    gan, _, _ = make_gan(64, 7, "bf16")
    real = O.synthetic_signatures(64, 64, seed=2).cuda()
    fixed_noise = torch.randn(16, 100, device="cuda")
    writer = AsyncCheckpointWriter(gan, tmp_path / "ckpt")
    for epoch in range(3):
        for _ in range(2):
            gan.train_step_async(real)                    # no host sync: the writer must order itself on the stream
        writer.save(epoch, global_step=2 * (epoch + 1), config={"latent_dim": 100, "image_size": 64, "image_channels": 1},
                    fixed_noise=fixed_noise, best_g_loss=1.5, is_best=(epoch == 1))
        gan.train_step_async(real)                        # already changing the parameters while the snapshot drains
    writer.wait()
    files = sorted(os.listdir(tmp_path / "ckpt"))
    assert files == ["checkpoint_best.pt", "checkpoint_epoch_0000.pt", "checkpoint_epoch_0001.pt",
                     "checkpoint_epoch_0002.pt", "checkpoint_latest.pt"], files
    ck = torch.load(tmp_path / "ckpt" / "checkpoint_latest.pt", map_location="cpu", weights_only=False)
    assert set(ck) == {"epoch", "global_step", "generator_state_dict", "discriminator_state_dict", "g_optimizer_state_dict",
                       "d_optimizer_state_dict", "config", "fixed_noise", "best_g_loss"}
    assert ck["epoch"] == 2 and ck["global_step"] == 6 and torch.equal(ck["fixed_noise"], fixed_noise.cpu())
    best = torch.load(tmp_path / "ckpt" / "checkpoint_best.pt", map_location="cpu", weights_only=False)
    assert best["epoch"] == 1
    # key order, shapes and dtypes are the modules' own
    for name, mod in (("generator_state_dict", gan.generator), ("discriminator_state_dict", gan.discriminator)):
        ref = mod.state_dict()
        assert list(ck[name].keys()) == list(ref.keys())
        for k in ref:
            assert ck[name][k].shape == ref[k].shape and ck[name][k].dtype == ref[k].dtype, k
    assert float(ck["g_optimizer_state_dict"]["state"][0]["step"]) == 8.0      # 3 x (2 + 1) steps run, snapshot after 8
    writer.close()


def test_snapshot_is_stream_ordered_and_loads_like_a_reference_checkpoint(tmp_path):
    """Capture synchronously right where save() is called, keep training, and compare after the writer finished."""
    from checkpoint_writer import AsyncCheckpointWriter
    from vanilla_gan_model import VanillaGAN
    gan, _, _ = make_gan(64, 8, "bf16")
    real = O.synthetic_signatures(64, 64, seed=3).cuda()
    writer = AsyncCheckpointWriter(gan, tmp_path / "ckpt", keep_epoch_files=False)
    gan.train_step_async(real)
    gan.train_step_async(real)
    writer.save(4, global_step=2)
    for _ in range(3):
        gan.train_step_async(real)      # enqueued BEFORE the synchronous capture below is even taken ...
    writer.wait()
    ck = torch.load(tmp_path / "ckpt" / "checkpoint_latest.pt", map_location="cpu", weights_only=False)
    # ... so replay: an identical run stopped after two steps is the ground truth
    import _siggan_lib as L
    gan_ref, _, _ = make_gan(64, 8, "bf16")
    gan3, _, _ = make_gan(64, 8, "bf16")

    # deterministic replay needs the same RNG / dropout stream: run both from a fixed state
    def run(g, n):
        L.DROPOUT.offset = 0
        torch.manual_seed(99)
        for _ in range(n):
            g.train_step_async(real)
        torch.cuda.synchronize()
    run(gan_ref, 2)
    w2 = AsyncCheckpointWriter(gan3, tmp_path / "ckpt2", keep_epoch_files=False)
    L.DROPOUT.offset = 0
    torch.manual_seed(99)
    gan3.train_step_async(real)
    gan3.train_step_async(real)
    w2.save(4, global_step=2)
    for _ in range(3):
        gan3.train_step_async(real)
    w2.close()
    ck2 = torch.load(tmp_path / "ckpt2" / "checkpoint_latest.pt", map_location="cpu", weights_only=False)
    for name, mod in (("generator_state_dict", gan_ref.generator), ("discriminator_state_dict", gan_ref.discriminator)):
        for k, v in mod.state_dict().items():
            assert torch.equal(ck2[name][k], v.cpu()), (name, k)          # the snapshot is the 2-step state, bit for bit
    for name, opt in (("g_optimizer_state_dict", gan_ref.g_optimizer), ("d_optimizer_state_dict", gan_ref.d_optimizer)):
        ref = opt.state_dict()
        assert ck2[name]["param_groups"] == ref["param_groups"]
        for i, st in ref["state"].items():
            assert float(ck2[name]["state"][i]["step"]) == float(st["step"]) == 2.0
            assert torch.equal(ck2[name]["state"][i]["exp_avg"], st["exp_avg"].cpu())
            assert torch.equal(ck2[name]["state"][i]["exp_avg_sq"], st["exp_avg_sq"].cpu())
    # the reference trainer's load_checkpoint statements (train…:466-470) on a fresh model, and a stock Adam
    gan4 = VanillaGAN(device="cuda")
    gan4.generator.load_state_dict(ck2["generator_state_dict"])
    gan4.discriminator.load_state_dict(ck2["discriminator_state_dict"])
    gan4.g_optimizer.load_state_dict(ck2["g_optimizer_state_dict"])
    gan4.d_optimizer.load_state_dict(ck2["d_optimizer_state_dict"])
    stock = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in gan4.discriminator.parameters()],
                             lr=2e-4, betas=(0.5, 0.999))
    stock.load_state_dict(ck2["d_optimizer_state_dict"])
    L.DROPOUT.offset = 0
    torch.manual_seed(5)
    m_a = gan4.train_step(real)
    L.DROPOUT.offset = 0
    torch.manual_seed(5)
    m_b = gan_ref.train_step(real)
    assert m_a == m_b, (m_a, m_b)       # the restored run continues exactly like the run the snapshot was taken from
    assert ck["epoch"] == 4 and ck["global_step"] == 2
    writer.close()
