"""The reference's own caller code over this repository's drop-in modules ("runs unchanged").

CPU part: state dicts interchange both ways between the reference's modules (oracle/_ref, imported under other names)
and the drop-ins, and the reference's `utils/inference.load_generator` (inference.py:57-104) builds and loads the drop-in
Generator from a checkpoint. GPU part: the reference's unchanged `GANTrainer.train()` (train…:134-640: its data loader,
`_train_discriminator` / `_train_generator`, sample grids, checkpoints, resume) and `generate_signatures_batch`
(inference.py:136-194) run over the drop-ins on cuda:0 in a process of its own (tests/run_reference_callers.py)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_ref  # noqa: E402

needs_ref = pytest.mark.skipif(not make_ref.check(), reason="oracle/_ref not built (python oracle/make_ref.py)")


def _load_ref(name):
    """Import a reference module from oracle/_ref/src under an alias, without putting that directory on sys.path."""
    path = os.path.join(make_ref.src_dir(), name + ".py")
    spec = importlib.util.spec_from_file_location("ref_" + name.replace("/", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_ref
def test_state_dicts_interchange_with_the_reference_modules():
    RG, RD = _load_ref("generator_vanilla_gan"), _load_ref("discriminator_vanilla_gan")
    from discriminator_vanilla_gan import Discriminator
    from generator_vanilla_gan import Generator
    for size in (64, 128):
        torch.manual_seed(size)
        rg, rd = RG.Generator(100, size), RD.Discriminator(size)
        g, d = Generator(100, size), Discriminator(size)
        g.load_state_dict(rg.state_dict(), strict=True)
        d.load_state_dict(rd.state_dict(), strict=True)
        for (k1, v1), (k2, v2) in zip(rg.state_dict().items(), g.state_dict().items()):
            assert k1 == k2 and v1.dtype == v2.dtype and torch.equal(v1, v2), k1
        torch.manual_seed(size + 1)
        g2, d2 = Generator(100, size), Discriminator(size)
        rg.load_state_dict(g2.state_dict(), strict=True)
        rd.load_state_dict(d2.state_dict(), strict=True)
        for (k1, v1), (k2, v2) in zip(rd.state_dict().items(), d2.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), k1
    rsn, sn = RD.Discriminator(64, use_spectral_norm=True), Discriminator(64, use_spectral_norm=True)
    sn.load_state_dict(rsn.state_dict(), strict=True)
    rsn.load_state_dict(sn.state_dict(), strict=True)


@needs_ref
def test_reference_load_generator_builds_the_dropin(tmp_path):
    """utils/inference.py resolves `from generator_vanilla_gan import Generator` to whatever is first on sys.path: here
    the drop-in. Every checkpoint flavour load_generator accepts (inference.py:76-92) must load."""
    RI = _load_ref("utils/inference")
    import generator_vanilla_gan as ours
    assert RI.Generator is ours.Generator
    g = ours.Generator(latent_dim=64, output_size=128)
    cases = {
        "full": {"config": {"latent_dim": 64, "image_size": 128, "image_channels": 1}, "generator_state_dict": g.state_dict()},
        "state_dict_key": {"state_dict": g.state_dict(), "config": {"latent_dim": 64, "image_size": 128}},
        "bare": g.state_dict(),            # architecture inferred from the key names / shapes (inference.py:21-55)
    }
    for name, ck in cases.items():
        path = tmp_path / f"{name}.pt"
        torch.save(ck, path)
        gen, latent = RI.load_generator(str(path), torch.device("cpu"))
        assert type(gen) is ours.Generator and latent == 64 and gen.output_size == 128 and not gen.training, name
        for (k, a), (_, b) in zip(g.state_dict().items(), gen.state_dict().items()):
            assert torch.equal(a, b), (name, k)
    assert RI.infer_architecture_from_state_dict(g.state_dict()) == (64, 128)


@needs_ref
@pytest.mark.gpu
def test_reference_trainer_and_inference_run_unchanged(tmp_path):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_callers.py"), str(tmp_path)],
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-3000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert r["model_module_is_dropin"] and r["trainer_module_is_reference"] and r["inference_module_is_reference"]
    assert r["inference_generator_is_dropin"] and r["trainer_model_is_dropin"] and r["loaded_generator_is_dropin"]
    # 48 images / batch 16, drop_last: 3 batches x 2 epochs
    assert r["global_step"] == 6 and r["adam_steps"] == [6.0, 6.0] and r["params_finite"]
    assert "epoch_0000.png" in r["samples"] and len(r["samples"]) >= 3
    assert "checkpoint_latest.pt" in r["checkpoints"]
    assert {"epoch", "global_step", "generator_state_dict", "discriminator_state_dict", "g_optimizer_state_dict",
            "d_optimizer_state_dict", "config", "fixed_noise", "best_g_loss"} <= set(r["checkpoint_keys"])
    # the reference saves 'epoch': epoch + 1 (train…:607) and resumes at checkpoint['epoch'] + 1 (train…:476): 3 after 2 epochs
    assert r["resume_epoch"] == 3 and r["resume_generator_equal"]
    assert all(v == v and abs(v) < 1e3 for v in r["resumed_step_metrics"].values())
    assert r["latent_dim"] == 100 and r["pil_count"] == 10 and r["pil_mode"] == "L" and r["pil_size"] == [64, 64]
    # the PIL images of the reference's per-image conversion loop == the fused uint8 egress on the same latents
    assert r["pil_equals_sample_uint8"], r["pil_max_abs_diff"]
