"""Executed by tests/test_reference_callers.py in a process of its own, with sys.path arranged so that
  * `generator_vanilla_gan`, `discriminator_vanilla_gan`, `vanilla_gan_model` resolve to THIS repository's drop-in modules,
  * `train_vanilla_gan_signatures`, `data_loader_signatures`, `utils.*` resolve to the UNMODIFIED reference (oracle/_ref).
It drives the reference's own caller code — GANTrainer.train() (train…:134-640) and utils/inference.py — over the drop-in
modules on cuda:0 and prints one JSON object with what the test asserts on."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_ref  # noqa: E402

ref_src = make_ref.src_dir()
sys.path[:0] = [os.path.join(ROOT, "signature-gan_b200"), os.path.join(ROOT, "tests", "stubs"), ref_src]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402


def main(work: str) -> None:
    out = {}
    import vanilla_gan_model
    import generator_vanilla_gan
    import train_vanilla_gan_signatures as T          # the reference's trainer, unchanged
    from utils import inference as RI                 # the reference's inference helpers, unchanged
    pkg = os.path.join(ROOT, "signature-gan_b200")
    out["model_module_is_dropin"] = os.path.realpath(vanilla_gan_model.__file__).startswith(os.path.realpath(pkg))
    out["trainer_module_is_reference"] = os.path.realpath(T.__file__).startswith(os.path.realpath(ref_src))
    out["inference_module_is_reference"] = os.path.realpath(RI.__file__).startswith(os.path.realpath(ref_src))
    out["inference_generator_is_dropin"] = RI.Generator is generator_vanilla_gan.Generator

    # ---- a small directory of signature-like PNGs for the reference's own data loader
    data = os.path.join(work, "data")
    os.makedirs(data, exist_ok=True)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siggan_oracle as O
    imgs = ((O.synthetic_signatures(48, 64, seed=3)[:, 0] + 1) * 127.5).clamp(0, 255).to(torch.uint8).numpy()
    for i, im in enumerate(imgs):
        Image.fromarray(im, mode="L").save(os.path.join(data, f"sig_{i:03d}.png"))

    # ---- GANTrainer.train(): data loader -> _train_discriminator / _train_generator -> samples -> checkpoints
    cfg = T.TrainingConfig(batch_size=16, epochs=2, num_workers=0, sample_interval=1, checkpoint_interval=1,
                           data_dir=data, checkpoint_dir=os.path.join(work, "ckpt"), sample_dir=os.path.join(work, "samples"),
                           log_dir=os.path.join(work, "logs"))
    torch.manual_seed(0)
    trainer = T.GANTrainer(cfg, device="cuda")
    out["trainer_model_is_dropin"] = type(trainer.model) is vanilla_gan_model.VanillaGAN
    summary = trainer.train()
    out["train_summary_keys"] = sorted(summary.keys()) if isinstance(summary, dict) else None
    out["global_step"] = trainer.global_step
    out["samples"] = sorted(os.listdir(cfg.sample_dir))
    out["checkpoints"] = sorted(os.listdir(cfg.checkpoint_dir))
    ck = torch.load(os.path.join(cfg.checkpoint_dir, "checkpoint_latest.pt"), map_location="cpu", weights_only=False)
    out["checkpoint_keys"] = sorted(ck.keys())
    out["adam_steps"] = [float(ck["g_optimizer_state_dict"]["state"][0]["step"]),
                         float(ck["d_optimizer_state_dict"]["state"][0]["step"])]
    out["params_finite"] = all(bool(torch.isfinite(v).all()) for v in ck["generator_state_dict"].values()
                               if v.is_floating_point())
    # resume through the reference's own loader of its own checkpoint
    trainer2 = T.GANTrainer(cfg, device="cuda")
    out["resume_epoch"] = trainer2.load_checkpoint()
    same = all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(trainer.model.generator.state_dict().values(),
                                                             trainer2.model.generator.state_dict().values()))
    out["resume_generator_equal"] = bool(same)
    m = trainer2._train_discriminator(torch.from_numpy(imgs[:16]).float().div(127.5).sub(1).unsqueeze(1))
    m.update(trainer2._train_generator(16))
    out["resumed_step_metrics"] = {k: float(v) for k, v in m.items() if v is not None}   # grad norms: None (clipping off)

    # ---- utils/inference.py: load_generator + generate_signatures_batch on the checkpoint the trainer wrote
    dev = torch.device("cuda")
    gen, latent = RI.load_generator(os.path.join(cfg.checkpoint_dir, "checkpoint_latest.pt"), dev)
    out["loaded_generator_is_dropin"] = type(gen) is generator_vanilla_gan.Generator
    out["latent_dim"] = latent
    pil = RI.generate_signatures_batch(gen, 10, latent, dev, seed=3, batch_size=4)
    got = np.stack([np.array(p) for p in pil])
    # the same latents through the fused uint8 egress (Generator.sample_uint8)
    torch.manual_seed(3)
    torch.cuda.manual_seed_all(3)
    zs = [torch.randn(n, latent, device=dev) for n in (4, 4, 2)]
    with torch.no_grad():
        u8 = torch.cat([gen.sample_uint8(z) for z in zs]).cpu().numpy()[:, 0]
    out["pil_count"], out["pil_mode"], out["pil_size"] = len(pil), pil[0].mode, list(pil[0].size)
    out["pil_equals_sample_uint8"] = bool((got == u8).all())
    out["pil_max_abs_diff"] = int(np.abs(got.astype(np.int32) - u8.astype(np.int32)).max())
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1])
