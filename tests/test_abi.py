"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol, the
module surface and state-dict contract match the reference's, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "siggan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import _siggan_lib as L
    lib = L.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in siggan.h but not exported"
        assert name in L.SYMBOLS, f"{name} has no ctypes signature"
    assert lib.sg_abi_version() == 3


def test_no_cpu_path():
    import _siggan_lib as L
    from vanilla_gan_model import VanillaGAN, BCELoss
    gan = VanillaGAN(device="cpu")
    with pytest.raises(RuntimeError, match="CUDA only"):
        gan.generator(torch.randn(2, 100))
    with pytest.raises(RuntimeError, match="CUDA only"):
        gan.discriminator(torch.randn(2, 1, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA only"):
        BCELoss()(torch.rand(2, 1), torch.ones(2, 1))
    with pytest.raises(RuntimeError):
        gan.train_step(torch.randn(2, 1, 64, 64))
    if not torch.cuda.is_available():
        cfg = L.SgConfig(64, 100, 0, 0.2, 1e-5, 0.1)
        h = ctypes.c_void_p()
        assert L.load_library().sg_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
        assert b"no CUDA device" in L.load_library().sg_last_error()


def test_state_dict_contract_matches_reference(golden_dir):
    """Keys, shapes and dtypes dumped from the reference modules (tests/golden/contract.pt)."""
    from generator_vanilla_gan import Generator
    from discriminator_vanilla_gan import Discriminator
    gold = torch.load(os.path.join(golden_dir, "contract.pt"), weights_only=False)
    for size in (64, 128):
        G, D = Generator(100, size), Discriminator(size)
        assert [(k, tuple(v.shape), str(v.dtype)) for k, v in G.state_dict().items()] == gold[f"g{size}"]
        assert [(k, tuple(v.shape), str(v.dtype)) for k, v in D.state_dict().items()] == gold[f"d{size}"]
        assert G.get_num_params() == gold[f"g{size}.nparams"] and D.get_num_params() == gold[f"d{size}.nparams"]
        assert G.get_output_shape() == (1, size, size) and D.get_input_shape() == (1, size, size)
    Dsn = Discriminator(64, use_spectral_norm=True)     # disc…:61-62, 201-202: weight_orig / weight_u / weight_v
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in Dsn.state_dict().items()] == gold["d64sn"]


def test_module_surface():
    import generator_vanilla_gan as g, discriminator_vanilla_gan as d, vanilla_gan_model as v
    for name in ("UpsampleBlock", "Generator", "create_generator"):
        assert hasattr(g, name)
    for name in ("DownsampleBlock", "Discriminator", "MinibatchDiscrimination", "create_discriminator"):
        assert hasattr(d, name)
    for name in ("VanillaGAN", "create_vanilla_gan"):
        assert hasattr(v, name)
    with pytest.raises(ValueError, match="output_size must be 64 or 128"):
        g.Generator(output_size=32)
    with pytest.raises(ValueError, match="input_size must be 64 or 128"):
        d.Discriminator(input_size=32)
    gan = v.VanillaGAN(device="cpu")
    for attr in ("generator", "discriminator", "criterion", "g_optimizer", "d_optimizer", "current_epoch",
                 "global_step", "d_losses", "g_losses", "latent_dim", "label_smoothing", "device"):
        assert hasattr(gan, attr)
    for m in ("train_discriminator_step", "train_generator_step", "train_step", "generate", "generate_interpolation",
              "get_config", "save", "load", "from_checkpoint", "set_learning_rates", "get_recent_losses", "summary"):
        assert callable(getattr(gan, m))
    assert isinstance(gan.criterion, torch.nn.BCELoss)
    assert isinstance(gan.g_optimizer, torch.optim.Adam)
    cfg = gan.get_config()
    assert cfg["g_params"] == 1127201 and cfg["d_params"] == 2762689
    assert gan.g_optimizer.param_groups[0]["betas"] == (0.5, 0.999) and gan.g_optimizer.param_groups[0]["lr"] == 2e-4
    mb = d.MinibatchDiscrimination(8, 4)
    assert mb(torch.randn(3, 8)).shape == (3, 12)


def test_inference_helper_contract():
    """utils/inference.py:20-55 infers the architecture from key substrings of the generator state dict."""
    from generator_vanilla_gan import Generator
    for size, blocks in ((64, 4), (128, 5)):
        sd = Generator(100, size).state_dict()
        assert sum(1 for k in sd if "upsample_blocks" in k and ".0.weight" in k) == blocks
        fc = [k for k in sd if "fc" in k and "weight" in k and sd[k].dim() == 2]
        assert fc and sd[fc[0]].shape[1] == 100
