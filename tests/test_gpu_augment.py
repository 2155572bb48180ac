"""GPU parity of the device input pipeline (sg_augment_batch through DeviceSignatureLoader) — bit-exact against the
fixtures produced by the reference's own `get_train_transforms` pipeline (src/data_loader_signatures.py:154-219) and
against the CPU oracle on random images / parameters, plus size-independent properties at a full batch."""
import os

import numpy as np
import pytest
import torch

import augment_oracle as A

pytestmark = pytest.mark.gpu


def _loader(images, **kw):
    from device_data_loader import DeviceSignatureLoader
    return DeviceSignatureLoader(images, device="cuda", **kw)


@pytest.mark.parametrize("name", ["augment_64.pt", "augment_64_flip.pt", "augment_128.pt"])
def test_kernel_equals_reference_pipeline_output(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name), weights_only=False)
    ld = _loader(gold["images"], batch_size=len(gold["angles"]))
    idx = torch.arange(len(gold["angles"]), dtype=torch.int32, device="cuda")
    flips = gold["flips"].numpy() if gold["flips"] is not None else None
    out = ld.batch(idx, gold["angles"].numpy(), gold["scales"].numpy(), flips)
    assert out.dtype == torch.float32 and out.shape == gold["out"].shape
    assert torch.equal(out.cpu(), gold["out"])


@pytest.mark.parametrize("size,n", [(64, 301), (128, 67)])
def test_kernel_equals_oracle_on_random_inputs(size, n):
    rng = np.random.default_rng(size + n)
    pool = (rng.random((n + 13, size, size)) * 256).astype(np.uint8)
    ang = rng.uniform(-5, 5, n)
    ang[:8] = [0, 90, 180, 270, -90, 360, 45, -133.7]
    ang[8:40] = rng.uniform(-400, 400, 32)
    sc = rng.uniform(0.9, 1.1, n)
    sc[:5] = [1.0, 0.9, 1.1, 0.4, 2.5]
    fl = (rng.random(n) < 0.5).astype(np.uint8)
    index = rng.permutation(n + 13)[:n].astype(np.int32)
    ld = _loader(torch.from_numpy(pool), batch_size=n)
    out = ld.batch(torch.from_numpy(index).cuda(), ang, sc, fl).cpu().numpy()
    for i in range(n):
        ref = A.augment(pool[index[i]], float(ang[i]), float(sc[i]), bool(fl[i]))
        assert np.array_equal(out[i, 0], ref), (i, ang[i], sc[i], fl[i])
    # batch of one, no flip table
    one = ld.batch(torch.tensor([5], dtype=torch.int32, device="cuda"), [3.3], [1.07]).cpu().numpy()
    assert np.array_equal(one[0, 0], A.augment(pool[5], 3.3, 1.07))


def test_full_batch_properties_and_epoch_coverage():
    from _siggan_lib import load_library  # noqa: F401  (the loader fails loudly if the library is missing)
    n, size, B = 8192, 64, 4096
    g = torch.Generator().manual_seed(3)
    pool = torch.randint(0, 256, (n, size, size), dtype=torch.uint8, generator=g)
    plain = (pool.float() / 255.0 - 0.5) / 0.5
    # augmentation off == ToTensor + Normalize of the pool, every image exactly once per epoch (drop_last keeps 2 batches)
    ld = _loader(pool, batch_size=B, augment=False, shuffle=True, seed=1)
    assert len(ld) == 2
    seen = torch.cat([b.cpu() for b in ld])
    assert seen.shape == (n, 1, size, size)
    key = lambda t: t.reshape(t.shape[0], -1).double() @ torch.linspace(1, 2, size * size, dtype=torch.float64)
    assert torch.equal(torch.sort(key(seen[:, 0]))[0], torch.sort(key(plain))[0])
    # identity parameters are the identity; a flip applied twice through the pipeline is the identity
    idx = torch.arange(B, dtype=torch.int32, device="cuda")
    same = ld.batch(idx, np.zeros(B), np.ones(B))
    assert torch.equal(same.cpu()[:, 0], plain[:B])
    flipped = ld.batch(idx, np.zeros(B), np.ones(B), np.ones(B, dtype=np.uint8))
    assert torch.equal(flipped.cpu()[:, 0], plain[:B].flip(-1))
    # augmented batches stay in range, keep the white fill at the corners of a rotated frame, and differ per draw
    ld2 = _loader(pool, batch_size=B, seed=2)
    a, b = ld2.batch(idx), ld2.batch(idx)
    assert a.min() >= -1 and a.max() <= 1 and not torch.equal(a, b)
    white = _loader(torch.full((16, size, size), 255, dtype=torch.uint8), batch_size=16).batch(idx[:16])
    assert torch.equal(white, torch.ones_like(white))
