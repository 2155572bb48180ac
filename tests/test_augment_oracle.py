"""Pins oracle/augment_oracle.py (the input-pipeline augmentation, reference src/data_loader_signatures.py:154-219) to
torchvision + Pillow themselves and to fixtures produced by the reference's own `get_train_transforms` pipeline, and
checks the library's host-side table builder `sg_augment_params` against it bit for bit (no GPU involved)."""
import os
import warnings

import numpy as np
import pytest
import torch

import augment_oracle as A


def test_oracle_matches_torchvision_and_pillow_bit_for_bit():
    from PIL import Image
    from torchvision.transforms import functional as F
    rng = np.random.default_rng(0)
    for size in (64, 128):
        for t in range(120):
            img = (rng.random((size, size)) * 256).astype(np.uint8)
            ang = float(rng.uniform(-5, 5)) if t % 10 else float(rng.choice([0, 90, 180, 270, -90, 360, 45.0, -180, 725.5]))
            sc = float(rng.uniform(0.9, 1.1)) if t % 7 else 1.0
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pil = Image.fromarray(img)
            rot = F.rotate(pil, ang, fill=255)                                   # RandomRotation.forward
            rot_o = A.rotate_nearest(img, A.rotation_fixed(ang, size))
            assert np.array_equal(np.asarray(rot), rot_o), (size, ang)
            aff = F.affine(rot, angle=0.0, translate=[0, 0], scale=sc, shear=[0.0, 0.0], fill=255)   # RandomAffine.forward
            aff_o = A.scale_nearest(rot_o, A.scale_params(sc, size))
            assert np.array_equal(np.asarray(aff), aff_o), (size, sc)
            ten = F.normalize(F.to_tensor(aff), [0.5], [0.5])[0].numpy()         # ToTensor, Normalize
            assert np.array_equal(ten, A.to_normalised(aff_o))
            assert np.array_equal(A.augment(img, ang, sc), ten)


@pytest.mark.parametrize("name", ["augment_64.pt", "augment_64_flip.pt", "augment_128.pt"])
def test_oracle_matches_reference_pipeline_fixture(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name), weights_only=False)
    imgs, angles, scales = gold["images"].numpy(), gold["angles"].tolist(), gold["scales"].tolist()
    flips = gold["flips"].tolist() if gold["flips"] is not None else [0] * len(angles)
    for i in range(len(angles)):
        got = A.augment(imgs[i], angles[i], scales[i], bool(flips[i]))
        assert np.array_equal(got, gold["out"][i, 0].numpy()), (name, i)


@pytest.mark.parametrize("size", [64, 128])
def test_host_table_builder_matches_oracle(size):
    import _siggan_lib as L
    lib = L.load_library()
    rng = np.random.default_rng(size)
    n = 5000
    ang = rng.uniform(-5, 5, n)
    ang[:9] = [0, 90, 180, 270, -90, 360, -180, 45, -1e-300]
    ang[9:300] = rng.uniform(-800, 800, 291)
    sc = rng.uniform(0.9, 1.1, n)
    sc[:4] = [1.0, 0.9, 1.1, 0.5]
    rot, sa = np.empty((n, 6), np.int32), np.empty((n, 4), np.float64)
    assert lib.sg_augment_params(ang.ctypes.data, sc.ctypes.data, n, size, rot.ctypes.data, sa.ctypes.data) == 0
    rot_o, sa_o = A.parameter_tables(ang, sc, size)
    assert np.array_equal(rot, rot_o)
    assert np.array_equal(sa, sa_o)
    # NULL angles / scales = identity tables; empty input is fine; bad sizes are refused without touching memory
    assert lib.sg_augment_params(None, None, 3, size, rot.ctypes.data, sa.ctypes.data) == 0
    ident_r, ident_s = A.parameter_tables([0, 0, 0], [1, 1, 1], size)
    assert np.array_equal(rot[:3], ident_r) and np.array_equal(sa[:3], ident_s)
    assert lib.sg_augment_params(ang.ctypes.data, sc.ctypes.data, 0, size, rot.ctypes.data, sa.ctypes.data) == 0
    assert lib.sg_augment_params(ang.ctypes.data, sc.ctypes.data, 1, 96, rot.ctypes.data, sa.ctypes.data) != 0


def test_loader_refuses_cpu():
    from device_data_loader import DeviceSignatureLoader
    imgs = torch.zeros(4, 64, 64, dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="CUDA"):
        DeviceSignatureLoader(imgs, device="cpu")
    with pytest.raises(TypeError):
        DeviceSignatureLoader(imgs.float(), device="cpu")
    with pytest.raises(ValueError):
        DeviceSignatureLoader(torch.zeros(4, 32, 32, dtype=torch.uint8), device="cpu")


def test_directory_pool_equals_reference_decode_and_resize(tmp_path):
    """DeviceSignatureLoader.load_directory_uint8 == per file Image.open(...).convert('L') + transforms.Resize((S, S))
    (data_loader_signatures.py:120-135, 176-177), with SignatureDataset's discovery rule (:88-103)."""
    from PIL import Image
    from torchvision import transforms
    from device_data_loader import DeviceSignatureLoader
    rng = np.random.default_rng(4)
    names = ["b.png", "a.PNG", "c.jpg", "d.bmp", "skip.txt", "e.gif"]
    for i, name in enumerate(names):
        arr = (rng.random((90 + 7 * i, 150 - 11 * i, 3)) * 255).astype(np.uint8)
        if name.endswith((".txt", ".gif")):
            (tmp_path / name).write_bytes(b"not an image the loader should look at")
        else:
            Image.fromarray(arr).save(tmp_path / name)
    (tmp_path / "sub").mkdir()
    Image.fromarray(np.zeros((20, 20), np.uint8)).save(tmp_path / "sub" / "nested.png")     # not recursive
    for size in (64, 128):
        pool = DeviceSignatureLoader.load_directory_uint8(tmp_path, size)
        expect = sorted(p for p in tmp_path.iterdir() if p.suffix.lower() in (".png", ".jpg", ".bmp"))
        assert pool.shape == (len(expect), size, size) and pool.dtype == torch.uint8
        for k, p in enumerate(expect):
            ref = transforms.Resize((size, size))(Image.open(p).convert("L"))
            assert np.array_equal(pool[k].numpy(), np.asarray(ref)), p.name
    with pytest.raises(ValueError, match="does not exist"):
        DeviceSignatureLoader.load_directory_uint8(tmp_path / "missing")
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(ValueError, match="No images"):
        DeviceSignatureLoader.load_directory_uint8(empty)
