"""Ink statistics (reference src/utils/metrics.py:118-174): the numpy oracle against fixtures produced by the reference's
own functions (CPU), and the device path (sg_ink_stats through device_metrics) against both (GPU, bit-exact counts)."""
import os

import numpy as np
import pytest
import torch

import augment_oracle as A
import siggan_oracle as O

THRESHOLDS = (0.5, 0.3)


def _same(a, b):
    if isinstance(a, dict):
        return set(a) == set(b) and all(_same(a[k], b[k]) for k in a)
    return a == b                    # floats produced by the same float32 operations: exact equality


def test_oracle_matches_reference_functions(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "metrics_64.pt"), weights_only=False)
    for name, x in O.metric_batches().items():
        for thr in THRESHOLDS:
            assert _same(A.stroke_density(x.numpy(), thr), gold[f"{name}.{thr}.stroke"]), (name, thr)
            assert _same(A.foreground_ratio(x.numpy(), thr), gold[f"{name}.{thr}.foreground"]), (name, thr)


@pytest.mark.gpu
def test_device_metrics_match_reference_and_oracle(golden_dir):
    import _siggan_lib as L
    import device_metrics as M
    gold = torch.load(os.path.join(golden_dir, "metrics_64.pt"), weights_only=False)
    lib = L.load_library()
    for name, x in O.metric_batches().items():
        xc = x.cuda()
        for thr in THRESHOLDS:
            assert _same(M.calculate_stroke_density(xc, threshold=thr), gold[f"{name}.{thr}.stroke"]), (name, thr)
            assert _same(M.calculate_foreground_ratio(xc, threshold=thr), gold[f"{name}.{thr}.foreground"]), (name, thr)
            n = x.shape[0]
            raw, res = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
            mn = torch.empty(n, device="cuda")
            assert lib.sg_ink_stats(xc.data_ptr(), n, 64 * 64, thr, raw.data_ptr(), res.data_ptr(), mn.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream) == 0
            o_raw, o_res, o_min = A.ink_counts(x.numpy(), thr)
            assert np.array_equal(raw.cpu().numpy(), o_raw) and np.array_equal(res.cpu().numpy(), o_res)
            assert np.array_equal(mn.cpu().numpy(), o_min)
    # full sampling batch, 128x128, ragged batch, single image
    for shape in ((16384, 1, 64, 64), (37, 1, 128, 128), (1, 1, 64, 64)):
        g = torch.Generator(device="cuda").manual_seed(shape[0])
        x = torch.rand(shape, device="cuda", generator=g) * 2 - 1
        frac = M._ink_fraction(x, 0.5)
        ref = (((x + 1) / 2) < 0.5).float().view(shape[0], -1).mean(dim=1).cpu().numpy()
        assert np.array_equal(frac, ref)
    with pytest.raises(NotImplementedError):
        M.calculate_stroke_density(torch.zeros(2, 3, 64, 64, device="cuda"))


def test_device_metrics_refuse_cpu():
    import device_metrics as M
    with pytest.raises(RuntimeError, match="CUDA only"):
        M.calculate_stroke_density(torch.zeros(2, 1, 64, 64))
