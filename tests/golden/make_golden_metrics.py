"""Golden fixture for the ink statistics (reference src/utils/metrics.py:118-174). Runs ONLY in the build container:
calls the reference's unmodified calculate_stroke_density / calculate_foreground_ratio on seeded batches in [-1, 1]
(generator output range) and in [0, 1] (no rescale) -> tests/golden/metrics_64.pt.

    python tests/golden/make_golden_metrics.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O  # noqa: E402  (also puts the reference's src/ on sys.path)


def main():
    import importlib.util
    # utils/metrics.py imports optional packages lazily; load the module file directly to avoid utils/__init__ side effects
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(os.environ.get(
        "SIGGAN_REFERENCE_SRC", "/root/reference/src"), "utils", "metrics.py"))
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)
    out = {}
    for name, x in O.metric_batches().items():
        for thr in (0.5, 0.3):
            out[f"{name}.{thr}.stroke"] = M.calculate_stroke_density(x.clone(), threshold=thr)
            out[f"{name}.{thr}.foreground"] = M.calculate_foreground_ratio(x.clone(), threshold=thr)
    torch.save(out, os.path.join(HERE, "metrics_64.pt"))
    print("metrics_64.pt:", {k: v for k, v in list(out.items())[:2]})


if __name__ == "__main__":
    main()
