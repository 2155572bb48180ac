#!/usr/bin/env python
"""A few sg_augment_batch launches on random 8-bit images, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:augment_kernel -c 2 -o gpurun_out/aug python tools/aug_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "signature-gan_b200"))
from device_data_loader import DeviceSignatureLoader  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
pool = torch.randint(0, 256, (4 * B, size, size), dtype=torch.uint8, device="cuda")
ld = DeviceSignatureLoader(pool, batch_size=B, seed=1)
idx = torch.randperm(4 * B, device="cuda")[:B].to(torch.int32)
for _ in range(4):
    out = ld.batch(idx)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.mean()))
