"""Profiling target for ncu: the SH9 projection of one 4096^2 RGBA32F cube (BASELINE config 5 on one GPU), a few times."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import datum_b200

w = int(os.environ.get("IBL_W", "4096"))
ctx = datum_b200.IblContext(0)
ctx.set_tuning("sh9_kernel", int(os.environ.get("IBL_SH9_KERNEL", "0")))
gen = torch.Generator(device="cuda:0")
gen.manual_seed(5)
cube = torch.rand((6 * w * w, 4), dtype=torch.float32, device="cuda:0", generator=gen)
out = torch.zeros(28, dtype=torch.float64, device="cuda:0")
for _ in range(int(os.environ.get("IBL_REPS", "3"))):
    ctx.sh9_partial_device(cube, datum_b200.FORMAT_F32, w, w, 0, 6 * w, out)
ctx.synchronize()
print("weight sum", float(out[27].item()))
