"""Profiling target for ncu: a few C2 chains (512^2 faces, 8 levels, 1024 spp) on cuda:0."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import datum_b200
from datum_b200 import synth

ws = int(os.environ.get("IBL_W", "512"))
levels = int(os.environ.get("IBL_LEVELS", "8"))
samples = int(os.environ.get("IBL_SAMPLES", "1024"))
chains = int(os.environ.get("IBL_CHAINS", "2"))
variant = int(os.environ.get("IBL_VARIANT", "0"))

ctx = datum_b200.IblContext(0)
ctx.set_prefilter_variant(variant)
bits = synth.synthetic_chain(ws, ws, levels)
d_bits = torch.from_numpy(bits.view(np.int32)).to("cuda:0")
for _ in range(chains):
    ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples)
ctx.synchronize()
print("prefilter ms", ctx.last_prefilter_ms())
