"""Target for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every shipped kernel once, small.

    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import datum_b200
from datum_b200 import synth

samples = int(os.environ.get("IBL_SAMPLES", "64"))
ctx = datum_b200.IblContext(0)

# 256^2 chain: level 1 = pair kernel with tile queues (4 warps), level 2 = pair kernel (8 warps), levels 3.. = tail kernel
w, levels = 256, 8
bits = synth.synthetic_chain(w, w, levels, probe=1)
ctx.image_buildmips_cube_ibl(w, w, levels, bits, samples)                       # pageable payload: staging path
pinned = torch.from_numpy(synth.synthetic_chain(w, w, levels, probe=2).view(np.int32).copy()).pin_memory()
ctx.image_buildmips_cube_ibl(w, w, levels, pinned, samples)

# a group of small probes: batched launches (pair kernel with queues over several probes, batched tail, batched SH9)
small = [synth.synthetic_chain(128, 128, 7, probe=10 + k) for k in range(6)]
sh = ctx.bake_probes(128, 128, 7, small, samples, sh9=True)

# SH9 of an RGBA32F cube, irradiance cube, LUTs, equirect pack, six-image ingest
cube = synth.synthetic_cube(96, 96, probe=3)
coeffs = ctx.project_sh9(cube, datum_b200.FORMAT_F32, 96, 96)
ctx.sh9_irradiance_cube(coeffs, 16, 16)
ctx.image_pack_envbrdf(32, 32, np.zeros(32 * 32, np.uint32), samples=64)
image = np.ones((64, 128, 4), np.float32)
image[..., :3] = np.random.default_rng(1).random((64, 128, 3), dtype=np.float32)
ctx.image_pack_cube_ibl(image, 32, 32, 4, np.zeros(datum_b200.image_datasize(32, 32, 6, 4) // 4, np.uint32), samples)
faces = np.random.default_rng(2).integers(0, 2**32, (6, 64, 64), dtype=np.uint64).astype(np.uint32)
ctx.skybox_from_argb32(faces, 5, np.zeros(datum_b200.image_datasize(64, 64, 6, 5) // 4, np.uint32), samples)
ctx.close()

# one probe shared by two contexts on this GPU: peer stores, last-CTA signal, stream waits
with datum_b200.MultiContext([0, 0]) as multi:
    shared = synth.synthetic_chain(192, 192, 7, probe=4)
    multi.image_buildmips_cube_ibl(192, 192, 7, shared, samples)
    multi.project_sh9(cube, datum_b200.FORMAT_F32, 96, 96)

print("sanitize target done", float(np.abs(sh).max()))
