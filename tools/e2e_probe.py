"""Where the end-to-end time of image_buildmips_cube_ibl goes (development probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
ws, levels, samples = 512, 8, 1024
bits = synth.synthetic_chain(ws, ws, levels)
offs = datum_b200.level_offsets(ws, ws, levels)
pinned = torch.from_numpy(bits.view(np.int32).copy()).pin_memory()
d = torch.empty_like(pinned, device="cuda:0")
s = ctx.torch_stream()
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s): d[:offs[1]].copy_(pinned[:offs[1]], non_blocking=True)
    s.synchronize()
def d2h():
    with torch.cuda.stream(s): pinned[offs[1]:].copy_(d[offs[1]:], non_blocking=True)
    s.synchronize()
def chain():
    ctx.buildmips_cube_ibl_device(ws, ws, levels, d, samples); ctx.synchronize()
def e2e():
    ctx.image_buildmips_cube_ibl(ws, ws, levels, pinned, samples)
pageable = bits.copy()
def e2e_pageable():
    ctx.image_buildmips_cube_ibl(ws, ws, levels, pageable, samples)
print("h2d level0 %.3f ms (%.1f GB/s)" % (t(h2d), offs[1]*4/t(h2d)/1e6))
print("d2h levels %.3f ms" % t(d2h))
print("chain on device (wall, synced) %.3f ms, events %.3f ms" % (t(chain), ctx.last_prefilter_ms()))
print("e2e pinned %.3f ms" % t(e2e))
print("e2e pageable %.3f ms" % t(e2e_pageable))
