import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, datum_b200, parity
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
ctx.set_tuning("table_order", 1)
for (w, levels, samples) in [(128, 8, 1024), (256, 8, 1024), (64, 4, 4096)]:
    bits = synth.synthetic_chain(w, w, levels, probe=11, sun=False)
    offs = datum_b200.level_offsets(w, w, levels)
    d_bits = torch.from_numpy(bits.view(np.int32).copy()).to("cuda:0")
    d_f32 = torch.zeros((offs[-1] - offs[1]) * 3, dtype=torch.float32, device="cuda:0")
    ctx.buildmips_cube_ibl_device(w, w, levels, d_bits, samples, d_f32); ctx.synchronize()
    got = d_bits.cpu().numpy().view(np.uint32); got_f32 = d_f32.cpu().numpy().reshape(-1, 3)
    for level in range(1, levels):
        ws = w >> (level - 1)
        hd = ws >> 1
        rng = [(0, 6 * hd)] if ws <= 128 else [(0, 4), (hd - 2, hd + 2), (6 * hd - 4, 6 * hd)]
        for a, b in rng:
            r = parity.check_level(got[offs[level]:offs[level + 1]], got_f32[offs[level] - offs[1]:offs[level + 1] - offs[1]], got[offs[level - 1]:offs[level]], ws, ws, level, levels, samples, a, b)
    print("order 1 parity ok", w, levels, samples, r["max_rel_clean"], r["identical"])
