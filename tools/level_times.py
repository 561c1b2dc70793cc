"""Per-level device times of the C2 chain for a list of kernel variants (development probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
ws, levels, samples = int(os.environ.get("IBL_W", "512")), int(os.environ.get("IBL_LEVELS", "8")), int(os.environ.get("IBL_SAMPLES", "1024"))
bits = synth.synthetic_chain(ws, ws, levels)
offs = datum_b200.level_offsets(ws, ws, levels)
d_bits = torch.from_numpy(bits.view(np.int32)).to("cuda:0")
variants = [int(v) for v in os.environ.get("IBL_VARIANTS", "0,10000").split(",")]
ctx.set_tuning("table_order", int(os.environ.get("IBL_TABLE_ORDER", "0")))
ctx.set_prefilter_variant(0)
ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples); ctx.synchronize()
for variant in variants:
    ctx.set_prefilter_variant(variant)
    row = []
    for level in range(1, levels):
        w_src = ws >> (level - 1)
        src = d_bits[offs[level - 1]:offs[level]]
        dst = d_bits[offs[level]:offs[level + 1]]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        with torch.cuda.stream(ctx.torch_stream()):
            for rep in range(4):
                ev0.record(); ctx.prefilter_level_device(src, w_src, w_src, level, levels, samples, 0, 6 * (w_src >> 1), dst); ev1.record(); ev1.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
        row.append(best)
    ctx.set_prefilter_variant(variant)
    for rep in range(3):
        ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples)
        chain = ctx.last_prefilter_ms()
    print("variant %2d levels(us) %s sum %.3f ms chain %.3f ms" % (variant, " ".join("%7.1f" % (1e3 * t) for t in row), sum(row), chain), flush=True)
