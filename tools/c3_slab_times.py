"""BASELINE config 3 on ONE GPU, slab by slab (development probe): for every level of the 2048^2 x 12-level x
4096-spp chain the time of the whole level and of the row slab rank r of `world` would own — where an
8-GPU bake loses against 1/8 of the one-GPU time, without needing 8 GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth, dist as ibl_dist

ctx = datum_b200.IblContext(0)
variant = int(os.environ.get("IBL_VARIANT", "0"))
w, levels, samples = int(os.environ.get("IBL_W", "2048")), int(os.environ.get("IBL_LEVELS", "12")), int(os.environ.get("IBL_SAMPLES", "4096"))
world = int(os.environ.get("IBL_WORLD", "8"))
offs = datum_b200.level_offsets(w, w, levels)
bits = synth.synthetic_chain(w, w, levels, probe=3)
d_bits = torch.from_numpy(bits.view(np.int32)).to("cuda:0")
ctx.buildmips_cube_ibl_device(w, w, levels, d_bits, samples); ctx.synchronize()
plan = ibl_dist.plan_single_probe(w, w, levels, world)


def timed(fn, reps=3):
    best = 1e9
    with torch.cuda.stream(ctx.torch_stream()):
        for _ in range(reps):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); fn(); ev1.record(); ev1.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
    return best


ctx.set_prefilter_variant(variant)
print("variant", variant)
total_full, total_slab = 0.0, 0.0
for step in plan:
    level, ws = step["level"], step["ws"]
    src, dst = d_bits[offs[level - 1]:offs[level]], d_bits[offs[level]:offs[level + 1]]
    full = timed(lambda: ctx.prefilter_level_device(src, ws, ws, level, levels, samples, 0, step["rows"], dst))
    slabs = []
    for r in sorted(set([0, world // 2 - 1, world - 1])):
        a, b = step["ranges"][r]
        slabs.append(timed(lambda: ctx.prefilter_level_device(src, ws, ws, level, levels, samples, a, b, dst)))
    worst = max(slabs)
    total_full += full
    total_slab += worst
    print("level %2d faces %4d^2 split %-5s full %9.1f us  /%d = %8.1f us   slab (worst of %d ranks) %8.1f us   loss %6.1f us" % (level, ws >> 1, step["split"], full * 1e3, world, full * 1e3 / world, len(slabs), worst * 1e3, (worst - full / world) * 1e3 if step["split"] else worst * 1e3 - full * 1e3 / world), flush=True)
print("sum: full %.3f ms, /%d = %.3f ms, slabs %.3f ms" % (total_full, world, total_full / world, total_slab))
