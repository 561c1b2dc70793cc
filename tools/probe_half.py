"""Development probe (GPU box): parity of the half-record prefilter variants against the oracle,
then per-level timings on C2.  Not a test, not the benchmark."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import datum_b200, parity
from datum_b200 import synth

ctx = datum_b200.IblContext(0)
dev = "cuda:0"

def check(variant, w, levels, samples, sun):
    ctx.set_prefilter_variant(variant)
    bits = synth.synthetic_chain(w, w, levels, probe=3, noise=True, sun=sun)
    offs = datum_b200.level_offsets(w, w, levels)
    d_bits = torch.from_numpy(bits.view(np.int32).copy()).to(dev)
    d_f32 = torch.zeros((offs[-1] - offs[1]) * 3, dtype=torch.float32, device=dev)
    ctx.buildmips_cube_ibl_device(w, w, levels, d_bits, samples, d_f32)
    ctx.synchronize()
    got = d_bits.cpu().numpy().view(np.uint32); gf = d_f32.cpu().numpy().reshape(-1, 3)
    for level in range(1, levels):
        ws = w >> (level - 1)
        try:
            rep = parity.check_level(got[offs[level]:offs[level + 1]], gf[offs[level] - offs[1]:offs[level + 1] - offs[1]],
                                     got[offs[level - 1]:offs[level]], ws, ws, level, levels, samples)
            print("variant", variant, "w", w, "sun", sun, "ok", {k: (round(v, 8) if isinstance(v, float) else v) for k, v in rep.items()}, flush=True)
        except AssertionError as e:
            print("variant", variant, "w", w, "sun", sun, "FAIL", str(e)[:400], flush=True)

variants = [int(v) for v in os.environ.get("IBL_VARIANTS", "50,51").split(",")]
if os.environ.get("IBL_PARITY", "1") == "1":
    for v in variants[:2]:
        check(v, 64, 5, 1024, False)
        check(v, 128, 4, 1024, True)

ws, levels, samples = 512, 8, 1024
bits = synth.synthetic_chain(ws, ws, levels)
offs = datum_b200.level_offsets(ws, ws, levels)
d_bits = torch.from_numpy(bits.view(np.int32)).to(dev)
for variant in [0] + [int(v) for v in os.environ.get("IBL_TIMED", "30,31,32,33,34,35,36,37,38,39,42").split(",")]:
    ctx.set_prefilter_variant(variant)
    row = []
    for level in range(1, 4):
        w_src = ws >> (level - 1)
        src = d_bits[offs[level - 1]:offs[level]]
        dst = torch.zeros(6 * (w_src >> 1) ** 2, dtype=torch.int32, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        with torch.cuda.stream(ctx.torch_stream()):
            for rep in range(4):
                ev0.record(); ctx.prefilter_level_device(src, w_src, w_src, level, levels, samples, 0, 6 * (w_src >> 1), dst); ev1.record(); ev1.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
        row.append(best)
    for rep in range(3):
        ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples)
        chain = ctx.last_prefilter_ms()
    print("variant %2d  levels 1-3 (us, incl. record build) %s  chain %.3f ms" % (variant, " ".join("%7.1f" % (1e3 * t) for t in row), chain), flush=True)
