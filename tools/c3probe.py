"""Development probe: chain times at the other BASELINE sizes (C3: 2048^2 x 12 levels x 4096 spp, C4: 256^2 x 8 x 1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
cases = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("IBL_CASES", "256,8,1024;2048,12,4096").split(";")]
variants = [int(v) for v in os.environ.get("IBL_VARIANTS", "0").split(",")]
for (w, levels, samples) in cases:
    n = sum(6 * (w >> i) ** 2 for i in range(levels))
    bits = synth.synthetic_chain(w, w, 1)
    d_bits = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    d_bits[: 6 * w * w] = torch.from_numpy(bits.view(np.int32)).to("cuda:0")
    ts = sum(6 * (w >> i) ** 2 for i in range(1, levels)) * samples
    for variant in variants:
        ctx.set_prefilter_variant(variant)
        best = 1e9
        for rep in range(3):
            ctx.buildmips_cube_ibl_device(w, w, levels, d_bits, samples); ctx.synchronize()
            best = min(best, ctx.last_prefilter_ms())
        print("w %d levels %d spp %d variant %d: %.3f ms  %.3e texel-samples/s" % (w, levels, samples, variant, best, ts / best * 1e3), flush=True)
