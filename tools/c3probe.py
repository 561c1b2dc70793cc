import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, datum_b200
variant = int(os.environ.get("IBL_VARIANT", "0"))
cases = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("IBL_CASES", "64,7,4096;128,8,4096;256,9,4096;512,10,4096").split(";")]
ctx = datum_b200.IblContext(0)
ctx.set_prefilter_variant(variant)
for (w, levels, samples) in cases:
    n = sum(6 * (w >> i) ** 2 for i in range(levels))
    d_bits = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda:0")
    try:
        ctx.buildmips_cube_ibl_device(w, w, levels, d_bits, samples); ctx.synchronize()
        print("ok", w, levels, samples, variant, "ms", ctx.last_prefilter_ms(), flush=True)
    except Exception as e:
        print("FAIL", w, levels, samples, variant, e, flush=True); break
