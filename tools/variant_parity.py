"""Development probe: oracle parity of a pinned kernel variant (IBL_VARIANT, default 0) on spot rows; needs the tools build for A/B variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, datum_b200, parity
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
variant = int(os.environ.get("IBL_VARIANT", "0"))
for (ws, level, levels, samples) in [(512, 1, 8, 1024), (256, 2, 8, 1024), (128, 3, 8, 1024), (64, 2, 4, 4096)]:
    src = synth.synthetic_chain(ws, ws, 1, probe=11, sun=False)
    d_src = torch.from_numpy(src.view(np.int32)).to("cuda:0")
    wd = ws // 2
    out = torch.zeros(6 * wd * wd, dtype=torch.int32, device="cuda:0")
    f32 = torch.zeros(6 * wd * wd * 3, dtype=torch.float32, device="cuda:0")
    ctx.set_prefilter_variant(variant)
    ctx.prefilter_level_device(d_src, ws, ws, level, levels, samples, 0, 6 * wd, out, f32)
    ctx.synchronize()
    ctx.set_prefilter_variant(0)
    words, vals = out.cpu().numpy().view(np.uint32), f32.cpu().numpy().reshape(-1, 3)
    rng = [(0, 6 * wd)] if ws <= 128 else [(0, 4), (wd - 2, wd + 2), (2 * wd + wd // 2, 2 * wd + wd // 2 + 4), (6 * wd - 4, 6 * wd)]
    for a, b in rng:
        r = parity.check_level(words, vals, src, ws, ws, level, levels, samples, a, b)
    print("variant", variant, "parity ok", ws, level, levels, samples, r["max_rel_clean"], r["identical"], flush=True)
