import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
ws, levels, samples = 512, 8, 1024
bits = synth.synthetic_chain(ws, ws, levels)
d_bits = torch.from_numpy(bits.view(np.int32)).to("cuda:0")
d_dst = torch.zeros(6*256*256, dtype=torch.int32, device="cuda:0")
for variant in (11, 10, 19, 1, 6, 11):
    ctx.set_prefilter_variant(variant)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream()):
        for rep in range(4):
            ev0.record(); ctx.prefilter_level_device(d_bits, ws, ws, 1, levels, samples, 0, 6*256, d_dst); ev1.record(); ev1.synchronize()
            ms = ev0.elapsed_time(ev1)
    print("variant", variant, "level1 ms", ms, flush=True)
