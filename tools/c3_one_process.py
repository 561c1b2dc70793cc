"""Development probe: BASELINE config 3 from ONE process over the GPUs of the box (datum_ibl_multi_*), end to end from a pinned host payload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth
world = int(os.environ.get("IBL_DEVICES", str(torch.cuda.device_count())))
w, levels, samples = 2048, 12, 4096
source = synth.synthetic_chain(w, w, levels, probe=3)
ctx = datum_b200.IblContext(0)
alone = torch.from_numpy(source.view(np.int32).copy()).pin_memory()
ctx.image_buildmips_cube_ibl(w, w, levels, alone, samples)
t0 = time.perf_counter(); ctx.image_buildmips_cube_ibl(w, w, levels, alone, samples); ms_alone = (time.perf_counter() - t0) * 1e3
shared = torch.from_numpy(source.view(np.int32).copy()).pin_memory()
with datum_b200.MultiContext(list(range(world))) as multi:
    multi.image_buildmips_cube_ibl(w, w, levels, shared, samples)
    times = []
    for _ in range(5):
        t0 = time.perf_counter(); multi.image_buildmips_cube_ibl(w, w, levels, shared, samples); times.append((time.perf_counter() - t0) * 1e3)
print("one GPU %.2f ms; %d GPUs from one process: %s ms; words identical %.6f" % (ms_alone, world, " ".join("%.2f" % t for t in times), float((shared.numpy() == alone.numpy()).mean())), flush=True)
