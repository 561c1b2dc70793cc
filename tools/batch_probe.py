"""Development probe: where a batched bake (datum_ibl_bake_probes) spends its time per probe."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
for (w, levels, samples, n) in [(16, 5, 1024, 256), (256, 8, 1024, 128), (512, 8, 1024, 64)]:
    distinct = [torch.from_numpy(synth.synthetic_chain(w, w, levels, probe=k).view(np.int32).copy()).pin_memory() for k in range(4)]
    payloads = [distinct[i % 4] for i in range(n)]
    dev = [d.cuda() for d in distinct]
    ctx.bake_probes(w, w, levels, payloads[:4], samples, sh9=True)
    for sh9 in (False, True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.bake_probes(w, w, levels, payloads, samples, sh9=sh9)
        dt = time.perf_counter() - t0
        print("w %d batch of %d sh9=%s: %.1f us per probe" % (w, n, sh9, dt / n * 1e6), flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        ctx.buildmips_cube_ibl_device(w, w, levels, dev[i % 4], samples)
    t_enq = time.perf_counter() - t0
    ctx.synchronize(); dt = time.perf_counter() - t0
    print("w %d device-resident loop: %.1f us per probe (enqueue alone %.1f us)" % (w, dt / n * 1e6, t_enq / n * 1e6), flush=True)
