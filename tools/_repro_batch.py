import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, datum_b200
from datum_b200 import synth
ctx = datum_b200.IblContext(0)
def run(w, levels, n, warm, sh9, distinct_n=4):
    distinct = [torch.from_numpy(synth.synthetic_chain(w, w, levels, probe=k).view(np.int32).copy()).pin_memory() for k in range(distinct_n)]
    payloads = [distinct[i % distinct_n] for i in range(n)]
    try:
        ctx.bake_probes(w, w, levels, payloads[:warm], 1024, sh9=sh9)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.bake_probes(w, w, levels, payloads, 1024, sh9=sh9)
        dt = time.perf_counter() - t0
        print("w %d n %d warm %d sh9 %s: %.1f us per probe" % (w, n, warm, sh9, dt / n * 1e6), flush=True)
    except Exception as e:
        print("w %d n %d warm %d sh9 %s: ERROR %s" % (w, n, warm, sh9, e), flush=True)
run(512, 8, 20, 4, False)
run(512, 8, 20, 4, False)
run(256, 8, 128, 8, True, 8)
run(256, 8, 128, 8, True, 8)
run(256, 8, 256, 8, True, 8)
run(512, 8, 64, 4, False)
