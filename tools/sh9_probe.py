"""Development probe: SH9 partial + combine device time by face size and texel format."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200
ctx = datum_b200.IblContext(0)
stream = ctx.torch_stream()
for w in (64, 256, 512, 1024, 2048):
    for fmt, name in ((datum_b200.FORMAT_RGBE, "rgbe"), (datum_b200.FORMAT_F32, "f32")):
        if fmt == datum_b200.FORMAT_RGBE:
            cube = torch.randint(0, 2**31 - 1, (6 * w * w,), dtype=torch.int32, device="cuda:0")
        else:
            cube = torch.rand((6 * w * w, 4), dtype=torch.float32, device="cuda:0")
        out = torch.zeros(28, dtype=torch.float64, device="cuda:0")
        ctx.sh9_partial_device(cube, fmt, w, w, 0, 6 * w, out); ctx.synchronize()
        best = 1e9
        for rep in range(5):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                ev0.record(); ctx.sh9_partial_device(cube, fmt, w, w, 0, 6 * w, out); ev1.record()
            ev1.synchronize(); best = min(best, ev0.elapsed_time(ev1))
        print("w %4d %s: %.1f us  (%.1f Gtexel/s)" % (w, name, best * 1e3, 6 * w * w / best / 1e6), flush=True)
