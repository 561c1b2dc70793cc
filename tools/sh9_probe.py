"""SH9 projection kernels on one GPU (development probe): the column-strip kernel against the row-segment
kernel, run lengths, the whole 4096^2 cube and the 3072-row slab one of eight GPUs owns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, datum_b200

ctx = datum_b200.IblContext(0)
w = int(os.environ.get("IBL_W", "4096"))
gen = torch.Generator(device="cuda:0"); gen.manual_seed(5)
cube = torch.rand((6 * w * w, 4), dtype=torch.float32, device="cuda:0", generator=gen)
out = torch.zeros(28, dtype=torch.float64, device="cuda:0")


def timed(begin, end, reps=10):
    best = 1e9
    with torch.cuda.stream(ctx.torch_stream()):
        for _ in range(reps):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); ctx.sh9_partial_device(cube, datum_b200.FORMAT_F32, w, w, begin, end, out); ev1.record(); ev1.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
    return best, out.cpu().numpy().copy()


ctx.set_tuning("sh9_kernel", 1)
base_ms, base = timed(0, 6 * w)
print("row segments   full %.4f ms  %.0f GB/s" % (base_ms, 6 * w * w * 16 / base_ms / 1e6), flush=True)
slab_ms, slab_base = timed(0, 6 * w // 8)
print("row segments   slab %.4f ms  %.0f GB/s" % (slab_ms, 6 * w * w * 2 / slab_ms / 1e6), flush=True)
ctx.set_tuning("sh9_kernel", 0)
for rows in (0, 8, 16, 32, 64, 128):
    ctx.set_tuning("sh9_rows_per_item", rows)
    ms, got = timed(0, 6 * w)
    err = np.abs(got - base).max() / np.abs(base).max()
    ms8, got8 = timed(0, 6 * w // 8)
    err8 = np.abs(got8 - slab_base).max() / np.abs(slab_base).max()
    print("columns rows/item %3d  full %.4f ms  %.0f GB/s  (vs row segments %.1e)   slab %.4f ms  %.0f GB/s (%.1e)" % (rows, ms, 6 * w * w * 16 / ms / 1e6, err, ms8, 6 * w * w * 2 / ms8 / 1e6, err8), flush=True)
