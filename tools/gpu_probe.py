"""Development probe run on the GPU box: parity of the prefilter against the oracle on
small cubes, then timings of every kernel variant on the C2 workload.  Not a test, not
the benchmark — output goes to gpurun_out/ for reading back."""

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datum_b200
from datum_b200 import synth
import oracle_lib

out = {}
ctx = datum_b200.IblContext(0)
dev = torch.device("cuda", 0)


def parity(ws, levels, samples, noise, variant):
    ctx.set_prefilter_variant(variant)
    bits = synth.synthetic_chain(ws, ws, levels, probe=1, noise=noise)
    offs = datum_b200.level_offsets(ws, ws, levels)
    d_bits = torch.from_numpy(bits.view(np.int32)).to(dev)
    d_f32 = torch.zeros((offs[-1] - offs[1]) * 3, dtype=torch.float32, device=dev)
    ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples, d_f32)
    ctx.synchronize()
    got = d_bits.cpu().numpy().view(np.uint32)
    got_f32 = d_f32.cpu().numpy().reshape(-1, 3)
    rows = []
    for level in range(1, levels):
        w_src = ws >> (level - 1)
        src = got[offs[level - 1]:offs[level]]          # same-source: oracle runs on OUR previous level
        ow, of = oracle_lib.prefilter_level(src, w_src, w_src, level, levels, samples)
        gw = got[offs[level]:offs[level + 1]]
        gf = got_f32[offs[level] - offs[1]:offs[level + 1] - offs[1]]
        rel = oracle_lib.relative_error(gf, of)
        amb = oracle_lib.edge_ambiguous_counts(w_src >> 1, w_src >> 1, level, levels, samples) > 0
        st = oracle_lib.word_stats(gw[~amb], ow[~amb]) if (~amb).any() else {}
        rows.append(dict(level=level, clean_max_rel=float(rel[~amb].max()) if (~amb).any() else None,
                         amb_frac=float(amb.mean()), amb_max_rel=float(rel[amb].max()) if amb.any() else 0.0, **st))
    return rows


for (ws, levels, samples, noise, variant) in [(32, 6, 1024, True, 0), (64, 7, 1024, False, 1), (64, 7, 256, True, 2), (128, 8, 1024, True, 0)]:
    key = "parity_%d_%d_%d_%s_v%d" % (ws, levels, samples, "noise" if noise else "smooth", variant)
    try:
        out[key] = parity(ws, levels, samples, noise, variant)
    except Exception as e:  # keep going: this is a probe
        out[key] = "ERROR " + repr(e)
    print(key, json.dumps(out[key]), flush=True)

# ---- timings on C2: 512^2, 8 levels, 1024 spp ----
ws, levels, samples = 512, 8, 1024
bits = synth.synthetic_chain(ws, ws, levels)
d_bits = torch.from_numpy(bits.view(np.int32)).to(dev)
ts = sum(6 * (ws >> i) ** 2 for i in range(1, levels)) * samples
out["fp32_peak_tflops"] = ctx.measure_fp32_peak()
print("fp32 peak", out["fp32_peak_tflops"], flush=True)
for variant in (1, 6, 10, 11, 12, 13, 14, 15, 0):
    ctx.set_prefilter_variant(variant)
    times = []
    for rep in range(6):
        ctx.buildmips_cube_ibl_device(ws, ws, levels, d_bits, samples)
        times.append(ctx.last_prefilter_ms())
    best = min(times[1:])
    out["c2_variant_%d" % variant] = dict(ms=best, all=times, texel_samples_per_s=ts / (best * 1e-3), frac_of_fp32=85 * ts / (best * 1e-3) / 1e12 / out["fp32_peak_tflops"])
    print("variant", variant, out["c2_variant_%d" % variant], flush=True)

# level-1 only timing per variant (75% of the work)
d_dst = torch.zeros(6 * 256 * 256, dtype=torch.int32, device=dev)
for variant in (1, 6, 10, 11, 12, 13, 15):
    ctx.set_prefilter_variant(variant)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream()):
        for rep in range(3):
            ev0.record()
            ctx.prefilter_level_device(d_bits, ws, ws, 1, levels, samples, 0, 6 * 256, d_dst)
            ev1.record()
            ev1.synchronize()
            ms = ev0.elapsed_time(ev1)
    out["level1_variant_%d_ms" % variant] = ms
    print("level1 variant", variant, ms, flush=True)

# e2e host call (pinned)
pinned = torch.from_numpy(bits.view(np.int32).copy()).pin_memory()
ctx.set_prefilter_variant(0)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.image_buildmips_cube_ibl(ws, ws, levels, pinned, samples)
    dt = time.perf_counter() - t0
out["c2_e2e_ms"] = dt * 1e3
print("e2e ms", dt * 1e3)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
