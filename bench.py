#!/usr/bin/env python
"""Benchmark of the IBL bake hot path (BASELINE.json: prefiltered texel-samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one GGX-prefilter chain bake of one synthetic HDR environment map
(BASELINE config 2: 512^2 faces, 8 levels, 1024 samples/texel, rgbe payload as
tools/assetbuilder.cpp hands it to image_buildmips_cube_ibl) per GPU.  Probes
are independent units, so N GPUs bake N probes per step with no data-path
collective (weak scaling); the only collectives are the barrier and the
max-over-ranks of the timings.

Prints ONE JSON line on rank 0.  `value` is timed on the device (CUDA events,
max over ranks) with inputs resident in HBM; `e2e` goes through the reference-
facing entry point image_buildmips_cube_ibl with pinned HOST buffers, host<->device
copies inside the timed region.  `roofline` is the level-1 prefilter launch
(75 % of the work) against the FP32 FMA peak measured in the same process;
`cpu_baseline` is the unmodified reference tools/ibl.cpp (compiled into
oracle/_ref with the reference's own -O2 -ffast-math) on the host cores.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, LEVELS, SAMPLES = 512, 8, 1024            # BASELINE.json configs[1]
POOL = 24                                        # distinct probes resident in HBM: 24 x 8.4 MB > 126 MB L2
FLOP_PER_TEXEL_SAMPLE = 85.0                     # SURVEY.md §8d, itemised in DESIGN.md
CPU_SAMPLE = (64, 7)                             # cpu_baseline leg: one 64^2-face 7-level reference bake per thread and round
CPU_STEP_SAMPLE = (32, 6)                        # --impl reference: one 32^2-face 6-level bake per thread and step (~0.3 s)

METRIC = "prefiltered texel-samples/sec"
UNIT = "texel-samples/s"


def texel_samples(width, levels, samples):
    return sum(6 * (width >> i) * (width >> i) for i in range(1, levels)) * samples


# ---- the reference's own CPU implementation (bounded sample) ---------------------------

def cpu_reference_run(rounds, threads=None, sample=CPU_SAMPLE):
    """`threads` concurrent calls of the UNMODIFIED reference image_buildmips_cube_ibl
    (oracle/_ref, the reference's -O2 -ffast-math flags) on 64^2 x 7-level synthetic
    chains, `rounds` times.  The reference function is single-threaded and has no global
    state; using every host core means one independent bake per core.  Falls back to the
    oracle port when oracle/_ref was not built.  Returns (texel-samples/s, info)."""
    import oracle_lib
    from datum_b200 import synth

    threads = threads or os.cpu_count() or 1
    w, levels = sample

    if oracle_lib.have_ref(fast=True):
        lib = oracle_lib.ref(fast=True)
        kind = "reference"

        def bake(bits):
            lib.ref_image_buildmips_cube_ibl(w, w, levels, bits.ctypes.data)
    else:
        kind = "port"

        def bake(bits):
            oracle_lib.buildmips_cube_ibl(w, w, levels, bits, samples=1024, threads=1)

    payloads = [synth.synthetic_chain(w, w, levels, probe=100 + t) for t in range(threads)]

    def worker(bits):
        for _ in range(rounds):
            bake(bits)          # ctypes releases the GIL for the duration of the call

    pool = [threading.Thread(target=worker, args=(p,)) for p in payloads]
    t0 = time.perf_counter()
    for t in pool:
        t.start()
    for t in pool:
        t.join()
    dt = time.perf_counter() - t0

    work = texel_samples(w, levels, 1024) * threads * rounds
    info = {
        "kind": kind,
        "cores": threads,
        "sample": "%d concurrent %s bakes of a %d^2-face %d-level rgbe chain at 1024 spp, x%d rounds (%.3g texel-samples, %.1f s)"
                  % (threads, "unmodified tools/ibl.cpp image_buildmips_cube_ibl" if kind == "reference" else "oracle port", w, levels, rounds, work, dt),
    }
    return work / dt, dt, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0

    for _ in range(args.warmup):
        cpu_reference_run(1, sample=CPU_STEP_SAMPLE)

    t0 = time.perf_counter()
    value, dt, info = cpu_reference_run(max(1, args.steps), sample=CPU_STEP_SAMPLE)
    ms_per_step = (time.perf_counter() - t0) * 1e3 / max(1, args.steps)

    info["value"] = value
    info["unit"] = UNIT
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 synthetic HDR cube 512^2 faces x 8 levels x 1024 spp (rgbe payload); CPU arm times a bounded sample of it",
                   "step": info["sample"]},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def hbm_side(bytes_per_launch, ms_per_launch):
    """The same launch against the HBM roof (it is nowhere near it: the prefilter is FP32-bound)."""
    peak, source = 6650.0, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, source = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch > 0 else 0.0
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": source}


# ---- clock sampling during the timed region ---------------------------------------------

class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the measurement runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
    }

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                util = int(nv.nvmlDeviceGetUtilizationRates(self._handle).gpu)
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                self.samples.append((mhz, util))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._nvml:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        clocks = [m for m, _ in self.samples]
        return {
            "sm_mhz": float(np.median(clocks)) if clocks else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(clocks),
        }


# ---- our arm -------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    import datum_b200
    from datum_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — datum_b200 has no CPU path to benchmark")

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = datum_b200.IblContext(local_rank)
    stream = ctx.torch_stream()

    w, levels, samples = WIDTH, LEVELS, SAMPLES
    offs = datum_b200.level_offsets(w, w, levels)
    step_work = texel_samples(w, levels, samples)

    # ---- inputs: a pool of distinct probes resident in HBM (larger than L2), and pinned host copies for e2e
    pool_host = [synth.synthetic_chain(w, w, levels, probe=rank * POOL + p) for p in range(POOL)]
    pool_dev = [torch.from_numpy(b.view(np.int32)).to(device) for b in pool_host]
    pinned = [torch.from_numpy(b.view(np.int32).copy()).pin_memory() for b in pool_host[:4]]
    torch.cuda.synchronize()

    fp32_peak = ctx.measure_fp32_peak()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident throughput: W warm-up + K timed steps, CUDA events on the launching stream
    for i in range(args.warmup):
        ctx.buildmips_cube_ibl_device(w, w, levels, pool_dev[i % POOL], samples)
    ctx.synchronize()
    ctx.dominant_kernel_stats(reset=True)
    launches_before = ctx.launch_count

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    with torch.cuda.stream(stream):
        ev0.record()
        for i in range(args.steps):
            ctx.buildmips_cube_ibl_device(w, w, levels, pool_dev[(args.warmup + i) % POOL], samples)
        ev1.record()
    ev1.synchronize()
    torch.cuda.synchronize()
    barrier()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launch_count - launches_before
    dom_n, dom_ms, dom_ts = ctx.dominant_kernel_stats(reset=True)

    # ---- end to end through the reference-facing call: pinned host payload in, baked payload out
    for i in range(min(3, max(1, args.warmup))):
        ctx.image_buildmips_cube_ibl(w, w, levels, pinned[i % len(pinned)], samples)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctx.image_buildmips_cube_ibl(w, w, levels, pinned[i % len(pinned)], samples)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    # ---- the same K bakes as ONE batched call (datum_ibl_bake_probes): copies of probe i+1 / i-1 under the kernels of probe i
    batch = [pinned[i % len(pinned)] for i in range(args.steps)]
    ctx.bake_probes(w, w, levels, batch[: min(4, len(batch))], samples)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    ctx.bake_probes(w, w, levels, batch, samples)
    e2e_batch_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    clocks = sampler.stop()

    total_launches = int(sum_over_ranks(launches))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * step_work * args.steps / (elapsed_ms * 1e-3)
    e2e_value = world * step_work * args.steps / e2e_s

    achieved = FLOP_PER_TEXEL_SAMPLE * dom_ts / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0

    cpu_info = None
    if world == 1:
        cpu_value, _, cpu_info = cpu_reference_run(rounds=2)
        cpu_info["value"] = cpu_value
        cpu_info["unit"] = UNIT

    line = {
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "C2: synthetic HDR cube map, 512^2 faces x 8 levels x 1024 GGX samples/texel, rgbe payload (one env map per GPU per step)",
            "texel_samples_per_step_per_gpu": step_work,
            "env_maps_per_s": world * args.steps / (elapsed_ms * 1e-3),
            "l2": "rotating pool of %d distinct probes per GPU (%d MB) > 126 MB L2; each step bakes the next one" % (POOL, POOL * offs[-1] * 4 // 2**20),
            "parallelism": "probe-sharded x%d, no data-path collective" % world,
        },
        "e2e": {
            "value": e2e_value, "unit": UNIT,
            "h2d_bytes_per_step": offs[1] * 4, "d2h_bytes_per_step": (offs[-1] - offs[1]) * 4,
            "api": "image_buildmips_cube_ibl(width, height, levels, bits) on a pinned host payload",
            "ms_per_step": e2e_s * 1e3 / args.steps,
        },
        "e2e_batched": {
            "value": world * step_work * args.steps / e2e_batch_s, "unit": UNIT,
            "api": "bake_probes: %d pinned host payloads in one datum_ibl_bake_probes call per GPU (uploads, kernels, downloads overlapped over two device payloads)" % args.steps,
            "ms_per_step": e2e_batch_s * 1e3 / args.steps,
        },
        "gpu_launches": total_launches,
        "roofline": {
            "bound": "fp32", "kernel": "prefilter_dp_kernel (level 1: 512^2 -> 256^2 faces)",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
            "peak_source": "FFMA-chain micro-benchmark run in this process (datum_ibl_measure_fp32_peak); MEASURED_PEAKS.json carries no FP32 figure",
            "flop_per_texel_sample": FLOP_PER_TEXEL_SAMPLE, "texel_samples_per_launch": dom_ts,
            "launches_timed": dom_n, "ms_per_launch": dom_ms,
            "traffic": 25.2e6, "traffic_source": "dram__bytes_read+write of one ncu --set full capture (profiles/)",
            "hbm": hbm_side(25.2e6, dom_ms),
        },
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
    }
    if cpu_info is not None:
        line["cpu_baseline"] = cpu_info

    print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=50)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = parser.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(0, args.warmup)
    args.steps = max(1, args.steps)

    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
