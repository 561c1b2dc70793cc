"""BASELINE configs 3, 4 and 5 on N GPUs of one node (development / evidence run, not bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/scale_configs.py

  C3  one 2048^2 x 12-level x 4096-spp probe shared by all ranks: (a) slabs exchanged with an NCCL
      all-gather per level, (b) slabs stored into the peers' chains by the kernel epilogue (PeerChain)
  C4  256 probes of 256^2 x 8 levels x 1024 spp (+ SH9), probe p on rank p % N, batched host entry
  C5  SH9 of one 4096^2 RGBA32F cube, rows split, 28 doubles all-reduced

Device times are CUDA events on the bake stream, max over ranks; C4 is host wall clock (it includes the
copies), max over ranks.  One JSON line per config on rank 0, also appended to gpurun_out/scale_configs.jsonl."""

import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import datum_b200
from datum_b200 import dist as ibl_dist
from datum_b200 import synth

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
configs = os.environ.get("IBL_CONFIGS", "c3,c4,c5").split(",")

torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=device)

ctx = datum_b200.IblContext(local)
engine = ibl_dist.CudaEngine(ctx)
stream = ctx.torch_stream()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def emit(line):
    if rank == 0:
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "scale_configs.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
            fn()
            ev1.record()
        ev1.synchronize()
        best = min(best, max_over_ranks(ev0.elapsed_time(ev1)))
    return best


if "c3" in configs:
    w, levels, samples = int(os.environ.get("IBL_C3_W", "2048")), int(os.environ.get("IBL_C3_LEVELS", "12")), int(os.environ.get("IBL_C3_SAMPLES", "4096"))
    offs = datum_b200.level_offsets(w, w, levels)
    work = sum(6 * (w >> i) ** 2 for i in range(1, levels)) * samples
    level0 = torch.from_numpy(synth.synthetic_chain(w, w, 1, probe=3).view(np.int32)).to(device)

    # every GPU alone (the reference point)
    alone = torch.zeros(offs[-1], dtype=torch.int32, device=device)
    alone[: offs[1]] = level0
    ms_alone = timed(lambda: ctx.buildmips_cube_ibl_device(w, w, levels, alone, samples), reps=2)

    chain = torch.zeros(offs[-1], dtype=torch.int32, device=device)
    chain[: offs[1]] = level0
    ms_nccl = timed(lambda: ibl_dist.bake_single_probe(engine, chain, w, w, levels, samples))

    shared = ibl_dist.PeerChain(ctx, w, w, levels)
    with torch.cuda.stream(stream):
        shared.chain[: offs[1]] = level0
    ms_peer = timed(lambda: shared.bake(samples))
    ctx.synchronize()

    same_nccl = float((chain == alone).float().mean().item())
    same_peer = float((shared.chain == alone).float().mean().item())
    peer_equals_nccl = bool(torch.equal(shared.chain, chain))
    shared.close()

    emit({"config": "C3", "workload": "one %d^2 x %d-level x %d-spp probe shared by %d GPU(s), rows of every level split" % (w, levels, samples, world),
          "n_gpus": world, "texel_samples": work,
          "ms_one_gpu_alone": ms_alone,
          "ms_nccl_all_gather": ms_nccl, "texel_samples_per_s_nccl": work / ms_nccl * 1e3,
          "ms_peer_stores": ms_peer, "texel_samples_per_s_peer_stores": work / ms_peer * 1e3,
          "speedup_vs_one_gpu_peer_stores": ms_alone / ms_peer, "speedup_vs_one_gpu_nccl": ms_alone / ms_nccl,
          "words_identical_to_one_gpu": {"nccl": same_nccl, "peer_stores": same_peer}, "peer_stores_equal_nccl_words": peer_equals_nccl})
    del alone, chain, level0

if "c4" in configs:
    w, levels, samples, probes = 256, 8, 1024, int(os.environ.get("IBL_C4_PROBES", "256"))
    mine = ibl_dist.shard_probes(probes, rank, world)
    distinct = [torch.from_numpy(synth.synthetic_chain(w, w, levels, probe=1000 + rank * 8 + k).view(np.int32).copy()).pin_memory() for k in range(8)]
    payloads = [distinct[i % len(distinct)] for i in range(len(mine))]
    work = sum(6 * (w >> i) ** 2 for i in range(1, levels)) * samples * probes
    ctx.bake_probes(w, w, levels, payloads[:4], samples, sh9=True)
    best = 1e30
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        sh = ctx.bake_probes(w, w, levels, payloads, samples, sh9=True)
        best = min(best, max_over_ranks(time.perf_counter() - t0))
    emit({"config": "C4", "workload": "%d probes of %d^2 x %d levels x %d spp, prefilter + SH9, probe p on rank p %% %d, one batched host call per rank (pinned payloads, copies included)" % (probes, w, levels, samples, world),
          "n_gpus": world, "texel_samples": work, "seconds": best, "texel_samples_per_s": work / best, "probes_per_s": probes / best,
          "sh9_finite": bool(np.isfinite(sh).all())})

if "c5" in configs:
    w = int(os.environ.get("IBL_C5_W", "4096"))
    gen = torch.Generator(device=device)
    gen.manual_seed(5)
    cube = torch.rand((6 * w * w, 4), dtype=torch.float32, device=device, generator=gen)
    begin, end = ibl_dist.split_rows(6 * w, world)[rank]
    out = torch.zeros(28, dtype=torch.float64, device=device)
    ctx.sh9_partial_device(cube, datum_b200.FORMAT_F32, w, w, begin, end, out)     # builds the solid-angle table
    ctx.synchronize()

    def project():
        ctx.sh9_partial_device(cube, datum_b200.FORMAT_F32, w, w, begin, end, out)
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)

    ms = timed(project, reps=5)
    ms_kernel = timed(lambda: ctx.sh9_partial_device(cube, datum_b200.FORMAT_F32, w, w, begin, end, out), reps=5)

    # the same without a collective: partial sums stored into the peers' arrays by the kernel + barrier kernel
    peer_sh = ibl_dist.PeerSh9(ctx)

    def project_peers():
        peer_sh.enqueue(cube, datum_b200.FORMAT_F32, w, w)

    ms_peers = timed(project_peers, reps=5)
    sh_peers = peer_sh.project(cube, datum_b200.FORMAT_F32, w, w)
    sh_nccl = ibl_dist.project_sh9_single_probe(engine, cube, datum_b200.FORMAT_F32, w, w)
    peers_match = float(np.abs(sh_peers - sh_nccl).max() / np.abs(sh_nccl).max())
    peer_sh.close()
    rows = end - begin
    emit({"config": "C5", "workload": "SH9 of one %d^2 RGBA32F cube, %d rows per GPU, all-reduce of 28 doubles" % (w, rows),
          "n_gpus": world, "texels": 6 * w * w, "ms": ms, "ms_kernels_only": ms_kernel, "ms_peer_stores": ms_peers, "peer_stores_vs_nccl_max_rel": peers_match,
          "texels_per_s": 6 * w * w / ms * 1e3, "hbm_gb_per_s_per_gpu_kernels_only": rows * w * 16 / (ms_kernel * 1e-3) / 1e9})

ctx.close()
if world > 1:
    dist.destroy_process_group()
