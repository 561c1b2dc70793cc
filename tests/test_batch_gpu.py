"""Batched bakes (datum_ibl_bake_probes): the results of `count` single calls, with uploads,
kernels and downloads of consecutive probes overlapped."""

import numpy as np
import pytest
import torch

import datum_b200
from datum_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = datum_b200.IblContext(0)
    yield c
    c.close()


def single_calls(ctx, w, levels, samples, probes):
    out = []
    for p in probes:
        bits = synth.synthetic_chain(w, w, levels, probe=p)
        ctx.image_buildmips_cube_ibl(w, w, levels, bits, samples)
        out.append(bits)
    return out


@pytest.mark.parametrize("count", [1, 2, 5])
@pytest.mark.parametrize("pinned", [False, True])
def test_batch_equals_single_calls(ctx, count, pinned):
    w, levels, samples = 64, 6, 256
    probes = list(range(40, 40 + count))
    want = single_calls(ctx, w, levels, samples, probes)
    payloads = [synth.synthetic_chain(w, w, levels, probe=p) for p in probes]
    if pinned:
        payloads = [torch.from_numpy(b.view(np.int32).copy()).pin_memory() for b in payloads]
    sh = ctx.bake_probes(w, w, levels, payloads, samples, sh9=True)
    assert sh.shape == (count, 9, 3)
    for i in range(count):
        got = payloads[i].numpy().view(np.uint32) if pinned else payloads[i]
        assert np.array_equal(got, want[i])
        n0 = 6 * w * w
        level0 = np.ascontiguousarray(want[i][:n0])
        assert np.array_equal(sh[i], ctx.project_sh9(level0, datum_b200.FORMAT_RGBE, w, w).reshape(9, 3))


def test_batch_at_the_c4_probe_size(ctx):
    """BASELINE config 4's unit (256^2 faces, 8 levels, 1024 spp), three probes."""
    w, levels, samples = 256, 8, 1024
    probes = [3, 4, 5]
    want = single_calls(ctx, w, levels, samples, probes)
    payloads = [synth.synthetic_chain(w, w, levels, probe=p) for p in probes]
    assert ctx.bake_probes(w, w, levels, payloads, samples) is None
    for got, ref in zip(payloads, want):
        assert np.array_equal(got, ref)
    # ... and a single call AFTER the batch (regression: the batch's launch shapes and a single call's
    # share shared-memory sizes; a per-size cache once skipped the second kernel's set-up)
    again = single_calls(ctx, w, levels, samples, probes[:1])
    assert np.array_equal(again[0], want[0])


def test_a_group_of_sixteen_small_probes_equals_single_calls(ctx):
    """Probes well below C2's size are baked up to 16 per launch (level L of the whole group in one grid):
    same words as single calls, SH9 included, also when the last group is short."""
    w, levels, samples = 64, 7, 128
    probes = list(range(200, 219))            # 19 probes: one group of 16 and a group of 3
    want = single_calls(ctx, w, levels, samples, probes)
    payloads = [torch.from_numpy(synth.synthetic_chain(w, w, levels, probe=p).view(np.int32).copy()).pin_memory() for p in probes]
    sh = ctx.bake_probes(w, w, levels, payloads, samples, sh9=True)
    for i in range(len(probes)):
        assert np.array_equal(payloads[i].numpy().view(np.uint32), want[i]), i
        level0 = np.ascontiguousarray(want[i][: 6 * w * w])
        assert np.array_equal(sh[i], ctx.project_sh9(level0, datum_b200.FORMAT_RGBE, w, w).reshape(9, 3))


def test_empty_batch_and_bad_arguments(ctx):
    assert ctx.bake_probes(16, 16, 4, [], 64) is None
    with pytest.raises(ValueError):
        ctx.bake_probes(16, 16, 4, [np.zeros(10, np.uint32)], 64)          # payload too small
    with pytest.raises(datum_b200.IblError):
        ctx.bake_probes(16, 16, 6, [np.zeros(6 * 400, np.uint32)], 64)     # 16 >> 5 == 0
    with pytest.raises(datum_b200.IblError):
        ctx.bake_probes(16, 16, 4, [np.zeros(6 * 400, np.uint32)], 0)      # samples < 1
