"""GPU parity of the SH9 projection (data/project.comp), the irradiance cube
(data/lighting.inc:351-371) and the two LUTs of tools/ibl.h through the C ABI."""

import os

import numpy as np
import pytest
import torch

import datum_b200
import oracle_lib
from datum_b200 import synth, FORMAT_F32, FORMAT_RGBE

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ibl_golden.npz"))
DEV = "cuda:0"
TOL_SH = 1e-4     # relative to the largest coefficient of the probe (north_star: SH <= 1e-4 relative)


def sh_error(got, want):
    want = np.asarray(want, np.float64)
    return np.abs(np.asarray(got, np.float64) - want).max() / np.abs(want).max()


@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (8, 8), (37, 23), (64, 64), (256, 256)])
@pytest.mark.parametrize("fmt", [FORMAT_F32, FORMAT_RGBE])
def test_sh9_matches_the_fp64_oracle(ctx, w, h, fmt):
    cube = synth.synthetic_cube(w, h, probe=21, sun=(w >= 64))
    level0 = cube if fmt == FORMAT_F32 else synth.rgbe_words(cube)
    want = oracle_lib.project_sh9(level0, fmt, w, h)
    got = ctx.project_sh9(level0, fmt, w, h)
    assert got.shape == (9, 3) and got.dtype == np.float32          # Irradiance::L, envmap.h:112-115
    assert sh_error(got, want) <= TOL_SH


def test_sh9_known_answers(ctx):
    w = 64
    c = np.array([0.5, 1.0, 2.0])
    level0 = np.ones((6, w, w, 4), np.float32)
    level0[..., :3] = c
    sh = ctx.project_sh9(level0, FORMAT_F32, w, w)
    assert np.allclose(sh[0], 0.282095 * 4 * np.pi * c, rtol=1e-5)
    assert np.abs(sh[1:]).max() <= 1e-4 * sh[0].max()
    level0[..., :3] = synth.cube_directions(w, w)[..., 1:2]
    sh = ctx.project_sh9(level0, FORMAT_F32, w, w)
    others = np.delete(np.arange(9), 1)
    assert abs(sh[1, 0]) > 1.0 and np.abs(sh[others]).max() <= 1e-3 * abs(sh[1, 0])


def test_sh9_row_partials_add_up_and_are_deterministic(ctx):
    """The multi-GPU split: partial sums over row slabs add to the full projection."""
    w = 128
    cube = synth.synthetic_cube(w, w, probe=22)
    d = torch.from_numpy(cube).to(DEV)
    out = torch.zeros(4, 28, dtype=torch.float64, device=DEV)
    ctx.sh9_partial_device(d, FORMAT_F32, w, w, 0, 6 * w, out[0])
    ctx.sh9_partial_device(d, FORMAT_F32, w, w, 0, 301, out[1])
    ctx.sh9_partial_device(d, FORMAT_F32, w, w, 301, 6 * w, out[2])
    ctx.sh9_partial_device(d, FORMAT_F32, w, w, 0, 6 * w, out[3])
    ctx.synchronize()
    p = out.cpu().numpy()
    assert np.array_equal(p[0], p[3])                                 # fixed-order reduction
    assert np.allclose(p[1] + p[2], p[0], rtol=1e-6, atol=1e-6 * np.abs(p[0]).max())
    want = oracle_lib.sh9_partial(cube, FORMAT_F32, w, w, 0, 301)
    assert np.abs(p[1] - want).max() <= TOL_SH * np.abs(want).max()
    assert np.isclose(p[0][27], 4 * np.pi, rtol=1e-6)
    assert sh_error(ctx.sh9_finish(p[1] + p[2]), oracle_lib.project_sh9(cube, FORMAT_F32, w, w)) <= TOL_SH


def test_sh9_large_faces_keep_their_precision(ctx):
    """2048^2 faces: the case where the reference shader's fp32 four-atan weight loses all its digits."""
    w = 2048
    cube = torch.ones((6, w, w, 4), dtype=torch.float32, device=DEV)
    cube[..., 0] = 0.25
    cube[..., 2] = 4.0
    out = torch.zeros(28, dtype=torch.float64, device=DEV)
    ctx.sh9_partial_device(cube, FORMAT_F32, w, w, 0, 6 * w, out)
    ctx.synchronize()
    sh = ctx.sh9_finish(out.cpu().numpy())
    assert np.isclose(out.cpu().numpy()[27], 4 * np.pi, rtol=1e-6)
    assert np.allclose(sh[0], 0.282095 * 4 * np.pi * np.array([0.25, 1.0, 4.0]), rtol=1e-5)
    assert np.abs(sh[1:]).max() <= 1e-4 * sh[0].max()


def oracle_sh9_partial_threads(cube, w, row_begin, row_end, threads=8):
    """fp64 oracle partial sums of rows [row_begin, row_end), row chunks on host threads."""
    import threading
    chunks = np.linspace(row_begin, row_end, threads + 1).astype(int)
    parts = [None] * threads

    def work(i):
        parts[i] = oracle_lib.sh9_partial(cube, FORMAT_F32, w, w, int(chunks[i]), int(chunks[i + 1]))

    pool = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in pool:
        t.start()
    for t in pool:
        t.join()
    return np.sum(parts, axis=0)


def test_sh9_noise_cube_at_2048_and_its_eight_row_slabs_against_the_oracle(ctx):
    """BASELINE config 5's shape (big RGBA32F faces, rows split over 8 GPUs) with per-texel noise over 6 stops
    — the mirror-indexed solid-angle table and the row-segment addressing at a size where they matter —
    against the fp64 oracle: the whole cube, and each of the eight slabs a GPU would own."""
    w = 2048
    rng = np.random.default_rng(55)
    cube = np.ones((6, w, w, 4), np.float32)
    cube[..., :3] = rng.random((6, w, w, 3), dtype=np.float32) * np.exp2(6 * rng.random((6, w, w, 1), dtype=np.float32) - 3)
    cube[2, 100:140, 300:360, :3] += 3000.0                                # a bright patch off-centre on one face
    d = torch.from_numpy(cube).to(DEV)
    out = torch.zeros(9, 28, dtype=torch.float64, device=DEV)
    ctx.sh9_partial_device(d, FORMAT_F32, w, w, 0, 6 * w, out[8])
    rows = 6 * w // 8
    for r in range(8):
        ctx.sh9_partial_device(d, FORMAT_F32, w, w, r * rows, (r + 1) * rows, out[r])
    ctx.synchronize()
    p = out.cpu().numpy()
    want_full = oracle_sh9_partial_threads(cube, w, 0, 6 * w)
    scale = np.abs(want_full[:27]).max()
    assert np.abs(p[8] - want_full).max() <= TOL_SH * scale
    for r in (0, 3, 7):
        want = oracle_sh9_partial_threads(cube, w, r * rows, (r + 1) * rows)
        assert np.abs(p[r] - want).max() <= TOL_SH * np.abs(want[:27]).max(), r
    assert np.abs(p[:8].sum(axis=0) - want_full).max() <= TOL_SH * scale
    assert sh_error(ctx.sh9_finish(p[:8].sum(axis=0)), oracle_lib.sh9_finish(want_full)) <= TOL_SH


def test_envbrdf_lut_at_the_shipped_size(ctx):
    """tools/assetbuilder.cpp:496-503 bakes the env-BRDF LUT at 256 x 256 (ibl.cpp:292-308): that size, straight
    through the C ABI, against the oracle's words and pre-quantisation values."""
    want_words, want_f32 = oracle_lib.pack_envbrdf(256, 256, 1024)
    got = np.zeros(256 * 256, np.uint32)
    ctx.image_pack_envbrdf(256, 256, got)
    dec = oracle_lib.rgbe_decode_array(got)[:, :3]
    assert oracle_lib.relative_error(dec, want_f32).max() <= 4e-3          # one 9-bit mantissa code
    stats = oracle_lib.word_stats(got, want_words)
    assert oracle_lib.words_within_one_code(stats, 0.98), stats


def test_irradiance_cube_matches_the_oracle(ctx):
    w = 32
    cube = synth.synthetic_cube(64, 64, probe=23, sun=False)
    sh = ctx.project_sh9(cube, FORMAT_F32, 64, 64)
    words, f32 = ctx.sh9_irradiance_cube(sh, w, w)
    normals = synth.cube_directions(w, w).reshape(-1, 3)
    want = oracle_lib.sh9_irradiance(sh.astype(np.float64), normals)
    assert oracle_lib.relative_error(f32, want).max() <= 1e-4
    stats = oracle_lib.word_stats(words, oracle_lib.rgbe_encode_array(want))
    assert oracle_lib.words_within_one_code(stats, 0.99), stats
    const = np.zeros((9, 3), np.float32)
    const[0] = 0.282095 * 4 * np.pi * np.array([0.2, 0.4, 0.8])
    _, e = ctx.sh9_irradiance_cube(const, 8, 8, want_words=False)
    assert np.allclose(e, np.pi * np.array([0.2, 0.4, 0.8]), rtol=1e-5)   # irradiance of constant radiance c is pi*c


def test_envbrdf_lut_matches_oracle_and_reference_golden(ctx):
    want_words, want_f32 = oracle_lib.pack_envbrdf(32, 32, 1024)
    got = np.zeros(32 * 32, np.uint32)
    ctx.image_pack_envbrdf(32, 32, got)
    dec = oracle_lib.rgbe_decode_array(got)[:, :3]
    assert oracle_lib.relative_error(dec, want_f32).max() <= 4e-3          # one 9-bit mantissa code
    stats = oracle_lib.word_stats(got, want_words)
    assert oracle_lib.words_within_one_code(stats, 0.98), stats
    got16 = np.zeros(16 * 16, np.uint32)
    datum_b200.image_pack_envbrdf(16, 16, got16)
    stats = oracle_lib.word_stats(got16, GOLDEN["envbrdf16"])
    assert oracle_lib.words_within_one_code(stats), stats


def test_watercolor_lut_matches_reference_golden(ctx):
    p = GOLDEN["water_params"]
    got = np.zeros(16 * 16, np.uint32)
    ctx.image_pack_watercolor(p[0:3], p[3:6], float(p[6]), p[7:10], float(p[10]), float(p[11]), 16, 16, got)
    stats = oracle_lib.word_stats(got, GOLDEN["water16"])
    assert oracle_lib.words_within_one_code(stats, 0.98), stats
    want = oracle_lib.pack_watercolor([0.1, 0.2, 0.3], [0.3, 0.5, 0.9], 0.5, [0.02, 0.03, 0.04], 0.2, 3.0, 64, 32)
    got = np.zeros(64 * 32, np.uint32)
    ctx.image_pack_watercolor([0.1, 0.2, 0.3], [0.3, 0.5, 0.9], 0.5, [0.02, 0.03, 0.04], 0.2, 3.0, 64, 32, got)
    stats = oracle_lib.word_stats(got, want)
    assert oracle_lib.words_within_one_code(stats, 0.98), stats
