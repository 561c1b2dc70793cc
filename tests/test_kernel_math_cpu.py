"""Host compilation of the CUDA kernels' math (datum_b200/csrc/ibl_math.cuh,
ibl_tables.cpp) against the oracle.  This catches algorithmic mistakes — the
table-driven reflected direction, the magic-add floor, quad-record addressing,
the biased-mantissa accumulation, the same-face fast path — on the CPU-only leg.
It is a check OF the kernel source, not a fallback for it."""

import os

import numpy as np
import pytest

import emu_lib
import oracle_lib
from datum_b200 import synth

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ibl_golden.npz"))


def test_device_codec_is_bit_exact_with_reference_words():
    emu = emu_lib.load()
    got = np.array([emu.emu_rgbe_encode(float(r), float(g), float(b)) for r, g, b in GOLDEN["codec_rgb"]], np.uint32)
    assert np.array_equal(got, GOLDEN["codec_words"])


def test_device_decode_is_bit_exact_with_reference():
    emu = emu_lib.load()
    out = np.zeros((len(GOLDEN["decode_words"]), 3), np.float32)
    for i, w in enumerate(GOLDEN["decode_words"]):
        emu.emu_rgbe_decode(int(w), out[i].ctypes.data)
    assert np.array_equal(out.view(np.uint32), GOLDEN["decode_rgba"][:, :3].view(np.uint32))


def test_texel_normals_are_exact_with_the_oracle():
    emu, orc = emu_lib.load(), oracle_lib.oracle()
    for f in range(6):
        for wd, hd in ((1, 1), (2, 2), (4, 4), (16, 8), (256, 256), (1024, 1024)):
            for x, y in {(0, 0), (wd - 1, hd - 1), (wd // 2, hd // 3), (wd // 3, hd // 2)}:
                a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
                orc.oracle_texel_direction(f, x, y, wd, hd, a.ctypes.data)
                emu.emu_texel_normal(f, x, y, wd, hd, b.ctypes.data)
                assert np.array_equal(a, b)   # value-exact; the sign of an exact zero may differ


@pytest.mark.parametrize("level,levels,samples", [(1, 8, 1024), (3, 8, 1024), (7, 8, 1024), (1, 12, 4096), (11, 12, 4096), (2, 5, 16)])
def test_sample_table_matches_reference_samples(level, levels, samples):
    """Table entries == the reference's per-sample (L in the tangent frame, NdotL), sorted by angle."""
    emu, orc = emu_lib.load(), oracle_lib.oracle()
    entries = np.zeros((samples, 4), np.float32)
    total = np.zeros(1, np.float64)
    n = emu.emu_table(level, levels, samples, entries.ctypes.data, total.ctypes.data)
    entries = entries[:n]

    # reference samples around N = +z... use a generic normal and project on its frame
    N = np.array([0.0, 0.0, -1.0], np.float32)   # front face centre: T = (0,-1,0)... any N works, frame from ibl.cpp:123-125
    dirs = np.zeros((samples, 3), np.float32)
    ndotl = np.zeros(samples, np.float32)
    roughness = np.float32(level) / np.float32(levels - 1)
    orc.oracle_trace_samples(float(roughness), samples, N.ctypes.data, dirs.ctypes.data, ndotl.ctypes.data)

    assert n == int((ndotl > 0).sum())
    assert np.isclose(total[0], ndotl[ndotl > 0].astype(np.float64).sum(), rtol=1e-6)
    assert np.all(np.diff(entries[:, 2]) <= 0)                      # decreasing NdotL
    assert np.allclose(entries[:, 3], 0.5 * entries[:, 2])
    # |L| == 1 and L.N == lz: compare the sorted weights with the reference's
    assert np.allclose(np.sort(ndotl[ndotl > 0])[::-1], entries[:, 2], atol=2e-6)
    assert np.allclose(np.linalg.norm(entries[:, :3], axis=1), 1.0, atol=2e-6)


@pytest.mark.parametrize("ws,levels,level,noise", [(16, 5, 1, True), (16, 5, 4, True), (32, 6, 1, False), (32, 6, 2, True), (32, 6, 5, False), (8, 4, 3, True)])
def test_kernel_math_matches_oracle(ws, levels, level, noise):
    emu = emu_lib.load()
    src = synth.synthetic_chain(ws, ws, 1, probe=6, noise=noise, sun=False)
    wd = ws // 2
    n = 6 * wd * wd
    want_words, want_f32 = oracle_lib.prefilter_level(src, ws, ws, level, levels, 1024)
    got_words, got_f32 = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    emu.emu_prefilter_level(src.ctypes.data, ws, ws, level, levels, 1024, got_words.ctypes.data, got_f32.ctypes.data)

    clean = oracle_lib.edge_ambiguous_counts(wd, wd, level, levels, 1024) == 0
    rel = oracle_lib.relative_error(got_f32, want_f32)
    assert rel[clean].max() <= 1e-4          # tolerance of the contract is 1e-3
    assert rel.max() <= 5e-2                 # texels with samples exactly on a cube edge: reference itself is ill-defined there
    stats = oracle_lib.word_stats(got_words[clean], want_words[clean])
    assert stats["max_code"] <= 1 and stats["exp_mismatch"] == 0 and stats["identical"] >= 0.99


@pytest.mark.parametrize("ws,levels,level,noise", [(16, 5, 1, True), (16, 5, 4, True), (32, 6, 1, False), (32, 6, 2, True), (32, 6, 5, False), (8, 4, 3, True)])
def test_dn_kernel_math_matches_oracle(ws, levels, level, noise):
    """prefilter_dn.cu's arithmetic: banded table, subnormal-mantissa taps, exponent in the weight."""
    emu = emu_lib.load()
    src = synth.synthetic_chain(ws, ws, 1, probe=6, noise=noise, sun=False)
    wd = ws // 2
    n = 6 * wd * wd
    want_words, want_f32 = oracle_lib.prefilter_level(src, ws, ws, level, levels, 1024)
    got_words, got_f32 = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    emu.emu_prefilter_level_dn(src.ctypes.data, ws, ws, level, levels, 1024, 16, got_words.ctypes.data, got_f32.ctypes.data)

    clean = oracle_lib.edge_ambiguous_counts(wd, wd, level, levels, 1024) == 0
    rel = oracle_lib.relative_error(got_f32, want_f32)
    assert rel[clean].max() <= 1e-4
    assert rel.max() <= 5e-2
    stats = oracle_lib.word_stats(got_words[clean], want_words[clean])
    assert stats["max_code"] <= 1 and stats["exp_mismatch"] == 0 and stats["identical"] >= 0.99


@pytest.mark.parametrize("ws,levels,level,samples,noise", [(16, 5, 1, 1024, True), (32, 6, 2, 1024, True), (8, 4, 3, 1024, True), (16, 5, 2, 7, True), (24, 4, 1, 100, False),
                                                           (64, 8, 1, 1024, True), (32, 12, 1, 4096, True)])
def test_pair_kernel_math_matches_the_one_sample_kernel_and_the_oracle(ws, levels, level, samples, noise):
    """prefilter_dp_kernel's arithmetic (ibl_math.cuh "projective form"): the table of (lx/lz, ly/lz) with its
    filled-up last band read two entries at a time, folded frame rows, the record index out of the fp32
    adder with the pointer moved back by its bias, right-hand weights by difference, blue summed per sample
    of a pair.  Same samples, same footprints as the one-sample kernel up to fp32 rounding of the
    coordinates."""
    emu = emu_lib.load()
    src = synth.synthetic_chain(ws, ws, 1, probe=8, noise=noise, sun=False)
    wd = ws // 2
    n = 6 * wd * wd
    a_w, a_f = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    b_w, b_f = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    emu.emu_prefilter_level_dn(src.ctypes.data, ws, ws, level, levels, samples, 16, a_w.ctypes.data, a_f.ctypes.data)
    emu.emu_prefilter_level_dp(src.ctypes.data, ws, ws, level, levels, samples, 16, b_w.ctypes.data, b_f.ctypes.data)
    # texels with a sample within rounding of a cube edge may send it to either face (tests/parity.py)
    clean = oracle_lib.edge_ambiguous_counts(wd, wd, level, levels, samples) == 0
    peak = np.maximum(a_f.max(axis=1, keepdims=True), 1e-30)
    assert (np.abs(a_f - b_f) / peak)[clean].max() <= 2e-5
    assert (np.abs(a_f - b_f) / peak).max() <= 5e-2

    want_words, want_f32 = oracle_lib.prefilter_level(src, ws, ws, level, levels, samples)
    assert oracle_lib.relative_error(b_f, want_f32)[clean].max() <= 1e-4
    stats = oracle_lib.word_stats(b_w[clean], want_words[clean])
    assert stats["max_code"] <= 1
    assert stats["exp_mismatch"] == 0 or stats["max_value_rel"] <= 4e-3      # a value that straddles a power of two: one code apart in value (tests/parity.py)


@pytest.mark.parametrize("level,levels,samples", [(1, 8, 1024), (7, 8, 1024), (2, 5, 16), (3, 4, 7), (1, 12, 4096)])
def test_paired_table_is_the_banded_table_filled_up_and_interleaved(level, levels, samples):
    emu = emu_lib.load()
    band = 16
    entries = np.zeros((samples, 4), np.float32)
    band_min = np.zeros(samples, np.float32)
    bands = np.zeros(1, np.int32)
    m = emu.emu_banded_table(level, levels, samples, band, entries.ctypes.data, band_min.ctypes.data, bands.ctypes.data)
    out = np.zeros(4 * (samples + band), np.float32)
    scale = np.float32(2.0 ** 64)
    padded = emu.emu_paired_table(level, levels, samples, band, scale, out.ctypes.data, out.size)
    assert padded == bands[0] * band and padded % 2 == 0
    q = out[: 4 * padded].reshape(padded // 2, 2, 2, 2)          # pair, half (xy | zw), component, sample
    flat = np.zeros((padded, 4), np.float32)
    flat[0::2] = np.stack([q[:, 0, 0, 0], q[:, 0, 1, 0], q[:, 1, 0, 0], q[:, 1, 1, 0]], axis=1)
    flat[1::2] = np.stack([q[:, 0, 0, 1], q[:, 0, 1, 1], q[:, 1, 0, 1], q[:, 1, 1, 1]], axis=1)
    assert np.array_equal(flat[:m], entries[:m] * scale)
    fill = flat[m:]
    assert np.all(fill[:, :2] == 0) and np.all(fill[:, 2] == scale * np.float32(2.0 ** -60)) and np.all(fill[:, 3] == 0.5 * fill[:, 2])


@pytest.mark.parametrize("ws,level,levels,samples,sectors", [
    (128, 1, 8, 1024, 4), (128, 2, 8, 1024, 4), (64, 3, 8, 1024, 8), (32, 6, 8, 1024, 8), (64, 1, 12, 4096, 4), (48, 2, 5, 100, 8), (16, 7, 8, 1024, 4)])
def test_sector_limits_never_admit_a_sample_that_leaves_the_face(ws, level, levels, samples, sectors):
    """The pair kernel's same-face decision (ibl_math.cuh sector_rho_limits + ibl_tables.h build_sector_entries):
    every sample a texel's sector limit sends down the path WITHOUT cube-face selection must lie strictly
    inside the texel's own face (checked in double precision for every texel, sector and sample); and the
    per-warp rule sends fewer warp-samples through the face selection than one isotropic count per tile."""
    emu = emu_lib.load()
    out = np.zeros(8)
    emu.emu_sector_study(ws, level, levels, samples, sectors, out.ctypes.data)
    general_per_tile, general_per_sector, violations, growth = out[:4]
    assert violations == 0
    assert general_per_sector <= general_per_tile
    # sectors hold (almost) equal shares of the accepted samples; a short table pays for whole shares of four
    assert 1.0 <= growth <= (1.2 if samples >= 1024 else 1.5)


@pytest.mark.parametrize("ws,hs", [(2, 2), (4, 4), (8, 8), (24, 12), (64, 64), (512, 512), (1370, 1370), (2048, 2048)])
def test_projective_footprint_addresses_the_plain_footprint(ws, hs):
    """cube_footprint_proj (index out of the fp32 adder, shrink folded into the scale, fraction as fma(q, hw, hwm - i))
    against cube_footprint on random directions and on the exact ties of the face selection: same face, inside
    the face, same texel coordinate to fp32 rounding — up to the largest size proj_usable admits (2048^2)."""
    emu = emu_lib.load()
    out = np.zeros(4)
    emu.emu_footprint_compare(ws, hs, 200000, 7, out.ctypes.data)
    worst, bad, count = out[:3]
    assert count > 190000 and bad == 0
    assert worst <= 4e-7 * max(ws, hs) + 1e-6          # a few ulps of a coordinate of that size


@pytest.mark.parametrize("level,levels,samples,sectors", [(1, 8, 1024, 4), (1, 8, 1024, 8), (7, 8, 1024, 8), (2, 5, 100, 4), (1, 12, 4096, 4), (3, 4, 7, 8)])
def test_sector_table_holds_every_accepted_sample_once(level, levels, samples, sectors):
    """build_sector_entries: the accepted samples of the level, each exactly once as (lx/lz, ly/lz, lz, wh) in its
    azimuth sector's place of a band, filled up with no-effect entries; rho_max covers every share and grows."""
    emu = emu_lib.load()
    band = 4 * sectors
    entries = np.zeros((samples, 4), np.float32)
    band_min = np.zeros(samples, np.float32)
    bands_plain = np.zeros(1, np.int32)
    m = emu.emu_banded_table(level, levels, samples, 16, entries.ctypes.data, band_min.ctypes.data, bands_plain.ctypes.data)
    out = np.zeros(4 * (2 * samples + 8 * band), np.float32)
    rho = np.zeros(sectors * (samples + 8), np.float32)
    scale = np.float32(2.0 ** 64)
    bands = emu.emu_sector_table(level, levels, samples, sectors, scale, out.ctypes.data, out.size, rho.ctypes.data)
    assert bands > 0
    q = out[: 4 * bands * band].reshape(bands * band // 2, 2, 2, 2)
    flat = np.zeros((bands * band, 4), np.float32)
    flat[0::2] = np.stack([q[:, 0, 0, 0], q[:, 0, 1, 0], q[:, 1, 0, 0], q[:, 1, 1, 0]], axis=1)
    flat[1::2] = np.stack([q[:, 0, 0, 1], q[:, 0, 1, 1], q[:, 1, 0, 1], q[:, 1, 1, 1]], axis=1)
    real = flat[flat[:, 2] > scale * np.float32(2.0 ** -30)]
    fill = flat[flat[:, 2] <= scale * np.float32(2.0 ** -30)]
    assert len(real) == m and np.all(fill[:, :2] == 0) and np.all(fill[:, 2] == scale * np.float32(2.0 ** -60))
    want = np.stack([(entries[:m, 0].astype(np.float64) / entries[:m, 2]).astype(np.float32), (entries[:m, 1].astype(np.float64) / entries[:m, 2]).astype(np.float32),
                     entries[:m, 2] * scale, entries[:m, 3] * scale], axis=1)
    key = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(real), key(want))
    # every entry sits in its sector's share, and rho_max bounds it
    per = band // sectors
    for i, e in enumerate(flat):
        if e[2] <= scale * np.float32(2.0 ** -30):
            continue
        k, w = i // band, (i % band) // per
        phi = np.arctan2(np.float64(e[1]), np.float64(e[0]))
        assert int(min(max(np.floor((phi + np.pi) / (2 * np.pi / sectors)), 0), sectors - 1)) == w
        assert np.hypot(np.float64(e[0]), np.float64(e[1])) <= rho[w * bands + k]
    r = rho[: sectors * bands].reshape(sectors, bands)
    assert np.all(np.diff(r, axis=1) >= 0)


def test_dn_tap_decodes_every_field_exactly():
    """One tap of weight w through the subnormal-mantissa path == w * rgbe decode, for the
    extreme exponents, mantissas and weights (products must stay in the normal range)."""
    emu = emu_lib.load()
    rng = np.random.default_rng(5)
    words = [0x00000000, 0xFFFFFFFF, 0x00000001, 0xF8000001, 0x0003FE00, 0x07FC0000, 0x08000200, 31 << 27 | 511 << 18 | 511 << 9 | 511]
    words += [int(w) for w in rng.integers(0, 2**32, 200, dtype=np.uint64)]
    for word in words:
        for w in (1.0, 0.5, 3.0e-5, 0.73, 1e-9):
            got = np.zeros(3, np.float32)
            want = np.zeros(3, np.float32)
            emu.emu_dn_tap(word, w, got.ctypes.data)
            emu.emu_rgbe_decode(word, want.ctypes.data)
            assert np.allclose(got, want * np.float32(w), rtol=3e-7, atol=0.0), (hex(word), w, got, want)
            raw = np.zeros(3, np.float32)
            emu.emu_raw_tap(word, w, raw.ctypes.data)          # the tail kernel's form, straight from the rgbe word
            assert np.array_equal(raw, got), (hex(word), w, raw, got)
    m = emu.emu_pack_dn_word(5 << 27 | 3 << 18 | 2 << 9 | 1)
    assert m == (1 << 23 | 2 << 14 | 3 << 5 | 5)


@pytest.mark.parametrize("level,levels,samples,band", [(1, 8, 1024, 16), (5, 8, 1024, 16), (1, 12, 4096, 16), (2, 5, 16, 16), (3, 4, 7, 16)])
def test_banded_table_is_the_sorted_table_cut_in_rings(level, levels, samples, band):
    emu = emu_lib.load()
    flat = np.zeros((samples, 4), np.float32)
    total = np.zeros(1, np.float64)
    n = emu.emu_table(level, levels, samples, flat.ctypes.data, total.ctypes.data)
    flat = flat[:n]
    entries = np.zeros((samples, 4), np.float32)
    band_min = np.zeros(samples, np.float32)
    bands = np.zeros(1, np.int32)
    m = emu.emu_banded_table(level, levels, samples, band, entries.ctypes.data, band_min.ctypes.data, bands.ctypes.data)
    entries, band_min = entries[:m], band_min[: bands[0]]
    assert m == n and bands[0] == (n + band - 1) // band
    for k in range(bands[0]):
        a = flat[k * band:(k + 1) * band]
        b = entries[k * band:(k + 1) * band]
        assert sorted(map(tuple, a)) == sorted(map(tuple, b))                   # same entries per band
        assert band_min[k] == a[:, 2].min()
        assert np.all(np.diff(np.arctan2(b[:, 1].astype(np.float64), b[:, 0].astype(np.float64))) >= 0)   # ring order
    assert np.all(np.diff(band_min) <= 0)


def test_fast_path_is_taken_and_agrees_with_general_path(monkeypatch):
    emu = emu_lib.load()
    ws, levels, level = 64, 8, 1
    src = synth.synthetic_chain(ws, ws, 1, probe=7, sun=False)
    n = 6 * (ws // 2) ** 2
    a_w, a_f = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    b_w, b_f = np.zeros(n, np.uint32), np.zeros((n, 3), np.float32)
    emu.emu_prefilter_level(src.ctypes.data, ws, ws, level, levels, 256, a_w.ctypes.data, a_f.ctypes.data)
    assert emu.emu_last_fast_fraction() > 0.5
    monkeypatch.setenv("EMU_NOFAST", "1")
    emu.emu_prefilter_level(src.ctypes.data, ws, ws, level, levels, 256, b_w.ctypes.data, b_f.ctypes.data)
    assert emu.emu_last_fast_fraction() == 0.0
    assert oracle_lib.relative_error(a_f, b_f).max() < 2e-5


@pytest.mark.parametrize("level,levels,samples", [(1, 8, 1024), (3, 8, 1024), (6, 8, 1024), (1, 12, 4096), (2, 5, 40)])
def test_patch_table_is_a_reordering_of_the_ring_table_with_monotone_band_floors(level, levels, samples):
    """ibl_tables.h order 1 (bands are compact patches of the lobe instead of rings): the same entries, the
    short band still last, per-band smallest lz non-increasing (the kernel's same-face search relies on it)
    and a true lower bound of every entry of the band and of all bands before it."""
    emu = emu_lib.load()
    band = 16
    ring = np.zeros((samples, 4), np.float32)
    patch = np.zeros((samples, 4), np.float32)
    floor_ring, floor_patch = np.zeros(samples, np.float32), np.zeros(samples, np.float32)
    nb0, nb1 = np.zeros(1, np.int32), np.zeros(1, np.int32)
    n0 = emu.emu_banded_table(level, levels, samples, band, ring.ctypes.data, floor_ring.ctypes.data, nb0.ctypes.data)
    n1 = emu.emu_patch_table(level, levels, samples, band, patch.ctypes.data, floor_patch.ctypes.data, nb1.ctypes.data)
    assert n0 == n1 and nb0[0] == nb1[0]
    key = [("a", "f4"), ("b", "f4"), ("c", "f4"), ("d", "f4")]
    assert np.array_equal(np.sort(ring[:n0].copy().view(key).ravel()), np.sort(patch[:n1].copy().view(key).ravel()))
    floors = floor_patch[: nb1[0]]
    assert np.all(np.diff(floors) <= 0)
    for k in range(nb1[0]):
        assert patch[k * band:min(n1, (k + 1) * band), 2].min() >= floors[k]
