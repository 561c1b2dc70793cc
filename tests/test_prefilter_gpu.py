"""GPU parity of the GGX prefilter chain (tools/ibl.cpp:242-279) through the C ABI."""

import os

import numpy as np
import pytest
import torch

import datum_b200
import oracle_lib
import parity
from datum_b200 import synth

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ibl_golden.npz"))
DEV = "cuda:0"


def run_chain_device(ctx, bits, w, h, levels, samples):
    """Chain on the device; returns (words, f32 of levels >= 1)."""
    offs = datum_b200.level_offsets(w, h, levels)
    d_bits = torch.from_numpy(bits.view(np.int32).copy()).to(DEV)
    d_f32 = torch.zeros(max(1, (offs[-1] - offs[1]) * 3), dtype=torch.float32, device=DEV)
    ctx.buildmips_cube_ibl_device(w, h, levels, d_bits, samples, d_f32)
    ctx.synchronize()
    return d_bits.cpu().numpy().view(np.uint32), d_f32.cpu().numpy().reshape(-1, 3)


@pytest.mark.parametrize("w,h,levels,samples,noise", [
    (32, 32, 6, 1024, True),       # all levels down to 1x1 faces
    (64, 64, 7, 1024, False),
    (128, 128, 8, 1024, True),
    (64, 64, 4, 4096, True),       # BASELINE config 3's sample count
    (16, 16, 5, 16, True),         # tiny sample count
    (24, 12, 3, 1024, True),       # non-square, not a power of two
    (2, 2, 2, 1024, True),         # smallest legal chain: 2x2 -> 1x1
])
def test_every_level_matches_the_oracle_on_the_same_source(ctx, w, h, levels, samples, noise):
    bits = synth.synthetic_chain(w, h, levels, probe=11, noise=noise, sun=False)
    got, got_f32 = run_chain_device(ctx, bits, w, h, levels, samples)
    offs = datum_b200.level_offsets(w, h, levels)
    assert np.array_equal(got[: offs[1]], bits[: offs[1]])          # level 0 untouched
    for level in range(1, levels):
        ws, hs = w >> (level - 1), h >> (level - 1)
        parity.check_level(got[offs[level]:offs[level + 1]], got_f32[offs[level] - offs[1]:offs[level + 1] - offs[1]],
                           got[offs[level - 1]:offs[level]], ws, hs, level, levels, samples)


def spot_ranges(hd, rows=4):
    """Row ranges of a 6*hd-row level: top edge of face 0, a face seam, the middle of face 2, the bottom edge of face 5."""
    return [(0, rows), (hd - rows // 2, hd + rows // 2), (2 * hd + hd // 2, 2 * hd + hd // 2 + rows), (6 * hd - rows, 6 * hd)]


def test_baseline_config_2_against_the_oracle_on_spot_rows(ctx):
    """BASELINE config 2 at full size (512^2 faces, 8 levels, 1024 spp): the launch shapes the
    benchmark times (per-SM tile queues, two samples at a time) against the oracle on rows at
    face edges, across a face seam and in a face centre of the three big levels; the small
    levels completely."""
    w, levels, samples = 512, 8, 1024
    bits = synth.synthetic_chain(w, w, levels, probe=0, noise=True, sun=False)
    got, got_f32 = run_chain_device(ctx, bits, w, w, levels, samples)
    offs = datum_b200.level_offsets(w, w, levels)
    for level in range(1, levels):
        ws = w >> (level - 1)
        hd = ws >> 1
        ranges = spot_ranges(hd) if level <= 3 else [(0, 6 * hd)]
        for a, b in ranges:
            parity.check_level(got[offs[level]:offs[level + 1]], got_f32[offs[level] - offs[1]:offs[level + 1] - offs[1]],
                               got[offs[level - 1]:offs[level]], ws, ws, level, levels, samples, a, b)


def test_baseline_config_3_level_1_against_the_oracle_on_spot_rows(ctx):
    """BASELINE config 3's dominant launch at full size: 2048^2 -> 1024^2 faces, 12 levels,
    4096 spp (the 4096-entry table is read through L1 instead of shared memory).  On 2048-wide
    faces the fp32 rounding of a footprint coordinate is 1e-4 of a texel; under the 6-stop
    per-texel noise of the synthetic input that moves values by 3e-5 (tolerance 1e-3) and
    therefore flips the 9-bit rounding of about 1 % of the words by one code: the identical-word
    floor is 98 % here instead of 99 % (DESIGN.md "Conditioning")."""
    ws, levels, samples = 2048, 12, 4096
    src = synth.synthetic_chain(ws, ws, 1, probe=3, noise=True, sun=False)
    d_src = torch.from_numpy(src.view(np.int32)).to(DEV)
    wd = ws // 2
    words = torch.zeros(6 * wd * wd, dtype=torch.int32, device=DEV)
    f32 = torch.zeros(6 * wd * wd * 3, dtype=torch.float32, device=DEV)
    ctx.prefilter_level_device(d_src, ws, ws, 1, levels, samples, 0, 6 * wd, words, f32)
    ctx.synchronize()
    got, got_f32 = words.cpu().numpy().view(np.uint32), f32.cpu().numpy().reshape(-1, 3)
    for a, b in spot_ranges(wd, rows=2):
        parity.check_level(got, got_f32, src, ws, ws, 1, levels, samples, a, b, min_identical=0.98)


def test_baseline_config_3_levels_2_to_11_against_the_oracle(ctx):
    """BASELINE config 3 below its first level: the whole 2048^2 x 12-level x 4096-spp chain is baked on the
    GPU, then every level is compared with the oracle run on the SAME source level — levels 2-4 (512^2,
    256^2 and 128^2 faces: the pair kernel reading its 4096-entry table through L1) on rows at a face edge,
    across a face seam, in a face centre and at the last edge; levels 5-11 (64^2 ... 1x1 faces, the tail
    kernel at 4096 spp) completely."""
    w, levels, samples = 2048, 12, 4096
    offs = datum_b200.level_offsets(w, w, levels)
    bits = synth.synthetic_chain(w, w, levels, probe=3, noise=True, sun=False)
    d_bits = torch.from_numpy(bits.view(np.int32)).to(DEV)
    d_f32 = torch.zeros((offs[-1] - offs[2]) * 3, dtype=torch.float32, device=DEV)      # levels >= 2 only (level 1 alone would be 75 MB)
    # level 1 through the level entry point, the rest through the chain on the sub-pyramid that starts at level 1:
    # roughness depends on (level, levels) only, so the sub-chain is run level by level
    for level in range(1, levels):
        ws = w >> (level - 1)
        f32 = d_f32[(offs[level] - offs[2]) * 3:(offs[level + 1] - offs[2]) * 3] if level >= 2 else None
        ctx.prefilter_level_device(d_bits[offs[level - 1]:offs[level]], ws, ws, level, levels, samples, 0, 6 * (ws >> 1), d_bits[offs[level]:offs[level + 1]], f32)
    ctx.synchronize()
    got = d_bits.cpu().numpy().view(np.uint32)
    got_f32 = d_f32.cpu().numpy().reshape(-1, 3)
    for level in range(2, levels):
        ws = w >> (level - 1)
        hd = ws >> 1
        ranges = spot_ranges(hd, rows=2) if level <= 4 else [(0, 6 * hd)]
        for a, b in ranges:
            parity.check_level(got[offs[level]:offs[level + 1]], got_f32[offs[level] - offs[2]:offs[level + 1] - offs[2]],
                               got[offs[level - 1]:offs[level]], ws, ws, level, levels, samples, a, b, min_identical=0.98 if level <= 3 else 0.99)


def test_a_config_4_probe_through_the_batch_entry_against_the_oracle(ctx):
    """BASELINE config 4's unit (256^2 faces, 8 levels, 1024 spp + SH9) through datum_ibl_bake_probes — the
    probe-batched launches — against the ORACLE (not only against single calls): prefilter rows of the two big
    levels, the small levels completely, and the SH9 coefficients."""
    w, levels, samples = 256, 8, 1024
    offs = datum_b200.level_offsets(w, w, levels)
    payloads = [synth.synthetic_chain(w, w, levels, probe=p, sun=False) for p in (70, 71, 72)]
    sources = [b.copy() for b in payloads]
    sh = ctx.bake_probes(w, w, levels, payloads, samples, sh9=True)
    for k in (0, 2):
        got = payloads[k]
        assert np.array_equal(got[: offs[1]], sources[k][: offs[1]])
        for level in range(1, levels):
            ws = w >> (level - 1)
            hd = ws >> 1
            for a, b in (spot_ranges(hd, rows=2) if level <= 2 else [(0, 6 * hd)]):
                parity.check_level(got[offs[level]:offs[level + 1]], None, got[offs[level - 1]:offs[level]], ws, ws, level, levels, samples, a, b)
        want = oracle_lib.project_sh9(sources[k][: offs[1]], datum_b200.FORMAT_RGBE, w, w)
        assert np.abs(sh[k] - want).max() <= 1e-4 * np.abs(want).max()


def test_hdr_sun_input_stays_within_tolerance(ctx):
    """A 2e4 sun disc next to 0.1-level sky: the ill-conditioned case."""
    w, levels = 128, 8
    bits = synth.synthetic_chain(w, w, levels, probe=2, noise=True, sun=True)
    got, got_f32 = run_chain_device(ctx, bits, w, w, levels, 1024)
    offs = datum_b200.level_offsets(w, w, levels)
    for level in range(1, levels):
        ws = w >> (level - 1)
        parity.check_level(got[offs[level]:offs[level + 1]], got_f32[offs[level] - offs[1]:offs[level + 1] - offs[1]],
                           got[offs[level - 1]:offs[level]], ws, ws, level, levels, 1024)


@pytest.mark.parametrize("name,w,h,levels", [("chain16_noise", 16, 16, 5), ("chain16_smooth", 16, 16, 5), ("chain32_noise", 32, 32, 6), ("chain24x12", 24, 12, 3)])
def test_own_chain_against_reference_golden(ctx, name, w, h, levels):
    """Whole chain from the reference's level 0, compared with the words the unmodified
    reference produced (each level here is built from OUR previous level, so one-code
    differences propagate: the bound is on decoded values)."""
    want = GOLDEN[name]
    bits = np.zeros_like(want)
    bits[: 6 * w * h] = GOLDEN[name + "_level0"]
    ctx.image_buildmips_cube_ibl(w, h, levels, bits, 1024)            # host entry point, pageable numpy buffer
    offs = datum_b200.level_offsets(w, h, levels)
    assert np.array_equal(bits[: offs[1]], want[: offs[1]])
    got = oracle_lib.rgbe_decode_array(bits[offs[1]:])[:, :3].astype(np.float64)
    ref = oracle_lib.rgbe_decode_array(want[offs[1]:])[:, :3].astype(np.float64)
    rel = np.abs(got - ref).max(axis=1) / np.maximum(ref.max(axis=1), 1e-30)
    assert np.quantile(rel, 0.99) <= 4e-3      # one mantissa code of a 9-bit mantissa
    assert rel.max() <= 1e-1                   # cube-edge samples, see parity.py
    assert (bits[offs[1]:] == want[offs[1]:]).mean() >= 0.97


def test_host_and_device_entry_points_agree(ctx):
    w, levels = 64, 7
    bits = synth.synthetic_chain(w, w, levels, probe=12)
    dev_words, _ = run_chain_device(ctx, bits, w, w, levels, 1024)
    host = bits.copy()
    ctx.image_buildmips_cube_ibl(w, w, levels, host, 1024)
    assert np.array_equal(host, dev_words)
    pinned = torch.from_numpy(bits.view(np.int32).copy()).pin_memory()
    ctx.image_buildmips_cube_ibl(w, w, levels, pinned, 1024)
    assert np.array_equal(pinned.numpy().view(np.uint32), dev_words)
    datum_b200.image_buildmips_cube_ibl(w, w, levels, host, 1024)     # reference-named free function
    assert np.array_equal(host, dev_words)


def test_row_slabs_reproduce_the_full_level(ctx):
    """Multi-GPU sharding primitive: a level computed as row slabs equals one full launch.  A
    slab may pick another kernel than the full level (here: slabs of at most 6144 texels go to the
    tail kernel, whose tangent frame never passes through face-local coordinates and whose sums
    associate differently), so the comparison is the packed-word criterion plus a 1e-4 bound on the
    fp32 values — a tenth of the tolerance against the reference."""
    ws, levels, level = 64, 7, 2
    src = synth.synthetic_chain(ws, ws, 1, probe=13)
    d_src = torch.from_numpy(src.view(np.int32)).to(DEV)
    wd = ws // 2
    cuts = [0, 7, 8, 50, 97, 6 * wd]

    def run(variant):
        ctx.set_prefilter_variant(variant)
        full = torch.zeros(6 * wd * wd, dtype=torch.int32, device=DEV)
        full_f = torch.zeros(6 * wd * wd * 3, dtype=torch.float32, device=DEV)
        ctx.prefilter_level_device(d_src, ws, ws, level, levels, 1024, 0, 6 * wd, full, full_f)
        slabs, slabs_f = torch.zeros_like(full), torch.zeros_like(full_f)
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.prefilter_level_device(d_src, ws, ws, level, levels, 1024, a, b, slabs, slabs_f)
        ctx.prefilter_level_device(d_src, ws, ws, level, levels, 1024, 5, 5, slabs, slabs_f)   # empty slab: legal, writes nothing
        ctx.synchronize()
        return full, full_f, slabs, slabs_f

    try:
        full, full_f, slabs, slabs_f = run(0)
        assert oracle_lib.relative_error(slabs_f.cpu().numpy(), full_f.cpu().numpy()).max() <= 1e-4
        stats = oracle_lib.word_stats(slabs.cpu().numpy().view(np.uint32), full.cpu().numpy().view(np.uint32))
        assert oracle_lib.words_within_one_code(stats, 0.995), stats
        # pinned variants (one-sample and pair kernel): same warp split; only the same-face sample count of the
        # re-cut tiles differs (the two paths round a footprint coordinate differently: up to 4e-5 when EVERY
        # sample changes path, tests/emu with EMU_NOFAST)
        for variant in (53, 72):
            full, full_f, slabs, slabs_f = run(variant)
            assert oracle_lib.relative_error(slabs_f.cpu().numpy(), full_f.cpu().numpy()).max() <= 5e-5
            assert (full == slabs).float().mean().item() >= 0.999
    finally:
        ctx.set_prefilter_variant(0)


def test_kernel_variants_agree(ctx):
    """Every tile shape / warp split computes the same level (fp32 sums differ only by
    association order: compare through the packed-word criterion).  The pair kernel scales its
    directions by 1/lz (projective form), so a sample within rounding of a cube edge may take the
    other face there: texels the oracle marks edge-ambiguous get the looser bound of tests/parity.py."""
    w, levels = 64, 6
    bits = synth.synthetic_chain(w, w, levels, probe=14, sun=False)
    offs = datum_b200.level_offsets(w, w, levels)
    clean = np.concatenate([oracle_lib.edge_ambiguous_counts(w >> level, w >> level, level, levels, 1024) == 0 for level in range(1, levels)])
    base = None
    try:
        for variant in (0, 51, 52, 53, 54, 66, 67, 70, 71, 72, 73, 80):     # every kernel shape the product library ships
            ctx.set_prefilter_variant(variant)
            words, f32 = run_chain_device(ctx, bits, w, w, levels, 1024)
            if base is None:
                base = (words, f32)
            else:
                rel = oracle_lib.relative_error(f32, base[1])
                assert rel[clean].max() <= 2e-4, (variant, float(rel[clean].max()))   # one-code word flips of level L feed level L+1
                assert rel.max() <= 2e-2, (variant, float(rel.max()), int(rel.argmax()))
                assert (words == base[0]).mean() >= 0.995, (variant, float((words == base[0]).mean()))
    finally:
        ctx.set_prefilter_variant(0)


def test_pair_kernel_hands_over_when_its_fp32_index_does_not_apply(ctx):
    """The two-samples-at-a-time kernel forms the record index in the fp32 adder (ibl_math.cuh, projective
    form): that needs an even source size (the align-corners offset must be an integer) of at most 2^22
    texels per face.  For an odd source the launcher must take the one-sample kernel instead: same words
    as that kernel pinned, and oracle parity.  1370 (even, not a power of two: round 2's integer index
    wrapped there) stays on the pair kernel and must agree with the oracle too."""
    levels, samples = 3, 8
    rows = (0, 16)
    for ws, handed_over in ((1371, True), (1370, False)):
        # smooth radiance: with 8 samples over 6-stop per-texel noise the 1e-7 coordinate rounding of a
        # 1370-wide face alone flips 8 % of the words by one code (DESIGN.md "Conditioning")
        src = synth.synthetic_chain(ws, ws, 1, probe=23, noise=False, sun=False)
        d_src = torch.from_numpy(src.view(np.int32)).to(DEV)
        wd = ws // 2
        outs = []
        for variant in (0, 51, 70, 53, 72):
            ctx.set_prefilter_variant(variant)
            try:
                out = torch.zeros(6 * wd * wd, dtype=torch.int32, device=DEV)
                f32 = torch.zeros(6 * wd * wd * 3, dtype=torch.float32, device=DEV)
                ctx.prefilter_level_device(d_src, ws, ws, 1, levels, samples, rows[0], rows[1], out, f32)
                ctx.synchronize()
            finally:
                ctx.set_prefilter_variant(0)
            outs.append((out.cpu().numpy().view(np.uint32), f32.cpu().numpy().reshape(-1, 3)))
        assert np.array_equal(outs[2][1], outs[1][1]) == handed_over     # 70 ran as 51
        assert np.array_equal(outs[4][1], outs[3][1]) == handed_over     # 72 ran as 53
        for k in (0, 2, 4):
            parity.check_level(outs[k][0], outs[k][1], src, ws, ws, 1, levels, samples, rows[0], rows[1])


def test_directions_on_cube_edges_stay_inside_the_face(ctx):
    """Regression: a reflected direction with two equal major components (|x| == |y| exactly)
    and a reciprocal rounded up put the footprint one texel outside its face.  With 2048^2 faces
    and 4096 samples some texel hits such a tie; the launch must complete and every word must be
    a finite colour."""
    ws, levels, samples = 2048, 3, 4096
    src = synth.synthetic_chain(ws, ws, 1, probe=21, sun=False)
    d_src = torch.from_numpy(src.view(np.int32)).to(DEV)
    wd = ws // 2
    for variant in (0, 52):
        ctx.set_prefilter_variant(variant)
        try:
            out = torch.zeros(6 * wd * wd, dtype=torch.int32, device=DEV)
            ctx.prefilter_level_device(d_src, ws, ws, 1, levels, samples, 0, 6 * wd, out)
            ctx.synchronize()
        finally:
            ctx.set_prefilter_variant(0)
        words = out.cpu().numpy().view(np.uint32)
        rgb = oracle_lib.rgbe_decode_array(words)[:, :3]
        assert np.isfinite(rgb).all() and rgb.max() <= 65408.0 and rgb.max() > 0.0


def test_constant_environment_stays_constant_at_full_size(ctx):
    """Size-independent property at BASELINE config 2 (512^2, 8 levels, 1024 spp)."""
    w, levels = 512, 8
    offs = datum_b200.level_offsets(w, w, levels)
    word = oracle_lib.rgbe_encode_array(np.array([[3.0, 1.5, 0.75]], np.float32))[0]
    bits = np.zeros(offs[-1], np.uint32)
    bits[: offs[1]] = word
    got, got_f32 = run_chain_device(ctx, bits, w, w, levels, 1024)
    value = oracle_lib.rgbe_decode_array(np.array([word], np.uint32))[0, :3]
    assert np.allclose(got_f32, value[None, :], rtol=2e-5)
    assert oracle_lib.word_stats(got[offs[1]:], np.full(offs[-1] - offs[1], word, np.uint32))["max_code"] <= 1


def test_linearity_in_radiance_at_full_size(ctx):
    """Scaling level 0 by 4 (an exponent shift: exact in rgbe) scales level 1 by exactly 4."""
    w, levels = 512, 2
    bits = synth.synthetic_chain(w, w, levels, probe=15, sun=False)
    scaled = bits.copy()
    n0 = 6 * w * w
    scaled[:n0] = bits[:n0] + np.uint32(2 << 27)
    a, a_f32 = run_chain_device(ctx, bits, w, w, levels, 1024)
    b, b_f32 = run_chain_device(ctx, scaled, w, w, levels, 1024)
    assert np.array_equal(b_f32, a_f32 * np.float32(4.0))
    assert np.array_equal(b[n0:], a[n0:] + np.uint32(2 << 27))


def test_bad_arguments_raise(ctx):
    bits = np.zeros(6 * (64 + 16 + 4 + 1), np.uint32)
    with pytest.raises(datum_b200.IblError):
        ctx.image_buildmips_cube_ibl(8, 8, 6, np.zeros(6 * 100, np.uint32))      # 8 >> 5 == 0: level without texels
    with pytest.raises(datum_b200.IblError):
        ctx.image_buildmips_cube_ibl(8, 8, 4, bits, samples=0)
    with pytest.raises(ValueError):
        ctx.image_buildmips_cube_ibl(8, 8, 4, np.zeros(10, np.uint32))             # payload too small
    d = torch.zeros(6 * 64, dtype=torch.int32, device=DEV)
    out = torch.zeros(6 * 16, dtype=torch.int32, device=DEV)
    with pytest.raises(datum_b200.IblError):
        ctx.prefilter_level_device(d, 8, 8, 1, 4, 1024, 0, 6 * 4 + 1, out)         # rows outside the level
    with pytest.raises(datum_b200.IblError):
        ctx.prefilter_level_device(d, 8, 8, 4, 4, 1024, 0, 6 * 4, out)             # level >= levels
    ctx.image_buildmips_cube_ibl(8, 8, 1, bits)                                    # one level: nothing to do, like ibl.cpp:247


def test_launch_counter_counts_our_kernels(ctx):
    before = ctx.launch_count
    bits = synth.synthetic_chain(16, 16, 5)
    ctx.image_buildmips_cube_ibl(16, 16, 5, bits)
    assert ctx.launch_count - before == 4          # tail levels (<= 6144 texels): one launch each, no record pass
    before = ctx.launch_count
    bits = synth.synthetic_chain(128, 128, 3)
    ctx.image_buildmips_cube_ibl(128, 128, 3, bits)
    assert ctx.launch_count - before == 2 + 1      # 64^2 faces: quad-record build + prefilter; 32^2 faces: tail kernel
