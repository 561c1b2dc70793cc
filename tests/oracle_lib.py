"""ctypes access to the CPU oracle (oracle/liboracle_ibl.so) and to the compiled
reference (oracle/_ref) — TEST INFRASTRUCTURE, plus the parity metrics shared by
the tests, smoke() and bench.py."""

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle_ibl.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libdatum_ref_ibl.so")
REF_FAST_PATH = os.path.join(ROOT, "oracle", "_ref", "libdatum_ref_ibl_fast.so")

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

_oracle = None
_ref = {}


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_PATH):
            raise RuntimeError("oracle not built: run `make -C oracle`")
        lib = ctypes.CDLL(ORACLE_PATH)
        lib.oracle_rgbe_encode.restype = ctypes.c_uint32
        lib.oracle_rgbe_encode.argtypes = [c_float] * 3
        lib.oracle_rgbe_decode.argtypes = [ctypes.c_uint32, c_void_p]
        lib.oracle_rgbe_encode_array.argtypes = [c_void_p, c_size_t, c_int, c_void_p]
        lib.oracle_rgbe_decode_array.argtypes = [c_void_p, c_size_t, c_void_p]
        lib.oracle_srgba_decode.argtypes = [ctypes.c_uint32, c_void_p]
        lib.oracle_ingest_cube_argb32.argtypes = [c_void_p, c_int, c_int, c_void_p]
        lib.oracle_radicalinverse.restype = c_float
        lib.oracle_radicalinverse.argtypes = [ctypes.c_uint32]
        lib.oracle_texel_direction.argtypes = [c_int] * 5 + [c_void_p]
        lib.oracle_cube_sample.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
        lib.oracle_convolve_dir.argtypes = [c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p]
        lib.oracle_trace_samples.argtypes = [c_float, c_int, c_void_p, c_void_p, c_void_p]
        lib.oracle_edge_ambiguous_counts.argtypes = [c_int, c_int, c_float, c_int, c_float, c_void_p, c_int]
        lib.oracle_edge_ambiguous_counts_rows.argtypes = [c_int, c_int, c_float, c_int, c_float, c_int, c_int, c_void_p, c_int]
        lib.oracle_prefilter_level.argtypes = [c_void_p, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_int]
        lib.oracle_buildmips_cube_ibl.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int]
        lib.oracle_sh9_partial.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        lib.oracle_sh9_finish.argtypes = [c_void_p, c_void_p]
        lib.oracle_project_sh9.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p]
        lib.oracle_sh9_irradiance.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        lib.oracle_pack_envbrdf.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_int]
        lib.oracle_pack_watercolor.argtypes = [c_void_p, c_void_p, c_float, c_void_p, c_float, c_float, c_int, c_int, c_void_p]
        lib.oracle_face_rotate.argtypes = [c_int, c_void_p, c_void_p]
        lib.oracle_image_pack_cube.argtypes = [c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
        lib.oracle_image_pack_cube_ibl.argtypes = [c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int]
        lib.oracle_image_blend_edges.argtypes = [c_int, c_int, c_int, c_void_p]
        lib.oracle_max_threads.restype = c_int
        _oracle = lib
    return _oracle


def have_ref(fast=False):
    return os.path.exists(REF_FAST_PATH if fast else REF_PATH)


def ref(fast=False):
    """The unmodified reference tools/ibl.cpp + tools/hdr.cpp (strict IEEE build, or the
    reference's own -O2 -ffast-math flags with fast=True)."""
    if fast not in _ref:
        lib = ctypes.CDLL(REF_FAST_PATH if fast else REF_PATH)
        lib.ref_image_buildmips_cube_ibl.argtypes = [c_int, c_int, c_int, c_void_p]
        lib.ref_image_pack_cube_ibl.argtypes = [c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
        lib.ref_image_pack_cube.argtypes = [c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
        lib.ref_image_pack_envbrdf.argtypes = [c_int, c_int, c_void_p]
        lib.ref_image_pack_watercolor.argtypes = [c_void_p, c_void_p, c_float, c_void_p, c_float, c_float, c_int, c_int, c_void_p]
        lib.ref_rgbe_encode.restype = ctypes.c_uint32
        lib.ref_rgbe_encode.argtypes = [c_float] * 3
        lib.ref_rgbe_decode.argtypes = [ctypes.c_uint32, c_void_p]
        lib.ref_srgba_decode.argtypes = [ctypes.c_uint32, c_void_p]
        lib.ref_face_rotate.argtypes = [c_int, c_void_p, c_void_p]
        _ref[fast] = lib
    return _ref[fast]


# ---- oracle wrappers -----------------------------------------------------------

def rgbe_encode_array(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    stride = rgb.shape[-1]
    flat = rgb.reshape(-1, stride)
    words = np.zeros(len(flat), np.uint32)
    oracle().oracle_rgbe_encode_array(flat.ctypes.data, len(flat), stride, words.ctypes.data)
    return words.reshape(rgb.shape[:-1])


def rgbe_decode_array(words):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.zeros(words.shape + (4,), np.float32)
    oracle().oracle_rgbe_decode_array(words.ctypes.data, words.size, out.ctypes.data)
    return out


def ingest_cube_argb32(faces):
    """tools/assetbuilder.cpp:443-462 on a (6, H, W) uint32 array of ARGB32 pixels -> level 0 words."""
    faces = np.ascontiguousarray(faces, dtype=np.uint32)
    height, width = faces.shape[1], faces.shape[2]
    out = np.zeros(6 * width * height, np.uint32)
    oracle().oracle_ingest_cube_argb32(faces.ctypes.data, width, height, out.ctypes.data)
    return out


def prefilter_level(src, ws, hs, level, levels, samples=1024, row_begin=0, row_end=None, threads=0):
    """(words, f32) of the (ws/2 x hs/2 x 6) level computed from rgbe level `src`."""
    src = np.ascontiguousarray(src, dtype=np.uint32)
    wd, hd = ws >> 1, hs >> 1
    if row_end is None:
        row_end = 6 * hd
    words = np.zeros(6 * hd * wd, np.uint32)
    f32 = np.zeros((6 * hd * wd, 3), np.float32)
    roughness = np.float32(level) / np.float32(levels - 1)
    oracle().oracle_prefilter_level(src.ctypes.data, ws, hs, float(roughness), samples, row_begin, row_end, words.ctypes.data, f32.ctypes.data, threads)
    return words, f32


def buildmips_cube_ibl(width, height, levels, bits, samples=1024, threads=0, want_f32=False):
    """In place on `bits` like tools/ibl.cpp:242; optionally returns the pre-quantisation fp32 of levels >= 1."""
    assert bits.dtype == np.uint32 and bits.flags["C_CONTIGUOUS"]
    total = sum((width >> i) * (height >> i) * 6 for i in range(levels))
    f32 = np.zeros((total - 6 * width * height, 3), np.float32) if want_f32 else None
    oracle().oracle_buildmips_cube_ibl(width, height, levels, samples, bits.ctypes.data, f32.ctypes.data if want_f32 else None, threads)
    return f32


def edge_ambiguous_counts(wd, hd, level, levels, samples=1024, eps=2e-6, threads=0, row_begin=0, row_end=None):
    """Per texel: samples lying within eps of a cube-face boundary.  With a row range only those
    rows are evaluated (the others stay 0): the full-size spot checks need a few rows of a big level."""
    counts = np.zeros(6 * hd * wd, np.int32)
    roughness = np.float32(level) / np.float32(levels - 1)
    if row_end is None:
        row_end = 6 * hd
    oracle().oracle_edge_ambiguous_counts_rows(wd, hd, float(roughness), samples, eps, row_begin, row_end, counts.ctypes.data, threads)
    return counts


def project_sh9(level0, fmt, w, h):
    level0 = np.ascontiguousarray(level0)
    sh = np.zeros((9, 3), np.float64)
    oracle().oracle_project_sh9(level0.ctypes.data, fmt, w, h, sh.ctypes.data)
    return sh


def sh9_partial(level0, fmt, w, h, row_begin, row_end):
    level0 = np.ascontiguousarray(level0)
    partial = np.zeros(28, np.float64)
    oracle().oracle_sh9_partial(level0.ctypes.data, fmt, w, h, row_begin, row_end, partial.ctypes.data)
    return partial


def sh9_finish(partial):
    partial = np.ascontiguousarray(partial, dtype=np.float64)
    sh = np.zeros((9, 3), np.float64)
    oracle().oracle_sh9_finish(partial.ctypes.data, sh.ctypes.data)
    return sh


def sh9_irradiance(sh, normals):
    sh = np.ascontiguousarray(sh, dtype=np.float64)
    normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((len(normals), 3), np.float32)
    oracle().oracle_sh9_irradiance(sh.ctypes.data, normals.ctypes.data, len(normals), out.ctypes.data)
    return out


def pack_envbrdf(width, height, samples=1024, threads=0):
    words = np.zeros(width * height, np.uint32)
    f32 = np.zeros((width * height, 3), np.float32)
    oracle().oracle_pack_envbrdf(width, height, samples, words.ctypes.data, f32.ctypes.data, threads)
    return words, f32


def pack_watercolor(deep, shallow, depthscale, fresnel, fresnelbias, fresnelpower, width, height):
    deep = np.ascontiguousarray(deep, dtype=np.float32)
    shallow = np.ascontiguousarray(shallow, dtype=np.float32)
    fresnel = np.ascontiguousarray(fresnel, dtype=np.float32)
    words = np.zeros(width * height, np.uint32)
    oracle().oracle_pack_watercolor(deep.ctypes.data, shallow.ctypes.data, depthscale, fresnel.ctypes.data, fresnelbias, fresnelpower, width, height, words.ctypes.data)
    return words


def image_pack_cube(image, width, height, levels=1):
    """tools/hdr.cpp:331-359 on an (H, W, 4) float32 equirect image."""
    image = np.ascontiguousarray(image, dtype=np.float32)
    total = sum(6 * (width >> i) * (height >> i) for i in range(levels))
    bits = np.zeros(total, np.uint32)
    oracle().oracle_image_pack_cube(image.shape[1], image.shape[0], image.ctypes.data, width, height, levels, bits.ctypes.data)
    return bits


def image_pack_cube_ibl(image, width, height, levels, samples=1024, threads=0):
    """tools/ibl.cpp:283-288"""
    image = np.ascontiguousarray(image, dtype=np.float32)
    total = sum(6 * (width >> i) * (height >> i) for i in range(levels))
    bits = np.zeros(total, np.uint32)
    oracle().oracle_image_pack_cube_ibl(image.shape[1], image.shape[0], image.ctypes.data, width, height, levels, samples, bits.ctypes.data, threads)
    return bits


# ---- parity metrics ----------------------------------------

def relative_error(got_f32, want_f32):
    """Per texel: max over channels of |got - want| / (max channel of want)."""
    got = np.asarray(got_f32, dtype=np.float64).reshape(-1, 3)
    want = np.asarray(want_f32, dtype=np.float64).reshape(-1, 3)
    scale = np.maximum(want.max(axis=1), 1e-30)
    return (np.abs(got - want).max(axis=1)) / scale


def word_stats(got, want):
    """identical fraction, max mantissa code difference (same-exponent words), exponent mismatches,
    max relative value difference over all words (covers exponent roll-over pairs)."""
    got = np.asarray(got, dtype=np.uint32).reshape(-1)
    want = np.asarray(want, dtype=np.uint32).reshape(-1)
    same_exp = (got >> 27) == (want >> 27)
    code = np.zeros(len(got), np.int64)
    for shift in (0, 9, 18):
        code = np.maximum(code, np.abs(((got >> shift) & 0x1FF).astype(np.int64) - ((want >> shift) & 0x1FF).astype(np.int64)))
    dg, dw = rgbe_decode_array(got)[:, :3].astype(np.float64), rgbe_decode_array(want)[:, :3].astype(np.float64)
    rel = np.abs(dg - dw).max(axis=1) / np.maximum(dw.max(axis=1), 1e-30)
    return {
        "identical": float((got == want).mean()) if len(got) else 1.0,
        "max_code": int(code[same_exp].max()) if same_exp.any() else 0,
        "exp_mismatch": int((~same_exp).sum()),
        "max_value_rel": float(rel.max()) if len(rel) else 0.0,
    }


def words_within_one_code(stats, min_identical=0.0):
    """Packed-word criterion: same-exponent words at most one mantissa code apart; words whose
    exponents differ (a value straddling a power of two) at most one code apart IN VALUE
    (one code of a 9-bit mantissa is at most 1/256 of the value)."""
    return stats["max_code"] <= 1 and stats["max_value_rel"] <= 4e-3 and stats["identical"] >= min_identical
