"""Regenerates tests/golden/ibl_golden.npz from the UNMODIFIED reference.

Runs in the authoring container only (needs /root/reference, compiled by
`make -C oracle ref` into oracle/_ref/libdatum_ref_ibl.so with strict IEEE
flags).  The .npz travels with the repo; the tests never read /root/reference.

    python tests/golden/make_golden.py
"""

import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from datum_b200 import synth  # noqa: E402


def main():
    ref = oracle_lib.ref()
    out = {}

    # ---- codec known answers (src/math/color.h:154-172) ----
    rng = np.random.default_rng(20261017)
    pows = np.exp2(np.arange(-20, 18)).astype(np.float32)
    cases = [
        np.zeros((1, 3), np.float32),
        np.ones((1, 3), np.float32),
        np.array([[65408.0, 1.0, 1.0], [1e9, 0.0, 0.0], [0.5, 0.25, 0.1], [255.99998, 0.0, 0.0], [-1.0, 2.0, 0.5]], np.float32),
        pows[:, None] * np.ones((1, 3), np.float32),
        np.nextafter(pows, np.float32(0))[:, None] * np.ones((1, 3), np.float32),
        (rng.random((2000, 3)) * np.exp2(rng.integers(-20, 18, (2000, 1)))).astype(np.float32),
    ]
    rgb = np.concatenate(cases).astype(np.float32)
    words = np.array([ref.ref_rgbe_encode(float(r), float(g), float(b)) for r, g, b in rgb], np.uint32)
    out["codec_rgb"] = rgb
    out["codec_words"] = words

    decode_words = np.concatenate([words[:200], rng.integers(0, 2**32, 500, dtype=np.uint64).astype(np.uint32), np.array([0, 0xFFFFFFFF, 0x84020100], np.uint32)])
    decoded = np.zeros((len(decode_words), 4), np.float32)
    for i, w in enumerate(decode_words):
        buf = (ctypes.c_float * 4)()
        ref.ref_rgbe_decode(ctypes.c_uint32(int(w)), buf)
        decoded[i] = list(buf)
    out["decode_words"] = decode_words
    out["decode_rgba"] = decoded

    # ---- sRGB decode used by the 6-image ingest (color.h:103-128) ----
    argb = rng.integers(0, 2**32, 300, dtype=np.uint64).astype(np.uint32)
    srgb = np.zeros((len(argb), 4), np.float32)
    for i, w in enumerate(argb):
        buf = (ctypes.c_float * 4)()
        ref.ref_srgba_decode(ctypes.c_uint32(int(w)), buf)
        srgb[i] = list(buf)
    out["srgba_argb"] = argb
    out["srgba_rgba"] = srgb

    # ---- face rotations (tools/ibl.cpp:253-261 through transform.h) ----
    vecs = rng.normal(size=(40, 3)).astype(np.float32)
    rotated = np.zeros((6, len(vecs), 3), np.float32)
    for f in range(6):
        for i, v in enumerate(vecs):
            vin = (ctypes.c_float * 3)(*map(float, v))
            vout = (ctypes.c_float * 3)()
            ref.ref_face_rotate(f, vin, vout)
            rotated[f, i] = list(vout)
    out["rotate_in"] = vecs
    out["rotate_out"] = rotated

    # ---- whole chains (tools/ibl.cpp:242-279, kSamples = 1024) ----
    for name, w, levels, noise in (("chain16_noise", 16, 5, True), ("chain16_smooth", 16, 5, False), ("chain32_noise", 32, 6, True)):
        bits = synth.synthetic_chain(w, w, levels, probe=3, noise=noise, sun=False)
        out[name + "_level0"] = bits[: 6 * w * w].copy()
        ref.ref_image_buildmips_cube_ibl(w, w, levels, bits.ctypes.data)
        out[name] = bits

    # non-square, non power of two: 24 x 12, 3 levels
    bits = synth.synthetic_chain(24, 12, 3, probe=4, noise=True, sun=False)
    out["chain24x12_level0"] = bits[: 6 * 24 * 12].copy()
    ref.ref_image_buildmips_cube_ibl(24, 12, 3, bits.ctypes.data)
    out["chain24x12"] = bits

    # ---- equirect -> cube (tools/hdr.cpp:331-359 incl. edge blend) and the full .hdr path (ibl.cpp:283-288) ----
    ih, iw = 40, 80
    yy, xx = np.meshgrid(np.arange(ih), np.arange(iw), indexing="ij")
    img = np.ones((ih, iw, 4), np.float32)
    img[..., 0] = 0.2 + 0.8 * np.abs(np.sin(0.17 * xx + 0.05 * yy))
    img[..., 1] = 0.1 + 0.9 * (yy / ih) ** 2
    img[..., 2] = 0.3 + 0.5 * np.abs(np.cos(0.11 * xx - 0.07 * yy))
    img[5:8, 20:24, :3] += 300.0
    img[..., :3] *= (1.0 + 0.3 * rng.random((ih, iw, 1))).astype(np.float32)
    img = np.ascontiguousarray(img.astype(np.float32))
    out["equirect"] = img
    cube = np.zeros(6 * 16 * 16, np.uint32)
    ref.ref_image_pack_cube(iw, ih, img.ctypes.data, 16, 16, 1, cube.ctypes.data)
    out["equirect_cube16"] = cube
    total = sum(6 * (16 >> i) ** 2 for i in range(4))
    chain = np.zeros(total, np.uint32)
    ref.ref_image_pack_cube_ibl(iw, ih, img.ctypes.data, 16, 16, 4, chain.ctypes.data)
    out["equirect_chain16"] = chain

    # ---- 2D LUTs (tools/ibl.cpp:292-329) ----
    lut = np.zeros(16 * 16, np.uint32)
    ref.ref_image_pack_envbrdf(16, 16, lut.ctypes.data)
    out["envbrdf16"] = lut

    deep = np.array([0.0, 0.007, 0.005], np.float32)
    shallow = np.array([0.1, 0.6, 0.7], np.float32)
    fresnel = np.array([0.0, 0.0, 0.0], np.float32)
    water = np.zeros(16 * 16, np.uint32)
    ref.ref_image_pack_watercolor(deep.ctypes.data, shallow.ctypes.data, 1.0, fresnel.ctypes.data, 0.328, 5.0, 16, 16, water.ctypes.data)
    out["water_params"] = np.concatenate([deep, shallow, [1.0], fresnel, [0.328, 5.0]]).astype(np.float32)
    out["water16"] = water

    # ---- per-pixel arithmetic of the six-image ingest (tools/assetbuilder.cpp:454): rgbe(srgba(pixel)),
    #      composed from the reference's own color.h functions (the loop around it needs Qt) ----
    gray = np.arange(256, dtype=np.uint32)
    pixels = np.concatenate([0xFF000000 | gray << 16 | gray << 8 | gray, rng.integers(0, 2**32, 3000, dtype=np.uint64).astype(np.uint32)]).astype(np.uint32)
    ingest = np.zeros(len(pixels), np.uint32)
    for i, w in enumerate(pixels):
        buf = (ctypes.c_float * 4)()
        ref.ref_srgba_decode(ctypes.c_uint32(int(w)), buf)
        ingest[i] = ref.ref_rgbe_encode(buf[0], buf[1], buf[2])
    out["ingest_argb"] = pixels
    out["ingest_words"] = ingest

    path = os.path.join(HERE, "ibl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
