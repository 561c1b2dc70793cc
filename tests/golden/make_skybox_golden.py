"""Regenerates tests/golden/skybox64.npz, skybox256.npz and skybox512.npz: BASELINE config 1 (the reference's
bundled data/skybox_{rt,lf,dn,up,fr,bk}.jpg cube) reduced 8x to 64^2 faces (the whole chain fits a
small fixture) and 2x to 256^2 faces (level 1 then runs the benchmark's dominant launch shape; the
fixture holds the faces as 8-bit RGB and the levels >= 1 the UNMODIFIED reference produced, ~25 s),
and at its NATIVE 512^2 faces x 8 levels (tools/assetbuilder.cpp:416-470 as shipped; same fixture form).

Runs in the authoring container only (needs /root/reference and PIL).  Steps, mirroring
write_skybox_asset(fout, id, paths) (tools/assetbuilder.cpp:416-470):
  1. decode the six JPEGs (PIL here, QImage in the reference), box-average 8x8 -> 64x64, pack as
     QImage::Format_ARGB32 pixels (0xFFRRGGBB) in the call-site order rt, lf, dn, up, fr, bk
     (assetbuilder.cpp:876);
  2. level 0 = rgbe(srgba(pixel)), mirrored — the oracle's restatement of :443-462 (its per-pixel
     arithmetic is pinned against the compiled reference in ibl_golden.npz; the loop needs Qt);
  3. levels 1..6 = the UNMODIFIED reference image_buildmips_cube_ibl (oracle/_ref), 1024 samples.

    python tests/golden/make_skybox_golden.py
"""

import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402

ORDER = ("rt", "lf", "dn", "up", "fr", "bk")   # tools/assetbuilder.cpp:876
W, LEVELS = 64, 7
W_BIG, LEVELS_BIG = 256, 8
W_NATIVE, LEVELS_NATIVE = 512, 8     # tools/assetbuilder.cpp:434-437: the images' own size, 8 levels


def reduced_faces(w):
    faces = np.zeros((6, w, w), np.uint32)
    for f, name in enumerate(ORDER):
        rgb = np.asarray(Image.open("/root/reference/data/skybox_%s.jpg" % name).convert("RGB"), np.float64)
        k = rgb.shape[0] // w
        small = rgb.reshape(w, k, w, k, 3).mean(axis=(1, 3)).round().astype(np.uint32)
        faces[f] = 0xFF000000 | small[..., 0] << 16 | small[..., 1] << 8 | small[..., 2]
    return faces


def reference_chain(faces, levels):
    w = faces.shape[2]
    total = sum(6 * (w >> i) ** 2 for i in range(levels))
    chain = np.zeros(total, np.uint32)
    chain[: 6 * w * w] = oracle_lib.ingest_cube_argb32(faces)
    oracle_lib.ref().ref_image_buildmips_cube_ibl(w, w, levels, chain.ctypes.data)
    return chain


def main():
    faces = reduced_faces(W)
    chain = reference_chain(faces, LEVELS)
    path = os.path.join(HERE, "skybox64.npz")
    np.savez_compressed(path, faces_argb=faces, chain=chain)
    print("wrote", path, os.path.getsize(path), "bytes")

    faces = reduced_faces(W_BIG)
    chain = reference_chain(faces, LEVELS_BIG)
    rgb = np.stack([(faces >> 16) & 0xFF, (faces >> 8) & 0xFF, faces & 0xFF], axis=-1).astype(np.uint8)
    path = os.path.join(HERE, "skybox256.npz")
    np.savez_compressed(path, faces_rgb=rgb, levels=chain[6 * W_BIG * W_BIG:])
    print("wrote", path, os.path.getsize(path), "bytes")

    # BASELINE config 1 at NATIVE size: the six 512^2 images as decoded, the whole bake by the unmodified
    # reference (one thread, ~100-130 s: the number BASELINE.md extrapolated); `seconds` records it
    import time
    faces = reduced_faces(W_NATIVE)
    t0 = time.perf_counter()
    chain = reference_chain(faces, LEVELS_NATIVE)
    seconds = time.perf_counter() - t0
    rgb = np.stack([(faces >> 16) & 0xFF, (faces >> 8) & 0xFF, faces & 0xFF], axis=-1).astype(np.uint8)
    path = os.path.join(HERE, "skybox512.npz")
    np.savez_compressed(path, faces_rgb=rgb, levels=chain[6 * W_NATIVE * W_NATIVE:], reference_seconds=np.float64(seconds))
    print("wrote", path, os.path.getsize(path), "bytes; reference bake (strict build, one thread) took %.1f s" % seconds)


if __name__ == "__main__":
    main()
