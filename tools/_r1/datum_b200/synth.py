"""Deterministic synthetic HDR cube maps for tests and bench.py.

Pure numpy, no RNG library: per texel a splitmix64 hash of (probe, face, y, x)
drives three colour factors and an exposure factor on top of a sky/ground
gradient of the texel's datum-space direction (tools/ibl.cpp:269 /
data/convolve.comp:85-100), plus a small sun disc.  `rgbe_words` packs the
fp32 cube with the reference codec's arithmetic (src/math/color.h:154-162) so
the benchmark's level 0 is what assetbuilder would hand to
image_buildmips_cube_ibl.
"""

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def cube_directions(width, height, dtype=np.float64):
    """Unit direction of every texel centre, shape (6, height, width, 3); faces in the
    reference order 0 right, 1 left, 2 down, 3 up, 4 forward, 5 back."""
    u = (2 * (np.arange(width, dtype=dtype) + 0.5) / width - 1)[None, :].repeat(height, 0)
    v = (2 * (np.arange(height, dtype=dtype) + 0.5) / height - 1)[:, None].repeat(width, 1)
    one = np.ones_like(u)
    faces = [(one, v, u), (-one, v, -u), (u, -one, -v), (u, one, v), (u, v, -one), (-u, v, one)]
    d = np.stack([np.stack(f, -1) for f in faces])
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def synthetic_cube(width, height, probe=0, noise=True, sun=True):
    """fp32 RGBA cube, shape (6, height, width, 4)."""
    d = cube_directions(width, height)

    t = np.clip((d[..., 1] + 0.2) / 0.8, 0.0, 1.0)
    gradient = 0.15 + 0.85 * (t * t * (3 - 2 * t))  # smoothstep(-0.2, 0.6, d.y)

    if noise:
        f = np.arange(6, dtype=np.uint64)[:, None, None]
        y = np.arange(height, dtype=np.uint64)[None, :, None]
        x = np.arange(width, dtype=np.uint64)[None, None, :]
        key = np.uint64(0x0DA7A1B1) ^ (np.uint64(probe) << np.uint64(40)) ^ (f << np.uint64(32)) ^ (y << np.uint64(16)) ^ x
        h = _splitmix64(key)
        u = [((h >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.float64) / 65536.0 for k in range(4)]
        base = gradient * np.exp2(6 * u[3] - 3)
        rgb = np.stack([base * (0.25 + 0.75 * u[k]) for k in range(3)], -1)
    else:
        tint = 0.5 + 0.5 * np.stack([np.sin(3 * d[..., 0] + probe), np.cos(2 * d[..., 2] - probe), np.sin(2 * d[..., 1] + 1.0)], -1)
        rgb = gradient[..., None] * (0.3 + 0.7 * tint)

    if sun:
        s = np.array([0.3, 0.8, -0.5])
        s /= np.linalg.norm(s)
        disc = (d @ s) > np.cos(np.deg2rad(2.0))
        rgb = rgb + disc[..., None] * np.array([2.0e4, 1.8e4, 1.5e4])

    out = np.ones((6, height, width, 4), np.float32)
    out[..., :3] = rgb.astype(np.float32)
    return out


def rgbe_words(rgb):
    """Pack fp32 rgb(a) texels (..., >=3) into E5B9G9R9 words with the arithmetic of
    src/math/color.h:154-162 (fp32 log2/floor/round-half-away)."""
    rgb = np.asarray(rgb, dtype=np.float32)
    r = np.clip(rgb[..., 0], 0.0, 65408.0).astype(np.float32)
    g = np.clip(rgb[..., 1], 0.0, 65408.0).astype(np.float32)
    b = np.clip(rgb[..., 2], 0.0, 65408.0).astype(np.float32)
    m = np.maximum(r, np.maximum(g, b))
    with np.errstate(divide="ignore"):
        e = np.maximum(np.float32(-16.0), np.floor(np.log2(m, dtype=np.float32))) + np.float32(1.0)
    scale = np.exp2(e).astype(np.float32)

    def mant(c):
        q = (c / scale * np.float32(511.0)).astype(np.float32)
        return np.floor(q + np.float32(0.5)).astype(np.uint32)  # q >= 0: round half away from zero

    return (((e + 15).astype(np.uint32) & np.uint32(0xFF)) << np.uint32(27)) | mant(r) | (mant(g) << np.uint32(9)) | (mant(b) << np.uint32(18))


def synthetic_chain(width, height, levels, probe=0, noise=True, sun=True):
    """uint32 payload of image_datasize(width, height, 6, levels) bytes with level 0 filled."""
    total = sum((width >> i) * (height >> i) * 6 for i in range(levels))
    bits = np.zeros(total, np.uint32)
    cube = synthetic_cube(width, height, probe, noise, sun)
    bits[: 6 * width * height] = rgbe_words(cube).reshape(-1)
    return bits
