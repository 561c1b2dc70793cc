"""Parity criteria shared by the GPU tests, smoke() and bench.py (TEST INFRASTRUCTURE).

Contract (DESIGN.md "Parity"):
  (i)  fp32 value handed to rgbe() (tools/ibl.cpp:269): max relative error per texel
       channel <= 1e-3 (relative to the texel's largest channel);
  (ii) packed words: <= 1 mantissa code apart, no exponent mismatch, >= 99 % identical.
Texels that own samples lying within 2e-6 of a cube-face boundary are held to a
looser bound: there the reference's own result is decided by the last bit of fp32
rounding (strict vs -ffast-math builds of the unmodified reference disagree on
exactly these texels; an exact tie is undefined behaviour in tools/ibl.cpp:51-85),
so no second implementation can match it to 1e-3.  Their error may not exceed the
weight share of those samples times the brightest radiance they could fetch.
"""

import numpy as np

import oracle_lib

TOL_F32 = 1e-3


def total_sample_weight(level, levels, samples):
    """Sum of NdotL over accepted samples (tools/ibl.cpp:182); the same for every texel up to rounding."""
    n = np.array([0.0, 0.0, -1.0], np.float32)
    dirs = np.zeros((samples, 3), np.float32)
    ndotl = np.zeros(samples, np.float32)
    roughness = np.float32(level) / np.float32(levels - 1)
    oracle_lib.oracle().oracle_trace_samples(float(roughness), samples, n.ctypes.data, dirs.ctypes.data, ndotl.ctypes.data)
    return float(ndotl[ndotl > 0].astype(np.float64).sum())


def check_level(got_words, got_f32, src_words, ws, hs, level, levels, samples, row_begin=0, row_end=None, tol=TOL_F32, min_identical=0.99):
    """Compare one level computed by the CUDA path from `src_words` with the oracle on the
    SAME source.  Returns a dict of measured figures; raises AssertionError on violation."""
    wd, hd = ws >> 1, hs >> 1
    if row_end is None:
        row_end = 6 * hd
    want_words, want_f32 = oracle_lib.prefilter_level(src_words, ws, hs, level, levels, samples, row_begin, row_end)
    sl = slice(row_begin * wd, row_end * wd)
    edge_counts = oracle_lib.edge_ambiguous_counts(wd, hd, level, levels, samples, row_begin=row_begin, row_end=row_end)[sl]
    edge = edge_counts > 0
    clean = ~edge

    report = {"level": level, "texels": int(clean.size), "edge_fraction": float(edge.mean())}

    if got_f32 is not None:
        got = np.asarray(got_f32, dtype=np.float64).reshape(-1, 3)[sl]
        rel = oracle_lib.relative_error(got, want_f32[sl])
        report["max_rel_clean"] = float(rel[clean].max()) if clean.any() else 0.0
        report["max_rel_edge"] = float(rel[edge].max()) if edge.any() else 0.0
        assert report["max_rel_clean"] <= tol, report
        if edge.any():
            brightest = float(oracle_lib.rgbe_decode_array(src_words)[:, :3].max())
            share = edge_counts[edge] / total_sample_weight(level, levels, samples)
            abs_err = np.abs(got[edge] - want_f32[sl][edge].astype(np.float64)).max(axis=1)
            allowed = share * brightest + tol * want_f32[sl][edge].max(axis=1)
            assert np.all(abs_err <= allowed), report

    if got_words is not None:
        stats = oracle_lib.word_stats(np.asarray(got_words).reshape(-1)[sl][clean], want_words[sl][clean])
        report.update(stats)
        assert stats["max_code"] <= 1, report
        assert stats["exp_mismatch"] == 0 or stats["max_value_rel"] <= 4e-3, report   # exponent roll-over pairs are one code apart in value
        if clean.sum() >= 512:
            assert stats["identical"] >= min_identical, report

    return report
