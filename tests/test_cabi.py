"""The C-ABI library loads and exports every symbol include/datum_ibl_cuda.h
declares; host-side helpers match the reference's formulas; without a GPU the
product fails loudly instead of falling back.  CPU only: no compute calls."""

import ctypes
import os
import re

import numpy as np
import pytest

import datum_b200
from datum_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "datum_ibl_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(datum_ibl_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_chain_size_matches_reference_formula():
    lib = _lib.load()
    # tools/assetpacker.cpp:488-497 with layers = 6; assetbuilder.cpp:479-484 bakes 512^2 x 8 levels
    assert lib.datum_ibl_chain_bytes(512, 512, 8) == datum_b200.image_datasize(512, 512, 6, 8) == 4 * 6 * sum((512 >> i) ** 2 for i in range(8))
    assert lib.datum_ibl_chain_bytes(24, 12, 3) == datum_b200.image_datasize(24, 12, 6, 3)
    assert datum_b200.image_maxlevels(512, 512) == 10 and datum_b200.image_maxlevels(2048, 2048) == 12   # assetpacker.cpp:472-484
    assert datum_b200.image_maxlevels(24, 12) == 4
    assert datum_b200.level_offsets(16, 16, 3) == [0, 1536, 1920, 2016]


def test_sh9_finish_is_host_arithmetic():
    lib = _lib.load()
    partial = np.arange(1, 29, dtype=np.float64)
    sh = np.zeros(27, np.float32)
    lib.datum_ibl_sh9_finish(partial.ctypes.data, sh.ctypes.data)
    assert np.allclose(sh, partial[:27] * 4 * np.pi / 28.0, rtol=1e-6)   # project.comp:99-105


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(datum_b200.IblError) as err:
        datum_b200.IblContext(0)
    assert "no CPU fallback" in str(err.value)
    with pytest.raises(datum_b200.IblError):
        datum_b200.image_buildmips_cube_ibl(8, 8, 2, np.zeros(6 * (64 + 16), np.uint32))


def test_product_package_never_touches_the_oracle():
    """Nothing under datum_b200/ may import, link or load oracle/ or the test emulation."""
    pkg = os.path.join(ROOT, "datum_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, name), errors="replace").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "prefilter_emu" not in text, os.path.join(dirpath, name)
