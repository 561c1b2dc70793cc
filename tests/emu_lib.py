"""TEST INFRASTRUCTURE: builds and loads the host compilation of the kernel math
(tests/emu/prefilter_emu.cpp).  Never imported by datum_b200."""

import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "emu", "prefilter_emu.cpp"), os.path.join(ROOT, "datum_b200", "csrc", "ibl_tables.cpp")]
DEPS = SRC + [os.path.join(ROOT, "datum_b200", "csrc", "ibl_math.cuh"), os.path.join(ROOT, "datum_b200", "csrc", "ibl_tables.h")]
LIB = os.path.join(ROOT, "tests", "emu", "libprefilter_emu.so")

_lib = None


def load():
    global _lib
    if _lib is None:
        stale = not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS)
        if stale:
            gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB] + SRC)
        lib = ctypes.CDLL(LIB)
        lib.emu_prefilter_level.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2
        lib.emu_prefilter_level_dn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
        lib.emu_prefilter_level_dp.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
        lib.emu_paired_table.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
        lib.emu_paired_table.restype = ctypes.c_int
        lib.emu_sector_study.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
        lib.emu_footprint_compare.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p]
        lib.emu_sector_table.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.emu_sector_table.restype = ctypes.c_int
        lib.emu_banded_table.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3
        lib.emu_banded_table.restype = ctypes.c_int
        lib.emu_patch_table.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3
        lib.emu_patch_table.restype = ctypes.c_int
        lib.emu_pack_dn_word.argtypes = [ctypes.c_uint32]
        lib.emu_pack_dn_word.restype = ctypes.c_uint32
        lib.emu_dn_tap.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
        lib.emu_raw_tap.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p]
        lib.emu_last_fast_fraction.restype = ctypes.c_double
        lib.emu_rgbe_encode.restype = ctypes.c_uint32
        lib.emu_rgbe_encode.argtypes = [ctypes.c_float] * 3
        lib.emu_rgbe_decode.argtypes = [ctypes.c_uint32, ctypes.c_void_p]
        lib.emu_texel_normal.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
        lib.emu_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.emu_table.restype = ctypes.c_int
        _lib = lib
    return _lib
