"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/tempest_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tempestsdr_b200 as tsdr
from tempestsdr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "tempest_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tsdr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.SO) if os.path.exists(_lib.SO) else _lib.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libtempest_b200.so does not export %s" % n
    # ... and the Python binding covers the same set
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.tsdr_version() == 100
    assert isinstance(lib.tsdr_last_error_string(), bytes)


def test_host_only_entry_points():
    lib = _lib.load()
    n = C.c_size_t(0)
    assert lib.tsdr_autocorr_out_len(4_000_000, 20e6, 0.0, 0.1, C.byref(n)) == 0 and n.value == 2_000_000
    assert lib.tsdr_autocorr_out_len(1000, 20e6, 0.0, 0.1, C.byref(n)) == -5  # BoundsError of the reference
    assert b"BoundsError" in lib.tsdr_last_error_string()
    assert lib.tsdr_autocorr_out_len(1000, 20e6, 0.0, 0.1, None) == -1


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    assert lib.tsdr_naive_resampler_f32(None, None, 10, 0) == -1
    assert lib.tsdr_am_demod_f32(None, None, 5) == -1
    assert lib.tsdr_chain_push_host(None, None, 0, None) == -1
    assert lib.tsdr_vsync_f32(None, None, None, None) == -1


@pytest.mark.skipif(tsdr.device_count() > 0, reason="needs a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(tsdr.TempestError) as e:
        tsdr.amDemod(np.ones(4, np.complex64))
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(tsdr.TempestError):
        tsdr.Chain(20e6, tsdr.VideoMode(2576, 1125, 60), max_samples=10 ** 6)
    with pytest.raises(tsdr.TempestError):
        tsdr.calculate_autocorrelation(np.ones(4096, np.float32), 4096.0, 0, 0.5)
    with pytest.raises(tsdr.TempestError):
        tsdr.SyncXY()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tempestsdr.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import orc" not in txt and "oracle_np" not in txt and "tsdr_oracle" not in txt, f
