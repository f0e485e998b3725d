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


def test_julia_wrapper_binds_declared_symbols_with_matching_arity():
    """The Julia ccall wrapper cannot be executed here (no Julia in the image): check statically that every
    symbol it binds is declared in include/tempest_b200.h and that each ccall passes as many argument types
    as the C prototype has parameters."""
    import re
    from tempestsdr_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "tempestsdr.jl_b200", "julia", "TempestSDRB200.jl"), encoding="utf-8").read()
    calls = list(re.finditer(r"\(:(tsdr_[a-z0-9_]+), LIB\),\s*([A-Za-z]+),\s*\(", src))
    assert len(calls) >= 25
    seen = set()
    for m in calls:
        name = m.group(1)
        assert name in _lib.SIGNATURES, "Julia wrapper binds %s, which the header does not declare" % name
        i, depth = m.end(), 1            # scan the argument-type tuple to its closing parenthesis
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        # top-level commas only (Ptr{Ptr{Cvoid}} has none, but stay safe with braces)
        parts, level, cur = [], 0, ""
        for ch in tup:
            if ch in "{(":
                level += 1
            elif ch in "})":
                level -= 1
            if ch == "," and level == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        n_julia = len([p for p in parts if p.strip()])
        n_c = len(_lib.SIGNATURES[name][1])
        assert n_julia == n_c, "%s: Julia passes %d argument types, the C prototype has %d" % (name, n_julia, n_c)
        ret = m.group(2)
        want = {"c_int": "Cint", "c_char_p": "Cstring", "c_ulong": "Csize_t", "c_size_t": "Csize_t"}.get(
            _lib.SIGNATURES[name][0].__name__, None)
        assert want is None or ret == want, "%s: return type %s vs %s" % (name, ret, want)
        seen.add(name)
    # the functions INTEGRATION.md promises are all bound
    for must in ("tsdr_am_demod_f32", "tsdr_sig_to_image_f32", "tsdr_downgrade_f32", "tsdr_autocorr_f32", "tsdr_vsync_f32",
                 "tsdr_chain_create", "tsdr_chain_push_host", "tsdr_chain_push_host_i16", "tsdr_chain_push_ring",
                 "tsdr_ring_create", "tsdr_ring_put", "tsdr_ring_take", "tsdr_get_spectrum_f32", "tsdr_get_welch_f32",
                 "tsdr_get_waterfall_f32", "tsdr_upsampler_create"):
        assert must in seen, must
