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


def _c_class(param):
    """class of one C parameter declaration"""
    p = param.strip()
    if "*" in p or "[" in p:
        return "ptr"
    for key, cls in (("uint64_t", "uint64"), ("size_t", "size_t"), ("double", "double"), ("float", "float"),
                     ("unsigned", "unsigned"), ("int", "int")):
        if re.search(r"\b%s\b" % key, p):
            return cls
    raise AssertionError("unclassified C parameter: %r" % param)


def _c_prototypes():
    """name -> (return class, [parameter classes]) parsed from include/tempest_b200.h"""
    txt = open(os.path.join(ROOT, "include", "tempest_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t|const char\s*\*)\s*(tsdr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret = "ptr" if "*" in m.group(1) else m.group(1)
        params = m.group(3).strip()
        out[m.group(2)] = (ret, [] if params in ("", "void") else [_c_class(p) for p in params.split(",")])
    return out


def _julia_class(t):
    t = t.strip()
    if t.startswith("Ptr{") or t in ("Cstring",):
        return "ptr"
    return {"Cint": "int", "Csize_t": "size_t", "Cdouble": "double", "Cfloat": "float", "Cuint": "unsigned",
            "UInt64": "uint64", "Culonglong": "uint64"}[t]


def test_header_prototypes_match_the_ctypes_table():
    """the ctypes signatures the GPU tests call through are the header's prototypes, class for class"""
    protos = _c_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    py = {"c_int": "int", "c_size_t": "size_t", "c_ulong": "size_t", "c_double": "double", "c_float": "float",
          "c_uint": "unsigned", "c_uint64": "uint64", "c_ulonglong": "uint64"}
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(params), name
        for k, (a, cc) in enumerate(zip(args, params)):
            got = py.get(getattr(a, "__name__", ""), "ptr")
            if cc == "uint64" and got == "size_t":
                got = "uint64"     # c_uint64 is c_ulong on LP64
            if cc == "size_t" and got == "uint64":
                got = "size_t"
            assert got == cc, "%s argument %d: ctypes %s vs header %s" % (name, k + 1, a, cc)


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
        jl_types = [p.strip() for p in parts if p.strip()]
        n_julia = len(jl_types)
        n_c = len(_lib.SIGNATURES[name][1])
        assert n_julia == n_c, "%s: Julia passes %d argument types, the C prototype has %d" % (name, n_julia, n_c)
        # ... and each argument has the C parameter's class: pointer / int / size_t / double / float / unsigned / uint64
        c_classes = _c_prototypes()[name][1]
        for k, (jt, cc) in enumerate(zip(jl_types, c_classes)):
            assert _julia_class(jt) == cc, "%s argument %d: Julia passes %s, the header declares a %s" % (name, k + 1, jt, cc)
        ret = m.group(2)
        want = {"c_int": "Cint", "c_char_p": "Cstring", "c_ulong": "Csize_t", "c_size_t": "Csize_t"}.get(
            _lib.SIGNATURES[name][0].__name__, None)
        assert want is None or ret == want, "%s: return type %s vs %s" % (name, ret, want)
        seen.add(name)
    # the functions INTEGRATION.md promises are all bound
    for must in ("tsdr_am_demod_f32", "tsdr_sig_to_image_f32", "tsdr_downgrade_f32", "tsdr_autocorr_f32", "tsdr_vsync_f32",
                 "tsdr_chain_create", "tsdr_chain_push_host", "tsdr_chain_push_host_i16", "tsdr_chain_push_ring",
                 "tsdr_ring_create", "tsdr_ring_put", "tsdr_ring_take", "tsdr_get_spectrum_f32", "tsdr_get_welch_f32",
                 "tsdr_get_waterfall_f32", "tsdr_upsampler_create", "tsdr_comm_get_unique_id", "tsdr_comm_init_rank",
                 "tsdr_chain_allreduce", "tsdr_chain_integrate_device", "tsdr_comm_destroy", "tsdr_sync_get_beta"):
        assert must in seen, must


def test_julia_use_adds_methods_in_the_owning_modules():
    """use!(TempestSDR) must evaluate its Float32 methods in the module that OWNS each function (the names reach TempestSDR
    through `@reexport using .X`, src/TempestSDR.jl:27-46, and cannot be extended from there).  Static check against the
    reference checkout when it is present (this container; the GPU box has no /root/reference)."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    src = open(os.path.join(ROOT, "tempestsdr.jl_b200", "julia", "TempestSDRB200.jl"), encoding="utf-8").read()
    body = src[src.index("function use!(ref::Module)"):]
    blocks = re.findall(r"Core\.eval\((ref(?:\.\w+)?), quote(.*?)\n    end\)", body, flags=re.S)
    assert len(blocks) == 5
    owner_file = {"ref": "Demodulation.jl", "ref.Resampler": "Resampler.jl", "ref.Autocorrelations": "Autocorrelations.jl",
                  "ref.FrameSynchronisation": "FrameSynchronisation.jl", "ref.GetSpectrum": "GetSpectrum.jl"}
    n = 0
    for owner, code in blocks:
        text = open(os.path.join(ref, owner_file[owner]), encoding="utf-8").read()
        for fname in re.findall(r"^\s+(\w+)\(", code, flags=re.M):
            assert re.search(r"function %s\(" % fname, text), "%s is not defined in %s" % (fname, owner_file[owner])
            assert "$B.%s" % ("vsync_into" if fname == "vsync" else fname) in code
            n += 1
    assert n == 11
    # Demodulation.jl is included at TempestSDR's top level, the others are submodules (the reason for the split above)
    top = open(os.path.join(ref, "TempestSDR.jl"), encoding="utf-8").read()
    assert 'include("Demodulation.jl")' in top and "@reexport using .Resampler" in top
