"""Consumes tests/golden/julia_v1/ -- outputs of the REAL reference produced by tools/julia_goldens.jl -- when it
exists, and compares the CPU oracle with them.  Without the directory (this image has no Julia) the tests are
skipped and parity stays "unpinned"; `pinned()` is what __graft_entry__.smoke() reports.

Bars: bit-exact for everything the oracle claims bit-exactly (envelope, both resizes, frames, EMA image, offsets,
beta tables); FFT based results within the tolerances stated in SURVEY.md 8(d).  A bit-level difference in the
projections (column / row sums) is reported by its own test: that is where Julia's @simd association -- which the
oracle has to fix by convention -- would show up first.
"""
import importlib.util
import os

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("julia_cases", os.path.join(HERE, "golden", "julia_cases.py"))
jc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(jc)

DTYPES = {"f32": np.float32, "f64": np.float64, "i32": np.int32}


def pinned():
    return os.path.exists(os.path.join(jc.OUTPUTS, "manifest.txt"))


def load_manifest():
    arrays, notes = {}, []
    with open(os.path.join(jc.OUTPUTS, "manifest.txt")) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith("#"):
                notes.append(line[2:])
                continue
            name, tag, *dims = line.split()
            a = np.fromfile(os.path.join(jc.OUTPUTS, name + ".bin"), dtype=DTYPES[tag])
            dims = [int(d) for d in dims]
            arrays[name] = a.reshape(dims[::-1]).T if len(dims) > 1 else a    # Julia is column-major
    return arrays, notes


needs_julia = pytest.mark.skipif(not pinned(), reason="no Julia goldens (tests/golden/julia_v1): parity unpinned; "
                                                      "run tools/julia_goldens.jl where Julia is installed")


def test_inputs_are_reproducible():
    """the seeded inputs hash to what was written for the Julia run (or, without one, are at least deterministic)"""
    a, b = jc.inputs(), jc.inputs()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    if pinned():
        _, notes = load_manifest()
        want = {n.split()[1]: n.split()[3] for n in notes if n.startswith("input ")}
        for k, z in a.items():
            assert jc.sha(z) == want[k], "input %s differs from the one the Julia goldens were computed on" % k


@needs_julia
def test_demodulation_and_resizes_bit_exact():
    G, _ = load_manifest()
    I = jc.inputs()
    assert np.array_equal(orc.amDemod(I["demod"]), G["amDemod"])
    assert np.array_equal(orc.invert_amDemod(I["demod"]), G["invert_amDemod"])
    assert np.array_equal(orc.abs2(I["demod"]), G["abs2"])
    assert np.max(np.abs(orc.fmDemod(I["demod"]) - G["fmDemod"])) <= 4 * np.spacing(np.float32(np.pi))
    sig = I["resize"].real.astype(np.float32)
    assert np.array_equal(orc.sig_to_image(sig, 45, 52), G["sig_to_image_45x52"])
    assert np.array_equal(orc.sig_to_image(sig, 70, 93), G["sig_to_image_70x93"])
    assert np.array_equal(orc.downgradeImage(orc.sig_to_image(sig, 70, 93)), G["downgrade_70x93"])
    big = orc.sig_to_image(np.concatenate([sig] * 4), 700, 900)
    assert np.array_equal(orc.downgradeImage(big), G["downgrade_700x900"])
    assert np.array_equal(orc.naiveResampler(sig[:100], 3), G["naiveResampler"])


@needs_julia
@pytest.mark.parametrize("case", jc.CHAIN_CASES, ids=[c[0] for c in jc.CHAIN_CASES])
def test_chain_frames_offsets_tables(case):
    name, Fs, (x_t, y_t, fv), frames = case
    G, _ = load_manifest()
    iq = jc.inputs()[name]
    so = orc.SyncXY()
    img, pub, sy, sx = orc.chain_buffer(iq, Fs, x_t, y_t, fv, 0.1, so, np.zeros((600, 800), np.float32))
    S = orc.frame_samples(Fs, fv)
    first = orc.downgradeImage(orc.sig_to_image(orc.amDemod(iq[:S]), y_t, x_t))
    assert np.array_equal(first, G[name + "_frame1"])
    assert np.array_equal(so.h, G[name + "_h"])
    assert list(sy) == list(G[name + "_sy"]) and list(sx) == list(G[name + "_sx"])      # the bit-exact bar of north_star
    assert np.array_equal(img, G[name + "_imageOut"])
    np.testing.assert_allclose(so.beta_x(), G[name + "_beta_x"], rtol=1e-5)
    np.testing.assert_allclose(so.beta_y(), G[name + "_beta_y"], rtol=1e-5)


@needs_julia
@pytest.mark.parametrize("case", jc.CHAIN_CASES, ids=[c[0] for c in jc.CHAIN_CASES])
def test_projection_association(case):
    """sum(image; dims=1) / dims=2 of the last frame: equal within 2 ulp always; bit-equality tells whether the
    oracle's fixed association happens to be the one this Julia build used"""
    name, Fs, (x_t, y_t, fv), frames = case
    G, _ = load_manifest()
    iq = jc.inputs()[name]
    S = orc.frame_samples(Fs, fv)
    last = orc.downgradeImage(orc.sig_to_image(orc.amDemod(iq[(frames - 1) * S: frames * S]), y_t, x_t))
    cs, rs = orc.proj_cols(last), orc.proj_rows(last)
    np.testing.assert_allclose(cs, G[name + "_colsum"], rtol=3e-7)
    np.testing.assert_allclose(rs, G[name + "_rowsum"], rtol=3e-7)
    assert np.array_equal(rs, G[name + "_rowsum"])      # dims=2 is sequential in Base: must be bit-equal


@needs_julia
def test_autocorrelation_and_picks():
    G, notes = load_manifest()
    a = jc.AUTOCORR
    x = jc.inputs()["autocorr"].real.astype(np.float32)
    got, lags = orc.calculate_autocorrelation(x, a["Fs"], 0, 0.15)
    ref = G["autocorr_log"]
    assert got.shape == ref.shape and np.array_equal(lags, G["autocorr_lags"])
    near = ref > ref.max() - 60
    assert np.max(np.abs(got[near] - ref[near])) <= 1e-2
    assert orc.findmax(got[1:])[1] == orc.findmax(ref[1:])[1]
    rates, gz = orc.zoom_autocorr(got, a["Fs"], 50, 90)
    assert np.array_equal(rates, G["zoom_rates"])
    pos = orc.findmax(gz)[1]
    assert [float(pos), 1 / (1 / rates[pos - 1])] == list(G["refresh_pick"])
    import tempestsdr_b200 as tsdr
    for n in notes:
        if n.startswith("closest "):
            lhs, rhs = n[len("closest "):].split(" => ")
            y_t, r = [float(v) for v in lhs.split()]
            assert sorted(tsdr.find_closest_configuration(y_t, r)) == rhs.split(" | ")
