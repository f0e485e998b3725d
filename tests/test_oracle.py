"""CPU tests of the oracle (test infrastructure): the C restatement against the independent
numpy restatement, against hand-derivable answers, and against the committed goldens.
PARITY UNPINNED by the reference itself: it ships no vectors on this path."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

import orc
import oracle_np as onp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, np.float32).tobytes()).digest(), np.uint8)


# ------------------------------------------------------------------ hand-derivable answers
def test_hypot_exact_cases():
    assert orc.hypot(3, 4) == 5 and orc.hypot(-5, 12) == 13 and orc.hypot(0, 0) == 0
    assert orc.hypot(np.inf, np.nan) == np.inf and np.isnan(orc.hypot(np.nan, 1))
    assert orc.hypot(1, 1e-9) == 1  # widely separated operands: returns the larger
    assert np.isfinite(orc.hypot(3e38, 1e38)) and orc.hypot(1e-30, 1e-30) > 0
    z = (np.random.default_rng(0).normal(size=20000) + 1j * np.random.default_rng(1).normal(size=20000)).astype(np.complex64)
    assert np.array_equal(orc.amDemod(z), onp.amDemod(z))  # correctly rounded: equals sqrt in float64 rounded once


def test_imresize_preserves_constants_and_ramps():
    c = np.full(1000, 0.375, np.float32)
    assert np.all(orc.imresize_1d(c, 2345) == np.float32(0.375))
    assert np.all(orc.imresize_1d(c, 300) == np.float32(0.375))
    ramp = np.arange(1, 1001, dtype=np.float32)  # a[i] = i (1-based): linear interpolation returns the coordinate itself
    out = orc.imresize_1d(ramp, 250)             # sf = 4, x(i) = 4i - 1.5
    assert np.allclose(out, 4 * np.arange(1, 251) - 1.5, rtol=0, atol=1e-4)
    up = orc.imresize_1d(ramp, 4000)             # sf = 0.25: clamped to [1, 1000] at both ends
    assert up[0] == 1 and up[-1] == 1000 and np.all(np.diff(up) >= 0)
    assert np.array_equal(orc.imresize_1d(ramp, 1000), ramp)  # same size: plain copy


def test_sig_to_image_layout():
    sig = np.arange(12, dtype=np.float32)
    img = orc.sig_to_image(sig, 3, 4)  # same size -> copy; row r is scan line r
    assert np.array_equal(img, sig.reshape(3, 4))


def test_autocorr_impulse_train():
    n, T = 4096, 64
    x = np.zeros(n, np.float32)
    x[::T] = 1.0
    lin, lags = orc.calculate_autocorrelation(x, float(n), 0, 0.5, scale="lin")
    assert lin.size == n // 2 and lags[1] == 1 / n
    r = np.sqrt(lin)
    k = np.arange(lin.size)
    assert np.allclose(r[k % T == 0], n / T / 1.0, rtol=1e-4)   # circular autocorrelation: n/T at every multiple of T
    assert np.all(r[k % T != 0] < 1e-2)
    with pytest.raises(IndexError):
        orc.calculate_autocorrelation(x[:100], float(n), 0, 0.5)


@pytest.mark.parametrize("n", [64, 1000, 3000, 4096, 30030, 3 * 10 ** 5])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    z = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    ref = np.fft.fft(z.astype(np.complex128))
    got = orc.fft(z)
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.sqrt(n) * np.max(np.abs(ref))
    back = orc.fft(got, inverse=True)
    assert np.max(np.abs(back - z)) <= 1e-5


def test_zoom_autocorr_keeps_reference_off_by_one():
    g = np.arange(1, 400001, dtype=np.float32)
    rates, sl = orc.zoom_autocorr(g, 20e6, rate_min=50, rate_max=90)
    assert sl[0] == 222222 and sl[-1] == 400000          # Gamma[222222:400000] (1-based, inclusive)
    assert rates[0] == 1 / (222222 / 20e6) and rates.size == sl.size
    r2, s2 = onp.zoom_autocorr(g, 20e6, rate_min=50, rate_max=90)
    assert np.array_equal(rates, r2) and np.array_equal(sl, s2)


def test_findmax_first_maximum_and_nan():
    v = np.array([1, 5, 3, 5, 2], np.float32)
    assert orc.findmax(v) == (5, 2)
    v[4] = np.nan
    assert orc.findmax(v)[1] == 5


def test_round_and_frame_samples():
    assert orc.frame_samples(20e6, 60.0) == 333333 and orc.frame_samples(200e6, 30.0) == 6666667
    assert orc.lib.orc_round_even(0.5) == 0 and orc.lib.orc_round_even(1.5) == 2 and orc.lib.orc_round_even(2.5) == 2


def test_syncxy_bounds_and_taps():
    s = orc.SyncXY()
    assert (s.wmin_y, s.wmax_y, s.wmin_x, s.wmax_x) == (6, 150, 40, 200)
    assert s.beta_x().shape == (161, 800) and s.beta_y().shape == (145, 600)
    assert np.array_equal(s.h, onp.gaussian_taps()) and abs(float(s.h.astype(np.float64).sum()) - 1) < 1e-6 and s.h[0] == s.h[4]


def test_vsync_stripe_frame_and_stale_beta_y():
    img = np.full((600, 800), 0.25, np.float32)
    img[300:330, :] = 1.0   # bright horizontal band  -> row projection peak
    img[:, 500:560] = 1.0   # bright vertical band    -> column projection peak
    s = orc.SyncXY()
    sy1, sx1 = orc.vsync(img, s)
    assert sy1 == 1                      # beta_y still zero on the first call (FrameSynchronisation.jl:66)
    sy2, sx2 = orc.vsync(img, s)
    assert sx2 == sx1 and 500 <= sx1 <= 565   # centre of the band, delayed by the causal 5-tap filter
    assert 300 <= sy2 <= 335
    sn = onp.SyncXY()
    assert onp.vsync(img, sn) == (sy1, sx1) and onp.vsync(img, sn) == (sy2, sx2)


def test_circshift_and_ema():
    img = np.arange(600 * 800, dtype=np.float32).reshape(600, 800)
    sh = orc.circshift(img, 3, 5)
    assert sh[0, 0] == img[3, 5] and sh[599, 799] == img[2, 4]
    assert np.array_equal(sh, onp.circshift(img, 3, 5))
    acc = np.full((4, 4), 2.0, np.float32)
    out = orc.ema(acc, np.full((4, 4), 4.0, np.float32), 0.1)
    assert np.all(out == np.float32(np.float32(0.1) * np.float32(2.0)) + np.float32(np.float32(1) - np.float32(0.1)) * np.float32(4.0))


# ------------------------------------------------------------------ C oracle vs numpy oracle
@pytest.mark.parametrize("n_in,n_out", [(3333, 28980), (3333, 1000), (1000, 1001), (1000, 999), (2, 7)])
def test_resize1d_c_vs_numpy(n_in, n_out):
    s = np.random.default_rng(n_in + n_out).random(n_in).astype(np.float32)
    assert np.array_equal(orc.imresize_1d(s, n_out), onp.imresize_1d(s, n_out))


@pytest.mark.parametrize("shape", [(225, 515), (90, 130), (600, 800), (700, 700)])
def test_downgrade_c_vs_numpy(shape):
    img = np.random.default_rng(shape[0]).random(shape).astype(np.float32)
    assert np.array_equal(orc.downgradeImage(img), onp.downgradeImage(img))


def test_chain_c_vs_numpy(synth):
    Fs, x_t, y_t, fv = 2.0e6, 800, 525, 60.0
    iq = synth.make_iq(orc.frame_samples(Fs, fv) * 3 + 10, Fs, x_t, y_t, fv, seed=3)
    a = orc.chain_buffer(iq, Fs, x_t, y_t, fv, 0.1, orc.SyncXY(), np.zeros((600, 800), np.float32))
    b = onp.chain_buffer(iq, Fs, x_t, y_t, fv, 0.1, onp.SyncXY(), np.zeros((600, 800), np.float32))
    assert list(a[2]) == list(b[2]) and list(a[3]) == list(b[3]) and a[2][0] == 1
    assert np.array_equal(a[0], b[0])
    # threads only change who renders which frame, never the result
    c = orc.chain_buffer(iq, Fs, x_t, y_t, fv, 0.1, orc.SyncXY(), np.zeros((600, 800), np.float32), nthreads=4)
    assert np.array_equal(a[0], c[0]) and list(a[2]) == list(c[2])


def test_autocorr_c_vs_scipy():
    x = (1.0 + np.random.default_rng(5).random(6000)).astype(np.float32)
    a, _ = orc.calculate_autocorrelation(x, 6000.0, 0, 0.5)
    b, _ = onp.calculate_autocorrelation(x, 6000.0, 0, 0.5)
    assert np.max(np.abs(a - b)) < 1e-3


# ------------------------------------------------------------------ committed goldens
def test_oracle_reproduces_goldens():
    z = GOLD["demod_in"]
    assert np.array_equal(orc.amDemod(z), GOLD["amDemod"])
    assert np.array_equal(orc.invert_amDemod(z), GOLD["invert_amDemod"])
    assert np.array_equal(orc.abs2(z), GOLD["abs2"])
    sig = GOLD["resize_in"]
    assert np.array_equal(orc.sig_to_image(sig, 45, 52), GOLD["sig_to_image_up"])
    assert np.array_equal(orc.sig_to_image(sig, 20, 33), GOLD["sig_to_image_down"])
    d = orc.downgradeImage(GOLD["downgrade_in"])
    assert np.array_equal(_sha(d), GOLD["downgrade_small_sha256"]) and np.array_equal(d[::37, ::41], GOLD["downgrade_small_sub"])
    a, _ = orc.calculate_autocorrelation(GOLD["autocorr_in"], 6000.0, 0, 0.5)
    assert np.array_equal(a, GOLD["autocorr_db"])


def test_oracle_chain_reproduces_golden(synth):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    c = mg.CHAIN_CASE
    iq = mg.chain_inputs()
    so = orc.SyncXY()
    acc, _, sy, sx = orc.chain_buffer(iq, c["Fs"], c["x_t"], c["y_t"], c["fv"], c["alpha"], so, np.zeros((600, 800), np.float32))
    assert np.array_equal(sy, GOLD["chain_sy"]) and np.array_equal(sx, GOLD["chain_sx"])
    assert np.array_equal(_sha(acc), GOLD["chain_image_sha256"])
    assert np.array_equal(_sha(so.beta_x()), GOLD["chain_beta_x_last_sha256"])


def test_oracle_reproduces_goldens_v2():
    G2 = np.load(os.path.join(HERE, "golden", "golden_v2.npz"))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    c = mg.CHAIN_CASE
    i16 = mg.int16_inputs()
    wide = (i16[:, 0].astype(np.float32) + 1j * i16[:, 1].astype(np.float32)).astype(np.complex64)
    so = orc.SyncXY()
    acc, _, sy, sx = orc.chain_buffer(wide, c["Fs"], c["x_t"], c["y_t"], c["fv"], c["alpha"], so,
                                      np.zeros((600, 800), np.float32), publish=False)
    assert np.array_equal(sy, G2["i16_chain_sy"]) and np.array_equal(sx, G2["i16_chain_sx"])
    assert np.array_equal(_sha(acc), G2["i16_chain_image_sha256"])
    x = G2["spectrum_in"]
    assert np.array_equal(x, mg.spectrum_input())
    assert np.array_equal(orc.getSpectrum(1.0, x, N=1000)[1], G2["getSpectrum_1000"])
    assert np.array_equal(orc.getSpectrum(1.0, x, N=1024)[1], G2["getSpectrum_1024"])
    assert np.array_equal(orc.getWelch(1.0, x, sizeFFT=256)[1], G2["getWelch_256"])
    assert np.array_equal(orc.getWaterfall(1.0, x, sizeFFT=64)[2].astype(np.float32), G2["getWaterfall_64"])


def test_get_spectrum_family_against_numpy():
    # src/GetSpectrum.jl:21-66 restated on the oracle's own FFT, checked here against numpy's double FFT
    rng = np.random.default_rng(21)
    n = 5000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    x += (3 * np.exp(2j * np.pi * 0.12 * np.arange(n))).astype(np.complex64)
    f, y = orc.getSpectrum(2.0e6, x, N=3000)                       # not a power of two
    X = np.fft.fftshift(np.fft.fft(x[:3000].astype(np.complex128)))
    assert y.dtype == np.float32 and y.shape == (3000,)
    assert np.allclose(f, (np.arange(3000) / 3000 - 0.5) * 2.0e6)
    assert np.abs(y - 10 * np.log10(np.abs(X) ** 2)).max() < 1e-3
    assert abs(f[np.argmax(y)] - 0.12 * 2.0e6) < 2.0e6 / 3000      # the carrier is where it was put
    f, y = orc.getWelch(1.0, x, sizeFFT=256)
    seg = x[: (n // 256) * 256].reshape(-1, 256).astype(np.complex128)
    S = (np.abs(np.fft.fft(seg, axis=1)) ** 2).sum(axis=0)
    assert np.abs(y - 10 * np.log10(np.fft.fftshift(S))).max() < 1e-3
    t, f, s = orc.getWaterfall(1.0, x, sizeFFT=256)
    assert s.dtype == np.float64 and s.shape == (256, n // 256) and np.allclose(t, np.arange(n // 256) * 256.0)
    ref = np.fft.fftshift(np.abs(np.fft.fft(seg, axis=1)) ** 2, axes=1).T
    assert np.abs(s - ref).max() <= 2e-6 * ref.max()


@pytest.mark.parametrize("shape", [(1300, 2100), (53, 37), (600, 800)])
def test_vsync_c_vs_numpy_any_size(shape):
    # generic SyncXY sizes, incl. projections longer than 1024 elements (Base.sum goes pairwise there): the C oracle and
    # the independent numpy restatement agree bit for bit on offsets and both beta tables
    import oracle_np as onp
    n_y, n_x = shape
    rng = np.random.default_rng(n_y)
    so, sn = orc.SyncXY(n_y, n_x), onp.SyncXY(n_y, n_x)
    for k in range(2):
        img = rng.random(shape).astype(np.float32)
        img[(n_y // 3 + 5 * k) % n_y, :] = 2.0
        img[:, (n_x // 2 + 9 * k) % n_x] = 2.0
        assert orc.vsync(img, so) == tuple(onp.vsync(img, sn))
        assert np.array_equal(so.beta_x(), sn.beta_x) and np.array_equal(so.beta_y(), sn.beta_y)


def test_base_sum_is_pairwise_beyond_1024():
    import oracle_np as onp
    rng = np.random.default_rng(2)
    for n in (5, 800, 1024, 1025, 4400):
        c = rng.random(n).astype(np.float32)
        beta = orc.fill_beta(c, 1, 1)          # beta[0, c] = ((Sigma - s)/(2(n-1)) + s/2)^2 exposes Sigma
        want = onp.fill_beta(c, 1, 1)
        assert np.array_equal(beta, want)
