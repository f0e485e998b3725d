"""Pinning the oracle as far as this image allows (CPU tests).

Julia cannot run here, so the third-party arithmetic the reference calls -- ImageTransformations.imresize
(src/Resampler.jl:119,125), DSP.filt / DSP.blackman (src/FrameSynchronisation.jl:63,73, src/Resampler.jl:93),
Base.hypot (src/Demodulation.jl:27), FFTW (src/Autocorrelations.jl:28-29) -- is restated in oracle/ from the
published algorithms.  These tests check each restatement against an INDEPENDENT implementation of the same
published semantics that ships in this image (OpenCV, scipy.ndimage, scipy.signal, numpy in Float64):

  imresize (pixel-centre aligned linear interpolation, no anti-aliasing, replicate at the borders)
      == cv2.resize(INTER_LINEAR) and scipy.ndimage.map_coordinates(order=1) on the same coordinate map
  DSP.filt(h, x)  == scipy.signal.lfilter(h, 1, x) (zero initial state, causal)
  blackman / initLPF == scipy.signal.windows.blackman and numpy.fft in Float64
  Base.hypot      == sqrt(x^2 + y^2) evaluated in Float64 and rounded once
  findmax         == numpy argmax with first-index ties

What stays unverifiable without Julia: the association of the @simd reductions
sum(A; dims=1) and sum(v), and whether DSP.filt fuses its multiply-adds.  They can only move a projection by an
ulp, i.e. only a near-tie of two beta maxima.
"""
import numpy as np
import pytest

import orc

cv2 = pytest.importorskip("cv2")
ndimage = pytest.importorskip("scipy.ndimage")
signal = pytest.importorskip("scipy.signal")


def _coords(n_in, n_out):
    """imresize's map for one dimension, 0-based, Float64: x = sf*(i+1) + 0.5 - 0.5*sf - 1, clamped to the array"""
    sf = n_in / n_out
    x = sf * np.arange(1, n_out + 1, dtype=np.float64) + (0.5 - 0.5 * sf) - 1.0
    return np.clip(x, 0.0, n_in - 1.0)


def _ulps(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))


@pytest.mark.parametrize("n_in,n_out", [(1000, 2898), (3333, 4028), (4028, 3333), (7, 50), (5000, 600), (600, 600)])
def test_imresize_1d_vs_scipy_map_coordinates(n_in, n_out):
    rng = np.random.default_rng(n_in + n_out)
    sig = rng.random(n_in).astype(np.float32) * 3.0
    got = orc.imresize_1d(sig, n_out)
    want = ndimage.map_coordinates(sig.astype(np.float64), [_coords(n_in, n_out)], order=1, mode="nearest").astype(np.float32)
    # both blend in Float64 and round once to Float32; the weights may differ in the last bit of the coordinate
    assert _ulps(got, want).max() <= 1.0
    assert np.mean(got == want) > 0.99


@pytest.mark.parametrize("shape,out", [((1125, 2576), (600, 800)), ((1481, 2720), (600, 800)), ((525, 800), (600, 800)),
                                       ((300, 400), (600, 800)), ((37, 53), (90, 41))])
def test_imresize_2d_vs_opencv_and_scipy(shape, out):
    rng = np.random.default_rng(shape[0])
    img = rng.random(shape).astype(np.float32)
    got = orc.imresize_2d(img, *out)
    # scipy: the same coordinate grid, linear in both axes, Float64 arithmetic
    yy, xx = np.meshgrid(_coords(shape[0], out[0]), _coords(shape[1], out[1]), indexing="ij")
    want = ndimage.map_coordinates(img.astype(np.float64), [yy, xx], order=1, mode="nearest").astype(np.float32)
    assert _ulps(got, want).max() <= 2.0
    # OpenCV: INTER_LINEAR is the same pixel-centre aligned, non-anti-aliased bilinear kernel with replicated borders;
    # it blends in Float32, hence a few ulps
    cv = cv2.resize(img, (out[1], out[0]), interpolation=cv2.INTER_LINEAR)
    assert np.abs(got - cv).max() <= 4e-6 * float(img.max())


def test_sig_to_image_then_downgrade_vs_opencv():
    # the reference's per-frame pipeline sig_to_image |> downgradeImage (src/GUI.jl:168) on a small mode
    rng = np.random.default_rng(3)
    S, y_t, x_t = 33333, 125, 286
    sig = rng.random(S).astype(np.float32)
    full = orc.sig_to_image(sig, y_t, x_t)
    # (OpenCV computes source coordinates in Float32: useless past ~10^4 output pixels, which is why the reference's
    #  Float64 coordinate map matters on 10^6-pixel frames -- the 1-D step is checked against scipy in Float64)
    flat = ndimage.map_coordinates(sig.astype(np.float64), [_coords(S, y_t * x_t)], order=1, mode="nearest")
    assert _ulps(full, flat.astype(np.float32).reshape(y_t, x_t)).max() <= 1.0
    small = orc.downgradeImage(full)
    cv = cv2.resize(full, (800, 600), interpolation=cv2.INTER_LINEAR)
    assert small.shape == (600, 800) and np.abs(small - cv).max() <= 4e-6


def test_filt5_vs_scipy_lfilter():
    rng = np.random.default_rng(9)
    s = orc.SyncXY()
    h = s.h
    for n in (5, 600, 800, 4400):
        x = (rng.random(n) * 600).astype(np.float32)
        got = orc.filt5(h, x).astype(np.float64)
        want = signal.lfilter(h.astype(np.float64), [1.0], x.astype(np.float64))
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    # causal, zero initial state: the first output is h[0]*x[0]
    x = np.zeros(8, np.float32); x[0] = 1.0
    assert np.array_equal(orc.filt5(h, x)[:5], h)


def test_gaussian_taps_are_the_normalised_window():
    s = orc.SyncXY()
    k = np.arange(-2, 3, dtype=np.float64)
    g = np.exp(-2.0 * k * k / 25.0)
    want = (g / g.sum()).astype(np.float32)          # init_gaussian_filter(5), src/FrameSynchronisation.jl:124-129
    assert np.array_equal(s.h, want)
    # scipy's gaussian window with std = 2.5 is the same shape
    assert np.allclose(signal.windows.gaussian(5, 2.5), g, rtol=1e-15)


def test_hypot_is_the_correctly_rounded_float64_result():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    y = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    got = orc.amDemod(x + 1j * y)
    want = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2).astype(np.float32)   # squares exact in Float64
    assert _ulps(got, want).max() <= 1.0
    assert np.mean(got == want) > 0.9999     # double rounding of the Float64 route can differ on exact ties only


def test_upsampler_filter_vs_numpy_float64():
    # initLPF (src/Resampler.jl:83-99) rebuilt with numpy.fft in Float64 and scipy's Blackman window
    for bufferSize, up in ((64, 4), (250, 2), (81, 3)):
        N = bufferSize * up
        r = orc.init_resampler(bufferSize, up)
        H0 = np.zeros(N, np.complex128)
        H0[: int(np.rint(N / up / 2))] = 1.0
        puls = 2 * np.pi * np.arange(N) / N
        e = H0 * np.exp(1j * (-(N - 1) / 2) * puls)
        Hr = np.rint(e.real) + 1j * np.rint(e.imag)                       # round.(complex)
        h = np.fft.ifft(Hr) * signal.windows.blackman(N, sym=True)        # DSP.blackman(N)
        want = np.fft.fft(h) * (-1.0) ** np.arange(N)
        # the oracle follows the reference and runs the ifft in Float32: agreement to single precision
        assert np.abs(r.H - want).max() <= 2e-6 * np.abs(want).max()


def test_resampler_vs_scipy_fft_route():
    # resampler! (src/Resampler.jl:42-60) = zero-stuff, FFT, * H, IFFT, 2*up*real -- redone in Float64
    rng = np.random.default_rng(4)
    for bufferSize, up in ((128, 4), (250, 2)):
        N = bufferSize * up
        r = orc.init_resampler(bufferSize, up)
        x = rng.standard_normal(bufferSize).astype(np.float32)
        out = np.empty(N, np.float32)
        r(out, x)
        z = np.zeros(N)
        z[::up] = x
        want = 2 * up * np.real(np.fft.ifft(np.fft.fft(z) * r.H))
        assert np.abs(out - want).max() <= 2e-5 * np.abs(want).max()


def test_findmax_vs_numpy_first_index():
    rng = np.random.default_rng(6)
    v = rng.integers(0, 50, 5000).astype(np.float32)    # many ties
    assert orc.findmax(v)[1] == int(np.argmax(v)) + 1
    v[1234] = np.nan; v[4000] = np.nan
    assert orc.findmax(v)[1] == 1235                    # first NaN dominates (isless)
    assert orc.findmax(np.array([-0.0, 0.0, 0.0], np.float32))[1] == 2   # isless(-0.0, 0.0)


def test_autocorr_vs_float64_fft():
    # circular autocorrelation (src/Autocorrelations.jl:27-33) of a 3*10^4-point (non power of two) power signal
    rng = np.random.default_rng(8)
    n = 30000
    x = (1.0 + rng.random(n)).astype(np.float32)
    got, _ = orc.calculate_autocorrelation(x, float(n), 0, 0.5)
    X = np.fft.fft(x.astype(np.float64))
    r = np.fft.ifft(X * np.conj(X))
    want = 10 * np.log10(np.abs(r[: n // 2]) ** 2)
    assert np.abs(got - want).max() <= 1e-2
    assert int(np.argmax(got[1:])) == int(np.argmax(want[1:]))
