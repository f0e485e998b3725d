"""oracle/orc.py -- ctypes binding of the C restatement (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED (see tsdr_oracle.h).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
All images cross this boundary in Julia layout (column-major, numpy order="F").
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import build as _build  # noqa: E402

RENDER_H, RENDER_W = 600, 800
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class _Sync(C.Structure):
    _fields_ = [("n_y", C.c_int), ("n_x", C.c_int), ("wmin_y", C.c_int), ("wmax_y", C.c_int),
                ("wmin_x", C.c_int), ("wmax_x", C.c_int), ("h", C.c_float * 5),
                ("beta_x", C.POINTER(C.c_float)), ("beta_y", C.POINTER(C.c_float))]


def _load():
    lib = C.CDLL(_build.build())
    sig = {
        "orc_hypotf": (C.c_float, [C.c_float, C.c_float]),
        "orc_am_demod": (None, [_f32p, _f32p, C.c_size_t]),
        "orc_invert_am_demod": (None, [_f32p, _f32p, C.c_size_t]),
        "orc_fm_demod": (None, [_f32p, _f32p, C.c_size_t]),
        "orc_abs2": (None, [_f32p, _f32p, C.c_size_t]),
        "orc_imresize_1d": (None, [_f32p, C.c_size_t, _f32p, C.c_size_t]),
        "orc_imresize_2d": (None, [_f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int]),
        "orc_sig_to_image": (None, [_f32p, C.c_size_t, C.c_int, C.c_int, _f32p]),
        "orc_downgrade": (None, [_f32p, C.c_int, C.c_int, _f32p]),
        "orc_naive_resampler": (None, [_f32p, _f32p, C.c_size_t, C.c_int]),
        "orc_fft_c2c": (C.c_int, [_f32p, _f32p, C.c_size_t, C.c_int]),
        "orc_autocorr": (C.c_int, [_f32p, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_int,
                                   _f32p, C.POINTER(C.c_size_t)]),
        "orc_zoom_window": (None, [C.c_size_t, C.c_double, C.c_double, C.c_double,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "orc_findmax": (C.c_size_t, [_f32p, C.c_size_t]),
        "orc_sync_create": (C.POINTER(_Sync), [C.c_int, C.c_int]),
        "orc_sync_destroy": (None, [C.POINTER(_Sync)]),
        "orc_proj_cols": (None, [_f32p, C.c_int, C.c_int, _f32p]),
        "orc_proj_rows": (None, [_f32p, C.c_int, C.c_int, _f32p]),
        "orc_filt5": (None, [_f32p, _f32p, _f32p, C.c_int]),
        "orc_fill_beta": (None, [_f32p, _f32p, C.c_int, C.c_int, C.c_int]),
        "orc_argmax_col": (C.c_int, [_f32p, C.c_int, C.c_int]),
        "orc_vsync": (None, [C.POINTER(_Sync), _f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_circshift": (None, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]),
        "orc_ema": (None, [_f32p, _f32p, C.c_size_t, C.c_float]),
        "orc_full_scale": (None, [_f32p, _f32p, C.c_size_t]),
        "orc_round_even": (C.c_int64, [C.c_double]),
        "orc_frame_samples": (C.c_int64, [C.c_double, C.c_double]),
        "orc_chain_buffer": (C.c_int, [_f32p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_double, C.c_float,
                                       C.POINTER(_Sync), _f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
        "orc_num_threads": (C.c_int, []),
        "orc_openmp": (C.c_int, []),
        "orc_upsampler_create": (C.c_void_p, [C.c_size_t, C.c_int]),
        "orc_upsampler_H": (None, [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]),
        "orc_upsampler_apply": (C.c_int, [C.c_void_p, _f32p, _f32p]),
        "orc_upsampler_destroy": (None, [C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def _iq(x):
    """complex64 vector -> interleaved float32 view (ComplexF32 memory layout)."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    return x.view(np.float32), x.size


def _cm(img):
    """2-D array -> flat float32 buffer in column-major (Julia) order."""
    return np.ascontiguousarray(np.asarray(img, dtype=np.float32).T).ravel()


def _from_cm(flat, h, w):
    return flat.reshape(w, h).T  # element (r, c) at r + h*c


def hypot(x, y):
    return np.float32(lib.orc_hypotf(np.float32(x), np.float32(y)))


def amDemod(sig):  # src/Demodulation.jl:26-28
    v, n = _iq(sig)
    out = np.empty(n, np.float32)
    lib.orc_am_demod(v, out, n)
    return out


def invert_amDemod(sig):  # src/Demodulation.jl:31-35
    v, n = _iq(sig)
    out = np.empty(n, np.float32)
    lib.orc_invert_am_demod(v, out, n)
    return out


def fmDemod(sig):  # src/Demodulation.jl:17-23
    v, n = _iq(sig)
    out = np.empty(n, np.float32)
    lib.orc_fm_demod(v, out, n)
    return out


def abs2(sig):  # src/GUI.jl:70
    v, n = _iq(sig)
    out = np.empty(n, np.float32)
    lib.orc_abs2(v, out, n)
    return out


def imresize_1d(sig, n_out):
    sig = np.ascontiguousarray(sig, np.float32)
    out = np.empty(n_out, np.float32)
    lib.orc_imresize_1d(sig, sig.size, out, n_out)
    return out


def imresize_2d(img, h_out, w_out):
    h, w = img.shape
    out = np.empty(h_out * w_out, np.float32)
    lib.orc_imresize_2d(_cm(img), h, w, out, h_out, w_out)
    return _from_cm(out, h_out, w_out)


def sig_to_image(sig, y_t, x_t):  # src/Resampler.jl:117-122 -> (y_t, x_t) array
    sig = np.ascontiguousarray(sig, np.float32)
    out = np.empty(y_t * x_t, np.float32)
    lib.orc_sig_to_image(sig, sig.size, y_t, x_t, out)
    return _from_cm(out, y_t, x_t)


def downgradeImage(img):  # src/Resampler.jl:124-126
    return imresize_2d(img, RENDER_H, RENDER_W)


def naiveResampler(sig, up):  # src/Resampler.jl:103-110
    sig = np.ascontiguousarray(sig, np.float32)
    out = np.empty(sig.size * up, np.float32)
    lib.orc_naive_resampler(out, sig, sig.size, up)
    return out


def init_resampler(bufferSize, upCoeff):  # src/Resampler.jl:26-62 (T = Float32) -> resampler(out, in)
    h = lib.orc_upsampler_create(bufferSize, upCoeff)
    if not h:
        raise ValueError("bad upsampler size")
    N = bufferSize * upCoeff

    def resampler(out, sig):
        sig = np.ascontiguousarray(sig, np.float32)
        assert sig.size == bufferSize, "Size of input should match size used during init"   # :47
        assert out.size == N and out.dtype == np.float32
        if lib.orc_upsampler_apply(h, out, sig):
            raise MemoryError
    Hbuf = np.empty(2 * N, np.float64)
    lib.orc_upsampler_H(h, Hbuf)
    resampler.H = Hbuf.view(np.complex128)
    resampler._handle = h
    return resampler


def fft(x, inverse=False):
    v, n = _iq(x)
    out = np.empty(2 * n, np.float32)
    if lib.orc_fft_c2c(v, out, n, int(inverse)):
        raise MemoryError
    return out.view(np.complex64)


def calculate_autocorrelation(x, Fs, minDelay, maxDelay, scale="log"):  # src/Autocorrelations.jl:23-37
    x = np.ascontiguousarray(x, np.float32)
    imin = 1 + int(lib.orc_round_even(minDelay * Fs))
    imax = int(lib.orc_round_even(maxDelay * Fs))
    out = np.empty(max(imax - imin + 1, 1), np.float32)
    n_out = C.c_size_t(0)
    rc = lib.orc_autocorr(x, x.size, Fs, minDelay, maxDelay, int(scale == "log"), out, C.byref(n_out))
    if rc == -1:
        raise IndexError("BoundsError: signal shorter than indexMax")
    if rc:
        raise ValueError("orc_autocorr rc=%d" % rc)
    lags = np.arange(0, imax - imin + 1, dtype=np.float64) * 1 / Fs
    return out[: n_out.value], lags


def _freq_axis(n, fs):
    return (np.arange(n, dtype=np.float64) / n - 0.5) * fs


def _fftshift(v):  # circshift(v, div(n, 2))
    return np.roll(v, v.shape[0] // 2, axis=0)


def _abs2(z):  # abs2(::ComplexF32) = re*re + im*im in Float32
    return (z.real * z.real + z.imag * z.imag).astype(np.float32)


def getSpectrum(fs, sig, N=None):  # src/GetSpectrum.jl:21-30
    sig = np.asarray(sig, np.complex64)
    N = sig.size if N is None else N
    y = np.float32(10) * np.log10(_abs2(_fftshift(fft(sig[:N])))).astype(np.float32)
    return _freq_axis(N, fs), y


def getWelch(fe, sig, sizeFFT=1024):  # src/GetSpectrum.jl:36-52
    sig = np.asarray(sig, np.complex64)
    S = np.zeros(sizeFFT, np.float32)
    for n in range(sig.size // sizeFFT):
        S += _abs2(fft(sig[n * sizeFFT:(n + 1) * sizeFFT]))     # S .+= abs2.(fft(ss)), segment order
    with np.errstate(divide="ignore"):
        y = np.float32(10) * np.log10(_fftshift(S)).astype(np.float32)
    return _freq_axis(sizeFFT, fe), y


def getWaterfall(fe, sig, sizeFFT=1024):  # src/GetSpectrum.jl:54-66
    sig = np.asarray(sig, np.complex64)
    nbSeg = sig.size // sizeFFT
    sMatrix = np.zeros((sizeFFT, nbSeg), np.float64)
    for iN in range(nbSeg):
        sMatrix[:, iN] = _abs2(_fftshift(fft(sig[iN * sizeFFT:(iN + 1) * sizeFFT])))
    return np.arange(nbSeg, dtype=np.float64) * (sizeFFT / fe), _freq_axis(sizeFFT, fe), sMatrix


def zoom_autocorr(gamma, Fs, rate_min=20, rate_max=100):  # src/Autocorrelations.jl:42-53
    a, b = C.c_int64(0), C.c_int64(0)
    lib.orc_zoom_window(len(gamma), Fs, float(rate_min), float(rate_max), C.byref(a), C.byref(b))
    idx = np.arange(a.value, b.value + 1, dtype=np.float64)
    rates = 1.0 / (idx / Fs)
    return rates, np.asarray(gamma)[a.value - 1: b.value]


def findmax(v):
    v = np.ascontiguousarray(v, np.float32)
    i = lib.orc_findmax(v, v.size)
    return v[i], i + 1  # 1-based like Julia


class SyncXY:  # src/FrameSynchronisation.jl:25-48
    def __init__(self, n_y=RENDER_H, n_x=RENDER_W):
        self._p = lib.orc_sync_create(n_y, n_x)
        s = self._p.contents
        self.n_y, self.n_x = s.n_y, s.n_x
        self.wmin_y, self.wmax_y, self.wmin_x, self.wmax_x = s.wmin_y, s.wmax_y, s.wmin_x, s.wmax_x
        self.h = np.array(list(s.h), np.float32)

    def beta_x(self):
        nw = 1 + self.wmax_x - self.wmin_x
        return np.ctypeslib.as_array(self._p.contents.beta_x, (self.n_x, nw)).T.copy()  # (nw, n_x)

    def beta_y(self):
        nw = 1 + self.wmax_y - self.wmin_y
        return np.ctypeslib.as_array(self._p.contents.beta_y, (self.n_y, nw)).T.copy()

    def __del__(self):
        if getattr(self, "_p", None):
            lib.orc_sync_destroy(self._p)
            self._p = None


def vsync(img, sync):  # src/FrameSynchronisation.jl:56-79 -> (s_y, s_x), 1-based
    sy, sx = C.c_int(0), C.c_int(0)
    lib.orc_vsync(sync._p, _cm(img), C.byref(sy), C.byref(sx))
    return sy.value, sx.value


def proj_cols(img):
    h, w = img.shape
    out = np.empty(w, np.float32)
    lib.orc_proj_cols(_cm(img), h, w, out)
    return out


def proj_rows(img):
    h, w = img.shape
    out = np.empty(h, np.float32)
    lib.orc_proj_rows(_cm(img), h, w, out)
    return out


def filt5(h, x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib.orc_filt5(np.ascontiguousarray(h, np.float32), x, y, x.size)
    return y


def fill_beta(c, wmin, wmax):
    c = np.ascontiguousarray(c, np.float32)
    nw = 1 + wmax - wmin
    beta = np.empty(nw * c.size, np.float32)
    lib.orc_fill_beta(beta, c, c.size, wmin, wmax)
    return beta.reshape(c.size, nw).T  # (nw, n): Julia's beta[cnt, c]


def circshift(img, s_y, s_x):  # src/GUI.jl:172 : circshift(img, (-s_y, -s_x))
    h, w = img.shape
    out = np.empty(h * w, np.float32)
    lib.orc_circshift(_cm(img), out, h, w, s_y, s_x)
    return _from_cm(out, h, w)


def ema(acc, img, alpha):  # src/GUI.jl:175, returns the new accumulator
    a = np.ascontiguousarray(acc, np.float32).ravel().copy()
    lib.orc_ema(a, np.ascontiguousarray(img, np.float32).ravel(), a.size, np.float32(alpha))
    return a.reshape(np.shape(acc))


def fullScale(mat):  # src/ScreenRenderer.jl:35-39
    m = np.ascontiguousarray(mat, np.float32)
    out = np.empty_like(m)
    lib.orc_full_scale(m.ravel(), out.ravel(), m.size)
    return out


def frame_samples(Fs, fv):  # src/GUI.jl:103-109
    return int(lib.orc_frame_samples(Fs, fv))


def chain_buffer(iq, Fs, x_t, y_t, fv, alpha, sync, image_out, publish=True, nthreads=1):
    """coreProcessing loop body for one buffer (src/GUI.jl:163-178).
    image_out: (600, 800) float32 EMA state; returns (new image_out, frames or None, sy, sx)."""
    v, n = _iq(iq)
    S = frame_samples(Fs, fv)
    nb = n // S
    acc = _cm(image_out).copy()
    frames = np.empty(nb * RENDER_H * RENDER_W, np.float32) if publish else None
    sy = np.zeros(max(nb, 1), np.int32)
    sx = np.zeros(max(nb, 1), np.int32)
    got = lib.orc_chain_buffer(v, n, Fs, x_t, y_t, fv, np.float32(alpha), sync._p, acc,
                               frames.ctypes.data if publish else None,
                               sy.ctypes.data, sx.ctypes.data, nthreads)
    assert got == nb
    fr = None
    if publish:
        fr = frames.reshape(nb, RENDER_W, RENDER_H).transpose(0, 2, 1)
    return _from_cm(acc, RENDER_H, RENDER_W), fr, sy[:nb], sx[:nb]


def chain_buffer_fullres(iq, Fs, x_t, y_t, fv, alpha, sync, image_out, do_align=True):
    """the loop body of coreProcessing (src/GUI.jl:163-178) WITHOUT `|> downgradeImage` (full-resolution mode,
    SURVEY 8(f) rank 4): frames, SyncXY (sync = SyncXY(y_t, x_t)), circshift and the EMA all at y_t x x_t.
    Returns (imageOut, [every intermediate imageOut], s_y list, s_x list)."""
    S = frame_samples(Fs, fv)
    env = amDemod(iq)
    acc = np.asarray(image_out, np.float32)
    sy, sx, pub = [], [], []
    for n in range(env.size // S):
        img = sig_to_image(env[n * S:(n + 1) * S], y_t, x_t)
        if do_align:
            t = vsync(img, sync)
            img = circshift(img, t[0], t[1])
            sy.append(t[0]); sx.append(t[1])
        acc = ema(acc, img, alpha)
        pub.append(acc.copy())
    return acc, pub, sy, sx


def investigate_capture(sigRx, Fs, find_closest_configuration, offset=420_000, N=500):
    """the headless recipe of production/investigate_data.jl:37-97,159-206 with the oracle's functions (checker for
    BASELINE configs[0]); find_closest_configuration is the host-side table lookup (src/VideoConfigurations.jl:117-124)"""
    sigId = amDemod(sigRx)
    G, _ = calculate_autocorrelation(sigId, Fs, 0, 1 / 10)
    rates, Gl = zoom_autocorr(G, Fs, 50, 90)
    posMax = findmax(Gl)[1]
    fv = float(np.round(1 / (1 / rates[posMax - 1]), 2))
    _, Gs = zoom_autocorr(G, Fs, fv, fv + 0.3)
    m = findmax(Gs[:N])[1]
    y_t = 1 / (fv * (m / Fs))
    found = find_closest_configuration(y_t, fv)
    name = list(found)[0]
    w, h = found[name].width, found[name].height
    d = frame_samples(Fs, fv)
    img = sig_to_image(sigId[offset: offset + d], h, w)
    so = SyncXY(h, w)
    tup = vsync(img, so)
    idx = int(np.floor((tup[1] * w + tup[0]) / (w * h) / fv * Fs))
    img2 = sig_to_image(sigId[offset + idx: offset + idx + d], h, w)
    return dict(fv=fv, posMax=posMax, m=m, y_t=y_t, name=name, vsync=tup, idx=idx, image=img, image_synced=img2)


def num_threads():
    """threads orc_chain_buffer can use: every core this process may run on when the library has OpenMP
    (launchers such as torchrun export OMP_NUM_THREADS=1, which the explicit nthreads argument overrides)"""
    if not lib.orc_openmp():
        return 1
    try:
        return max(lib.orc_num_threads(), len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(lib.orc_num_threads(), os.cpu_count() or 1)
