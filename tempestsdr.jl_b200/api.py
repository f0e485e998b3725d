"""Host-side mirror of the reference's exported DSP functions (src/TempestSDR.jl:21-47).

Same names, argument meaning and error behaviour as the Julia functions; every
compute call goes through the C ABI of libtempest_b200.so (hand-written sm_100a
CUDA).  There is no CPU implementation in this module: without the library or
without a GPU the calls raise.  Matrices are numpy arrays indexed [row, col] and
cross the ABI in Julia's column-major layout.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import TempestError, check
from .dat_files import readComplexBinary, readComplexBinaryRaw, writeComplexBinary  # noqa: F401
from .video_configurations import (VideoMode, allVideoConfigurations, find_closest_configuration,  # noqa: F401
                                   find_configuration, get_refresh_rates, dict2video)

__all__ = [
    "amDemod", "invert_amDemod", "fmDemod", "abs2", "sig_to_image", "downgradeImage", "naiveResampler", "init_resampler",
    "calculate_autocorrelation", "zoom_autocorr", "getSpectrum", "getWelch", "getWaterfall", "findmax", "findmax_device", "findmax_windows_device", "sweep_refresh_hypotheses", "SyncXY", "vsync", "fullScale",
    "VideoMode", "allVideoConfigurations", "find_closest_configuration", "find_configuration",
    "get_refresh_rates", "dict2video", "getImageDuration", "delay2yt", "yt2index", "yt2delay",
    "readComplexBinary", "readComplexBinaryRaw", "writeComplexBinary", "toImage", "investigate_capture", "Chain", "Comm", "comm_available", "AtomicCircularBuffer", "circ_put", "circ_take", "AutocorrPlan", "extract_configuration", "estimate_lines", "search_configuration", "blanking_contrast", "auto_configure", "TempestError", "RENDERING_SIZE",
    "device_count", "set_device",
]

RENDERING_SIZE = (600, 800)  # src/GUI.jl:10


def device_count():
    return _lib.device_count()


def set_device(device):
    """device used by this thread's per-function (host-pointer) calls; handles (Chain, AutocorrPlan) carry their own"""
    check(_lib.load().tsdr_set_device(int(device)))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _iq(sig):
    z = np.ascontiguousarray(sig, dtype=np.complex64)
    return z, z.size


def _demod(fn, sig):
    z, n = _iq(sig)
    out = np.empty(n, np.float32)
    check(fn(_ptr(z), _ptr(out), n))
    return out


def amDemod(sig):
    """abs.(sig) -- src/Demodulation.jl:26-28"""
    return _demod(_lib.load().tsdr_am_demod_f32, sig)


def invert_amDemod(sig):
    """1 .- abs.(sig) ./ maximum(abs.(sig)) -- src/Demodulation.jl:31-35"""
    return _demod(_lib.load().tsdr_invert_am_demod_f32, sig)


def fmDemod(sig):
    """angle(sig[n+1] * conj(sig[n])) -- src/Demodulation.jl:17-23"""
    return _demod(_lib.load().tsdr_fm_demod_f32, sig)


def abs2(sig):
    """abs2.(sig), what extract_configuration feeds the autocorrelation -- src/GUI.jl:70"""
    return _demod(_lib.load().tsdr_abs2_f32, sig)


def sig_to_image(sig, y_t, x_t):
    """imresize(sig, y_t*x_t) |> reshape(x_t, y_t) |> transpose -- src/Resampler.jl:117-122
    Returns a (y_t, x_t) matrix whose row r is scan line r."""
    s = np.ascontiguousarray(sig, dtype=np.float32)
    out = np.empty((int(y_t), int(x_t)), np.float32, order="F")
    check(_lib.load().tsdr_sig_to_image_f32(_ptr(s), s.size, int(y_t), int(x_t), _ptr(out)))
    return out


def downgradeImage(image):
    """imresize(image, (600, 800)) -- src/Resampler.jl:124-126"""
    img = np.asfortranarray(image, dtype=np.float32)
    out = np.empty(RENDERING_SIZE, np.float32, order="F")
    check(_lib.load().tsdr_downgrade_f32(_ptr(img), img.shape[0], img.shape[1], _ptr(out)))
    return out


def naiveResampler(sigOut, sigId, upCoeff):
    """sample-and-hold upsampler writing into sigOut -- src/Resampler.jl:103-110"""
    s = np.ascontiguousarray(sigId, dtype=np.float32)
    if not (isinstance(sigOut, np.ndarray) and sigOut.dtype == np.float32 and sigOut.flags.c_contiguous):
        raise TypeError("sigOut must be a contiguous float32 array")
    if sigOut.size < s.size * int(upCoeff):
        raise IndexError("BoundsError: sigOut shorter than upCoeff*length(sigId)")
    check(_lib.load().tsdr_naive_resampler_f32(_ptr(sigOut), _ptr(s), s.size, int(upCoeff)))
    return None


def init_resampler(T, bufferSize, upCoeff=None):
    """init_resampler(T, bufferSize, upCoeff) or init_resampler(x, upCoeff) -> resampler_(out, in)
    -- src/Resampler.jl:26-68.  Only T = Float32 exists on the GPU; the closure enforces the
    reference's type / size assertions (:44, :47)."""
    if upCoeff is None:  # init_resampler(x::Vector{T}, upCoeff)  (:65-68)
        x, upCoeff = T, bufferSize
        T, bufferSize = np.asarray(x).dtype.type, len(x)
    if np.dtype(T) != np.float32:
        raise TypeError("libtempest_b200 implements init_resampler for Float32 only (got %s)" % np.dtype(T))
    h = C.c_void_p()
    check(_lib.load().tsdr_upsampler_create(int(bufferSize), int(upCoeff), C.byref(h)))
    N = int(bufferSize) * int(upCoeff)

    class _Closure:
        def __init__(self):
            self._h = h
            Hb = np.empty(2 * N, np.float64)
            check(_lib.load().tsdr_upsampler_get_filter(h, _ptr(Hb)))
            self.H = Hb.view(np.complex128)

        def __call__(self, out, sig):
            if not isinstance(out, np.ndarray) or out.dtype != np.float32 or np.asarray(sig).dtype != np.float32:
                raise AssertionError("Type of input should match type used during init (Float32)")   # :44
            s = np.ascontiguousarray(sig)
            if s.size != bufferSize:
                raise AssertionError("Size of input %d should match size used during init %d" % (s.size, bufferSize))  # :47
            check(_lib.load().tsdr_upsampler_apply_f32(self._h, _ptr(out), out.size, _ptr(s), s.size))

        def __del__(self):
            if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "load", None):
                _lib.load().tsdr_upsampler_destroy(self._h)
                self._h = None

    return _Closure()


def _round(x):  # Base.round, ties to even
    return int(np.rint(np.float64(x)))


def calculate_autocorrelation(x, Fs, minDelay, maxDelay, scale="log"):
    """(Gamma, lags) -- src/Autocorrelations.jl:23-37.  scale: "log" (:log) or anything else (abs2)."""
    xs = np.ascontiguousarray(x, dtype=np.float32)
    n_out = C.c_size_t(0)
    lib = _lib.load()
    rc = lib.tsdr_autocorr_out_len(xs.size, float(Fs), float(minDelay), float(maxDelay), C.byref(n_out))
    if rc == -5:
        raise IndexError("BoundsError: " + lib.tsdr_last_error_string().decode())
    check(rc)
    out = np.empty(n_out.value, np.float32)
    check(lib.tsdr_autocorr_f32(_ptr(xs), xs.size, float(Fs), float(minDelay), float(maxDelay),
                                1 if scale in ("log", ":log") else 0, _ptr(out), C.byref(n_out)))
    nbS = _round(maxDelay * Fs) - (1 + _round(minDelay * Fs))
    lags = np.arange(0, nbS + 1, dtype=np.float64) * 1 / Fs
    return out, lags


def zoom_autocorr(Gamma, Fs, rate_min=20, rate_max=100):
    """(rates, Gamma[pos_rate_min:pos_rate_max]) -- src/Autocorrelations.jl:42-53 (host-side slicing).
    Keeps the reference's convention that index k stands for lag k/Fs."""
    N = len(Gamma)
    pos_rate_min = min(_round(1 / rate_max * Fs), N)
    pos_rate_max = min(_round(1 / rate_min * Fs), N)
    if pos_rate_min < 1:   # the reference indexes Gamma[0:...] here: BoundsError
        raise IndexError("BoundsError: zoom_autocorr window starts at index %d (rate_max %g too large for Fs %g)"
                         % (pos_rate_min, rate_max, Fs))
    xAx = np.arange(pos_rate_min, pos_rate_max + 1, dtype=np.float64) / Fs
    return 1.0 / xAx, np.asarray(Gamma)[pos_rate_min - 1: max(pos_rate_max, pos_rate_min - 1)]


def _freq_axis(n, fs):  # collect(((0:N-1)./N .- 0.5)*fs)
    return (np.arange(n, dtype=np.float64) / n - 0.5) * fs


def getSpectrum(fs, sig=None, N=None):
    """(freqAx, y) -- src/GetSpectrum.jl:21-30; getSpectrum(sig) = getSpectrum(1, sig) (:31)"""
    if sig is None:
        fs, sig = 1, fs
    z, n = _iq(sig)
    if N is None:
        N = n
    if N > n:
        raise IndexError("BoundsError: sig[1:%d] of a %d-sample signal" % (N, n))
    y = np.empty(N, np.float32)
    check(_lib.load().tsdr_get_spectrum_f32(_ptr(z), int(N), 1, _ptr(y)))
    return _freq_axis(N, fs), y


def getWelch(fe, sig, sizeFFT=1024):
    """(freqAx, y) -- src/GetSpectrum.jl:36-52"""
    z, n = _iq(sig)
    y = np.empty(sizeFFT, np.float32)
    check(_lib.load().tsdr_get_welch_f32(_ptr(z), n, int(sizeFFT), _ptr(y)))
    return _freq_axis(sizeFFT, fe), y


def getWaterfall(fe, sig=None, sizeFFT=1024):
    """(tAx, fAx, sMatrix[sizeFFT, nbSeg]) -- src/GetSpectrum.jl:54-67; the matrix is Float64 like the reference's"""
    if sig is None:
        fe, sig = 1, fe
    z, n = _iq(sig)
    nbSeg = n // sizeFFT
    s = np.empty((nbSeg, sizeFFT), np.float32)   # column-major sizeFFT x nbSeg
    check(_lib.load().tsdr_get_waterfall_f32(_ptr(z), n, int(sizeFFT), _ptr(s)))
    tAx = np.arange(nbSeg, dtype=np.float64) * (sizeFFT / fe)
    return tAx, _freq_axis(sizeFFT, fe), s.T.astype(np.float64)


def findmax(v):
    """(value, 1-based index) of the first maximum, NaN dominating (Base.findmax), on the GPU."""
    a = np.ascontiguousarray(v, dtype=np.float32)
    val, idx = C.c_float(0), C.c_size_t(0)
    check(_lib.load().tsdr_findmax_f32(_ptr(a), a.size, C.byref(val), C.byref(idx)))
    return np.float32(val.value), idx.value


def findmax_device(ptr, n, stream=None):
    """findmax over n floats at device pointer `ptr` -> (value, 1-based index)"""
    val, idx = C.c_float(0), C.c_size_t(0)
    check(_lib.load().tsdr_findmax_dev_f32(C.c_void_p(ptr), int(n), C.byref(val), C.byref(idx),
                                           C.c_void_p(stream) if stream else None))
    return np.float32(val.value), idx.value


def findmax_windows_device(ptr, starts0, lengths, stream=None):
    """findmax of several windows of one device vector at once -> [(value, 1-based index inside the window)]"""
    n = len(starts0)
    lo = (C.c_size_t * n)(*[int(v) for v in starts0])
    ln = (C.c_size_t * n)(*[int(v) for v in lengths])
    vals = (C.c_float * n)()
    idx = (C.c_size_t * n)()
    check(_lib.load().tsdr_findmax_windows_dev_f32(C.c_void_p(ptr), n, lo, ln, vals, idx, C.c_void_p(stream) if stream else None))
    return [(np.float32(vals[i]), int(idx[i])) for i in range(n)]


def fullScale(mat):
    """(mat .- min) / (max - min) -- src/ScreenRenderer.jl:35-39"""
    m = np.ascontiguousarray(mat, dtype=np.float32)
    out = np.empty_like(m)
    check(_lib.load().tsdr_full_scale_f32(_ptr(m), _ptr(out), m.size))
    return out


class SyncXY:
    """SyncXY(image) -- src/FrameSynchronisation.jl:25-48.  Owns beta_x / beta_y on the device."""

    def __init__(self, image=None):
        shape = RENDERING_SIZE if image is None else np.shape(image)
        h = C.c_void_p()
        check(_lib.load().tsdr_sync_create(int(shape[0]), int(shape[1]), C.byref(h)))
        self._h = h
        self.n_y, self.n_x = int(shape[0]), int(shape[1])
        b = [C.c_int(0) for _ in range(4)]
        check(_lib.load().tsdr_sync_bounds(self._h, *[C.byref(v) for v in b]))
        self.wmin_y, self.wmax_y, self.wmin_x, self.wmax_x = [v.value for v in b]

    def _betas(self):
        bx = np.empty((1 + self.wmax_x - self.wmin_x, self.n_x), np.float32, order="F")
        by = np.empty((1 + self.wmax_y - self.wmin_y, self.n_y), np.float32, order="F")
        check(_lib.load().tsdr_sync_get_beta(self._h, _ptr(bx), _ptr(by)))
        return bx, by

    @property
    def beta_x(self):
        return self._betas()[0]

    @property
    def beta_y(self):
        return self._betas()[1]

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tsdr_sync_destroy(self._h)
            self._h = None

    __del__ = close


def vsync(image, sync):
    """(s_y, s_x), 1-based -- src/FrameSynchronisation.jl:56-79 (s_y comes from the previous call's beta_y)."""
    img = np.asfortranarray(image, dtype=np.float32)
    if img.shape != (sync.n_y, sync.n_x):
        raise ValueError("image shape %s does not match SyncXY %s" % (img.shape, (sync.n_y, sync.n_x)))
    sy, sx = C.c_int(0), C.c_int(0)
    check(_lib.load().tsdr_vsync_f32(sync._h, _ptr(img), C.byref(sy), C.byref(sx)))
    return sy.value, sx.value


def getImageDuration(theConfig, Fs):
    """round(Fs / refresh) -- src/GUI.jl:103-109"""
    return _round(Fs / theConfig.refresh)


def delay2yt(tau_or_index, *args):  # src/GUI.jl:238-243
    if len(args) == 1:
        return float(np.rint(1 / (args[0] * tau_or_index)))
    Fs, fv = args
    return float(np.rint(1 / (fv * (tau_or_index / Fs))))


def yt2index(yt, Fs, fv):  # src/GUI.jl:247-249
    return float(np.rint(Fs / (fv * yt)))


def yt2delay(yt, fv):  # src/GUI.jl:250-252
    return 1 / (fv * yt)


class Chain:
    """The loop body of coreProcessing (src/GUI.jl:163-178) as one device-resident object:
    amDemod -> sig_to_image -> downgradeImage -> vsync -> circshift -> EMA for every
    frame of a recv! buffer.

        ch = Chain(Fs, VideoMode(2576, 1125, 60), alpha=0.1, max_samples=10_000_000)
        n_frames = ch.push(iq)            # numpy complex64 (host) ...
        n_frames = ch.push_device(ptr, n) # ... or a device pointer to interleaved float32
        img = ch.image()                  # imageOut, (600, 800)
    """

    def __init__(self, Fs, config, alpha=0.1, max_samples=None, device=0, publish_all=False, do_align=True,
                 sum_mode=False, stream=None, overlap=True, full_res=False):
        if max_samples is None:
            max_samples = getImageDuration(config, Fs)
        flags = (_lib.TSDR_CHAIN_PUBLISH_ALL if publish_all else 0) | (0 if do_align else _lib.TSDR_CHAIN_NO_ALIGN) \
            | (_lib.TSDR_CHAIN_SUM if sum_mode else 0) | (0 if overlap else _lib.TSDR_CHAIN_NO_OVERLAP) \
            | (_lib.TSDR_CHAIN_FULLRES if full_res else 0)
        h = C.c_void_p()
        check(_lib.load().tsdr_chain_create(C.byref(h), int(device), float(Fs), int(config.width), int(config.height),
                                            float(config.refresh), float(alpha), int(max_samples), flags,
                                            C.c_void_p(stream) if stream else None))
        self._h = h
        self.Fs, self.config, self.device = float(Fs), config, int(device)
        self.S = getImageDuration(config, Fs)
        self.max_samples = int(max_samples)
        self.publish_all = publish_all

    def configure(self, Fs, config):  # FLAG_CONFIG_UPDATE, src/GUI.jl:151-158
        check(_lib.load().tsdr_chain_configure(self._h, float(Fs), int(config.width), int(config.height),
                                               float(config.refresh)))
        self.Fs, self.config, self.S = float(Fs), config, getImageDuration(config, Fs)

    def set_alpha(self, alpha):  # OBS_alpha, src/GUI.jl:160
        check(_lib.load().tsdr_chain_set_alpha(self._h, float(alpha)))

    def reset(self):
        check(_lib.load().tsdr_chain_reset(self._h))

    def push(self, iq):
        """one recv! buffer from host memory (numpy complex64, or a pinned torch tensor's numpy view)"""
        z, n = _iq(iq)
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_host(self._h, _ptr(z), n, C.byref(nf)))
        # the H2D copy is asynchronous and double buffered: keep the last two host buffers alive
        self._keep = [z] + list(getattr(self, "_keep", []))[:1]
        return nf.value

    def push_i16(self, iq16):
        """one buffer of interleaved Int16 (re, im) samples as a `:short` .dat file stores them
        (src/DatBinaryFiles.jl:47-49): numpy int16 of shape (n, 2) or (2n,)"""
        z = np.ascontiguousarray(iq16, dtype=np.int16).reshape(-1)
        if z.size % 2:
            raise ValueError("Int16 IQ buffer needs an even number of values (re, im pairs)")
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_host_i16(self._h, _ptr(z), z.size // 2, C.byref(nf)))
        self._keep = [z] + list(getattr(self, "_keep", []))[:1]
        return nf.value

    def prime(self, iq):
        """advance only the SyncXY state with these frames (halo frame of a sharded integration)"""
        z, n = _iq(iq)
        check(_lib.load().tsdr_chain_prime_host(self._h, _ptr(z), n))
        self._keep = [z] + list(getattr(self, "_keep", []))[:1]

    def prime_device(self, ptr, n):
        """prime() for samples already in device memory"""
        check(_lib.load().tsdr_chain_prime_device(self._h, C.c_void_p(ptr), int(n)))

    def push_host_ptr(self, ptr, n):
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_host(self._h, C.c_void_p(ptr), int(n), C.byref(nf)))
        return nf.value

    def push_deliver_ptr(self, iq_ptr, n, image_out_ptr):
        """push a (pinned) host buffer and deliver this buffer's imageOut asynchronously to a (pinned) host image"""
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_host_deliver(self._h, C.c_void_p(iq_ptr), int(n), C.byref(nf), C.c_void_p(image_out_ptr)))
        return nf.value

    def push_i16_deliver_ptr(self, iq16_ptr, n, image_out_ptr):
        """push_deliver_ptr for a (pinned) buffer of Int16 (re, im) pairs"""
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_host_i16_deliver(self._h, C.c_void_p(iq16_ptr), int(n), C.byref(nf), C.c_void_p(image_out_ptr)))
        return nf.value

    def push_ring(self, ring, timeout_ms=-1):
        """recv! + loop body: take the ring's next buffer and push it straight from its page-locked slot"""
        nf = C.c_int(0)
        fmt = 0 if ring.dtype == np.complex64 else 1
        check(_lib.load().tsdr_chain_push_ring(self._h, ring._h, fmt, int(timeout_ms), C.byref(nf)))
        return nf.value

    def wait_delivery(self, age=0):
        check(_lib.load().tsdr_chain_wait_delivery(self._h, int(age)))

    def push_device(self, ptr, n):
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_device(self._h, C.c_void_p(ptr), int(n), C.byref(nf)))
        return nf.value

    def push_device_i16(self, ptr, n):
        """Int16 (re, im) pairs already in device memory: 16-byte aligned, allocation rounded up to 16 bytes"""
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_push_device_i16(self._h, C.c_void_p(ptr), int(n), C.byref(nf)))
        return nf.value

    def sync(self):
        check(_lib.load().tsdr_chain_sync(self._h))

    def flush(self):
        """primary stream waits (device side) for the chain's auxiliary stream"""
        check(_lib.load().tsdr_chain_flush(self._h))

    def image_size(self):
        """(rows, cols) of imageOut: RENDERING_SIZE, or (y_t, x_t) for a full_res chain"""
        a, b = C.c_int(0), C.c_int(0)
        check(_lib.load().tsdr_chain_image_size(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def image(self):
        out = np.empty(self.image_size(), np.float32, order="F")
        check(_lib.load().tsdr_chain_read_image(self._h, _ptr(out)))
        return out

    def image_downgraded(self):
        """downgradeImage(imageOut): the 600 x 800 view of a full_res chain's accumulator"""
        out = np.empty(RENDERING_SIZE, np.float32, order="F")
        check(_lib.load().tsdr_chain_read_image_downgraded(self._h, _ptr(out)))
        return out

    def offsets(self, max_frames=65536):
        sy = np.zeros(max_frames, np.int32)
        sx = np.zeros(max_frames, np.int32)
        n = C.c_int(0)
        check(_lib.load().tsdr_chain_read_offsets(self._h, _ptr(sy), _ptr(sx), max_frames, C.byref(n)))
        k = min(n.value, max_frames)
        return sy[:k].copy(), sx[:k].copy()

    def scores(self, max_frames=65536):
        """per frame of the last buffer: (max beta_x, max beta_y, Sigma_x, Sigma_y) -- the maxima of the two sync tables
        (src/FrameSynchronisation.jl:66,76) and the sums of the filtered projections (:96)"""
        bx, by = np.zeros(max_frames, np.float32), np.zeros(max_frames, np.float32)
        sx, sy = np.zeros(max_frames, np.float32), np.zeros(max_frames, np.float32)
        n = C.c_int(0)
        check(_lib.load().tsdr_chain_read_scores(self._h, _ptr(bx), _ptr(by), _ptr(sx), _ptr(sy), max_frames, C.byref(n)))
        k = min(n.value, max_frames)
        return bx[:k].copy(), by[:k].copy(), sx[:k].copy(), sy[:k].copy()

    def published(self, max_frames=None):
        """every intermediate imageOut of the last buffer (non_blocking_put!, src/GUI.jl:177)"""
        if max_frames is None:
            max_frames = self.max_samples // self.S
        h, w = self.image_size()
        buf = np.empty((max_frames, w, h), np.float32)
        n = C.c_int(0)
        check(_lib.load().tsdr_chain_read_published(self._h, _ptr(buf), max_frames, C.byref(n)))
        k = min(n.value, max_frames)
        return buf[:k].transpose(0, 2, 1)  # each frame column-major -> [row, col] view

    def integrate_device(self, halo_ptr, halo_samples, buf_ptrs, buf_samples, comm=None, weight=1.0):
        """one rank's share of a sharded integration in ONE call: reset, prime with the halo frame (0: none), push the
        device buffers in order, combine over `comm` (None: scale only).  Returns the frames pushed."""
        n = len(buf_ptrs)
        ptrs = (C.c_void_p * max(n, 1))(*[int(p) for p in buf_ptrs])
        cnt = (C.c_size_t * max(n, 1))(*[int(v) for v in buf_samples])
        nf = C.c_int(0)
        check(_lib.load().tsdr_chain_integrate_device(self._h, C.c_void_p(halo_ptr) if halo_ptr else None, int(halo_samples),
                                                      ptrs, cnt, n, comm._h if comm is not None else None, float(weight),
                                                      C.byref(nf)))
        return nf.value

    def accumulator_ptr(self):
        p, n = C.c_void_p(), C.c_size_t(0)
        check(_lib.load().tsdr_chain_accumulator(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def scale_accumulator(self, factor):
        check(_lib.load().tsdr_chain_scale_accumulator(self._h, float(factor)))

    def stream(self):
        p = C.c_void_p()
        check(_lib.load().tsdr_chain_stream(self._h, C.byref(p)))
        return p.value or 0

    def launch_count(self):
        v = C.c_uint64(0)
        check(_lib.load().tsdr_chain_launch_count(self._h, C.byref(v)))
        return v.value

    def set_profiling(self, enable=True):
        check(_lib.load().tsdr_chain_set_profiling(self._h, 1 if enable else 0))

    def kernel_times(self):
        """(ms per stage [render, project+sync, accumulate], profiled pushes) since the last call"""
        ms = (C.c_float * 3)()
        n = C.c_uint64(0)
        check(_lib.load().tsdr_chain_kernel_times(self._h, ms, C.byref(n)))
        return [float(v) for v in ms], n.value

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tsdr_chain_destroy(self._h)
            self._h = None

    __del__ = close


def comm_available():
    """NCCL version the library could bind at run time (dlopen), or 0"""
    v = C.c_int(0)
    rc = _lib.load().tsdr_comm_available(C.byref(v))
    return v.value if rc == 0 else 0


class Comm:
    """One NCCL communicator rank behind the C ABI (tsdr_comm_*): combines the partial imageOut accumulators of a
    sharded integration (src/GUI.jl:175 is linear in the frames).  One process per GPU:

        uid = Comm.unique_id() on rank 0  ->  distribute the 128 bytes  ->  Comm(uid, world, rank, device)
    """

    ID_BYTES = 128

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * Comm.ID_BYTES)()
        check(_lib.load().tsdr_comm_get_unique_id(buf))
        return bytes(buf)

    def __init__(self, uid, nranks, rank, device=0):
        if len(uid) != Comm.ID_BYTES:
            raise ValueError("unique id must be %d bytes" % Comm.ID_BYTES)
        h = C.c_void_p()
        buf = (C.c_ubyte * Comm.ID_BYTES).from_buffer_copy(uid)
        check(_lib.load().tsdr_comm_init_rank(C.byref(h), int(device), int(nranks), int(rank), buf))
        self._h, self.nranks, self.rank, self.device = h, int(nranks), int(rank), int(device)

    @classmethod
    def from_torch_distributed(cls, device, group=None):
        """bootstrap over an initialised torch.distributed group (plumbing only: 128 bytes, once)"""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(box[0], world, rank, device)

    def allreduce_chain(self, chain, weight=1.0):
        """imageOut <- sum over ranks of weight_rank * imageOut_rank, asynchronous on the chain's stream"""
        check(_lib.load().tsdr_chain_allreduce(chain._h, self._h, float(weight)))

    def allreduce(self, ptr, n_floats, weight=1.0, stream=None):
        check(_lib.load().tsdr_comm_allreduce_f32(self._h, C.c_void_p(ptr), int(n_floats), float(weight),
                                                  C.c_void_p(stream) if stream else None))

    def allgather(self, send_ptr, recv_ptr, bytes_per_rank, stream=None):
        check(_lib.load().tsdr_comm_allgather(self._h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr), int(bytes_per_rank),
                                              C.c_void_p(stream) if stream else None))

    def collectives(self):
        v = C.c_uint64(0)
        check(_lib.load().tsdr_comm_info(self._h, None, None, None, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tsdr_comm_destroy(self._h)
            self._h = None

    __del__ = close


class AtomicCircularBuffer:
    """AtomicCircularBuffer{T}(nEch, depth) (src/AtomicAbstractSDRs.jl:67-79) in page-locked host memory.

    dtype: numpy complex64 (one recv! buffer of nEch ComplexF32 samples per slot) or int16 (nEch (re, im) Int16
    pairs per slot).  pinned=False allocates ordinary memory (hosts without a GPU)."""

    def __init__(self, nEch, depth, dtype=np.complex64, pinned=True):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.complex64), np.dtype(np.int16)):
            raise TypeError("ring slots hold complex64 or int16 (re, im) samples")
        self.nEch, self.depth = int(nEch), int(depth)
        self.sample_bytes = 8 if self.dtype == np.complex64 else 4
        h = C.c_void_p()
        check(_lib.load().tsdr_ring_create(C.byref(h), self.nEch * self.sample_bytes, self.depth, 1 if pinned else 0))
        self._h = h

    def _flat(self, data):
        z = np.ascontiguousarray(data, dtype=self.dtype).reshape(-1)
        if z.nbytes != self.nEch * self.sample_bytes:  # the reference asserts equal lengths (:113)
            raise ValueError("buffer of %d bytes does not match the ring's slot of %d samples" % (z.nbytes, self.nEch))
        return z

    def put(self, data):
        """circ_put! (:159-170): never waits for the consumer"""
        z = self._flat(data)
        check(_lib.load().tsdr_ring_put(self._h, _ptr(z), z.nbytes))

    def take(self, out=None, timeout_ms=-1):
        """circ_take! (:176-189): waits for a new buffer, copies it out"""
        if out is None:
            out = np.empty(self.nEch if self.dtype == np.complex64 else 2 * self.nEch, self.dtype)
        if not (isinstance(out, np.ndarray) and out.flags.c_contiguous and out.dtype == self.dtype
                and out.nbytes == self.nEch * self.sample_bytes):
            raise ValueError("out must be a contiguous %s array of one slot" % self.dtype)
        check(_lib.load().tsdr_ring_take(self._h, _ptr(out), out.nbytes, int(timeout_ms)))
        return out

    def stats(self):
        n = C.c_int(0)
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(_lib.load().tsdr_ring_stats(self._h, C.byref(n), C.byref(a), C.byref(b), C.byref(c)))
        return {"available": n.value, "produced": a.value, "consumed": b.value, "overwritten": c.value}

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tsdr_ring_destroy(self._h)
            self._h = None

    __del__ = close


def circ_put(circ_buff, data):
    circ_buff.put(data)


def circ_take(buffer, circ_buff, timeout_ms=-1):
    return circ_buff.take(buffer, timeout_ms)


class AutocorrPlan:
    """Device-resident calculate_autocorrelation for a fixed length n (the measured M2 path)."""

    def __init__(self, n, device=0, stream=None):
        h = C.c_void_p()
        check(_lib.load().tsdr_autocorr_plan_create(C.byref(h), int(device), int(n),
                                                    C.c_void_p(stream) if stream else None))
        self._h, self.n = h, int(n)

    def exec(self, x_dev_ptr, index_min, index_max, out_dev_ptr, log_scale=True):
        check(_lib.load().tsdr_autocorr_plan_exec(self._h, C.c_void_p(x_dev_ptr), int(index_min), int(index_max),
                                                  1 if log_scale else 0, C.c_void_p(out_dev_ptr)))

    def launch_count(self):
        v = C.c_uint64(0)
        check(_lib.load().tsdr_autocorr_plan_launch_count(self._h, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "load", None):
            _lib.load().tsdr_autocorr_plan_destroy(self._h)
            self._h = None

    __del__ = close


def extract_configuration(sig_corr, Fs, delayRate=1 / 10, rate_min=50, rate_max=90):
    """Numerical part of extract_configuration (src/GUI.jl:49-88): sig_corr is the
    abs2 power signal the reference assembles from nbBuffer recv! calls (:67-71).
    Returns (rates_refresh, Gamma_refresh, fv)."""
    Gamma, _ = calculate_autocorrelation(sig_corr, Fs, 0, delayRate)
    rates_refresh, Gamma_refresh = zoom_autocorr(Gamma, Fs, rate_min=rate_min, rate_max=rate_max)
    _, posMax = findmax(Gamma_refresh)
    posMax_time = 1 / rates_refresh[posMax - 1]
    fv = 1 / posMax_time
    return rates_refresh, Gamma_refresh, fv


def sweep_refresh_hypotheses(gamma_ptr, n_gamma, Fs, hypotheses=None, half_width_hz=0.5, stream=None, rank=0, world=1):
    """BASELINE cfg 4: score every refresh-rate hypothesis of allVideoConfigurations against a DEVICE-resident
    Gamma (output of AutocorrPlan.exec with index_min = 1).  For refresh rate r the window is the
    zoom_autocorr window [r - hw, r + hw] (same index convention as src/Autocorrelations.jl:42-53); the score is the
    window's first maximum (findmax on the GPU).  Hypotheses are sharded round-robin over `world` ranks (no exchange
    needed on the data path; gather the small result lists on the host).  Returns [(rate, score_dB, fv_hat, lag_index)]."""
    if hypotheses is None:
        hypotheses = sorted(get_refresh_rates(allVideoConfigurations))
    mine = []
    for i, r in enumerate(hypotheses):
        if i % world != rank:
            continue
        lo = min(_round(1 / (r + half_width_hz) * Fs), n_gamma)
        hi = min(_round(1 / (r - half_width_hz) * Fs), n_gamma)
        if hi < lo or lo < 1:
            continue
        mine.append((r, lo, hi))
    if not mine:
        return []
    # every window in one call: two launches and one synchronise instead of one of each per hypothesis
    found = findmax_windows_device(gamma_ptr, [lo - 1 for _, lo, _ in mine], [hi - lo + 1 for _, lo, hi in mine], stream)
    out = []
    for (r, lo, _), (val, idx) in zip(mine, found):
        k = lo + idx - 1                      # 1-based index into Gamma; the reference reads it as lag k/Fs
        out.append((float(r), float(val), 1.0 / (k / Fs), int(k)))
    return out


def blanking_contrast(beta_max, sigma, n):
    """beta = ((Sigma - S_w)/(2(n-w)) + S_w/(2w))^2 at its maximum, relative to the value the same expression takes for a
    flat projection (mean^2): 1 for a featureless image, larger the more a blanking band stands out"""
    mean = np.asarray(sigma, np.float64) / n
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.asarray(beta_max, np.float64) / (mean * mean)


def search_configuration(iq, Fs, candidates, frames=3, y_nudges=(0,), device=0, rank=0, world=1):
    """render-and-score configuration search, replacing the click loop of the GUI (the list clicks of
    src/GUI.jl:450-459 and the +-1 line nudges of :526-537).  Every hypothesis -- a VideoMode of `candidates`, its height
    moved by each of `y_nudges` -- renders the first `frames` frames of the capture through the chain and is scored by the
    blanking contrast of its sync tables: a raster with the right line count keeps the horizontal blanking bar vertical,
    so the column projection has a sharp band and max(beta_x) stands far above its flat-image value; one line off and the
    bar shears across the whole width.  Hypotheses are sharded round-robin over `world` ranks (no exchange on the data
    path).  Returns [(score, VideoMode, contrast_x, contrast_y)] sorted best first.  There is no reference function to
    compare with: tests check that synthetic captures of known modes are recovered."""
    z, n = _iq(iq)
    hyps = []
    for cfg in candidates:
        for dy in y_nudges:
            hyps.append(VideoMode(cfg.width, cfg.height + dy, cfg.refresh))
    out = []
    for i, cfg in enumerate(hyps):
        if i % world != rank:
            continue
        S = getImageDuration(cfg, Fs)
        need = frames * S
        if need > n or cfg.height < 2:
            continue
        ch = Chain(Fs, cfg, alpha=0.0, max_samples=need, device=device)
        try:
            ch.push(z[:need])
            bx, by, sx, sy = ch.scores()
        finally:
            ch.close()
        cx = float(np.median(blanking_contrast(bx, sx, RENDERING_SIZE[1])))
        cy = float(np.median(blanking_contrast(by, sy, RENDERING_SIZE[0])))
        out.append((cx * cy, cfg, cx, cy))
    out.sort(key=lambda t: -t[0])
    return out


def auto_configure(iq, Fs, frames=3, nudge=3, refresh_tol=1.0, delayRate=1 / 10, rate_min=50, rate_max=90, device=0,
                   rank=0, world=1):
    """The whole configuration workflow of the GUI without the clicks: the refresh rate from the autocorrelation peak
    (extract_configuration, src/GUI.jl:49-88), the line count from the line-lag peak (production/investigate_data.jl:69-82)
    and the closest table entry (find_closest_configuration) as the reference computes them; then search_configuration
    at the measured refresh rate over that estimate AND over every table entry within refresh_tol Hz (the list the GUI
    user clicks through, src/GUI.jl:450-459), each with its line count moved by -nudge..nudge lines (the manual +-1
    clicks of :526-537) -- the line-lag heuristic alone locks onto a sub-multiple on some captures.
    Returns (best VideoMode, fv, y_t estimate, table entry name, ranking)."""
    z, n = _iq(iq)
    power = abs2(z)
    _, _, fv = extract_configuration(power, Fs, delayRate=delayRate, rate_min=rate_min, rate_max=rate_max)
    Gamma, _ = calculate_autocorrelation(power, Fs, 0, delayRate)
    y_hat = estimate_lines(Gamma, Fs, fv)
    name = list(find_closest_configuration(y_hat, fv))[0]
    entry = allVideoConfigurations[name]
    y0 = _round(y_hat)
    bases = [(entry.width, y0)] + [(c.width, c.height) for c in allVideoConfigurations.values() if abs(c.refresh - fv) < refresh_tol]
    seen, cands = set(), []
    for w, h in bases:
        for d in range(-nudge, nudge + 1):
            if h + d >= 2 and (w, h + d) not in seen:
                seen.add((w, h + d))
                cands.append(VideoMode(w, h + d, float(fv)))
    ranking = search_configuration(z, Fs, cands, frames=frames, device=device, rank=rank, world=world)
    best = ranking[0][1] if ranking else VideoMode(entry.width, y0, float(fv))
    return best, float(fv), float(y_hat), name, ranking


def toImage(sigId, offset, Fs, finalConfig):
    """toImage(sigId, offset, Fs, finalConfig) of production/investigate_data.jl:159-169: the frame starting `offset`
    samples into the (demodulated) capture, as a height x width image -- the same arithmetic as sig_to_image"""
    d = getImageDuration(finalConfig, Fs)
    s = np.asarray(sigId)[offset: offset + d]
    if s.size < d:
        raise IndexError("BoundsError: frame of %d samples at offset %d exceeds the capture" % (d, offset))
    return sig_to_image(s, int(finalConfig.height), int(finalConfig.width))


def investigate_capture(sigRx, Fs, offset=420_000, rate_min=50, rate_max=90, N=500):
    """The headless replay recipe of production/investigate_data.jl (BASELINE configs[0]) on the GPU, step by step with
    the reference's own function names: amDemod (:37), calculate_autocorrelation(sigId, Fs, 0, 1/10) (:52), refresh peak
    (:55-62, fv rounded to 2 digits), line peak (:69-82), find_closest_configuration (:92), toImage (:194), SyncXY / vsync
    on the FULL-SIZE frame (:196-197), sample-offset correction (:200-201) and the re-rendered frame (:206).
    Returns a dict with every intermediate the recipe names."""
    sigId = amDemod(sigRx)
    Gamma, _ = calculate_autocorrelation(sigId, Fs, 0, 1 / 10)
    rates_large, Gamma_large = zoom_autocorr(Gamma, Fs, rate_min=rate_min, rate_max=rate_max)
    _, posMax = findmax(Gamma_large)
    posMax_time = 1 / rates_large[posMax - 1]
    fv = float(np.round(1 / posMax_time, 2))           # round(1/posMax_time; digits=2)
    _, Gamma_short = zoom_autocorr(Gamma, Fs, rate_min=fv, rate_max=fv + 0.3)
    m = findmax(Gamma_short[:N])[1]
    tau = m / Fs
    y_t = 1 / (fv * tau)
    found = find_closest_configuration(y_t, fv)
    name = list(found)[0]                                # first(find_closest_configuration(y_t, fv))
    est = found[name]
    finalConfig = VideoMode(est.width, est.height, fv)
    anImage = toImage(sigId, offset, Fs, finalConfig)
    sync = SyncXY(anImage)
    tup = vsync(anImage, sync)
    sync.close()
    tau_px = tup[1] * finalConfig.width + tup[0]         # tup[2] * width + tup[1]
    idx = int(np.floor(tau_px / (finalConfig.width * finalConfig.height) / fv * Fs))
    d = getImageDuration(finalConfig, Fs)
    anImage2 = sig_to_image(sigId[offset + idx: offset + idx + d], int(finalConfig.height), int(finalConfig.width))
    return {"fv": fv, "posMax": int(posMax), "m": int(m), "y_t": float(y_t), "name": name, "config": finalConfig,
            "vsync": tuple(tup), "idx": idx, "image": anImage, "image_synced": anImage2}


def estimate_lines(Gamma, Fs, fv, N=500):
    """Headless line-count pick of production/investigate_data.jl:69-82:
    zoom to [fv, fv+0.3] Hz, first N lags, findmax -> y_t = 1/(fv * m/Fs)."""
    _, Gamma_short = zoom_autocorr(Gamma, Fs, rate_min=fv, rate_max=fv + 0.3)
    Gamma_short = Gamma_short[:N]
    m = findmax(Gamma_short)[1]
    tau = m / Fs
    return 1 / (fv * tau)
