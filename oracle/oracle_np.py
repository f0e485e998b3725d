"""oracle/oracle_np.py -- second, independent restatement in numpy (TEST INFRASTRUCTURE ONLY).

Written separately from tsdr_oracle.c so that the two can be diffed against
each other (tests/test_oracle.py); neither is pinned by the reference, whose
tests hold no vector on this path (PARITY UNPINNED, see tsdr_oracle.h).
Citations are relative to /root/reference.  Images are (rows, cols) numpy
arrays; Julia's column-major linear order is arr.T.ravel().
"""
import numpy as np

RENDER = (600, 800)  # src/GUI.jl:10


def round_even(x):  # Base.round: ties to even
    return int(np.rint(np.float64(x)))


def amDemod(sig):  # src/Demodulation.jl:26-28 (correctly rounded hypot through float64)
    z = np.asarray(sig, np.complex64)
    re, im = z.real.astype(np.float64), z.imag.astype(np.float64)
    return np.sqrt(re * re + im * im).astype(np.float32)


def invert_amDemod(sig):  # src/Demodulation.jl:31-35
    d = amDemod(sig)
    d = d / d.max()
    return (np.float32(1) - d).astype(np.float32)


def fmDemod(sig):  # src/Demodulation.jl:17-23
    z = np.asarray(sig, np.complex64)
    out = np.zeros(z.size, np.float32)
    out[1:] = np.angle(z[1:] * np.conj(z[:-1])).astype(np.float32)
    return out


def _coords(n_in, n_out, clamp):
    """ImageTransformations.imresize! index map + Interpolations Linear() position."""
    sf = np.float64(n_in) / np.float64(n_out)
    off = np.float64(0.5) - np.float64(0.5) * sf
    i = np.arange(1, n_out + 1, dtype=np.float64)
    x = sf * i
    x = x + off
    if clamp:
        x = np.clip(x, 1.0, np.float64(n_in))
    f = np.floor(x)
    f = f - (f > n_in - 1)
    return f.astype(np.int64), x - f


def imresize_1d(sig, n_out):
    a = np.asarray(sig, np.float32)
    if a.size == n_out:
        return a.copy()
    f, d = _coords(a.size, n_out, clamp=not (a.size / n_out >= 1))
    a64 = a.astype(np.float64)
    return ((1.0 - d) * a64[f - 1] + d * a64[f]).astype(np.float32)


def imresize_2d(img, h_out, w_out):
    a = np.asarray(img, np.float32)
    h, w = a.shape
    if (h, w) == (h_out, w_out):
        return a.copy()
    clamp = not (h / h_out >= 1 and w / w_out >= 1)
    fy, dy = _coords(h, h_out, clamp)
    fx, dx = _coords(w, w_out, clamp)
    a64 = a.astype(np.float64)
    fy, dy = fy[:, None], dy[:, None]
    fx, dx = fx[None, :], dx[None, :]
    r0 = (1.0 - dx) * a64[fy - 1, fx - 1] + dx * a64[fy - 1, fx]
    r1 = (1.0 - dx) * a64[fy, fx - 1] + dx * a64[fy, fx]
    return ((1.0 - dy) * r0 + dy * r1).astype(np.float32)


def sig_to_image(sig, y_t, x_t):  # src/Resampler.jl:117-122
    return imresize_1d(sig, y_t * x_t).reshape(y_t, x_t)


def downgradeImage(img):  # src/Resampler.jl:124-126
    return imresize_2d(img, *RENDER)


def naiveResampler(sig, up):  # src/Resampler.jl:103-110
    return np.repeat(np.asarray(sig, np.float32), up)


def calculate_autocorrelation(x, Fs, minDelay, maxDelay, scale="log"):  # src/Autocorrelations.jl:23-37
    import scipy.fft as sfft
    x = np.asarray(x, np.float32)
    imin = 1 + round_even(minDelay * Fs)
    imax = round_even(maxDelay * Fs)
    n = min(2 * imax, x.size)
    if imax > n:
        raise IndexError("BoundsError")
    X = sfft.fft(x[:n].astype(np.complex64))
    r = sfft.ifft(X * np.conj(X))
    p = (r.real * r.real + r.imag * r.imag)[imin - 1: imax].astype(np.float32)
    lags = np.arange(0, imax - imin + 1) / Fs
    return ((np.float32(10) * np.log10(p)).astype(np.float32) if scale == "log" else p), lags


def zoom_autocorr(gamma, Fs, rate_min=20, rate_max=100):  # src/Autocorrelations.jl:42-53
    N = len(gamma)
    a = min(round_even(1 / rate_max * Fs), N)
    b = min(round_even(1 / rate_min * Fs), N)
    idx = np.arange(a, b + 1, dtype=np.float64)
    return 1.0 / (idx / Fs), np.asarray(gamma)[a - 1: b]


def findmax(v):
    v = np.asarray(v)
    nan = np.flatnonzero(np.isnan(v))
    i = int(nan[0]) if nan.size else int(np.argmax(v))
    return v[i], i + 1


def gaussian_taps():  # src/FrameSynchronisation.jl:124-129, then new{Float32} (:46)
    k = np.arange(-2, 3, dtype=np.float64)
    h = np.exp(-2 * k * k / 25.0)
    s = 0.0
    for v in h:
        s += v
    return (h / s).astype(np.float32)


def sync_bounds(n_y=600, n_x=800):  # src/FrameSynchronisation.jl:36-41
    return (int(np.ceil(1 / 100 * n_y)), int(np.floor(n_y / 4)),
            int(np.ceil(5 / 100 * n_x)), int(np.floor(n_x / 4)))


def seq_sum(a, axis=0):
    """strictly sequential Float32 sum (np.cumsum never uses pairwise blocking)."""
    return np.cumsum(np.asarray(a, np.float32), axis=axis, dtype=np.float32).take(-1, axis=axis)


def col_sum_blocked(img):
    """sum(image;dims=1) with the oracle's fixed association: bands of 32 rows, folded in order."""
    img = np.asarray(img, np.float32)
    rows_per = 32
    tot = None
    for r0 in range(0, img.shape[0], rows_per):
        part = seq_sum(img[r0:r0 + rows_per], axis=0)
        tot = part if tot is None else (tot + part).astype(np.float32)
    return tot


def filt5(h, x):
    """DSP.filt transposed direct form; each muladd done in float64 then rounded
    (double rounding differs from a true fmaf in ~1e-8 of the cases)."""
    h = np.asarray(h, np.float32).astype(np.float64)
    x = np.asarray(x, np.float32).astype(np.float64)
    n = x.size
    xp = np.concatenate([np.zeros(4), x])
    f32 = lambda v: v.astype(np.float32).astype(np.float64)
    acc = f32(h[4] * xp[0:n])
    acc = f32(xp[1:n + 1] * h[3] + acc)
    acc = f32(xp[2:n + 2] * h[2] + acc)
    acc = f32(xp[3:n + 3] * h[1] + acc)
    acc = f32(xp[4:n + 4] * h[0] + acc)
    return acc.astype(np.float32)


def base_sum(c):
    """Base.sum(::Vector{Float32}): pairwise halving down to runs of fewer than 1025 elements (reduce.jl,
    mapreduce_impl, blksize 1024); each run fixed to a 32-lane SIMD-shaped order (see tsdr_oracle.c:orc_sum_base)"""
    c = np.asarray(c, np.float32)
    if c.size - 1 < 1024:
        lanes = [seq_sum(c[l::32]) for l in range(min(32, c.size))]
        tot = lanes[0]
        for v in lanes[1:]:
            tot = np.float32(tot + v)
        return tot
    mid = (c.size - 1) >> 1          # imid = ifirst + (ilast - ifirst) >> 1, 0-based inclusive
    return np.float32(base_sum(c[: mid + 1]) + base_sum(c[mid + 1:]))


def fill_beta(c, wmin, wmax):  # src/FrameSynchronisation.jl:94-112 -> (nw, n)
    c = np.asarray(c, np.float32)
    n = c.size
    Sigma = base_sum(c)
    ctr = np.arange(n)  # 0-based centres
    acc = np.zeros(n, np.float32)
    for k in range(-(wmin - 1), wmin):
        acc = acc + c[(ctr + k) % n]
    s = np.float32(2) * acc
    beta = np.empty((1 + wmax - wmin, n), np.float32)
    for w in range(wmin, wmax + 1):
        s = s + np.float32(2) * c[(ctr - w) % n]
        s = s + np.float32(2) * c[(ctr + w) % n]
        v = (Sigma - s) / np.float32(2 * (n - w)) + s / np.float32(2 * w)
        beta[w - wmin] = v * v
    return beta


def argmax_col(beta):  # findmax(beta)[2][2] with column-major scan order, 1-based
    _, i = findmax(beta.T.ravel())
    return (i - 1) // beta.shape[0] + 1


class SyncXY:
    def __init__(self, n_y=600, n_x=800):
        self.n_y, self.n_x = n_y, n_x
        self.h = gaussian_taps()
        self.wmin_y, self.wmax_y, self.wmin_x, self.wmax_x = sync_bounds(n_y, n_x)
        self.beta_y = np.zeros((1 + self.wmax_y - self.wmin_y, n_y), np.float32)
        self.beta_x = np.zeros((1 + self.wmax_x - self.wmin_x, n_x), np.float32)


def vsync(img, s):  # src/FrameSynchronisation.jl:56-79
    img = np.asarray(img, np.float32)
    c_v = filt5(s.h, col_sum_blocked(img))   # column sums: 32-row bands, see tsdr_oracle.c:orc_proj_cols
    s.beta_x = fill_beta(c_v, s.wmin_x, s.wmax_x)
    s_y = argmax_col(s.beta_y)               # stale beta_y (previous call)
    c_h = filt5(s.h, seq_sum(img, axis=1))   # row sums, columns added in order
    s.beta_y = fill_beta(c_h, s.wmin_y, s.wmax_y)
    s_x = argmax_col(s.beta_x)
    return s_y, s_x


def circshift(img, s_y, s_x):  # src/GUI.jl:172
    return np.roll(np.asarray(img), (-s_y, -s_x), axis=(0, 1))


def ema(acc, img, alpha):  # src/GUI.jl:175
    a = np.float32(alpha)
    return (a * np.asarray(acc, np.float32) + (np.float32(1) - a) * np.asarray(img, np.float32)).astype(np.float32)


def fullScale(m):  # src/ScreenRenderer.jl:35-39
    m = np.asarray(m, np.float32)
    return (m - m.min()) / (m.max() - m.min())


def frame_samples(Fs, fv):  # src/GUI.jl:103-109
    return round_even(Fs / fv)


def chain_buffer(iq, Fs, x_t, y_t, fv, alpha, sync, image_out):  # src/GUI.jl:163-178
    S = frame_samples(Fs, fv)
    sig_abs = amDemod(iq)
    nb = sig_abs.size // S
    frames, sy, sx = [], [], []
    out = np.asarray(image_out, np.float32).copy()
    for n in range(nb):
        img = downgradeImage(sig_to_image(sig_abs[n * S:(n + 1) * S], y_t, x_t))
        t = vsync(img, sync)
        img = circshift(img, t[0], t[1])
        out = ema(out, img, alpha)
        frames.append(out.copy()); sy.append(t[0]); sx.append(t[1])
    return out, frames, sy, sx
