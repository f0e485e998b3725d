"""Seeded synthetic TEMPEST captures.

A test-card raster of the TOTAL video mode (active area + blanking, the
convention of VideoMode in src/VideoConfigurations.jl:5-9) is scanned at the
pixel clock fv*x_t*y_t, linearly resampled to the SDR rate Fs with a frame
period of exactly Fs/fv samples, used as the AM envelope of a slightly
detuned carrier, and buried in complex white noise.  Host version (numpy) for
tests and goldens, device version (torch, plumbing only) for bench.py.
"""
import numpy as np


def test_card(y_t, x_t, active_frac=(0.93, 0.78), seed=0):
    """(y_t, x_t) float32 raster: gradient + checker + bars in [0.2, 1]; the last
    rows/columns are the blanking interval at a distinct (bright) level."""
    rng = np.random.default_rng(0xB200 + seed)
    ya, xa = int(y_t * active_frac[0]), int(x_t * active_frac[1])
    yy, xx = np.mgrid[0:y_t, 0:x_t]
    img = 0.2 + 0.5 * (xx / max(xa, 1)) * (yy < ya // 2) + 0.35 * (((xx // 64) + (yy // 48)) % 2) * (yy >= ya // 2)
    bars = rng.uniform(0.2, 1.0, size=24)
    sel = (yy > ya // 3) & (yy < ya // 3 + ya // 10)
    img = np.where(sel, bars[(xx * 24 // max(xa, 1)) % 24], img)
    img = np.clip(img, 0.2, 1.0)
    img[:, xa:] = 1.25   # H-blank
    img[ya:, :] = 1.25   # V-blank
    return img.astype(np.float32)


def envelope(n, Fs, fv, raster, phase_px=0.0, t0=0):
    """m(t) for samples t0 .. t0+n-1 (float64): linear interpolation of the scanned raster."""
    P = raster.size
    flat = raster.reshape(-1).astype(np.float64)
    t = np.arange(t0, t0 + n, dtype=np.float64)
    u = t * (P * fv / Fs) + phase_px
    k = np.floor(u)
    d = u - k
    k = k.astype(np.int64) % P
    return (1.0 - d) * flat[k] + d * flat[(k + 1) % P]


def make_iq(n, Fs, x_t, y_t, fv, seed=0, sigma=0.05, df=1e3, phase_px=None, t0=0, raster=None):
    """n complex64 samples of the synthetic capture; deterministic in (seed, t0)."""
    rng = np.random.default_rng([0xB200, seed, t0])
    if raster is None:
        raster = test_card(y_t, x_t, seed=seed)
    if phase_px is None:
        phase_px = float(np.random.default_rng(0xB200 + seed).integers(0, x_t * y_t))
    m = envelope(n, Fs, fv, raster, phase_px, t0)
    t = np.arange(t0, t0 + n, dtype=np.float64)
    carrier = np.exp(1j * (0.3 + 2 * np.pi * df * t / Fs))
    noise = rng.normal(0.0, sigma, n) + 1j * rng.normal(0.0, sigma, n)
    return ((0.1 + m) * carrier + noise).astype(np.complex64)


def make_iq_torch(n, Fs, x_t, y_t, fv, device, seed=0, sigma=0.05, df=1e3, phase_px=None, t0=0, raster=None):
    """Device-side generator for bench.py: returns an (n, 2) float32 tensor
    (interleaved re, im = ComplexF32 memory layout).  Same model as make_iq, its own
    random stream (bench parity is checked on these exact samples, not on make_iq's)."""
    import torch
    if raster is None:
        raster = test_card(y_t, x_t, seed=seed)
    if phase_px is None:
        phase_px = float(np.random.default_rng(0xB200 + seed).integers(0, x_t * y_t))
    P = raster.size
    flat = torch.from_numpy(raster.reshape(-1).astype(np.float32)).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(0xB200 * 1000003 + seed * 7919 + (t0 % 1000003))
    out = torch.empty((n, 2), dtype=torch.float32, device=device)
    chunk = 1 << 24
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        t = torch.arange(t0 + a, t0 + b, dtype=torch.float64, device=device)
        u = t * (P * fv / Fs) + phase_px
        k = torch.floor(u)
        d = (u - k).to(torch.float32)
        k = k.to(torch.int64) % P
        m = (1.0 - d) * flat[k] + d * flat[(k + 1) % P]
        ph = 0.3 + 2 * np.pi * df * t / Fs
        amp = 0.1 + m
        out[a:b, 0] = amp * torch.cos(ph).to(torch.float32)
        out[a:b, 1] = amp * torch.sin(ph).to(torch.float32)
        out[a:b] += sigma * torch.randn((b - a, 2), dtype=torch.float32, device=device, generator=g)
    return out
