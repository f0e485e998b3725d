"""tools/sanitize.py -- one small invocation of every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
Results are compared with nothing here (the parity tests do that); the point is the tool's report."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tempestsdr_b200 as tsdr  # noqa: E402

rng = np.random.default_rng(0)


def iq(n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


which = set(sys.argv[1:]) or {"demod", "resample", "vsync", "chain", "fullres", "autocorr", "spectrum", "upsampler"}
if "demod" in which:
    x = iq(10007)
    tsdr.amDemod(x); tsdr.invert_amDemod(x); tsdr.fmDemod(x); tsdr.abs2(x); tsdr.fullScale(np.abs(x)); tsdr.findmax(np.abs(x))
if "resample" in which:
    img = tsdr.sig_to_image(np.abs(iq(5000)), 70, 90)
    tsdr.downgradeImage(img)
    tsdr.downgradeImage(rng.random((700, 900), dtype=np.float32))
    out = np.zeros(300, np.float32); tsdr.naiveResampler(out, np.arange(100, dtype=np.float32), 3)
if "vsync" in which:
    im = rng.random((600, 800), dtype=np.float32)
    s = tsdr.SyncXY(im); tsdr.vsync(im, s); tsdr.vsync(im, s)
if "chain" in which:
    for (Fs, mode, frames) in [(20e6, (1056, 628, 60.0), 2), (8e6, (800, 600, 70.0), 2), (30e6, (832, 445, 85.0), 2)]:
        cfg = tsdr.VideoMode(*mode)
        S = tsdr.getImageDuration(cfg, Fs)
        n = S * frames + 3
        ch = tsdr.Chain(Fs, cfg, alpha=0.2, max_samples=n, publish_all=True)
        for k in range(3):
            z = iq(n)
            ch.push(z)
            ch.push_i16(np.stack([z.real * 100, z.imag * 100], axis=1).astype(np.int16))
        ch.image(); ch.offsets(); ch.published()
        ch.close()
if "fullres" in which:     # TSDR_CHAIN_FULLRES: k_render_full, k_project_p<true> (tiled TMA ring), k_beta<true>, k_accumulate_rows (bulk-copy ring)
    for (Fs, mode, frames) in [(2e6, (400, 300, 60.0), 7), (2e6, (1053, 627, 60.0), 2)]:   # the second: rows that are not 16-byte multiples
        cfg = tsdr.VideoMode(*mode)
        S = tsdr.getImageDuration(cfg, Fs)
        n = S * frames + 3
        for pub in (False, True):
            ch = tsdr.Chain(Fs, cfg, alpha=0.2, max_samples=n, publish_all=pub, full_res=True)
            for k in range(2):
                ch.push(iq(n))
            ch.image(); ch.offsets(); ch.image_downgraded()
            ch.close()
if "autocorr" in which:
    for n in (1 << 12, 5000, 1 << 20):
        x = rng.random(n, dtype=np.float32) + 1
        tsdr.calculate_autocorrelation(x, 1.0e6, 0.0, (n // 2) / 1.0e6)
if "autocorr3" in which:      # the three-level kernels (n = 2^23): slow under the sanitizer
    x = rng.random(1 << 23, dtype=np.float32) + 1
    tsdr.calculate_autocorrelation(x, 1.0e6, 0.0, (1 << 22) / 1.0e6)
if "spectrum" in which:
    x = iq(9000)
    tsdr.getSpectrum(1.0, x, N=4096); tsdr.getSpectrum(1.0, x, N=1000); tsdr.getWelch(1.0, x, 256); tsdr.getWaterfall(1.0, x, 64)
if "upsampler" in which:
    r = tsdr.init_resampler(np.float32, 1024, 4)
    out = np.zeros(4096, np.float32); r(out, rng.random(1024, dtype=np.float32))
print("sanitize.py done:", sorted(which))
