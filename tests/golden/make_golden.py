"""tests/golden/make_golden.py -- regenerate the committed golden vectors.

The reference (Julia) cannot run in this image and ships no vectors for this path
(PARITY UNPINNED), so these goldens are produced by the C restatement in oracle/ on
small seeded inputs.  They pin the oracle against accidental change and give the GPU
tests fixed targets that do not depend on rebuilding the oracle.

    python tests/golden/make_golden.py
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

spec = importlib.util.spec_from_file_location("tsdr_synth", os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

CHAIN_CASE = dict(Fs=1.0e6, x_t=832, y_t=445, fv=85.0, frames=3, alpha=0.25, seed=7)


def chain_inputs():
    c = CHAIN_CASE
    S = orc.frame_samples(c["Fs"], c["fv"])
    return synth.make_iq(S * c["frames"] + 9, c["Fs"], c["x_t"], c["y_t"], c["fv"], seed=c["seed"])


def main():
    rng = np.random.default_rng(0xB200)
    out = {}

    def big(name, arr):
        """large arrays are pinned by their SHA-256 plus a coarse sub-sample (keeps the fixture small)"""
        arr = np.ascontiguousarray(arr, np.float32)
        out[name + "_sha256"] = np.frombuffer(hashlib.sha256(arr.tobytes()).digest(), np.uint8).copy()
        out[name + "_sub"] = arr[::37, ::41].copy()

    z = (rng.normal(size=257) + 1j * rng.normal(size=257)).astype(np.complex64)
    out["demod_in"] = z
    out["amDemod"] = orc.amDemod(z)
    out["invert_amDemod"] = orc.invert_amDemod(z)
    out["abs2"] = orc.abs2(z)
    sig = rng.random(1000).astype(np.float32)
    out["resize_in"] = sig
    out["sig_to_image_up"] = orc.sig_to_image(sig, 45, 52)       # 1000 -> 2340 (clamped)
    out["sig_to_image_down"] = orc.sig_to_image(sig, 20, 33)     # 1000 -> 660
    img = rng.random((90, 130)).astype(np.float32)
    out["downgrade_in"] = img
    big("downgrade_small", orc.downgradeImage(img))              # 90x130 -> 600x800 (clamped up-sampling)
    iq = chain_inputs()
    c = CHAIN_CASE
    so = orc.SyncXY()
    acc, frames, sy, sx = orc.chain_buffer(iq, c["Fs"], c["x_t"], c["y_t"], c["fv"], c["alpha"], so,
                                           np.zeros((600, 800), np.float32))
    out["chain_sy"], out["chain_sx"] = sy, sx
    big("chain_image", acc)
    big("chain_beta_x_last", so.beta_x())
    # autocorrelation of a periodic power signal: n = 6000 (= 2^4 * 3 * 5^3), lags 1..3000
    x = (1.0 + np.tile(rng.random(125).astype(np.float32), 48)).astype(np.float32)
    out["autocorr_in"] = x
    out["autocorr_db"], _ = orc.calculate_autocorrelation(x, 6000.0, 0, 0.5)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz:", {k: v.shape for k, v in out.items()})


def int16_inputs():
    """the chain case quantised to Int16 pairs, as a `:short` recording stores it (src/DatBinaryFiles.jl:47-49)"""
    iq = chain_inputs()
    return np.stack([np.rint(iq.real * 6000.0), np.rint(iq.imag * 6000.0)], axis=1).astype(np.int16)


def spectrum_input():
    rng = np.random.default_rng(0xB201)
    n = 1500
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    return x + (5.0 * np.exp(2j * np.pi * 0.3125 * np.arange(n))).astype(np.complex64)


def main_v2():
    """golden_v2.npz: the rows added after v1 -- Int16 ingest and the GetSpectrum.jl functions"""
    out = {}
    i16 = int16_inputs()
    wide = (i16[:, 0].astype(np.float32) + 1j * i16[:, 1].astype(np.float32)).astype(np.complex64)
    c = CHAIN_CASE
    so = orc.SyncXY()
    acc, _, sy, sx = orc.chain_buffer(wide, c["Fs"], c["x_t"], c["y_t"], c["fv"], c["alpha"], so,
                                      np.zeros((600, 800), np.float32), publish=False)
    out["i16_chain_sy"], out["i16_chain_sx"] = sy, sx
    acc = np.ascontiguousarray(acc, np.float32)
    out["i16_chain_image_sha256"] = np.frombuffer(hashlib.sha256(acc.tobytes()).digest(), np.uint8).copy()
    out["i16_chain_image_sub"] = acc[::37, ::41].copy()
    x = spectrum_input()
    out["spectrum_in"] = x
    out["getSpectrum_1000"] = orc.getSpectrum(1.0, x, N=1000)[1]     # chirp-z route on the GPU
    out["getSpectrum_1024"] = orc.getSpectrum(1.0, x, N=1024)[1]     # direct power-of-two route
    out["getWelch_256"] = orc.getWelch(1.0, x, sizeFFT=256)[1]
    out["getWaterfall_64"] = orc.getWaterfall(1.0, x, sizeFFT=64)[2].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **out)
    print("wrote golden_v2.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--v1" in sys.argv:
        main()
    main_v2()
