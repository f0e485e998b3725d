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


if __name__ == "__main__":
    main()
