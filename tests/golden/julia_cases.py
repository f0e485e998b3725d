"""Seeded inputs of the Julia golden run (tools/julia_goldens.jl) -- shared by the writer below and by
tests/test_julia_goldens.py, which regenerates the same arrays in memory.

    python tests/golden/julia_cases.py            # writes tests/golden/julia_inputs/*.dat + inputs.sha256
    julia tools/julia_goldens.jl                  # (needs Julia + the reference's deps) writes tests/golden/julia_v1/

The .dat files follow src/DatBinaryFiles.jl:15-31 (`:single`: interleaved little-endian Float32 re, im), so the
Julia side reads them with the reference's own readComplexBinary.  Real vectors are stored as complex with a zero
imaginary part.  The inputs are git-ignored (regenerated from the seeds; their sha256 travels in the manifest).
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INPUTS = os.path.join(HERE, "julia_inputs")
OUTPUTS = os.path.join(HERE, "julia_v1")

# name, Fs, (x_t, y_t, fv), frames
CHAIN_CASES = [
    ("chain_up", 1.0e6, (800, 525, 60.0), 3),        # 1-D upsampling (P > S), y_t < 600: clamped 2-D resize
    ("chain_down", 1.0e6, (176, 120, 40.0), 3),      # 1-D downsampling (S > P)
    ("chain_typ", 2.0e6, (1056, 628, 60.0), 2),      # a table entry, both dimensions shrink
]
AUTOCORR = dict(n=60000, Fs=200000.0, period=3333)   # n = 2*indexMax with maxDelay 0.15: not a power of two


def _synth():
    spec = importlib.util.spec_from_file_location("tsdr_synth", os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs():
    """name -> complex64 vector"""
    synth = _synth()
    rng = np.random.default_rng(0xB200)
    out = {}
    z = (rng.standard_normal(4096) + 1j * rng.standard_normal(4096)).astype(np.complex64)
    z[:4] = [3 + 4j, 0, 1e-3 - 2e-3j, -5 + 12j]
    out["demod"] = z
    out["resize"] = rng.random(3333).astype(np.float32).astype(np.complex64)
    for name, Fs, (x_t, y_t, fv), frames in CHAIN_CASES:
        S = int(np.rint(Fs / fv))
        out[name] = synth.make_iq(frames * S + 7, Fs, x_t, y_t, fv, seed=31)
    a = AUTOCORR
    base = rng.random(a["period"]).astype(np.float32)
    x = np.tile(base, a["n"] // a["period"] + 1)[: a["n"]] + 0.3 * rng.random(a["n"]).astype(np.float32)
    out["autocorr"] = (1.0 + x).astype(np.float32).astype(np.complex64)
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(INPUTS, exist_ok=True)
    lines = []
    for name, z in inputs().items():
        z.view(np.float32).tofile(os.path.join(INPUTS, name + ".dat"))   # writeComplexBinary(z, file, :single)
        lines.append("%s %d %s" % (name, z.size, sha(z)))
    with open(os.path.join(INPUTS, "inputs.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote %d inputs to %s" % (len(lines), INPUTS))


if __name__ == "__main__":
    sys.exit(main())
