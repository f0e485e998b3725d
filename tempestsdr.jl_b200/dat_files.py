"""readComplexBinary / writeComplexBinary -- the .dat container of src/DatBinaryFiles.jl:15-66 (raw interleaved
little-endian I,Q words, GNU Radio file-sink compatible), host side.  The reference's reader stays what it is; this
mirror exists so that the Python host, the tests and bench.py can write and replay captures in the same three formats
(`:short` Int16, `:single` Float32, `:double` Float64).  File I/O only -- no arithmetic on the hot path lives here."""
import numpy as np

_FORMATS = {"short": np.dtype("<i2"), "single": np.dtype("<f4"), "double": np.dtype("<f8")}


def _fmt(format):
    key = str(format).lstrip(":")
    if key not in _FORMATS:
        raise ValueError("Unsupported format for readComplexBinary. Only support :short, :single, :double and got %s" % format)
    return key, _FORMATS[key]


def readComplexBinary(file, format="single", nbSeg=None):
    """z = y[1:2:end] + 1im*y[2:2:end] (:44-66).  nbSeg counts scalar words, as in the reference.  :short samples are
    widened without scaling.  Returns complex64 (complex128 for :double, as Julia's promotion gives)."""
    key, dt = _fmt(format)
    y = np.fromfile(file, dtype=dt, count=-1 if nbSeg is None else int(nbSeg))
    re, im = y[0::2], y[1::2]
    if re.size != im.size:   # odd word count: Julia's broadcast of unequal lengths throws
        raise ValueError("DimensionMismatch: odd number of words in %s" % file)
    out = np.empty(re.size, np.complex128 if key == "double" else np.complex64)
    out.real, out.imag = re, im
    return out


def readComplexBinaryRaw(file, format="short", nbSeg=None):
    """the words as they lie in the file, shape (n, 2): what Chain.push_i16 / the pinned ring ingest for `:short`
    recordings without the host-side widening (4 bytes per sample across PCIe instead of 8)"""
    _, dt = _fmt(format)
    y = np.fromfile(file, dtype=dt, count=-1 if nbSeg is None else int(nbSeg))
    return y[: y.size // 2 * 2].reshape(-1, 2)


def writeComplexBinary(x, fileID, format="single"):
    """(:15-31) :short scales each component to 2^14 at its own maximum, ties to even, as the reference does"""
    key, dt = _fmt(format)
    x = np.asarray(x)
    out = np.zeros(2 * x.size, dt)
    if key == "short":
        scale = 1 << 14
        out[0::2] = np.rint(scale * x.real / np.max(x.real))
        out[1::2] = np.rint(scale * x.imag / np.max(x.imag))
    else:
        out[0::2] = x.real
        out[1::2] = x.imag
    out.tofile(fileID)
