"""Build libtempest_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python tempestsdr.jl_b200/build.py [--force] [--verbose]

One object per .cu (the exact-rounding files with -fmad=false, the FFT with
fused multiply-adds), linked into tempestsdr.jl_b200/libtempest_b200.so.  The
.so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libtempest_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--threads", "2", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
          "-Xptxas", "-v"]
# (source, extra flags)
UNITS = [
    ("tsdr_core.cu", ["-fmad=false"]),
    ("tsdr_fft.cu", ["-fmad=true"]),
    ("tsdr_ring.cu", []),
    ("tsdr_comm.cu", []),
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tempest_b200.h")]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in _deps() + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    os.makedirs(OBJDIR, exist_ok=True)
    env = dict(os.environ)
    # nvcc's host compiler: the system gcc (the image's CC wrapper lacks some specs)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    log = []
    for src, extra in UNITS:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc()] + ccbin + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-1])
        objs.append(obj)
    cmd = [nvcc()] + ccbin + ARCH + ["-shared", "-o", SO + ".tmp"] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    os.replace(SO + ".tmp", SO)
    with open(os.path.join(OBJDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
