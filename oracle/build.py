"""oracle/build.py -- compile the C restatement (TEST INFRASTRUCTURE ONLY).

Builds oracle/libtsdr_oracle.so from oracle/tsdr_oracle.c with floating-point
contraction disabled.  Called by __graft_entry__.build() and lazily by
oracle/orc.py; nothing in the shipped CUDA library depends on it.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtsdr_oracle.so")
SRCS = ["tsdr_oracle.c", "tsdr_oracle.h", "orc_fft.inc"]
BASE = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=gnu11", "-shared"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in SRCS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return SO
    errors = []
    for cc in ("/usr/bin/gcc", shutil.which("gcc"), shutil.which("cc")):
        if not cc or not os.path.exists(cc):
            continue
        for omp in (["-fopenmp"], []):
            cmd = [cc] + BASE + omp + ["-o", SO + ".tmp", os.path.join(HERE, "tsdr_oracle.c"), "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                os.replace(SO + ".tmp", SO)
                if verbose:
                    print("[oracle] built with", " ".join(cmd))
                return SO
            errors.append(r.stderr.strip()[-400:])
    raise RuntimeError("could not build the oracle:\n" + "\n---\n".join(errors))


if __name__ == "__main__":
    print(build(force=True, verbose=True))
