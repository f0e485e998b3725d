"""Build libtempest_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python tempestsdr.jl_b200/build.py [--force] [--verbose]

One object per .cu (the exact-rounding files with -fmad=false, the FFT with
fused multiply-adds), linked into tempestsdr.jl_b200/libtempest_b200.so.  The
.so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import json
import os
import re
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


HASHFILE = os.path.join(HERE, "libtempest_b200.srchash")   # beside the .so: git-ignored, travels with gpurun
PROFILES = os.path.join(HERE, "..", "profiles")


def _deps():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "tempest_b200.h")]


def source_hash():
    """sha256 over every source the library is built from and the flags it is built with"""
    h = hashlib.sha256()
    for p in _deps():
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(repr((ARCH, COMMON, UNITS)).encode())
    return h.hexdigest()


def stale():
    """the binary is current iff the hash recorded beside it equals the hash of the sources: file times do not
    survive a checkout or the copy to the GPU box, and a stale .so must never be reused silently"""
    if not os.path.exists(SO) or not os.path.exists(HASHFILE):
        return True
    with open(HASHFILE) as f:
        return f.read().strip() != source_hash()


MNEMONICS = ["UBLKCP", "SYNCS", "DADD.RM", "DADD", "DMUL", "DFMA", "MUFU.RSQ", "MUFU.RCP", "MUFU.LG2", "F2F", "FFMA", "FADD", "FMUL",
             "SHFL", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "BAR", "HMMA", "UTCHMMA", "UTMALDG"]


def sass_summary(so=SO, write=True):
    """per kernel: instruction count, the mnemonics DESIGN.md cites (UBLKCP = TMA bulk copy, SYNCS = mbarrier,
    DADD.RM = the round-down floor trick, MUFU.* ...) and a sha256 of the instruction stream, from cuobjdump -sass.
    Tracked evidence under profiles/ (sass_summary.txt / .json); bench.py ties ncu traffic captures to these hashes."""
    cuobjdump = os.path.join(os.path.dirname(nvcc()), "cuobjdump")
    r = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cuobjdump failed: " + r.stderr)
    funcs, cur = {}, None
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            funcs[cur].append(re.sub(r"\s+", " ", m.group(1).strip()))
    filt = shutil.which("cu++filt") or os.path.join(os.path.dirname(nvcc()), "cu++filt")
    names = list(funcs)
    try:
        dem = subprocess.run([filt] + names, capture_output=True, text=True).stdout.splitlines()
        dem = [d.replace("tsdr::", "") for d in dem]
        dem = [re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d) for d in dem]
        dem = [re.sub(r"^void ", "", d) for d in dem]
    except Exception:
        dem = names
    if len(dem) != len(names):
        dem = names
    out = {}
    for nm, d in zip(names, dem):
        ins = funcs[nm]
        ops = [i.split(" ", 2)[1] if i.startswith("@") and " " in i else i.split(" ", 1)[0] for i in ins]
        counts = {k: sum(1 for o in ops if o == k or o.startswith(k + ".")) for k in MNEMONICS}
        counts["DADD.RM"] = sum(1 for o in ops if o.startswith("DADD") and ".RM" in o)
        out[d] = {"instructions": len(ins), "sha256": hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16],
                  "mnemonics": {k: v for k, v in counts.items() if v}}
    if write and os.path.isdir(PROFILES):
        with open(os.path.join(PROFILES, "sass_summary.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        with open(os.path.join(PROFILES, "sass_summary.txt"), "w") as f:
            f.write("# cuobjdump -sass of libtempest_b200.so (sm_100a), written by tempestsdr.jl_b200/build.py at every build\n")
            f.write("# kernel | SASS instructions | sha256[:16] of the instruction stream | mnemonic counts\n")
            for d in sorted(out):
                f.write("%s | %d | %s | %s\n" % (d, out[d]["instructions"], out[d]["sha256"],
                                               " ".join("%s=%d" % kv for kv in sorted(out[d]["mnemonics"].items()))))
    return out


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
    with open(HASHFILE, "w") as f:
        f.write(source_hash() + "\n")
    with open(os.path.join(OBJDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    try:
        sass_summary()
    except Exception as exc:   # evidence, not a build product
        log.append("sass summary skipped: %r" % (exc,))
    if verbose:
        print("\n".join(log))
    return SO


def build_variant(tag, defines):
    """libtempest_b200_<tag>.so with extra -D flags (same-box A/B of compile-time shapes with tools/ab_render.py):
        python tempestsdr.jl_b200/build.py --variant s4 -DTSDR_PROJP_STAGES=4
    Not the product library: no hash file, no SASS summary."""
    vdir = os.path.join(OBJDIR, "variant_" + tag)
    os.makedirs(vdir, exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(vdir, src.replace(".cu", ".o"))
        cmd = [nvcc()] + ccbin + ARCH + [f for f in COMMON if f not in ("-Xptxas", "-v")] + extra + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        objs.append(obj)
    out = os.path.join(HERE, "libtempest_b200_%s.so" % tag)
    cmd = [nvcc()] + ccbin + ARCH + ["-shared", "-o", out] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
