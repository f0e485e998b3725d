"""A/B of library builds on one box: python tools/ab_render.py a.so b.so ... prints the k_render event time
and the pipelined step time of each build for cfg2 and cfg3, three rounds, alternating (run-to-run noise between
boxes is ~2 %, larger than most kernel tweaks)."""
import ctypes as C, sys, os, importlib.util
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
WL = {"cfg2": (20e6, 2576, 1125, 60.0, 10_000_000, 4), "cfg3": (200e6, 2720, 1481, 60.0, 100_000_000, 2)}
dev = torch.device("cuda:0")
rings = {}
for name, (Fs, x, y, fv, n, r) in WL.items():
    rings[name] = [synth.make_iq_torch(n, Fs, x, y, fv, dev, seed=900 + i, t0=i * n) for i in range(r)]
torch.cuda.synchronize()
def run(so, name):
    L = C.CDLL(so)
    Fs, x, y, fv, n, r = WL[name]
    ring = rings[name]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        h = C.c_void_p()
        L.tsdr_chain_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_float, C.c_size_t, C.c_int, C.c_void_p]
        assert L.tsdr_chain_create(C.byref(h), 0, Fs, x, y, fv, 0.1, n, int(os.environ.get("AB_FLAGS", "0")), C.c_void_p(st.cuda_stream)) == 0
        L.tsdr_chain_push_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.tsdr_chain_flush.argtypes = [C.c_void_p]; L.tsdr_chain_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.tsdr_chain_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.tsdr_chain_destroy.argtypes = [C.c_void_p]
        push = lambda i: L.tsdr_chain_push_device(h, C.c_void_p(ring[i % r].data_ptr()), n, None)
        for i in range(5): push(i)
        L.tsdr_chain_flush(h); torch.cuda.synchronize()
        steps = 40 if name == "cfg2" else 12
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): push(i)
        L.tsdr_chain_flush(h); e1.record(); torch.cuda.synchronize()
        step = e0.elapsed_time(e1) / steps
        L.tsdr_chain_set_profiling(h, 1)
        for i in range(steps): push(i)
        ms = (C.c_float * 3)(); p = (C.c_uint64 * 1)()
        L.tsdr_chain_kernel_times(h, ms, p)
        L.tsdr_chain_set_profiling(h, 0)
        L.tsdr_chain_destroy(h)
    return step * 1e3, ms[0] / p[0] * 1e3, ms[1] / p[0] * 1e3
sos = sys.argv[1:]
envs = os.environ.get("AB_ENVS", "").split(";")   # e.g. AB_ENVS="TSDR_X=1;TSDR_X=2": each build is run under each setting
for rep in range(3):
    for name in WL:
        for ev in envs:
            if ev:
                k, v = ev.split("=")
                os.environ[k] = v
            print(rep, name, ev, "  ".join("%s: step %.1f render %.1f sync %.1f us" % ((os.path.basename(s),) + run(s, name)) for s in sos), flush=True)
