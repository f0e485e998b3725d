#!/usr/bin/env python
"""bench.py -- headline benchmark of the raw-IQ -> image chain (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg4] [--impl native|reference]

Headline (default) workload: BASELINE.json configs[2], the 200 MS/s stream the north-star target is stated on
(VideoMode(2720,1481,60), 10^8-sample buffers = 30 frames, 800 MB each).  A step = one pass of the fused chain
(amDemod -> sig_to_image -> downgradeImage -> vsync -> circshift -> EMA, src/GUI.jl:163-178) over one recv!
buffer.  `value` is whole-job complex MS/s with the buffers already in HBM; `e2e` is the same metric through the
public host API (pinned host buffer -> H2D -> chain -> D2H image).  One rank per GPU; buffers are independent, so
N ranks process N buffers per step with no data-path collective (weak scaling).

Beside the headline the same JSON line carries, under "also":
  cfg2  configs[1] (20 MS/s, 1920x1080@60) -- device-resident value + k_render roofline
  cfg5  configs[4] (1000-frame integration at 3840x2160@30): frame blocks per rank + ONE all-reduce behind the
        C ABI (tsdr_chain_allreduce, NCCL over NVLink); strong scaling; `matches_sequential` checks a sharded
        integration of the same shape against the CPU oracle run sequentially
  cfg4  configs[3] (autocorrelation of 2^26-sample buffers + refresh-rate sweep), one buffer per rank
so that a multi-GPU run of the default command exercises the collective and carries a correctness flag for it.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib.util
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "complex MS/s demod->resample->render (whole job)"
WORKLOADS = {
    # BASELINE.json configs[1]: synthetic 20 MS/s complex-Float32 IQ, 1920x1080@60 (total raster 2576x1125)
    "cfg2": dict(name="cfg2: 20 MS/s, VideoMode(2576,1125,60) '1920x1080 @ 60Hz', 10^7-sample buffers (30 frames)",
                 Fs=20e6, x_t=2576, y_t=1125, fv=60.0, n_ech=10_000_000, ring=4),
    # BASELINE.json configs[2]: synthetic 200 MS/s stream, 2560x1440@60 (CVT-RB total raster 2720x1481)
    "cfg3": dict(name="cfg3: 200 MS/s, VideoMode(2720,1481,60) 2560x1440@60, 10^8-sample buffers (30 frames)",
                 Fs=200e6, x_t=2720, y_t=1481, fv=60.0, n_ech=100_000_000, ring=2),
    # BASELINE.json configs[4]: 1000-frame averaging at 3840x2160@30 (CTA-861 total raster 4400x2250), frames sharded
    # over the ranks, partial accumulators combined with ONE all-reduce (measure_integration)
    "cfg5": dict(name="cfg5: 200 MS/s, VideoMode(4400,2250,30) 3840x2160@30, 1000-frame integration, "
                      "frame blocks per GPU + one all-reduce",
                 Fs=200e6, x_t=4400, y_t=2250, fv=30.0, frames_per_buf=25, ring=2, total_frames=1000),
    # BASELINE.json configs[3]: autocorrelation refresh-rate sweep over every refresh rate of allVideoConfigurations,
    # 2^26-sample power buffers, one buffer per GPU per step (measure_sweep)
    "cfg4": dict(name="cfg4: autocorrelation of 2^26 power samples (200 MS/s, 2560x1440@60 capture) + refresh-rate sweep "
                      "over all VideoConfigurations hypotheses, one buffer per GPU",
                 Fs=200e6, x_t=2720, y_t=1481, fv=60.0, n_ech=1 << 26, ring=2, sweep=True),
}
R = 600 * 800


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel="k_render<(bool)0>"):
    """dram bytes of one launch of `kernel` from the committed ncu capture of this workload -- only if the capture
    was taken on the SAME machine code: profiles/traffic_<cfg>.json records the SASS hash of the kernel it measured,
    profiles/sass_summary.json (written by build.py at every build) holds the hash of the kernel in the built .so."""
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
    sp = os.path.join(ROOT, "profiles", "sass_summary.json")
    if not (os.path.exists(tp) and os.path.exists(sp)):
        return None, "no ncu capture committed for this workload"
    with open(tp) as f:
        t = json.load(f)
    with open(sp) as f:
        sass = json.load(f)
    have = sass.get(kernel, {}).get("sha256")
    if not have or t.get("kernel_sass_sha256") != have:
        return None, "stale: capture was taken on kernel %s, the built library holds %s" % (t.get("kernel_sass_sha256"), have)
    return t.get("k_render_dram_bytes_per_launch"), t.get("source")


class ClockSampler:
    """nvidia-smi style clock / throttle-reason samples during the timed region (via NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_numa(index):
    """pin this rank's host threads (and therefore the pages of the pinned buffers it allocates next) to the NUMA node
    of its GPU, when the host has more than one node; returns what was found for the e2e record"""
    info = {"numa_node": None, "numa_nodes_on_host": None, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes_on_host"] = len(nodes)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = os.path.join("/sys/bus/pci/devices", bus[-12:].lower())
        with open(os.path.join(dev, "numa_node")) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node >= 0 and len(nodes) > 1:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            info["bound"] = True
    except Exception as exc:
        info["error"] = repr(exc)[:120]
    return info


def cpu_chain_baseline(orc, iq_host, wl, frames, threads):
    """time the oracle's coreProcessing body on `frames` frames of iq_host with `threads` OpenMP threads"""
    import numpy as np
    S = orc.frame_samples(wl["Fs"], wl["fv"])
    z = iq_host[: frames * S]
    so = orc.SyncXY()
    t0 = time.perf_counter()
    img, _, sy, sx = orc.chain_buffer(z, wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], 0.1, so,
                                      np.zeros((600, 800), np.float32), publish=False, nthreads=threads)
    dt = time.perf_counter() - t0
    return frames * S / dt / 1e6, dt, img, sy, sx


def run_reference(args, wl, rank):
    """--impl reference: the reference's CPU path.  Julia is not installed, so this is the C
    restatement (oracle/, kind "port") with every host thread it can use (OpenMP over frames)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")
    threads = orc.num_threads()
    S = orc.frame_samples(wl["Fs"], wl["fv"])
    n_frames_wl = wl.get("n_ech", 30 * S) // S
    probe = synth.make_iq(2 * S, wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], seed=2)
    _, dt1, *_ = cpu_chain_baseline(orc, probe, wl, 2, threads)
    per_frame = dt1 / 2
    budget = 150.0 / max(args.steps + args.warmup, 1)
    frames = int(max(1, min(n_frames_wl, budget / max(per_frame, 1e-6))))
    frames = max(frames, min(threads, n_frames_wl))  # at least one frame per thread
    iq = synth.make_iq(frames * S, wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], seed=2)
    for _ in range(args.warmup):
        cpu_chain_baseline(orc, iq, wl, frames, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_chain_baseline(orc, iq, wl, frames, threads)
    dt = time.perf_counter() - t0
    val = args.steps * frames * S / dt / 1e6
    sample = "%d frames (%d samples) of the workload per step" % (frames, frames * S)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32 (f64 coordinates)", "data": "synthetic",
           "config": {"workload": wl["name"], "note": "Julia is not installed: oracle C port of the reference path, "
                      "OpenMP over frames, sleep(0.1) of GUI.jl:179 excluded"},
           "cpu_baseline": {"value": val, "unit": "MS/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


class Ctx:
    """what every measurement needs: the rank layout, the device, one work stream, torch.distributed for barriers and
    max-over-ranks timing (plumbing), and -- for world > 1 -- the library's own NCCL communicator (tsdr.Comm), which is
    the only thing that ever touches the data path between GPUs"""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        import tempestsdr_b200 as tsdr
        self.torch, self.dist, self.tsdr = torch, dist, tsdr
        self.synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")
        self.rank, self.local_rank, self.world = rank, local_rank, world
        if tsdr.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: libtempest_b200 has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.affinity0 = os.sched_getaffinity(0)
        self.numa = bind_to_gpu_numa(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
        self.comm = None
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.comm = tsdr.Comm.from_torch_distributed(local_rank)
        self.hbm_peak, self.peak_src = peaks()
        # a dedicated (non-default) stream shared by the handles and the timing events: the legacy default stream
        # has handle 0, which the C ABI reads as "create a private stream"
        torch.cuda.synchronize()
        self.work_stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.work_stream)
        self.stream = self.work_stream.cuda_stream
        assert self.stream != 0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def bench_autocorr(ctx):
    """M2: autocorrelation ms per 2^24 samples (device resident).  Two comparators: the eager PyTorch route
    (rfft -> |X|^2 -> irfft -> slice -> log10: >= 6 kernels with temporaries) and the two cuFFT transforms alone."""
    torch, tsdr, dev = ctx.torch, ctx.tsdr, ctx.dev
    n = 1 << 24
    L = n // 2
    ring = [torch.rand(n, device=dev, dtype=torch.float32) + 1.0 for _ in range(4)]
    out = torch.empty(L, device=dev, dtype=torch.float32)
    plan = tsdr.AutocorrPlan(n, device=dev.index, stream=ctx.stream)
    for i in range(4):
        plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms = timed(lambda i: plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr()))

    def torch_route(x):
        X = torch.fft.rfft(x)
        r = torch.fft.irfft(X.real * X.real + X.imag * X.imag, n=n)[:L]
        return 10.0 * torch.log10(r * r)
    ref = torch_route(ring[(iters - 1) % 4])
    for i in range(3):
        torch_route(ring[i])
    ms_torch = timed(lambda i: torch_route(ring[i % 4]))
    spec = torch.fft.rfft(ring[0])
    for i in range(3):
        torch.fft.irfft(torch.fft.rfft(ring[i]), n=n)
    ms_cufft = timed(lambda i: torch.fft.irfft(torch.fft.rfft(ring[i % 4]), n=n))
    del spec
    err = float((out - ref).abs().max())
    algo = 4.0 * n + 4.0 * L
    launches = plan.launch_count()
    kernels_per_call = launches // (iters + 4)   # 4 warm-up + iters timed calls
    plan.close()
    return {"metric": "autocorr ms per 2^24 samples", "value": ms, "unit": "ms", "n": n, "lags": L,
            "torch_eager_route_ms": ms_torch, "cufft_transforms_only_ms": ms_cufft,
            "comparators": "torch_eager_route = rfft, |X|^2, irfft, slice, log10 as eager PyTorch ops; "
                           "cufft_transforms_only = torch.fft.rfft + irfft (the two cuFFT execs, no pointwise work)",
            "max_abs_dB_diff_vs_torch": err,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / ctx.hbm_peak, "traffic": None,
                         "algorithmic_bytes": algo, "kernels_per_call": kernels_per_call},
            "launches_total": launches}


def h2d_bare_gbs(ctx, host_tensor, reps=4):
    """bare pinned H2D bandwidth of this rank, every rank copying at once (what PCIe / the host memory system gives
    N concurrent streams): the ceiling of any end-to-end number"""
    torch = ctx.torch
    dst = torch.empty(host_tensor.shape, dtype=host_tensor.dtype, device=ctx.dev)
    dst.copy_(host_tensor, non_blocking=True)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(host_tensor, non_blocking=True)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    del dst
    return reps * host_tensor.numel() * host_tensor.element_size() / dt / 1e9


def measure_chain(ctx, key, steps, warmup, headline):
    """device-resident value + k_render roofline (+ for the headline: clocks, e2e, Int16 ingest, CPU baseline)"""
    import numpy as np
    torch, tsdr, synth, dev = ctx.torch, ctx.tsdr, ctx.synth, ctx.dev
    wl = WORKLOADS[key]
    Fs, x_t, y_t, fv, n_ech = wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], wl["n_ech"]
    cfg = tsdr.VideoMode(x_t, y_t, fv)
    S = tsdr.getImageDuration(cfg, Fs)
    frames = n_ech // S
    world = ctx.world
    # ring of distinct device buffers larger than L2 (126 MB): no buffer is L2-resident when its step starts
    ring = [synth.make_iq_torch(n_ech, Fs, x_t, y_t, fv, dev, seed=100 * ctx.rank + i, t0=i * n_ech) for i in range(wl["ring"])]
    torch.cuda.synchronize()
    ch = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=n_ech, device=ctx.local_rank, stream=ctx.stream)

    for i in range(warmup):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if headline:
        sampler.start()
    l0 = ch.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    ch.flush()  # the timed region ends when the last buffer's sync/accumulate kernels have finished too
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ch.launch_count() - l0
    clocks = None
    if headline:
        # keep the sampler running over a few more steps when the timed region was shorter than its period
        if elapsed_ms < 200:
            t_end = time.perf_counter() + 0.25
            i = 0
            while time.perf_counter() < t_end:
                ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
                i += 1
                if i % 8 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
        clocks = sampler.stop()
    ctx.barrier()
    elapsed_ms = ctx.max_over_ranks(elapsed_ms)
    value = world * steps * frames * S / (elapsed_ms * 1e-3) / 1e6  # whole-job MS/s (samples of complete frames)

    # ---- per-kernel event timing for the roofline (same steps, events between kernels, serial mode) ----
    ch.set_profiling(True)
    for i in range(steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    stage_ms, pushes = ch.kernel_times()
    ch.set_profiling(False)
    render_ms = stage_ms[0] / max(pushes, 1)
    algo_bytes = (8.0 * S + 4.0 * R) * frames  # k_render: every complex64 sample once, one 600x800 frame out
    achieved = algo_bytes / (render_ms * 1e-3) / 1e9
    chain_bytes = (8.0 * S + 12.0 * R) * frames  # SURVEY 8(d) B_chain, whole step
    traffic, traffic_src = ncu_traffic(key)
    step_ms = elapsed_ms / steps
    roofline = {"bound": "hbm", "kernel": "k_render (amDemod+sig_to_image+downgradeImage fused)", "achieved": achieved,
                "peak": ctx.hbm_peak, "peak_source": ctx.peak_src, "unit": "GB/s", "frac": achieved / ctx.hbm_peak,
                "frac_of_spec_8000": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                # what actually crosses the DRAM pins (ncu dram__bytes_read + write of one launch) over the same
                # event-timed launch: the kernel skips source lines no output row touches and its 600x800 frames
                # mostly stay in L2, so this is lower than the algorithmic fraction
                "dram_frac": (traffic / (render_ms * 1e-3) / 1e9 / ctx.hbm_peak) if traffic else None,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_per_launch": render_ms,
                "stage_ms_per_step": {"k_render": render_ms, "k_project+k_beta": stage_ms[1] / max(pushes, 1),
                                      "k_accumulate+carry": stage_ms[2] / max(pushes, 1)},
                "kernel_share_of_step": render_ms / step_ms,
                "chain_step": {"algorithmic_bytes": chain_bytes, "achieved": chain_bytes / (step_ms * 1e-3) / 1e9,
                               "frac": chain_bytes / (step_ms * 1e-3) / 1e9 / ctx.hbm_peak}}
    out = {"workload": wl["name"], "value": value, "unit": "MS/s", "ms_per_step": step_ms, "steps": steps,
           "gpu_launches": int(launches), "roofline": roofline, "frames_per_step": frames, "samples_per_frame": S,
           "l2_policy": "ring of %d distinct %.0f MB device buffers (> 126 MB L2 in total), no flush" % (len(ring), n_ech * 8 / 1e6)}
    if not headline:
        ch.close()
        del ring
        torch.cuda.empty_cache()
        return out
    out["clocks"] = clocks

    # ---- end to end through the host API: pinned host buffer -> H2D -> chain -> D2H image ----
    host_ring = [torch.empty((n_ech, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host_ring):
        h.copy_(ring[i % len(ring)])
    img_host = [torch.empty((800, 600), dtype=torch.float32).pin_memory() for _ in range(2)]  # column-major 600x800
    e2e_steps = max(3, min(steps, 20))
    for i in range(2):
        ch.push_deliver_ptr(host_ring[i % 2].data_ptr(), n_ech, img_host[i % 2].data_ptr())
    ch.wait_delivery(0)
    ctx.barrier()
    # every step: H2D of the step's pinned buffer, the chain, D2H of that buffer's imageOut into pinned memory.
    # The host waits for delivery i-1 after queueing step i, so the copy of step i overlaps step i-1's kernels.
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ch.push_deliver_ptr(host_ring[i % 2].data_ptr(), n_ech, img_host[i % 2].data_ptr())
        if i:
            ch.wait_delivery(1)
    ch.wait_delivery(0)
    ch.sync()
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * e2e_steps * frames * S / dt / 1e6
    bare = h2d_bare_gbs(ctx, host_ring[0])
    h2d_rate = e2e_steps * frames * S * 8 / dt / 1e9
    out["e2e"] = {"value": e2e_val, "unit": "MS/s", "h2d_bytes_per_step": frames * S * 8, "d2h_bytes_per_step": R * 4,
                  "steps": e2e_steps, "h2d_gbs_per_rank": h2d_rate, "h2d_bare_gbs_per_rank": bare,
                  "frac_of_bare_h2d": h2d_rate / bare, "limiter": "pinned host->device copy (PCIe / host memory fan-out): "
                  "the chain runs at %.0f%% of what a bare cudaMemcpyAsync of the same buffers reaches with all %d rank(s) copying at once"
                  % (100 * h2d_rate / bare, world), "numa": ctx.numa,
                  "note": "pinned host buffer -> tsdr_chain_push_host_deliver (H2D, chain, D2H of imageOut into pinned memory every step), wall clock"}

    if ctx.rank == 0:
        # ---- CPU baseline on a bounded sample + parity of the same frames ----
        os.sched_setaffinity(0, ctx.affinity0)   # the CPU baseline may use every host core
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        threads = orc.num_threads()
        cpu_frames = min(frames, max(2, threads))
        iq_host = host_ring[0].numpy().view(np.complex64).reshape(-1)
        v1, dt1, *_ = cpu_chain_baseline(orc, iq_host, wl, min(cpu_frames, 2), 1)
        vN, dtN, img_ref, sy_ref, sx_ref = cpu_chain_baseline(orc, iq_host, wl, cpu_frames, threads)
        chk = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=cpu_frames * S, device=ctx.local_rank)
        chk.push(iq_host[: cpu_frames * S])
        sy, sx = chk.offsets()
        same = bool(np.array_equal(chk.image(), img_ref) and np.array_equal(sy, sy_ref) and np.array_equal(sx, sx_ref))
        chk.close()
        out["cpu_baseline"] = {"value": vN, "unit": "MS/s", "cores": threads, "kind": "port",
                               "sample": "%d frames (%d samples) of the workload, OpenMP over frames; 1 thread: %.2f MS/s"
                                         % (cpu_frames, cpu_frames * S, v1),
                               "single_thread_value": v1, "gpu_matches_oracle_bit_exact": same,
                               "checked": "imageOut and (s_y, s_x) of %d %s frames, GPU chain vs oracle" % (cpu_frames, key)}
    # ---- the same workload delivered as `:short` samples (Int16 pairs, src/DatBinaryFiles.jl:47-49) ----
    try:
        out["int16_ingest"] = int16_ingest_measure(ctx, ch, ring, host_ring, wl, S, frames, min(steps, 20), 3)
    except Exception as exc:
        out["int16_ingest"] = {"error": repr(exc)}

    ch.close()
    del ring, host_ring
    torch.cuda.empty_cache()
    return out


def int16_ingest_measure(ctx, ch, ring, host_ring, wl, S, frames, steps, warmup):
    """device-resident value through k_render<Int16> and the end-to-end rate with half the PCIe bytes"""
    torch = ctx.torch
    n_ech = wl["n_ech"]
    q = []
    for r in ring:   # quantise the synthetic stream to 12 significant bits, padded to whole 4-sample groups
        t = torch.zeros(2 * n_ech + 8, dtype=torch.int16, device=r.device)
        t[: 2 * n_ech] = torch.clamp(torch.round(r.reshape(-1) * 2048.0), -32768, 32767).to(torch.int16)
        q.append(t)
    for i in range(warmup):
        ch.push_device_i16(q[i % len(q)].data_ptr(), n_ech)
    ch.flush()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ch.push_device_i16(q[i % len(q)].data_ptr(), n_ech)
    ch.flush()
    e1.record()
    torch.cuda.synchronize()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1) / steps)
    # reuse the pinned Float32 buffers' memory for the Int16 stream (half of each)
    host = [h.view(torch.int16).reshape(-1)[: 2 * n_ech] for h in host_ring]
    for i, h in enumerate(host):
        h.copy_(q[i % len(q)][: 2 * n_ech])
    img = [torch.empty((800, 600), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        ch.push_i16_deliver_ptr(host[i % 2].data_ptr(), n_ech, img[i % 2].data_ptr())
    ch.wait_delivery(0)
    k = max(3, min(steps, 20))
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(k):
        ch.push_i16_deliver_ptr(host[i % 2].data_ptr(), n_ech, img[i % 2].data_ptr())
        if i:
            ch.wait_delivery(1)
    ch.wait_delivery(0)
    ch.sync()
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    return {"workload": wl["name"] + " as Int16 (re, im) pairs", "value": ctx.world * frames * S / (ms * 1e-3) / 1e6, "unit": "MS/s",
            "ms_per_step": ms, "steps": steps,
            "e2e": {"value": ctx.world * k * frames * S / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": frames * S * 4,
                    "d2h_bytes_per_step": R * 4, "steps": k}}


def measure_integration(ctx, steps, warmup, check=True, full_res=False):
    """cfg 5: a step = one 1000-frame integration.  Rank g takes a contiguous block of frames (parallel.shard_contiguous),
    primes the sync state with its halo frame, runs the chain over its block from a zero accumulator, and ONE all-reduce
    behind the C ABI (tsdr_chain_allreduce: NCCL PreMulSum, the tail weight alpha^(frames after the block) folded into
    the collective) sums the partials.  The whole block is queued by one C call (tsdr_chain_integrate_device).  Total
    work is fixed as N grows: strong scaling."""
    import numpy as np
    torch, tsdr, synth, dev = ctx.torch, ctx.tsdr, ctx.synth, ctx.dev
    from tempestsdr_b200 import parallel
    wl = WORKLOADS["cfg5"]
    world, rank = ctx.world, ctx.rank
    Fs, x_t, y_t, fv, alpha = wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], 0.1
    cfg = tsdr.VideoMode(x_t, y_t, fv)
    S = tsdr.getImageDuration(cfg, Fs)
    per_buf = wl["frames_per_buf"]                    # frames per device buffer = frames per push
    n_ech = per_buf * S
    total = int(os.environ.get("TSDR_BENCH_CFG5_FRAMES", wl["total_frames"]))   # override: shorter blocks per rank on few GPUs
    k0, k1 = parallel.shard_contiguous(total, world, rank)
    ring = [synth.make_iq_torch(n_ech, Fs, x_t, y_t, fv, dev, seed=500 + 10 * rank + i, t0=i * n_ech) for i in range(wl["ring"])]
    torch.cuda.synchronize()
    ch = tsdr.Chain(Fs, cfg, alpha=alpha, max_samples=n_ech, device=ctx.local_rank, stream=ctx.stream, full_res=full_res)
    img_floats = ch.accumulator_ptr()[1]
    weight = parallel.ema_tail_weight(alpha, total - k1)
    # this rank's block as a list of (device pointer, samples): whole buffers of the ring, then the remainder
    bufs, cnts, done, i = [], [], 0, 0
    while done < k1 - k0:
        f = min(per_buf, k1 - k0 - done)
        bufs.append(ring[i % len(ring)].data_ptr())
        cnts.append(f * S)
        done += f
        i += 1
    halo = ring[-1].data_ptr() if k0 > 0 else 0        # the frame before this rank's block

    def one_integration():
        ch.integrate_device(halo, S, bufs, cnts, comm=ctx.comm, weight=weight)

    for _ in range(warmup):
        one_integration()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    l0 = ch.launch_count()
    c0 = ctx.comm.collectives() if ctx.comm else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_integration()
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ch.launch_count() - l0
    collectives = (ctx.comm.collectives() - c0) if ctx.comm else 0
    elapsed_ms = ctx.max_over_ranks(elapsed_ms)
    step_ms = elapsed_ms / steps
    # keep the clock sampler running for ~0.2 s more.  Every integration contains a collective, so the number of extra
    # iterations must be the SAME on every rank: it is derived from the max-over-ranks step time, never from a local
    # wall clock (a rank leaving a wall-clock loop one iteration early leaves the others inside the all-reduce for good)
    for _ in range(max(1, min(200, int(200.0 / max(step_ms, 1e-3))))):
        one_integration()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ctx.barrier()
    value = total * S / (step_ms * 1e-3) / 1e6

    # ---- where a step goes: the pieces timed one by one (CUDA events on the chain's stream, a barrier in front) ----
    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            ctx.barrier()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
        return ctx.max_over_ranks(best)

    def block_only():
        ch.integrate_device(halo, S, bufs, cnts, comm=None, weight=1.0)

    def prime_only():
        ch.reset()
        if halo:
            ch.prime_device(halo, S)
        ch.flush()

    breakdown = {"reset+halo_prime_ms": timed(prime_only), "block_without_collective_ms": timed(block_only)}
    if ctx.comm:
        t_ar = timed(lambda: ctx.comm.allreduce_chain(ch, weight))
        breakdown["allreduce_alone_ms"] = t_ar
        # bus bandwidth of the all-reduce (2(N-1)/N x bytes per rank / time): NVLink-bound only for the full-resolution accumulator
        breakdown["allreduce_busbw_gbs"] = 2.0 * (world - 1) / world * img_floats * 4 / (t_ar * 1e-3) / 1e9
    ch.set_profiling(True)
    block_only()
    stage_ms, pushes = ch.kernel_times()
    ch.set_profiling(False)
    breakdown["kernels_serial_ms"] = {"k_render": stage_ms[0], "k_project+k_beta": stage_ms[1], "k_accumulate+carry": stage_ms[2],
                                      "pushes": int(pushes)}
    render_ms = stage_ms[0] / max(pushes, 1)
    frames_per_launch = (k1 - k0 + (1 if halo else 0)) / max(pushes, 1)
    Rimg = img_floats                                  # pixels of imageOut: 600*800, or y_t*x_t at full resolution
    algo = (8.0 * S + 4.0 * Rimg) * frames_per_launch
    chain_bytes = (8.0 * S + 12.0 * Rimg) * total
    roofline = {"bound": "hbm", "kernel": "k_render_full" if full_res else "k_render", "achieved": algo / (render_ms * 1e-3) / 1e9, "peak": ctx.hbm_peak,
                "unit": "GB/s", "frac": algo / (render_ms * 1e-3) / 1e9 / ctx.hbm_peak,
                "algorithmic_bytes_per_launch": algo, "kernel_ms_per_launch": render_ms,
                "chain_step": {"algorithmic_bytes": chain_bytes, "achieved": chain_bytes / (step_ms * 1e-3) / 1e9,
                               "frac_per_gpu": chain_bytes / (step_ms * 1e-3) / 1e9 / ctx.hbm_peak / world}}
    out = {"workload": wl["name"] + (" -- FULL RESOLUTION (no downgradeImage: frames, SyncXY and imageOut at 2250x4400)" if full_res else ""),
           "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": steps,
           "ms_per_step": step_ms, "scaling": "strong", "accumulator_bytes": img_floats * 4, "frames_per_step": total, "samples_per_frame": S,
           "frames_this_rank": k1 - k0, "frames_per_push": per_buf, "pushes_this_rank": len(bufs),
           "gpu_launches": int(launches), "collectives": int(collectives),
           "collective": "tsdr_chain_allreduce: NCCL all-reduce (PreMulSum) of the %.2f MB accumulator, bound by the library " % (img_floats * 4 / 1e6) +
                         "itself behind the C ABI; torch.distributed only carried the communicator id" if ctx.comm else "none (1 rank)",
           "clocks": clocks, "breakdown": breakdown, "roofline": roofline}
    ch.close()
    del ring
    torch.cuda.empty_cache()
    if check:
        try:
            out.update(check_integration(ctx, cfg, Fs, alpha, full_res))
        except Exception as exc:
            out["matches_sequential"] = None
            out["check_error"] = repr(exc)[:300]
    return out


def check_integration(ctx, cfg, Fs, alpha, full_res=False):
    """correctness of the sharded integration at the cfg 5 shape: T frames, the same capture on every rank (rank 0
    generates it, torch.distributed carries it: test data, not the data path), each rank integrates its block and the
    all-reduce combines; rank 0 runs the CPU oracle over all T frames SEQUENTIALLY and compares image and offsets"""
    import numpy as np
    torch, tsdr, synth, dev, dist = ctx.torch, ctx.tsdr, ctx.synth, ctx.dev, ctx.dist
    from tempestsdr_b200 import parallel
    world, rank = ctx.world, ctx.rank
    S = tsdr.getImageDuration(cfg, Fs)
    T = max(4, world) if full_res else max(8, 2 * world)
    if rank == 0:
        iq = synth.make_iq_torch(T * S, Fs, cfg.width, cfg.height, cfg.refresh, dev, seed=4242)
    else:
        iq = torch.empty((T * S, 2), dtype=torch.float32, device=dev)
    if world > 1:
        dist.broadcast(iq, src=0)
    torch.cuda.synchronize()
    k0, k1 = parallel.shard_contiguous(T, world, rank)
    ch = tsdr.Chain(Fs, cfg, alpha=alpha, max_samples=(k1 - k0) * S, device=ctx.local_rank, stream=ctx.stream, full_res=full_res)
    base = iq.data_ptr()
    halo = base + (k0 - 1) * S * 8 if k0 > 0 else 0
    ch.integrate_device(halo, S, [base + k0 * S * 8], [(k1 - k0) * S], comm=ctx.comm, weight=parallel.ema_tail_weight(alpha, T - k1))
    img = ch.image()
    sy, sx = ch.offsets()
    ch.close()
    offs = [(k0, [int(v) for v in sy], [int(v) for v in sx])]
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, offs[0])
        offs = box
    res = {}
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        z = iq.cpu().numpy().view(np.complex64).reshape(-1)
        if full_res:
            ref, _, sy_ref, sx_ref = orc.chain_buffer_fullres(z, Fs, cfg.width, cfg.height, cfg.refresh, alpha,
                                                              orc.SyncXY(cfg.height, cfg.width),
                                                              np.zeros((cfg.height, cfg.width), np.float32))
        else:
            ref, _, sy_ref, sx_ref = orc.chain_buffer(z, Fs, cfg.width, cfg.height, cfg.refresh, alpha, orc.SyncXY(),
                                                      np.zeros((600, 800), np.float32), publish=False, nthreads=orc.num_threads())
        sy_all = [v for _, a, _ in sorted(offs) for v in a]
        sx_all = [v for _, _, b in sorted(offs) for v in b]
        offsets_ok = sy_all == [int(v) for v in sy_ref] and sx_all == [int(v) for v in sx_ref]
        err = float(np.max(np.abs(img.astype(np.float64) - ref) / (np.abs(ref) * 2e-6 + 1e-7)))
        res = {"matches_sequential": bool(offsets_ok and (np.array_equal(img, ref) if world == 1 else err <= 1.0)),
               "check": {"frames": T, "offsets_equal": bool(offsets_ok), "bit_exact": bool(np.array_equal(img, ref)),
                         "max_err_over_tolerance": err, "tolerance": "rtol 2e-6 + atol 1e-7 (Float32 rounding order of the "
                         "recombined EMA); bit-exact required on 1 GPU", "against": "CPU oracle, sequential over all frames"}}
    del iq
    torch.cuda.empty_cache()
    return res


def measure_sweep(ctx, steps, warmup):
    """cfg 4: a step = one 2^26-sample power buffer per GPU: FFT autocorrelation (device resident) and the score of every
    refresh-rate hypothesis of allVideoConfigurations (first maximum of Gamma in each hypothesis' window).  Buffers are
    independent: no collective on the data path (weak scaling); the 13 (rate, score, lag) triples per buffer stay on the
    host of their rank."""
    torch, tsdr, synth, dev = ctx.torch, ctx.tsdr, ctx.synth, ctx.dev
    wl = WORKLOADS["cfg4"]
    world, rank = ctx.world, ctx.rank
    Fs, n = wl["Fs"], wl["n_ech"]
    L = n // 2
    ring = []
    for i in range(wl["ring"]):   # abs2.(IQ) as extract_configuration feeds it (src/GUI.jl:70)
        z = synth.make_iq_torch(n, Fs, wl["x_t"], wl["y_t"], wl["fv"], dev, seed=700 + 10 * rank + i, t0=i * n)
        ring.append((z[:, 0] * z[:, 0] + z[:, 1] * z[:, 1]).contiguous())
        del z
    torch.cuda.empty_cache()
    st = ctx.stream
    plan = tsdr.AutocorrPlan(n, device=ctx.local_rank, stream=st)
    gamma = torch.empty(L, device=dev, dtype=torch.float32)
    rates = sorted(tsdr.get_refresh_rates(tsdr.allVideoConfigurations))

    def step(x_ptr):
        plan.exec(x_ptr, 1, L, gamma.data_ptr())
        return tsdr.sweep_refresh_hypotheses(gamma.data_ptr(), L, Fs, rates, stream=st)

    for i in range(warmup):
        res = step(ring[i % len(ring)].data_ptr())
    ctx.barrier()
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        res = step(ring[i % len(ring)].data_ptr())
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = plan.launch_count() - l0 + 2 * steps   # + the two launches of the batched window search per step
    ctx.barrier()
    elapsed_ms = ctx.max_over_ranks(elapsed_ms)
    per_2_24 = elapsed_ms / (world * steps * (n >> 24))      # whole job: ms per 2^24 samples
    # autocorrelation kernels alone, for the roofline
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        plan.exec(ring[i % len(ring)].data_ptr(), 1, L, gamma.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    fft_ms = e0.elapsed_time(e1) / steps
    algo = 4.0 * n + 4.0 * L
    best = max(res, key=lambda r: r[1])
    # end to end: pinned host power buffer -> H2D -> autocorrelation -> sweep -> the triples on the host
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host):
        h.copy_(ring[i % len(ring)])
    xdev = [torch.empty(n, device=dev, dtype=torch.float32) for _ in range(2)]
    k = max(2, min(steps, 6))
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(k):
        xdev[i % 2].copy_(host[i % 2], non_blocking=True)
        step(xdev[i % 2].data_ptr())
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    out = {"workload": wl["name"], "metric": "autocorr ms per 2^24 samples (whole job, incl. the refresh-hypothesis sweep)",
           "value": per_2_24, "unit": "ms", "n_gpus": world, "steps": steps, "ms_per_step": elapsed_ms / steps,
           "higher_is_better": False, "scaling": "weak", "samples_per_step_per_gpu": n, "lags": L, "hypotheses": len(rates),
           "detected": {"rate_hypothesis": best[0], "fv_hat": best[2], "lag_index": best[3], "true_refresh": wl["fv"]},
           "recovers_refresh": bool(abs(best[2] - wl["fv"]) < 0.05 and best[0] == wl["fv"]),
           "e2e": {"value": dt * 1e3 / (world * k * (n >> 24)), "unit": "ms", "h2d_bytes_per_step": n * 4,
                   "d2h_bytes_per_step": len(rates) * 12, "steps": k},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k3_p1..p5 (three-level autocorrelation, 5 launches)", "achieved": algo / (fft_ms * 1e-3) / 1e9,
                        "peak": ctx.hbm_peak, "unit": "GB/s", "frac": algo / (fft_ms * 1e-3) / 1e9 / ctx.hbm_peak,
                        "traffic": None, "algorithmic_bytes_per_launch": algo, "kernel_ms_per_launch": fft_ms}}
    plan.close()
    del ring, host, xdev, gamma
    torch.cuda.empty_cache()
    return out


def measure_replay(ctx):
    """cfg 1 (BASELINE configs[0]): headless replay of a recorded capture through the reference's own recipe
    (production/investigate_data.jl:37-97,159-206): .dat read -> amDemod -> autocorrelation -> refresh and line peaks ->
    find_closest_configuration -> toImage -> SyncXY/vsync on the full-size frame -> offset correction.  The bundled
    dumpIQ_0.dat is missing from the checkout: a seeded 10^7-sample stand-in of the mode the docs name for it
    (VideoMode(2800,1589,60.14), docs/src/gui.md:29) is written in the `:single` .dat format and read back.  GPU path =
    the per-function (host pointer) entry points, as a Julia caller would use them; CPU = the oracle, same recipe."""
    import tempfile
    import numpy as np
    tsdr, synth = ctx.tsdr, ctx.synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    Fs, mode = 20e6, (2800, 1589, 60.14)
    iq = synth.make_iq(10_000_000, Fs, *mode, seed=314)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "dumpIQ_standin.dat")
        tsdr.writeComplexBinary(iq, path, "single")
        t0 = time.perf_counter()
        sigRx = tsdr.readComplexBinary(path, "single")
        t_read = time.perf_counter() - t0
    tsdr.investigate_capture(sigRx[:4_000_000 + 10], Fs, offset=1000)    # warm-up (plans, scratch)
    t0 = time.perf_counter()
    got = tsdr.investigate_capture(sigRx, Fs)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = orc.investigate_capture(sigRx, Fs, tsdr.find_closest_configuration)
    t_cpu = time.perf_counter() - t0
    same = all(got[k] == ref[k] for k in ("fv", "posMax", "m", "y_t", "name", "idx")) and tuple(got["vsync"]) == tuple(ref["vsync"]) \
        and bool(np.array_equal(got["image"], ref["image"])) and bool(np.array_equal(got["image_synced"], ref["image_synced"]))
    return {"workload": "cfg1: headless replay (production/investigate_data.jl) of a 10^7-sample .dat stand-in for dumpIQ_0.dat, "
                        "Fs 20 MS/s, VideoMode(2800,1589,60.14)",
            "gpu_ms": t_gpu * 1e3, "cpu_oracle_ms": t_cpu * 1e3, "dat_read_ms": t_read * 1e3,
            "value": sigRx.size / t_gpu / 1e6, "unit": "MS/s (capture samples / wall time of the whole recipe, host arrays in and out)",
            "cpu_value": sigRx.size / t_cpu / 1e6,
            "detected": {"fv": got["fv"], "y_t": got["y_t"], "name": got["name"], "vsync": list(got["vsync"]), "idx": got["idx"]},
            "matches_oracle": bool(same),
            "checked": "fv, refresh / line peak positions, y_t, table entry, full-size (1589x2800) vsync offsets, offset "
                       "correction: equal; both rendered frames: bit-exact"}


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout carries exactly one JSON line: everything else libraries print there (NCCL's version banner, torchrun
    notices) is sent to stderr; emit() writes the line to the real stdout"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj)
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.write(line + "\n")
        _REAL_STDOUT.flush()
    else:
        print(line, flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg5_fullres"])
    ap.add_argument("--no-extras", action="store_true", help="skip the also / autocorr extras")
    args = ap.parse_args()
    wl = WORKLOADS["cfg5" if args.workload == "cfg5_fullres" else args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, wl if "n_ech" in wl and not wl.get("sweep") else WORKLOADS["cfg3"], rank)
        return

    ctx = Ctx(rank, local_rank, world)
    base = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "vs_baseline": None, "data": "synthetic"}
    if args.workload in ("cfg5", "cfg5_fullres"):
        m = measure_integration(ctx, args.steps, args.warmup, full_res=args.workload == "cfg5_fullres")
        out = dict(base, metric=METRIC, value=m["value"], unit="MS/s", ms_per_step=m["ms_per_step"], higher_is_better=True,
                   scaling="strong", dtype="f32 (f64 coordinates)",
                   config={"workload": m["workload"], "frames_per_step": m["frames_per_step"], "frames_per_push": m["frames_per_push"],
                           "parallelism": "contiguous frame blocks, halo frame primed, one all-reduce of %.2f MB per integration (C ABI, NCCL)" % (m["accumulator_bytes"] / 1e6),
                           "l2_policy": "ring of 2 distinct 1.3 GB device buffers (> 126 MB L2), no flush"},
                   clocks=m["clocks"], gpu_launches=m["gpu_launches"], roofline=m["roofline"], integration=m,
                   e2e={"value": None, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "device-resident workload (53 GB per integration does not come from one host buffer)"})
    elif args.workload == "cfg4":
        m = measure_sweep(ctx, args.steps, args.warmup)
        out = dict(base, metric=m["metric"], value=m["value"], unit="ms", ms_per_step=m["ms_per_step"], higher_is_better=False,
                   scaling="weak", dtype="f32", config={"workload": m["workload"], "hypotheses": m["hypotheses"],
                                                         "parallelism": "one buffer per GPU per step, no collective"},
                   e2e=m["e2e"], gpu_launches=m["gpu_launches"], roofline=m["roofline"], sweep=m)
    else:
        m = measure_chain(ctx, args.workload, args.steps, args.warmup, headline=True)
        out = dict(base, metric=METRIC, value=m["value"], unit="MS/s", ms_per_step=m["ms_per_step"], higher_is_better=True,
                   scaling="weak", dtype="f32 (f64 coordinates)",
                   config={"workload": m["workload"], "frames_per_step": m["frames_per_step"], "samples_per_frame": m["samples_per_frame"],
                           "l2_policy": m["l2_policy"],
                           "parallelism": "one buffer per GPU per step, no collective" if world > 1 else "single GPU"},
                   clocks=m["clocks"], e2e=m["e2e"], gpu_launches=m["gpu_launches"], roofline=m["roofline"])
        for k in ("cpu_baseline", "int16_ingest"):
            if k in m:
                out[k] = m[k]
        if not args.no_extras:
            also = {}
            # The headline above is complete.  The extras below contain collectives (cfg 5); should one of them ever hang
            # (a rank lost, a collective mismatch) the run must still end with its line: after TSDR_BENCH_EXTRAS_DEADLINE
            # seconds (default 240) every rank leaves, rank 0 printing the headline with whatever extras had finished.
            deadline = float(os.environ.get("TSDR_BENCH_EXTRAS_DEADLINE", "240"))

            def _give_up():
                if rank == 0:
                    out["also"] = dict(also, timed_out="extras did not finish within %.0f s; legs present had completed" % deadline)
                    emit(out)
                os._exit(0)

            watchdog = threading.Timer(deadline, _give_up)
            watchdog.daemon = True
            watchdog.start()
            for name, fn in (("cfg2" if args.workload == "cfg3" else "cfg3",
                              lambda: measure_chain(ctx, "cfg2" if args.workload == "cfg3" else "cfg3", min(args.steps, 20), 3, headline=False)),
                             ("cfg5", lambda: measure_integration(ctx, max(2, min(args.steps, 5)), 3)),
                             ("cfg5_fullres", lambda: measure_integration(ctx, 2, 3, full_res=True)),
                             ("cfg4", lambda: measure_sweep(ctx, max(2, min(args.steps, 8)), 3))):
                try:
                    also[name] = fn()
                except Exception as exc:  # the headline line must still print
                    also[name] = {"error": repr(exc)[:300]}
            if world == 1:
                try:
                    also["cfg1"] = measure_replay(ctx)
                except Exception as exc:
                    also["cfg1"] = {"error": repr(exc)[:300]}
            out["also"] = also
            if rank == 0:
                try:
                    out["autocorr"] = bench_autocorr(ctx)
                except Exception as exc:
                    out["autocorr"] = {"error": repr(exc)[:300]}
    try:
        watchdog.cancel()
    except NameError:
        pass
    ctx.close()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
