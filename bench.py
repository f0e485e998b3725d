#!/usr/bin/env python
"""bench.py -- headline benchmark of the raw-IQ -> image chain (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3] [--impl native|reference]

A step = one pass of the fused chain (amDemod -> sig_to_image -> downgradeImage ->
vsync -> circshift -> EMA, src/GUI.jl:163-178) over one recv! buffer of synthetic IQ.
`value` is whole-job complex MS/s with the buffers already in HBM; `e2e` is the same
metric through the public host API (pinned host buffer -> H2D -> chain -> D2H image).
One rank per GPU; buffers are independent, so N ranks process N buffers per step with
no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
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
    # over the ranks, partial accumulators combined with ONE NCCL all-reduce (handled by run_integration)
    "cfg5": dict(name="cfg5: 200 MS/s, VideoMode(4400,2250,30) 3840x2160@30, 1000-frame integration, "
                      "frame blocks per GPU + one all-reduce",
                 Fs=200e6, x_t=4400, y_t=2250, fv=30.0, n_ech=10 * 6_666_667, ring=2, total_frames=1000),
    # BASELINE.json configs[3]: autocorrelation refresh-rate sweep over every refresh rate of allVideoConfigurations,
    # 2^26-sample power buffers, one buffer per GPU per step (handled by run_sweep)
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
    import numpy as np
    import orc
    synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")
    threads = orc.num_threads()
    S = orc.frame_samples(wl["Fs"], wl["fv"])
    probe = synth.make_iq(2 * S, wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], seed=2)
    _, dt1, *_ = cpu_chain_baseline(orc, probe, wl, 2, threads)
    per_frame = dt1 / 2
    budget = 150.0 / max(args.steps + args.warmup, 1)
    frames = int(max(1, min(wl["n_ech"] // S, budget / max(per_frame, 1e-6))))
    frames = max(frames, min(threads, wl["n_ech"] // S))  # at least one frame per thread
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


def bench_autocorr(tsdr, torch, dev, hbm_peak):
    """M2: autocorrelation ms per 2^24 samples (device resident), cuFFT (torch.fft) timed beside it."""
    n = 1 << 24
    L = n // 2
    ring = [torch.rand(n, device=dev, dtype=torch.float32) + 1.0 for _ in range(4)]
    out = torch.empty(L, device=dev, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    plan = tsdr.AutocorrPlan(n, device=dev.index, stream=st)
    for i in range(4):
        plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters

    def cufft(x):
        X = torch.fft.rfft(x)
        r = torch.fft.irfft(X.real * X.real + X.imag * X.imag, n=n)[:L]
        return 10.0 * torch.log10(r * r)
    ref = cufft(ring[(iters - 1) % 4])
    for i in range(3):
        cufft(ring[i])
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        cufft(ring[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms_cufft = e0.elapsed_time(e1) / iters
    err = float((out - ref).abs().max())
    algo = 4.0 * n + 4.0 * L
    launches = plan.launch_count()
    plan.close()
    return {"metric": "autocorr ms per 2^24 samples", "value": ms, "unit": "ms", "n": n, "lags": L,
            "cufft_torch_ms": ms_cufft, "max_abs_dB_diff_vs_cufft": err,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "algorithmic_bytes": algo, "kernels_per_call": launches // (iters + 4)},
            "launches_total": launches}


def quick_chain_measure(tsdr, torch, synth, wl, dev, local_rank, stream, steps, warmup, hbm_peak):
    """device-resident value + k_render roofline for another workload (reported under "also")"""
    Fs, x_t, y_t, fv, n_ech = wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], wl["n_ech"]
    cfg = tsdr.VideoMode(x_t, y_t, fv)
    S = tsdr.getImageDuration(cfg, Fs)
    frames = n_ech // S
    ring = [synth.make_iq_torch(n_ech, Fs, x_t, y_t, fv, dev, seed=900 + i, t0=i * n_ech) for i in range(wl["ring"])]
    torch.cuda.synchronize()
    ch = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=n_ech, device=local_rank, stream=stream)
    for i in range(warmup):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    ch.flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    ch.flush()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ch.set_profiling(True)
    for i in range(steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    stage_ms, pushes = ch.kernel_times()
    ch.set_profiling(False)
    ch.close()
    del ring
    torch.cuda.empty_cache()
    render_ms = stage_ms[0] / max(pushes, 1)
    algo = (8.0 * S + 4.0 * R) * frames
    chain_bytes = (8.0 * S + 12.0 * R) * frames
    return {"workload": wl["name"], "value": frames * S / (ms * 1e-3) / 1e6, "unit": "MS/s", "ms_per_step": ms, "steps": steps,
            "roofline": {"bound": "hbm", "kernel": "k_render", "achieved": algo / (render_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "frac": algo / (render_ms * 1e-3) / 1e9 / hbm_peak, "kernel_ms_per_launch": render_ms,
                         "algorithmic_bytes_per_launch": algo,
                         "chain_step": {"algorithmic_bytes": chain_bytes, "achieved": chain_bytes / (ms * 1e-3) / 1e9,
                                        "frac": chain_bytes / (ms * 1e-3) / 1e9 / hbm_peak}}}


def int16_ingest_measure(tsdr, torch, ch, ring, wl, S, frames, steps, warmup):
    """the same workload delivered as `:short` samples (Int16 pairs, src/DatBinaryFiles.jl:47-49): device-resident
    value through k_render<Int16> and the end-to-end rate with half the PCIe bytes"""
    n_ech = wl["n_ech"]
    q = []
    for r in ring:   # quantise the synthetic stream to 12 significant bits, padded to whole 4-sample groups
        t = torch.zeros(2 * n_ech + 8, dtype=torch.int16, device=r.device)
        t[: 2 * n_ech] = torch.clamp(torch.round(r.reshape(-1) * 2048.0), -32768, 32767).to(torch.int16)
        q.append(t)
    for i in range(warmup):
        ch.push_device_i16(q[i % len(q)].data_ptr(), n_ech)
    ch.flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ch.push_device_i16(q[i % len(q)].data_ptr(), n_ech)
    ch.flush()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    host = [torch.empty(2 * n_ech, dtype=torch.int16).pin_memory() for _ in range(2)]
    for i, h in enumerate(host):
        h.copy_(q[i % len(q)][: 2 * n_ech])
    img = [torch.empty((800, 600), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        ch.push_i16_deliver_ptr(host[i % 2].data_ptr(), n_ech, img[i % 2].data_ptr())
    ch.wait_delivery(0)
    k = max(3, min(steps, 20))
    t0 = time.perf_counter()
    for i in range(k):
        ch.push_i16_deliver_ptr(host[i % 2].data_ptr(), n_ech, img[i % 2].data_ptr())
        if i:
            ch.wait_delivery(1)
    ch.wait_delivery(0)
    ch.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"workload": wl["name"] + " as Int16 (re, im) pairs", "value": frames * S / (ms * 1e-3) / 1e6, "unit": "MS/s",
            "ms_per_step": ms, "steps": steps,
            "e2e": {"value": k * frames * S / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": frames * S * 4,
                    "d2h_bytes_per_step": R * 4, "steps": k}}


def run_integration(args, wl, rank, local_rank, world):
    """cfg 5: a step = one 1000-frame integration.  Rank g takes a contiguous block of frames (parallel.shard_contiguous),
    primes the sync state with its halo frame, runs the chain over its block from a zero accumulator, scales the
    partial image by alpha^(frames after the block) and ONE all-reduce sums the partials (parallel.py).  Total work is
    fixed as N grows: strong scaling."""
    import torch
    import torch.distributed as dist
    import tempestsdr_b200 as tsdr
    from tempestsdr_b200 import parallel
    synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()
    Fs, x_t, y_t, fv, alpha = wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], 0.1
    cfg = tsdr.VideoMode(x_t, y_t, fv)
    S = tsdr.getImageDuration(cfg, Fs)
    per_buf = wl["n_ech"] // S                       # frames per device buffer
    n_ech = per_buf * S
    total = wl["total_frames"]
    k0, k1 = parallel.shard_contiguous(total, world, rank)
    ring = [synth.make_iq_torch(n_ech, Fs, x_t, y_t, fv, dev, seed=500 + 10 * rank + i, t0=i * n_ech) for i in range(wl["ring"])]
    torch.cuda.synchronize()
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    ch = tsdr.Chain(Fs, cfg, alpha=alpha, max_samples=n_ech, device=local_rank, stream=work_stream.cuda_stream)
    acc = parallel.accumulator_tensor(ch)
    weight = parallel.ema_tail_weight(alpha, total - k1)

    def one_integration(push):
        ch.reset()
        if k0 > 0:
            ch.prime_device(ring[-1].data_ptr(), S)   # halo: the frame before this rank's block
        done, i = 0, 0
        while done < k1 - k0:
            f = min(per_buf, k1 - k0 - done)
            push(i, f * S)
            done += f
            i += 1
        ch.flush()
        acc.mul_(weight)
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    dev_push = lambda i, n: ch.push_device(ring[i % len(ring)].data_ptr(), n)
    for _ in range(args.warmup):
        one_integration(dev_push)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ch.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_integration(dev_push)
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ch.launch_count() - l0
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        one_integration(dev_push)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = args.steps * total * S / (elapsed_ms * 1e-3) / 1e6

    ch.set_profiling(True)
    one_integration(dev_push)
    stage_ms, pushes = ch.kernel_times()
    ch.set_profiling(False)
    render_ms = stage_ms[0] / max(pushes, 1)
    algo = (8.0 * S + 4.0 * R) * per_buf
    chain_bytes = (8.0 * S + 12.0 * R) * total
    roofline = {"bound": "hbm", "kernel": "k_render", "achieved": algo / (render_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": algo / (render_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_launch": algo, "kernel_ms_per_launch": render_ms,
                "chain_step": {"algorithmic_bytes": chain_bytes,
                               "achieved": chain_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9,
                               "frac": chain_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9 / hbm_peak / world}}

    # end to end: the same integration fed from pinned host buffers, the combined image read back every step
    host = [torch.empty((n_ech, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host):
        h.copy_(ring[i % len(ring)])
    img = torch.empty(R, dtype=torch.float32).pin_memory()
    host_push = lambda i, n: ch.push_host_ptr(host[i % 2].data_ptr(), n)
    one_integration(host_push)
    barrier()
    k = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        one_integration(host_push)
        img.copy_(acc, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": k * total * S / float(te.item()) / 1e6, "unit": "MS/s", "h2d_bytes_per_step": (k1 - k0) * S * 8,
           "d2h_bytes_per_step": R * 4, "steps": k,
           "note": "per rank: its frame block from pinned host buffers (tsdr_chain_push_host), all-reduce, image to pinned memory"}
    out = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 (f64 coordinates)", "data": "synthetic",
           "config": {"workload": wl["name"], "frames_per_step": total, "samples_per_frame": S,
                      "frames_this_rank": k1 - k0, "frames_per_push": per_buf,
                      "l2_policy": "ring of %d distinct %.0f MB device buffers (> 126 MB L2), no flush" % (len(ring), n_ech * 8 / 1e6),
                      "parallelism": "contiguous frame blocks, halo frame primed, one NCCL all-reduce of 1.92 MB per integration"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    ch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def run_sweep(args, wl, rank, local_rank, world):
    """cfg 4: a step = one 2^26-sample power buffer per GPU: FFT autocorrelation (device resident) and the score of every
    refresh-rate hypothesis of allVideoConfigurations (first maximum of Gamma in each hypothesis' window).  Buffers are
    independent: no collective on the data path (weak scaling); the 13 (rate, score, lag) triples per buffer stay on the
    host of their rank."""
    import torch
    import torch.distributed as dist
    import tempestsdr_b200 as tsdr
    synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()
    Fs, n = wl["Fs"], wl["n_ech"]
    L = n // 2
    ring = []
    for i in range(wl["ring"]):   # abs2.(IQ) as extract_configuration feeds it (src/GUI.jl:70)
        z = synth.make_iq_torch(n, Fs, wl["x_t"], wl["y_t"], wl["fv"], dev, seed=700 + 10 * rank + i, t0=i * n)
        ring.append((z[:, 0] * z[:, 0] + z[:, 1] * z[:, 1]).contiguous())
        del z
    torch.cuda.empty_cache()
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    st = work_stream.cuda_stream
    plan = tsdr.AutocorrPlan(n, device=local_rank, stream=st)
    gamma = torch.empty(L, device=dev, dtype=torch.float32)
    rates = sorted(tsdr.get_refresh_rates(tsdr.allVideoConfigurations))

    def step(x_ptr):
        plan.exec(x_ptr, 1, L, gamma.data_ptr())
        return tsdr.sweep_refresh_hypotheses(gamma.data_ptr(), L, Fs, rates, stream=st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for i in range(args.warmup):
        res = step(ring[i % len(ring)].data_ptr())
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        res = step(ring[i % len(ring)].data_ptr())
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = plan.launch_count() - l0 + 2 * args.steps   # + the two launches of the batched window search per step
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        step(ring[0].data_ptr())
    clocks = sampler.stop()
    barrier()
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    per_2_24 = elapsed_ms / (world * args.steps * (n >> 24))      # whole job: ms per 2^24 samples
    # autocorrelation kernels alone, for the roofline
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        plan.exec(ring[i % len(ring)].data_ptr(), 1, L, gamma.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    fft_ms = e0.elapsed_time(e1) / args.steps
    algo = 4.0 * n + 4.0 * L
    best = max(res, key=lambda r: r[1])
    # end to end: pinned host power buffer -> H2D -> autocorrelation -> sweep -> the triples on the host
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host):
        h.copy_(ring[i % len(ring)])
    xdev = [torch.empty(n, device=dev, dtype=torch.float32) for _ in range(2)]
    k = max(2, min(args.steps, 6))
    barrier()
    t0 = time.perf_counter()
    for i in range(k):
        xdev[i % 2].copy_(host[i % 2], non_blocking=True)
        step(xdev[i % 2].data_ptr())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    out = {"metric": "autocorr ms per 2^24 samples (whole job, incl. the refresh-hypothesis sweep)", "value": per_2_24, "unit": "ms",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
           "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl["name"], "samples_per_step_per_gpu": n, "lags": L, "hypotheses": len(rates),
                      "l2_policy": "ring of %d distinct %.0f MB device buffers (> 126 MB L2), no flush" % (len(ring), n * 4 / 1e6),
                      "parallelism": "one buffer per GPU per step, no collective" if world > 1 else "single GPU",
                      "detected": {"rate_hypothesis": best[0], "fv_hat": best[2], "lag_index": best[3]}},
           "clocks": clocks,
           "e2e": {"value": float(te.item()) * 1e3 / (world * k * (n >> 24)), "unit": "ms", "h2d_bytes_per_step": n * 4,
                   "d2h_bytes_per_step": len(rates) * 12, "steps": k},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k3_p1..p5 (three-level autocorrelation, 5 launches)", "achieved": algo / (fft_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": algo / (fft_ms * 1e-3) / 1e9 / hbm_peak,
                        "traffic": None, "algorithmic_bytes_per_launch": algo, "kernel_ms_per_launch": fft_ms}}
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


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
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip the autocorr / cfg3 / cpu_baseline extras")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if "total_frames" in wl:
        run_integration(args, wl, rank, local_rank, world)
        return
    if wl.get("sweep"):
        run_sweep(args, wl, rank, local_rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import tempestsdr_b200 as tsdr
    synth = _load(os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"), "_tsdr_synth")

    if tsdr.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libtempest_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()

    Fs, x_t, y_t, fv, n_ech = wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], wl["n_ech"]
    cfg = tsdr.VideoMode(x_t, y_t, fv)
    S = tsdr.getImageDuration(cfg, Fs)
    frames = n_ech // S
    # ring of distinct device buffers larger than L2 (126 MB): no buffer is L2-resident when its step starts
    ring = [synth.make_iq_torch(n_ech, Fs, x_t, y_t, fv, dev, seed=100 * rank + i, t0=i * n_ech) for i in range(wl["ring"])]
    # a dedicated (non-default) stream shared by the chain and the timing events: the
    # legacy default stream has handle 0, which the C ABI reads as "create a private stream"
    torch.cuda.synchronize()
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0
    ch = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=n_ech, device=local_rank, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing ------------------------------------------------------------
    for i in range(args.warmup):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ch.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    ch.flush()  # the timed region ends when the last buffer's sync/accumulate kernels have finished too
    e1.record()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ch.launch_count() - l0
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
    barrier()
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * args.steps * frames * S / (elapsed_ms * 1e-3) / 1e6  # whole-job MS/s (samples of complete frames)

    # ---- per-kernel event timing for the roofline (same steps, events between kernels) ------
    ch.set_profiling(True)
    for i in range(args.steps):
        ch.push_device(ring[i % len(ring)].data_ptr(), n_ech)
    stage_ms, pushes = ch.kernel_times()
    ch.set_profiling(False)
    render_ms = stage_ms[0] / max(pushes, 1)
    algo_bytes = (8.0 * S + 4.0 * R) * frames  # k_render: read every complex64 sample once, write one 600x800 frame
    achieved = algo_bytes / (render_ms * 1e-3) / 1e9
    chain_bytes = (8.0 * S + 12.0 * R) * frames  # SURVEY 8(d) B_chain, whole step
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("k_render_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_render (amDemod+sig_to_image+downgradeImage fused)", "achieved": achieved,
                "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                "frac_of_spec_8000": achieved / 8000.0, "traffic": traffic,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_per_launch": render_ms,
                "stage_ms_per_step": {"k_render": render_ms, "k_project+k_sync": stage_ms[1] / max(pushes, 1),
                                      "k_accumulate+carry": stage_ms[2] / max(pushes, 1)},
                "chain_step": {"algorithmic_bytes": chain_bytes,
                               "achieved": chain_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9,
                               "frac": chain_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9 / hbm_peak}}

    # ---- end to end through the host API: pinned host buffer -> H2D -> chain -> D2H image ----
    host_ring = [torch.empty((n_ech, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(host_ring):
        h.copy_(ring[i % len(ring)])
    img_host = [torch.empty((800, 600), dtype=torch.float32).pin_memory() for _ in range(2)]  # column-major 600x800
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(2):
        ch.push_deliver_ptr(host_ring[i % 2].data_ptr(), n_ech, img_host[i % 2].data_ptr())
    ch.wait_delivery(0)
    barrier()
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
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps * frames * S / float(te.item()) / 1e6
    e2e = {"value": e2e_val, "unit": "MS/s", "h2d_bytes_per_step": frames * S * 8, "d2h_bytes_per_step": R * 4,
           "steps": e2e_steps, "note": "pinned host buffer -> tsdr_chain_push_host_deliver (H2D, chain, D2H of imageOut into pinned memory every step), wall clock"}

    out = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (f64 coordinates)", "data": "synthetic",
           "config": {"workload": wl["name"], "frames_per_step": frames, "samples_per_frame": S,
                      "l2_policy": "ring of %d distinct %.0f MB device buffers (> 126 MB L2 in total), no flush"
                                   % (len(ring), n_ech * 8 / 1e6),
                      "parallelism": "one buffer per GPU per step, no collective" if world > 1 else "single GPU"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    if rank == 0 and not args.no_extras:
        # ---- CPU baseline on a bounded sample + parity of the same frames --------------------
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        threads = orc.num_threads()
        cpu_frames = min(frames, max(2, threads))
        iq_host = host_ring[0].numpy().view(np.complex64).reshape(-1)
        v1, dt1, *_ = cpu_chain_baseline(orc, iq_host, wl, min(cpu_frames, 4), 1)
        vN, dtN, img_ref, sy_ref, sx_ref = cpu_chain_baseline(orc, iq_host, wl, cpu_frames, threads)
        chk = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=cpu_frames * S, device=local_rank)
        chk.push(iq_host[: cpu_frames * S])
        sy, sx = chk.offsets()
        same = bool(np.array_equal(chk.image(), img_ref) and np.array_equal(sy, sy_ref) and np.array_equal(sx, sx_ref))
        chk.close()
        out["cpu_baseline"] = {"value": vN, "unit": "MS/s", "cores": threads, "kind": "port",
                               "sample": "%d frames (%d samples) of the workload, OpenMP over frames; 1 thread: %.2f MS/s"
                                         % (cpu_frames, cpu_frames * S, v1),
                               "single_thread_value": v1, "gpu_matches_oracle_bit_exact": same}
        try:
            out["autocorr"] = bench_autocorr(tsdr, torch, dev, hbm_peak)
        except Exception as exc:  # the headline line must still print
            out["autocorr"] = {"error": repr(exc)}
        if world == 1:
            try:
                out["int16_ingest"] = int16_ingest_measure(tsdr, torch, ch, ring, wl, S, frames, min(args.steps, 20), 3)
            except Exception as exc:
                out["int16_ingest"] = {"error": repr(exc)}
        if args.workload != "cfg3" and world == 1:
            # the north-star target is stated on the 200 MS/s stream (BASELINE.json configs[2]): report it beside the headline
            try:
                del ring, host_ring
                torch.cuda.empty_cache()
                out["also"] = {"cfg3": quick_chain_measure(tsdr, torch, synth, WORKLOADS["cfg3"], dev, local_rank, stream,
                                                           steps=10, warmup=3, hbm_peak=hbm_peak)}
            except Exception as exc:
                out["also"] = {"error": repr(exc)}
    ch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


if __name__ == "__main__":
    main()
