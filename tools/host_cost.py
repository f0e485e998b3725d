"""tools/host_cost.py -- host-side cost of one tsdr_chain_push_device (5 launches + the stream events of the two-stream
pipeline) next to the device time per push, for 30-frame and 3-frame buffers.  B200: 24 us of host time per push against
103 us / 42 us on the device, so the chain stays device-bound without CUDA graphs."""
import sys, os, time, importlib.util
sys.path.insert(0, os.getcwd())
import torch, tempestsdr_b200 as tsdr
spec = importlib.util.spec_from_file_location("synth", "tempestsdr.jl_b200/synth.py"); synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
dev = torch.device("cuda:0")
Fs, x, y, fv = 20e6, 2576, 1125, 60.0
cfg = tsdr.VideoMode(x, y, fv); S = tsdr.getImageDuration(cfg, Fs)
for frames in (30, 3):
    n = frames * S
    buf = synth.make_iq_torch(n, Fs, x, y, fv, dev, seed=1)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    ch = tsdr.Chain(Fs, cfg, alpha=0.1, max_samples=n, stream=st.cuda_stream)
    for _ in range(5): ch.push_device(buf.data_ptr(), n)
    torch.cuda.synchronize()
    K = 200
    t0 = time.perf_counter()
    for _ in range(K): ch.push_device(buf.data_ptr(), n)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("frames/push %d: host enqueue %.1f us per push, total %.1f us per push" % (frames, (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6))
    ch.close()
