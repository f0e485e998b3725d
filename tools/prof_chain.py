"""tools/prof_chain.py -- a few pushes of one workload through the chain, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof python tools/prof_chain.py cfg3 4
    python tools/prof_chain.py cfg3 4 [i16|full]"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import tempestsdr_b200 as tsdr

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "tempestsdr.jl_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

key = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
pushes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
i16 = len(sys.argv) > 3 and sys.argv[3] == "i16"
full = len(sys.argv) > 3 and sys.argv[3] == "full"   # TSDR_CHAIN_FULLRES (cfg5: 25 frames of 4400 x 2250 per push)
wl = dict(bench.WORKLOADS[key])
cfg = tsdr.VideoMode(wl["x_t"], wl["y_t"], wl["fv"])
S = tsdr.getImageDuration(cfg, wl["Fs"])
n_ech = wl.get("n_ech") or wl["frames_per_buf"] * S
dev = torch.device("cuda", 0)
ring = [synth.make_iq_torch(n_ech, wl["Fs"], wl["x_t"], wl["y_t"], wl["fv"], dev, seed=i, t0=i * n_ech) for i in range(2)]
if i16:
    q = []
    for r in ring:
        t = torch.zeros(2 * n_ech + 8, dtype=torch.int16, device=dev)
        t[: 2 * n_ech] = torch.clamp(torch.round(r.reshape(-1) * 2048.0), -32768, 32767).to(torch.int16)
        q.append(t)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ch = tsdr.Chain(wl["Fs"], cfg, alpha=0.1, max_samples=n_ech, stream=st.cuda_stream, full_res=full)
for i in range(pushes):
    if i16:
        ch.push_device_i16(q[i % 2].data_ptr(), n_ech)
    else:
        ch.push_device(ring[i % 2].data_ptr(), n_ech)
ch.sync()
print("pushed", pushes, key, "i16" if i16 else "cf32", "frames/push", n_ech // S)
ch.close()
