"""tools/run_autocorr.py -- run the device-resident autocorrelation a few times (for ncu / quick timing).
    python tools/run_autocorr.py [log2n] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tempestsdr_b200 as tsdr

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 1 << k
L = n // 2
dev = torch.device("cuda", 0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ring = [torch.rand(n, device=dev) + 1.0 for _ in range(4)]
out = torch.empty(L, device=dev)
plan = tsdr.AutocorrPlan(n, device=0, stream=s.cuda_stream)
for i in range(3):
    plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(iters):
    plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
e1.record()
torch.cuda.synchronize()
print("n=2^%d: %.4f ms per autocorrelation" % (k, e0.elapsed_time(e1) / iters))
