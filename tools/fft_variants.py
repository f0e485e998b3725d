"""tools/fft_variants.py -- three-level vs two-level autocorrelation kernels at n = 2^LOG2N (default 24), each checked
against a Float64 FFT of the same input.  TSDR_FFT_TWO_LEVEL=1 makes a plan skip the three-level kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tempestsdr_b200 as tsdr

k = int(os.environ.get("LOG2N", "24"))
n = 1 << k
L = n // 2
dev = torch.device("cuda", 0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ring = [torch.rand(n, device=dev) + 1.0 for _ in range(4)]
out = torch.empty(L, device=dev)
X = torch.fft.rfft(ring[0].double())            # Float64 reference: the difference below is this library's own error
r = torch.fft.irfft(X.real * X.real + X.imag * X.imag, n=n)[:L]
ref = (10.0 * torch.log10(r * r)).float()
for rep in range(2):
    for two_level in (0, 1):
        os.environ.pop("TSDR_FFT_TWO_LEVEL", None)
        if two_level:
            os.environ["TSDR_FFT_TWO_LEVEL"] = "1"
        plan = tsdr.AutocorrPlan(n, device=0, stream=s.cuda_stream)
        for i in range(4):
            plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
        torch.cuda.synchronize()
        err = float((out - ref).abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            plan.exec(ring[i % 4].data_ptr(), 1, L, out.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        print("2^%d rep %d %s: %.4f ms  max|dB diff vs Float64 FFT| %.3g"
              % (k, rep, "two-level" if two_level else "default  ", e0.elapsed_time(e1) / 20, err), flush=True)
        plan.close()
