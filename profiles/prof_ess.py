"""ESS kernel timing on bench-shaped data (CUDA events on the launching stream).
usage: python profiles/prof_ess.py [series] [S] [reps]
AR(1) series with per-series autocorrelation 0.9 .. 0.995 (first negative lag ~ hundreds, like the bench chains).
Compares the shared-memory FFT kernel (default for S <= 1024) with the direct-summation kernels (ARP_ESS_DIRECT=1)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from autoreparam_b200 import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384 * 51
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
C, D = n // 51, 51
n = C * D
g = torch.Generator(device="cuda").manual_seed(0)
phi = 0.9 + 0.095 * torch.rand(n, device="cuda", generator=g)
x = torch.empty((S, C, D), device="cuda")
cur = torch.randn(n, device="cuda", generator=g) / torch.sqrt(1 - phi * phi)
for t in range(S):
    cur = phi * cur + torch.randn(n, device="cuda", generator=g)
    x[t] = (cur + 3.0).view(C, D)
res = {}
for tag, env in (("fft", "0"), ("direct", "1")):
    os.environ["ARP_ESS_DIRECT"] = env
    out = engine.ess(x, want_moments=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = engine.ess(x, want_moments=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[tag] = out
    gb = x.numel() * 4 / 1e9
    print("%s: %d series x %d samples: min %.3f ms  median %.3f ms  (%.0f GB/s of algorithmic sample bytes)" % (
        tag, n, S, min(ts), float(np.median(ts)), gb / (min(ts) * 1e-3)))
e_f, e_d = res["fft"][0], res["direct"][0]
rel = ((e_f - e_d).abs() / e_d.abs())
print("fft vs direct: ESS rel diff median %.2e  max %.2e;  mean diff %.2e  var rel diff %.2e;  ESS mean %.1f" % (
    rel.median().item(), rel.max().item(), (res["fft"][1] - res["direct"][1]).abs().max().item(),
    ((res["fft"][2] - res["direct"][2]).abs() / res["direct"][2]).max().item(), e_f.mean().item()))
