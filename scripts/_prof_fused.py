import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import ProfilerActivity, profile
import problems as P
import torchdiffeq_b200 as tdq
dev = torch.device("cuda:0")
f = tdq.LinearField(P.skew_matrix(128, torch.float32).to(dev))
y0 = torch.randn(65536, 128, generator=torch.Generator().manual_seed(1)).to(dev)
t = torch.tensor([0., 10.], device=dev)
for opts in ({"graph": False}, {"graph": False, "fused_linear": False}):
    def fn():
        with torch.no_grad():
            return tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-7, options=dict(opts))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        fn()
        torch.cuda.synchronize()
    ev = pr.key_averages()
    rows = sorted(((e.key, e.device_time_total / 1e3, e.count) for e in ev if e.device_time_total > 0), key=lambda r: -r[1])
    print(opts, "device ms", sum(r[1] for r in rows))
    for k, tt, c in rows[:12]:
        print("    %-100s %8.3f ms  x%d  avg %.2f us" % (k[:100], tt, c, tt / c * 1e3))
