"""Per-phase host/device timeline of the small-state configs (C1 rk4 spiral, C3 adjoint MLP): wall clock per call,
summed device time, and the kernels that make it up (torch.profiler / CUPTI).  Evidence for profiles/, not a bench."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                     # noqa: E402
from torch.profiler import ProfilerActivity, profile   # noqa: E402
import problems as P             # noqa: E402
import torchdiffeq_b200 as tdq   # noqa: E402

dev = torch.device("cuda:0")


def wall(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def prof(name, fn, reps=3):
    ms = wall(fn)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    ev = pr.key_averages()
    rows = sorted(((e.key, e.device_time_total / reps / 1e3, e.count / reps) for e in ev if e.device_time_total > 0
                   and e.device_type == torch.autograd.DeviceType.CUDA), key=lambda r: -r[1])
    dev_ms = sum(r[1] for r in rows)
    print(json.dumps({"config": name, "wall_ms_per_call": ms, "device_ms_per_call": dev_ms,
                      "kernels_per_call": sum(r[2] for r in rows)}), flush=True)
    for k, t, c in rows[:14]:
        print("    %-90s %8.3f ms  x%.0f" % (k[:90], t, c), flush=True)


def host_profile(name, fn, n=50):
    """Where the host time of a warm call goes (cProfile, cumulative)."""
    import cProfile
    import io
    import pstats
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(28)
    print("---- host profile: %s (%d calls) ----" % (name, n))
    print("\n".join(l for l in buf.getvalue().splitlines() if l.strip())[:6000], flush=True)


def c3():
    f = P.MLPField(dim=64, hidden=256, seed=0).to(dev)
    y0 = torch.randn(8192, 64, generator=torch.Generator().manual_seed(1)).to(dev)
    t = torch.tensor([0., 1.], device=dev)

    def fwd_only():
        with torch.no_grad():
            tdq.odeint(f, y0, t, method="dopri5", rtol=1e-4, atol=1e-6)

    def step():
        y = y0.clone().requires_grad_(True)
        f.zero_grad()
        out = tdq.odeint_adjoint(f, y, t, method="dopri5", rtol=1e-4, atol=1e-6)
        out[-1].pow(2).mean().backward()
    prof("C3 forward only (no_grad odeint)", fwd_only)
    prof("C3 odeint_adjoint fwd+bwd", step)
    host_profile("C3 forward only", fwd_only)
    host_profile("C3 odeint_adjoint fwd+bwd", step)


def c1():
    f = P.Spiral().to(dev)
    g = torch.Generator().manual_seed(0)
    y0 = (torch.tensor([[2., 0.]]) * (1 + 0.1 * torch.rand(1024, 1, generator=g))).to(dev)
    t = torch.linspace(0., 25., 1000).to(dev)

    def run():
        with torch.no_grad():
            tdq.odeint(f, y0, t, method="rk4")
    prof("C1 rk4 spiral 999 steps", run)


if __name__ == "__main__":
    c3()
    c1()
