"""Device-side while loop (tdq_loop_create) smoke check: the same solve in lock step, host-replayed graph and
device loop must give bitwise identical results and step counts.  Run under `timeout`: a loop whose condition is
never cleared would spin on the GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                     # noqa: E402
import problems as P             # noqa: E402
import torchdiffeq_b200 as tdq   # noqa: E402

dev = torch.device("cuda:0")
f = P.BatchedLinear(128).to(dev)
y0 = torch.randn(4096, 128, generator=torch.Generator().manual_seed(1)).to(dev)
out = {}
for t in (torch.tensor([0., 2.], device=dev), torch.linspace(0, 2, 9).to(dev)):
    for name, opts in (("lockstep", {"run_ahead": 0, "graph": False}),
                       ("replay", {"graph": True, "device_loop": False}),
                       ("loop", {"graph": True, "device_loop": True})):
        st = {}
        for rep in range(3):         # first call captures, later calls reuse the cached engine + loop
            t0 = time.perf_counter()
            with torch.no_grad():
                y = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-7, options=dict(opts), _stats=st)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
        out[name] = (y, dict(st))
        print(len(t), name, "ms=%.2f" % ms, st, flush=True)
    a = out["lockstep"]
    for name in ("replay", "loop"):
        b = out[name]
        assert torch.equal(a[0], b[0]), name
        assert (a[1]["n_accept"], a[1]["n_reject"]) == (b[1]["n_accept"], b[1]["n_reject"]), name
    assert out["loop"][1]["nfe"] == 6 * out["loop"][1]["attempts"] + 2, out["loop"][1]
# the adjoint's backward loop
fm = P.MLPField(dim=8, hidden=16, seed=0).to(dev)
res = {}
for name, opts in (("lockstep", {"run_ahead": 0, "graph": False}), ("loop", {"graph": True, "device_loop": True})):
    yy = torch.randn(32, 8, generator=torch.Generator().manual_seed(1)).to(dev).requires_grad_(True)
    fm.zero_grad()
    o = tdq.odeint_adjoint(fm, yy, torch.tensor([0., 0.5, 1.], device=dev), method="dopri5", rtol=1e-6, atol=1e-8,
                           options=dict(opts))
    (o[-1].pow(2).mean() + 0.01 * o[1].sum()).backward()
    res[name] = (yy.grad.clone(), [q.grad.clone() for q in fm.parameters()])
assert torch.allclose(res["lockstep"][0], res["loop"][0], rtol=1e-6, atol=1e-9)
for a, b in zip(res["lockstep"][1], res["loop"][1]):
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)
print("loop_check OK")
