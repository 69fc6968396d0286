"""Pin the CPU oracle (oracle/ode_oracle.py) to vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

import problems as P
from oracle import ode_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ld = lambda name: torch.load(os.path.join(G, name), weights_only=False)


def _close(a, b, dtype, scale=1.0):
    # The oracle sums stage terms in index order, the reference through torch.sum's blocked order
    # (which also depends on the host's SIMD width).  Wherever the embedded error estimate is itself
    # rounding noise (first tiny step, tolerances near the dtype's precision) that changes the step
    # sequence, so solutions agree to the solver's accuracy, not bitwise.
    tol = (5e-4 if dtype == torch.float32 else 1e-6) * scale
    return torch.allclose(a, b, rtol=tol, atol=tol * 1e-2)


def _nfe_close(a, b, dtype=torch.float64):
    if dtype == torch.float32:          # tolerances at/below float32 precision: step count is noise-driven
        return b / 2 - 30 <= a <= 2 * b + 30
    return abs(a - b) <= max(30, (4 * b) // 10)


ZOO = ld("zoo.pt")
FIXED = ("rk4", "euler", "midpoint", "heun2", "heun3")


@pytest.mark.parametrize("key", sorted(ZOO))
def test_zoo(key):
    ode, method, dt, direction = key.split("/")
    dtype = getattr(torch, dt)
    case = ZOO[key]
    f, y0, t, _ = P.construct_problem("cpu", ode=ode, reverse=direction == "rev", dtype=dtype)
    cf = O.Counter(f)
    with torch.no_grad():
        if method in FIXED:
            y = O.odeint_fixed(cf, y0, t, method)
        else:
            rec = {}
            kw = {"rtol": case["kw"].get("rtol", 1e-7), "atol": case["kw"].get("atol", 1e-9)}
            y = O.odeint_adaptive(cf, y0, t, method, record=rec, **kw)
    assert _close(y, case["y"], dtype), (y - case["y"]).abs().max()
    if method in FIXED:
        assert torch.equal(y, case["y"])          # fixed grid, same op order: bitwise
        assert cf.nfe == case["nfe"]
    else:
        assert _nfe_close(cf.nfe, case["nfe"], dtype)
        # the reference's own acceptance test (odeint_tests.py:45-58): relative error against the exact solution
        eps = {"constant": 3e-4, "sine": 3e-4, "linear": 2e-3, "exp": 5e-2}[ode]
        if method in ("adaptive_heun", "fehlberg2", "bosh3"):
            eps = {"constant": 1e-3, "sine": 5e-3, "linear": 2e-3, "exp": 5e-2}[ode]
        rel = ((case["exact"] - y) / case["exact"]).abs().max()
        assert rel < eps, rel


@pytest.mark.parametrize("key", sorted(ld("linear_batch.pt")))
def test_linear_batch(key):
    case = ld("linear_batch.pt")[key]
    dtype = getattr(torch, key.split("/")[1])
    f = P.BatchedLinear(128, dtype)
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(64, 128, generator=g).to(dtype)
    rec = {}
    cf = O.Counter(f)
    with torch.no_grad():
        y = O.odeint_adaptive(cf, y0, case["t"], "dopri5", rtol=1e-5, atol=1e-7, record=rec)
    assert _close(y, case["y"], dtype)
    assert _nfe_close(cf.nfe, case["nfe"])
    if dtype == torch.float64:
        assert cf.nfe == case["nfe"] and rec["accepted"] == case["acc"]
        assert torch.allclose(torch.tensor(rec["dts"]), torch.tensor(case["dts"]), rtol=1e-4, atol=0)


def test_spiral_rk4():
    case = ld("spiral_rk4.pt")
    f = P.Spiral()
    with torch.no_grad():
        y = O.odeint_rk4(f, case["y0"], torch.linspace(0., 25., 1000))
        y2 = O.odeint_rk4(f, case["y0"][:16], case["t2"], grid=_grid(case["t2"], 0.03))
    # same op order as the reference: bitwise
    assert torch.equal(y[case["rows"]], case["y_rows"])
    assert torch.equal(y2, case["y2"])
    for key, want in case["fixed"].items():       # every explicit fixed-grid method, with and without perturb
        method, perturb = key.split("/")
        with torch.no_grad():
            got = O.odeint_fixed(f, case["y0"][:16], case["t2"], method, grid=_grid(case["t2"], 0.03),
                                 perturb=bool(int(perturb)))
        # explicit Euler blows up on the cubic spiral at this step size: NaNs must match too
        assert torch.allclose(got, want, rtol=0, atol=0, equal_nan=True), key


def _grid(t, step_size):
    niters = torch.ceil((t[-1] - t[0]) / step_size + 1).item()
    g = torch.arange(0, niters, dtype=t.dtype) * step_size + t[0]
    g[-1] = t[-1]
    return g


@pytest.mark.parametrize("key", sorted(ld("adjoint_mlp.pt")))
def test_adjoint(key):
    case = ld("adjoint_mlp.pt")[key]
    name, norm, dt = key.split("/")
    dtype = getattr(torch, dt)
    f = P.MLPField(dim=8, hidden=16, seed=0, dtype=dtype)
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(32, 8, generator=g).to(dtype)
    t = case["t"]
    # dL/dy(t_i) of loss = mean(y[-1]^2) + 0.01*sum(y[1]) (multi)
    gy = torch.zeros(len(t), 32, 8, dtype=dtype)
    gy[-1] = 2 * case["y"][-1] / case["y"][-1].numel()
    if len(t) > 2:
        gy[1] += 0.01
    ys, gy0, gp = O.adjoint_gradients(f, list(f.parameters()), y0, t, gy, "dopri5", rtol=1e-6, atol=1e-8,
                                      seminorm=norm == "seminorm")
    tol = 2e-4 if dtype == torch.float32 else 1e-8
    assert torch.allclose(ys, case["y"], rtol=tol, atol=tol)
    assert torch.allclose(gy0, case["gy0"], rtol=tol, atol=tol * 1e-2)
    for a, b in zip(gp, case["gp"]):
        assert torch.allclose(a, b, rtol=tol, atol=tol * 1e-2), (a - b).abs().max()


DET = ld("detest.pt")


DET_KEYS = sorted(k for k in DET if not k.endswith("/truth"))


@pytest.mark.parametrize("key", DET_KEYS)
def test_detest(key):
    """All 25 DETEST problems (tests/DETEST/detest.py:8-315, run.py:22-55): NFE against the reference's, y(20) against
    the reference's, and the RMS error against dopri5 @ 1e-12 that run.py:47 reports."""
    name, method, tol = key.split("/")
    tol = float(tol)
    f, y0, t0 = P.detest(name)
    cf = O.Counter(f)
    with torch.no_grad():
        y = O.odeint_adaptive(cf, y0, torch.tensor([t0, 20.0], dtype=torch.float64), method, rtol=tol, atol=tol)
    # BASELINE.md's known-answer NFE table: equal up to one or two noise-decided accept/reject flips
    S = 6 if method == "dopri5" else 13
    band = max(2 * S, DET[key]["nfe"] // 20)
    if tol <= 1e-12:          # the embedded error estimate sits in float64 rounding noise: the sum order decides steps
        band = max(4 * S, DET[key]["nfe"] // 8)
    assert abs(cf.nfe - DET[key]["nfe"]) <= band, (cf.nfe, DET[key]["nfe"])
    # y(20) is INTERPOLATED inside the last step (rk_common.py:250) by a 4th-order polynomial, so for
    # dopri8's long steps it is only as accurate as that interpolant and moves with the step sequence
    ytol = max(100 * tol, 1e-3 if method == "dopri8" else 1e-6)
    scale = max(1.0, float(DET[key]["y"].abs().max()))           # C5 carries one coordinate of 1.7e11 (detest.py:219)
    assert torch.allclose(y[-1], DET[key]["y"], rtol=ytol, atol=ytol * scale)
    err = float(torch.sqrt(torch.mean((DET[name + "/truth"]["y"] - y[-1]) ** 2)))
    assert err <= 10 * DET[key]["err"] + 1e-9 * scale, (err, DET[key]["err"])


@pytest.mark.parametrize("key", ["min_step", "max_step", "first_step", "step_t", "factors"])
def test_options(key):
    case = ld("options.pt")[key]
    f, y0, t, _ = P.construct_problem("cpu", ode="linear", dtype=torch.float64)
    cf = O.Counter(f)
    rec = {}
    with torch.no_grad():
        y = O.odeint_adaptive(cf, y0, t, "dopri5", record=rec, **case["opts"])
    assert _nfe_close(cf.nfe, case["nfe"])
    if key in ("min_step", "max_step", "step_t"):
        assert cf.nfe == case["nfe"] and rec["accepted"] == case["acc"]
    assert torch.allclose(y, case["y"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("key", sorted(k for k in ld("options.pt") if k.startswith("jump/")))
def test_jump_t(key):
    """rk_common.py:302-308, :346-351 against the reference (odeint_tests.py:126-161)."""
    case = ld("options.pt")[key]
    _, method, dt = key.split("/")
    dtype = getattr(torch, dt)
    x0 = torch.tensor([1.0, 2.0], dtype=dtype)
    tj = torch.tensor([0., 1.0])
    f = P.JumpField()
    with torch.no_grad():
        y = O.odeint_adaptive(f, x0, tj, method, rtol=1e-6, atol=1e-6, jump_t=torch.tensor([0.5]))
    # float32 at rtol 1e-6 sits on the precision floor: a noise-decided reject may differ (see _close)
    assert f.nfe == case["nfe_jump"] if dtype == torch.float64 else abs(f.nfe - case["nfe_jump"]) <= 24
    assert f.nfe < case["nfe_plain"]
    assert torch.allclose(y, case["y_jump"], rtol=1e-5 if dtype == torch.float32 else 1e-9, atol=1e-6)


EV = ld("events.pt")


@pytest.mark.parametrize("key", sorted(k for k in EV if k.count("/") == 3))
def test_events(key):
    """solvers.py:41-49 + rk_common.py:252-264 + event_handling.py:5-20 against the reference (event_tests.py:14-49)."""
    ode, method, dt, direction = key.split("/")
    dtype = getattr(torch, dt)
    case = EV[key]
    f, y0, t, sol = P.construct_problem("cpu", ode=ode, reverse=direction == "rev", dtype=dtype)
    target = case["target"]
    cf = O.Counter(f)
    with torch.no_grad():
        et, ys = O.odeint_adaptive(cf, y0, t[0:2], method, event_fn=lambda t_, y_: torch.sum(y_ - target).real)
    tol = 1e-4                                                     # event_tests.py:33
    assert ((case["t2"] - et) / case["t2"]).abs() < tol and ((target - ys[-1]) / target).abs().max() < tol
    close = 1e-4 if dtype == torch.float32 else 1e-7
    assert abs(float(et) - float(case["event_t"])) <= close * abs(float(case["event_t"]))
    assert torch.allclose(ys, case["y"], rtol=close, atol=close)
    if dtype == torch.float64 and method != "dopri8":
        assert cf.nfe == case["nfe"]


FX = ld("fixed_extra.pt")


@pytest.mark.parametrize("key", sorted(k for k in FX if k.startswith("cubic")))
def test_fixed_cubic_interp(key):
    """interp='cubic' (solvers.py:120-125, :166-173): same op order as the reference => bitwise, same NFE."""
    case = FX[key]
    parts = key.split("/")
    if parts[0] == "cubic":
        f = P.Spiral()
        g = torch.Generator().manual_seed(0)
        y0 = (torch.tensor([[2., 0.]]) * (1 + 0.1 * torch.rand(1024, 1, generator=g)))[:16]
        t = torch.linspace(0., 5., 7)
        cf = O.Counter(f)
        with torch.no_grad():
            y = O.odeint_fixed(cf, y0, t, parts[1], grid=_grid(t, 0.03), perturb=bool(int(parts[2])), interp="cubic")
    else:
        dtype = getattr(torch, parts[1])
        f, y0, t, _ = P.construct_problem("cpu", ode="constant", reverse=parts[2] == "rev", dtype=dtype)
        cf = O.Counter(f)
        ta = -t if parts[2] == "rev" else t
        with torch.no_grad():
            y = O.odeint_fixed(cf, y0, t, "rk4", grid=(-_grid(ta, 0.1) if parts[2] == "rev" else _grid(ta, 0.1)),
                               interp="cubic")
    assert torch.allclose(y, case["y"], rtol=0, atol=0, equal_nan=True), (y - case["y"]).abs().max()
    assert cf.nfe == case["nfe"]


@pytest.mark.parametrize("key", sorted(k for k in FX if k.startswith("event/")))
def test_fixed_event(key):
    """Event handling with the fixed-grid methods (solvers.py:130-164, event_tests.py:14-49)."""
    _, ode, method, dt, direction, interp = key.split("/")
    dtype = getattr(torch, dt)
    case = FX[key]
    f, y0, t, sol = P.construct_problem("cpu", ode=ode, reverse=direction == "rev", dtype=dtype)
    target = sol[2]
    cf = O.Counter(f)
    with torch.no_grad():
        et, ys = O.odeint_fixed_event(cf, y0, t[0], lambda t_, y_: torch.sum(y_ - target).real, method, 0.01, interp=interp,
                                      atol=1e-9, reverse=direction == "rev")
    assert torch.equal(ys, case["y"]) and float(et) == float(case["event_t"]), (et, case["event_t"])
    assert cf.nfe == case["nfe"]


AD = ld("adams.pt")


@pytest.mark.parametrize("key", sorted(k for k in AD if k.count("/") == 4))
def test_adams(key):
    """explicit_adams / implicit_adams (fixed_adams.py:164-228): same op order as the reference => bitwise, same NFE."""
    ode, method, dt, direction, name = key.split("/")
    dtype = getattr(torch, dt)
    case = AD[key]
    f, y0, t, _ = P.construct_problem("cpu", ode=ode, reverse=direction == "rev", dtype=dtype)
    opts = case["opts"] or {}
    grid = None
    if "step_size" in opts:
        ta = -t if direction == "rev" else t
        grid = _grid(ta, opts["step_size"])
        grid = -grid if direction == "rev" else grid
    cf = O.Counter(f)
    with torch.no_grad():
        y = O.odeint_adams(cf, y0, t, implicit=method == "implicit_adams", grid=grid, interp=opts.get("interp", "linear"))
    assert torch.equal(y, case["y"]), (y - case["y"]).abs().max()
    assert cf.nfe == case["nfe"]


def test_adams_weights_match_reference():
    """The product generates the Adams weights as exact rationals (torchdiffeq_b200/_adams.py); they must be the
    reference's float64 values bit for bit (tests/golden/adams.json, dumped from fixed_adams.py:10-141)."""
    from torchdiffeq_b200 import _adams as A
    tabs = O.adams_tables()
    for k in range(2, 13):                       # order 1's Moulton entry is 1/11 in the reference (a typo it never uses)
        assert A._BASHFORTH[k] == tabs["bashforth"][k], k
        assert A._MOULTON[k] == tabs["moulton"][k], k
    assert A._BASHFORTH[1] == tabs["bashforth"][1]
