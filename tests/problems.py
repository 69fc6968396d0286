"""Problem zoo for the parity tests.

The four analytic problems and their parameters follow the reference's tests/problems.py:7-61
(ConstantODE a=0.2 b=3.0, SineODE, LinearODE with the seeded skew construction, ExpODE); the DETEST
entries are the 25 published non-stiff test problems of Hull, Enright, Fellen & Sedgwick (1972), classes
A-E, which the reference carries in tests/DETEST/detest.py:8-315 (checked value-for-value against that file
by tests/golden/make_golden.py).  They take an optional trailing batch dimension, so the same definition
serves the oracle on CPU and BASELINE config 4's batched form on the CUDA path.
"""
import math

import torch


class ConstantODE(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.2))
        self.b = torch.nn.Parameter(torch.tensor(3.0))

    def forward(self, t, y):
        return self.a + (y - (self.a * t + self.b)) ** 5

    def y_exact(self, t):
        return self.a * t + self.b


class SineODE(torch.nn.Module):
    def forward(self, t, y):
        return 2 * y / t + t ** 4 * torch.sin(2 * t) - t ** 2 + 4 * t ** 3

    def y_exact(self, t):
        return (-0.5 * t ** 4 * torch.cos(2 * t) + 0.5 * t ** 3 * torch.sin(2 * t) + 0.25 * t ** 2 * torch.cos(2 * t)
                - t ** 3 + 2 * t ** 4 + (math.pi - 0.25) * t ** 2)


def skew_matrix(dim, dtype=torch.float32):
    """The matrix of the reference's LinearODE (problems.py:35-38): seed 0, U = 0.1*randn, A = 2U - (U + U^T)."""
    g = torch.Generator().manual_seed(0)
    U = torch.randn(dim, dim, generator=g) * 0.1
    return (2 * U - (U + U.transpose(0, 1))).to(dtype)


class LinearODE(torch.nn.Module):
    """dy/dt = A y for one trajectory of dimension `dim` (problems.py:31-53); counts evaluations."""

    def __init__(self, dim=10):
        super().__init__()
        self.dim = dim
        self.A = torch.nn.Parameter(skew_matrix(dim))
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        return torch.mm(self.A, y.reshape(self.dim, 1)).reshape(-1)

    def y_exact(self, t):
        A = self.A.detach().double().cpu()
        y0 = torch.ones(self.dim, 1, dtype=torch.float64)
        return torch.stack([torch.matrix_exp(A * float(ti)) @ y0 for ti in t.detach().cpu()]).reshape(
            len(t), self.dim).to(t)


class ExpODE(torch.nn.Module):
    def forward(self, t, y):
        return -0.1 * self.y_exact(t)

    def y_exact(self, t):
        return torch.exp(-0.1 * t)


class JumpField:
    """A vector field with a discontinuity at t = 0.5 (the reference's _JumpF, odeint_tests.py:113-123);
    branches on the host, so it is only usable in lock-step mode."""

    def __init__(self):
        self.nfe = 0

    def __call__(self, t, x):
        self.nfe += 1
        if t < 0.5:
            return -0.5 * x
        return x ** 2


PROBLEMS = {"constant": ConstantODE, "linear": LinearODE, "sine": SineODE, "exp": ExpODE}


def construct_problem(device, npts=10, ode="constant", reverse=False, dtype=torch.float64):
    """Same recipe as problems.py:79-95: t = linspace(1, 8, npts) float64, y0 = exact(t[0])."""
    f = PROBLEMS[ode]().to(dtype=dtype, device=device)
    t_points = torch.linspace(1, 8, npts, dtype=torch.float64, device=device)
    sol = f.y_exact(t_points).to(dtype)
    if reverse:
        t_points = t_points.flip(0).clone()
        sol = sol.flip(0).clone()
    return f, sol[0].detach().clone(), t_points, sol


# ---- batched benchmark problems ------------------------------------------
class BatchedLinear(torch.nn.Module):
    """C2: dy/dt = y A^T for a batch of trajectories [B, D] with the skew matrix above (norm preserving)."""

    def __init__(self, dim=128, dtype=torch.float32):
        super().__init__()
        self.register_buffer("At", skew_matrix(dim, dtype).t().contiguous())

    def forward(self, t, y):
        return y @ self.At


class Spiral(torch.nn.Module):
    """C1: the cubic spiral of examples/ode_demo.py:31-37, dy/dt = (y**3) A."""

    def __init__(self, dtype=torch.float32):
        super().__init__()
        self.register_buffer("A", torch.tensor([[-0.1, 2.0], [-2.0, -0.1]], dtype=dtype))

    def forward(self, t, y):
        return torch.mm(y ** 3, self.A)


class MLPField(torch.nn.Module):
    """C3-shaped Neural-ODE vector field: Linear-Tanh-Linear-Tanh-Linear, default torch init, seed given."""

    def __init__(self, dim=64, hidden=256, seed=0, dtype=torch.float32):
        super().__init__()
        torch.manual_seed(seed)
        self.net = torch.nn.Sequential(
            torch.nn.Linear(dim, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, hidden), torch.nn.Tanh(),
            torch.nn.Linear(hidden, dim)).to(dtype)

    def forward(self, t, y):
        return self.net(y)


class DeepField(torch.nn.Module):
    """A vector field with MANY parameter tensors (2*depth of them): the adjoint's default norm then has one segment
    per tensor (adjoint.py:247-250) -- far more than any by-value descriptor holds."""

    def __init__(self, dim=6, depth=40, seed=0, dtype=torch.float64):
        super().__init__()
        torch.manual_seed(seed)
        self.layers = torch.nn.ModuleList([torch.nn.Linear(dim, dim) for _ in range(depth)]).to(dtype)

    def forward(self, t, y):
        h = y
        for lin in self.layers:
            h = h + 0.1 * torch.tanh(lin(h))
        return h - y


# ---- DETEST (Hull, Enright, Fellen & Sedgwick 1972), all 25 problems of tests/DETEST/detest.py:8-315 --------------
# Written so that ONE definition serves the unbatched form the reference uses (y of shape [d], [] for class A,
# [2, 3, 5] for C5) and BASELINE config 4's batched form with a TRAILING batch dimension (y of shape [d, B]):
# they index y[i] and stack on dim 0, matrices act from the left.  Constants are float64 (run.py:8 makes float64
# the default dtype) and follow the state to its device.
def _banded(n, diag, lower, upper=None):
    A = torch.zeros(n, n, dtype=torch.float64)
    A.view(-1)[::n + 1] = diag
    A.view(-1)[n::n + 1] = lower
    if upper is not None:
        A.view(-1)[1::n + 1] = upper
    return A


class _MatField:
    """dy/dt = A y (detest.py:58-167: torch.mv(A, y)); A @ y also takes y [d, B]."""

    def __init__(self, A):
        self.A, self._dev = A, {}

    def __call__(self, t, y):
        A = self._dev.get(y.device)
        if A is None:
            A = self._dev[y.device] = self.A.to(y.device)
        return A @ y


class _FiveBody:
    """C5, the five outer planets (detest.py:170-214), for y of shape [2, 3, 5] or [2, 3, 5, B]."""
    k2 = 2.95912208286
    m0 = 1.00000597682
    m = [0.000954786104043, 0.000285583733151, 0.0000437273164546, 0.0000517759138449, 0.00000277777777778]

    def __init__(self):
        self._dev = {}

    def __call__(self, t, y):
        shape = y.shape
        yb = y.reshape(2, 3, 5, -1)
        m = self._dev.get(y.device)
        if m is None:
            m = self._dev[y.device] = torch.tensor(self.m, dtype=torch.float64, device=y.device)
        dy, p = yb[1], yb[0]                                            # [3, 5, B]
        r = torch.sqrt(torch.sum(p ** 2, 0))                            # [5, B]
        d = torch.sqrt(torch.sum((p[:, :, None] - p[:, None, :]) ** 2, 0))      # [5, 5, B]
        F = m.view(1, 1, 5, 1) * ((p[:, None, :] - p[:, :, None]) / (d * d * d)[None] + p[:, None, :] / (r * r * r)[None, None])
        eye = torch.eye(5, dtype=torch.bool, device=y.device).view(1, 5, 5, 1)
        F = torch.where(eye, torch.zeros((), dtype=F.dtype, device=F.device), F)     # F.view(3, 25)[:, ::6] = 0
        ddy = self.k2 * (-(self.m0 + m.view(1, 5, 1)) * p / (r * r * r)[None]) + F.sum(2)
        return torch.stack([dy, ddy], 0).reshape(shape)


def _c5_init():
    y0 = torch.tensor([
        3.42947415189, 3.35386959711, 1.35494901715, 6.64145542550, 5.97156957878, 2.18231499728, 11.2630437207,
        14.6952576794, 6.27960525067, -30.1552268759, 165699966404, 1.43785752721, -21.1238353380, 28.4465098142,
        15.388265967], dtype=torch.float64).view(5, 3).transpose(0, 1)          # the 165699966404 is the reference's (detest.py:219)
    dy0 = torch.tensor([
        -.557160570446, .505696783289, .230578543901, -.415570776342, .365682722812, .169143213293, -.325325669158,
        .189706021964, .0877265322780, -.0240476254170, -.287659532608, -.117219543175, -.176860753121,
        -.216393453025, -.0148647893090], dtype=torch.float64).view(5, 3).transpose(0, 1)
    return torch.stack([y0, dy0], 0).contiguous()


def _orbit(eps):
    def f(t, y):
        r = (y[0] ** 2 + y[1] ** 2) ** (3 / 2)
        return torch.stack([y[2], y[3], -y[0] / r, -y[1] / r])
    return f, [1 - eps, 0, 0, math.sqrt((1 + eps) / (1 - eps))]


def detest(name):
    """Returns (func, y0 float64 tensor in the reference's shape, t0).  Integrate to t = 20 (run.py:22-55)."""
    T = lambda v: torch.tensor(v, dtype=torch.float64)
    if name == "A1":
        return (lambda t, y: -y), T(1.0), 0.0
    if name == "A2":
        return (lambda t, y: -y ** 3 / 2), T(1.0), 0.0
    if name == "A3":
        return (lambda t, y: y * torch.cos(t)), T(1.0), 0.0
    if name == "A4":
        return (lambda t, y: y / 4 * (1 - y / 20)), T(1.0), 0.0
    if name == "A5":
        return (lambda t, y: (y - t) / (y + t)), T(4.0), 0.0
    if name == "B1":
        def f(t, y):
            return torch.stack([2 * (y[0] - y[0] * y[1]), -(y[1] - y[0] * y[1])])
        return f, T([1.0, 3.0]), 0.0
    if name == "B2":
        return _MatField(T([[-1., 1., 0.], [1., -2., 1.], [0., 1., -1.]])), T([2., 0., 1.]), 0.0
    if name == "B3":
        def f(t, y):
            return torch.stack([-y[0], y[0] - y[1] * y[1], y[1] * y[1]])
        return f, T([1.0, 0.0, 0.0]), 0.0
    if name == "B4":
        def f(t, y):
            a = torch.sqrt(y[0] * y[0] + y[1] * y[1])
            return torch.stack([-y[1] - y[0] * y[2] / a, y[0] - y[1] * y[2] / a, y[0] / a])
        return f, T([3.0, 0.0, 0.0]), 0.0
    if name == "B5":
        def f(t, y):
            return torch.stack([y[1] * y[2], -y[0] * y[2], -0.51 * y[0] * y[1]])
        return f, T([0.0, 1.0, 1.0]), 0.0
    if name in ("C1", "C2", "C3", "C4"):
        n = 51 if name == "C4" else 10
        if name == "C1":
            A = torch.zeros(10, 10, dtype=torch.float64)
            A.view(-1)[:-1:11] = -1
            A.view(-1)[10::11] = 1
        elif name == "C2":
            A = torch.zeros(10, 10, dtype=torch.float64)
            A.view(-1)[:-1:11] = torch.linspace(-1, -9, 9, dtype=torch.float64)
            A.view(-1)[10::11] = torch.linspace(1, 9, 9, dtype=torch.float64)
        else:
            A = _banded(n, -2, 1, 1)
        y0 = torch.zeros(n, dtype=torch.float64)
        y0[0] = 1
        return _MatField(A), y0, 0.0
    if name == "C5":
        return _FiveBody(), _c5_init(), 0.0
    if name in ("D1", "D2", "D3", "D4", "D5"):
        f, y0 = _orbit({"D1": 0.1, "D2": 0.3, "D3": 0.5, "D4": 0.7, "D5": 0.9}[name])
        return f, T(y0), 0.0
    if name == "E1":
        def f(t, y):
            return torch.stack([y[1], -(y[1] / (t + 1) + (1 - 0.25 / (t + 1) ** 2) * y[0])])
        return f, T([.671396707141803, .0954005144474744]), 0.0
    if name == "E2":
        def f(t, y):
            return torch.stack([y[1], (1 - y[0] ** 2) * y[1] - y[0]])
        return f, T([2., 0.]), 0.0
    if name == "E3":
        def f(t, y):
            return torch.stack([y[1], y[0] ** 3 / 6 - y[0] + 2 * torch.sin(2.78535 * t)])
        return f, T([0., 0.]), 0.0
    if name == "E4":
        def f(t, y):
            return torch.stack([y[1], .32 - .4 * y[1] ** 2])
        return f, T([30., 0.]), 0.0
    if name == "E5":
        def f(t, y):
            return torch.stack([y[1], torch.sqrt(1 + y[1] ** 2) / (25 - t)])
        return f, T([0., 0.]), 0.0
    raise KeyError(name)


DETEST_NAMES = tuple(c + i for c in "ABCDE" for i in "12345")
