"""Problem zoo for the parity tests.

The four analytic problems and their parameters follow the reference's tests/problems.py:7-61
(ConstantODE a=0.2 b=3.0, SineODE, LinearODE with the seeded skew construction, ExpODE); the DETEST
entries are the published non-stiff test problems of Hull, Enright, Fellen & Sedgwick (1972), classes
A and B, which the reference carries in tests/DETEST/detest.py:8-117.  They are written batch-first so the
same definition serves the reference on CPU (golden generation), the oracle and the CUDA path.
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


# ---- DETEST classes A and B, trailing batch dimension: y has shape [d, B] (or [B] for class A) -----------
def detest(name):
    """Returns (func, y0 [d] float64 list, t0).  Integrate to t = 20 (run.py:22-55)."""
    if name == "A1":
        return (lambda t, y: -y), [1.0], 0.0
    if name == "A2":
        return (lambda t, y: -y ** 3 / 2), [1.0], 0.0
    if name == "A3":
        return (lambda t, y: y * torch.cos(t)), [1.0], 0.0
    if name == "A4":
        return (lambda t, y: y / 4 * (1 - y / 20)), [1.0], 0.0
    if name == "A5":
        return (lambda t, y: (y - t) / (y + t)), [4.0], 0.0
    if name == "B1":
        def f(t, y):
            return torch.stack([2 * (y[0] - y[0] * y[1]), -(y[1] - y[0] * y[1])])
        return f, [1.0, 3.0], 0.0
    if name == "B3":
        def f(t, y):
            return torch.stack([-y[0], y[0] - y[1] * y[1], y[1] * y[1]])
        return f, [1.0, 0.0, 0.0], 0.0
    if name == "B4":
        def f(t, y):
            a = torch.sqrt(y[0] * y[0] + y[1] * y[1])
            return torch.stack([-y[1] - y[0] * y[2] / a, y[0] - y[1] * y[2] / a, y[0] / a])
        return f, [3.0, 0.0, 0.0], 0.0
    if name == "B5":
        def f(t, y):
            return torch.stack([y[1] * y[2], -y[0] * y[2], -0.51 * y[0] * y[1]])
        return f, [0.0, 1.0, 1.0], 0.0
    raise KeyError(name)


DETEST_NAMES = ("A1", "A2", "A3", "A4", "A5", "B1", "B3", "B4", "B5")
