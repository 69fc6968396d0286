"""Differentiable (non-adjoint) odeint: gradients of the DISCRETE solve, as the reference obtains them by letting
autograd record every solver operation (rk_common.py:31-90 with _UncheckedAssign, interp.py:1-48, solvers.py:102-128;
pinned by tests/gradient_tests.py:13-23 and tests/api_tests.py:28-39).

Recording ~570 ATen ops per attempt is exactly what the B200 path exists to avoid, so the same gradient is computed
differently (discretise-then-differentiate, with checkpoints instead of a recorded graph):

  forward   the ordinary device-resident solve, in lock step, keeping a TAPE of the accepted steps: start time, step
            size, the state y0 and derivative k_0 the step started from (2 N elements per accepted step), and which
            output rows the step produced.  Rejected attempts leave no trace -- in the reference their graph is
            unreachable from the outputs too.
  backward  the accepted steps in reverse.  For one step the stage values are recomputed (same formulas), each
            func evaluation is re-run under autograd to get its vector-Jacobian products w.r.t. (t, y, parameters),
            and the adjoints of the Runge-Kutta recurrences and of the dense-output polynomial are propagated by hand:
                Y_i = y0 + sum_j beta_ij dt k_j          =>  y0_bar += Y_i_bar ;  k_j_bar += beta_ij dt Y_i_bar
                k_{i+1} = f(t_i, Y_i)                    =>  (Y_i_bar, theta_bar, t_i_bar) += vjp_f(k_{i+1}_bar)
                y(t_j) = sum_p c_p x^p,  x = (t_j - t0)/(t1 - t0),  c = interp.py:17-22 of (y0, y1, f0, f1, y_mid)

What is differentiated is what the reference differentiates: step sizes after the first are constants
(misc.py:85 `@torch.no_grad()` on _optimal_step_size); output times enter through x and through every stage time
(t_i = t0 + alpha_i dt with t0 = t[0] + constants); for fixed grids dt = grid[k+1] - grid[k] is differentiated as
well and the grid constructor is differentiated by autograd itself.  One documented difference: the reference's FIRST
step size comes from _select_initial_step, whose value depends differentiably on y0 and t[0] (misc.py:36-77); that
dependence -- a derivative of the discretisation error, not of the solution -- is not propagated here.
"""
import torch

from . import _lib
from ._engine import on_solver_stream


def discover_params(func):
    """Tensors requiring grad that func can reach without being run: nn.Module parameters (also of modules found in a
    plain function's closure cells, its __self__ or its attributes) and bare tensors in those places."""
    seen, out = set(), []

    def add(x):
        if isinstance(x, torch.Tensor):
            if x.requires_grad and id(x) not in seen:
                seen.add(id(x))
                out.append(x)
        elif isinstance(x, torch.nn.Module):
            for q in x.parameters():
                add(q)
    add(func)
    owner = getattr(func, "__self__", None)
    add(owner)
    for holder in (func, owner):
        if holder is not None and hasattr(holder, "__dict__") and not isinstance(holder, torch.nn.Module):
            for v in vars(holder).values():
                add(v)
    for cell in getattr(func, "__closure__", None) or ():
        try:
            add(cell.cell_contents)
        except ValueError:
            pass
    inner = getattr(func, "func", None)                    # functools.partial
    if inner is not None and inner is not func:
        for q in discover_params(inner):
            add(q)
        for a in getattr(func, "args", ()) or ():
            add(a)
        for a in (getattr(func, "keywords", None) or {}).values():
            add(a)
    return tuple(out)


class Tableau:
    """Dense Python-side copy of an explicit tableau for the reverse sweep."""

    def __init__(self, alpha, beta, c_sol, fsal, c_mid=None):
        self.alpha, self.beta, self.c_sol, self.fsal, self.c_mid = alpha, beta, c_sol, fsal, c_mid
        self.S = len(alpha)


def adaptive_tableau(method):
    d = _lib.tableau_as_dict(method)
    return Tableau(d["alpha"], d["beta"], d["c_sol"], d["fsal"], d["c_mid"])


# The fixed-grid step functions written as tableaus (fixed_grid.py:6-60, rk_common.py:110-158): alpha, beta rows,
# weights of (k_1 .. k_S) in dy.  k_1 = f(t0, y0) is a fresh evaluation every step (no FSAL carry).
_THIRD = 1 / 3
FIXED_TABLEAUS = {
    "euler": ([], [], [1.0]),
    "midpoint": ([0.5], [[0.5]], [0.0, 1.0]),
    "heun2": ([1.0], [[1.0]], [0.5, 0.5]),
    "heun3": ([_THIRD, 2 / 3], [[_THIRD], [0.0, 2 / 3]], [0.25, 0.0, 0.75]),
    "rk4": ([_THIRD, 2 / 3, 1.0], [[_THIRD], [-_THIRD, 1.0], [1.0, -1.0, 1.0]], [0.125, 0.375, 0.375, 0.125]),
}


def _acc(a, b, alpha=None):
    """a + alpha*b for adjoint accumulators that start as None."""
    if b is None:
        return a
    if alpha is not None:
        b = b * alpha
    return b if a is None else a + b


class StepAdjoint:
    """Reverse sweep through ONE explicit Runge-Kutta step with constant coefficients c_ij = beta_ij * dt."""

    def __init__(self, F, params, need_t):
        self.F, self.params, self.need_t = F, tuple(params), need_t
        self.pbar = [None] * len(self.params)

    def vjp(self, t_val, y_val, g):
        """(y_bar, t_bar) of f(t, y) against g; parameter gradients accumulate in self.pbar."""
        with torch.enable_grad():
            tr = t_val.detach().clone().requires_grad_(self.need_t)
            yr = y_val.detach().requires_grad_(True)
            out = self.F(tr, yr)
            if not out.requires_grad:
                return None, None
            inputs = [yr] + ([tr] if self.need_t else []) + list(self.params)
            grads = torch.autograd.grad(out, inputs, g, allow_unused=True)
        off = 2 if self.need_t else 1
        for i, gq in enumerate(grads[off:]):
            self.pbar[i] = _acc(self.pbar[i], gq)
        return grads[0], (grads[1] if self.need_t else None)

    def stages(self, times, y0, k_first, coefs):
        """Recompute stage values Y_i and slopes k_{i+1} = F(t_i, Y_i).  coefs[i][j] multiplies k_j in Y_i
        (already including dt and, for the fixed-grid tableaus, the offset of k_1).  Returns (Ys, ks)."""
        ks, Ys = [k_first], []
        with torch.no_grad():
            for i, row in enumerate(coefs):
                acc = None
                for j, c in enumerate(row):
                    if c != 0.0:
                        term = ks[j] * c
                        acc = term if acc is None else acc + term
                Yi = y0 if acc is None else y0 + acc
                Ys.append(Yi)
                ks.append(self.F(times[i], Yi))
        return Ys, ks

    def sweep(self, times, Ys, coefs, kbar, ybar0, Ybar_last=None):
        """Adjoint of `stages`: kbar[j] holds what later computations contributed to k_j (kbar[0]: k_first).  Returns
        (ybar0, kbar0, sum_i t_i_bar, sum_i alpha-free dt-sensitivity list) -- the per-stage (Ybar_i, tbar_i) are
        returned for callers that differentiate dt."""
        S = len(coefs)
        per_stage = [None] * S
        tsum = None
        for i in reversed(range(S)):
            Yb = Ybar_last if i == S - 1 else None
            tb = None
            if kbar[i + 1] is not None:
                gy, tb = self.vjp(times[i], Ys[i], kbar[i + 1])
                Yb = _acc(Yb, gy)
            if Yb is not None:
                ybar0 = _acc(ybar0, Yb)
                for j, c in enumerate(coefs[i]):
                    if c != 0.0:
                        kbar[j] = _acc(kbar[j], Yb, c)
            tsum = _acc(tsum, tb)
            per_stage[i] = (Yb, tb)
        return ybar0, kbar[0], tsum, per_stage


def _T(x, dtype, device=None):
    """x rounded to the state dtype, as a 0-dim CPU tensor (scalar arithmetic stays on the host: no syncs)."""
    return torch.as_tensor(x, dtype=torch.float64).to(dtype)


def _dev(x, device):
    """0-dim device tensor holding the CPU scalar x (a fill kernel, not a synchronous copy)."""
    return torch.full((), float(x), dtype=x.dtype, device=device)


def _prev(t):
    return torch.nextafter(t, t - 1)


def _next(t):
    return torch.nextafter(t, t + 1)


def adaptive_backward(p, tab, tape, t, grad_sol, params, need_t):
    """Reverse sweep over the taped accepted steps of an adaptive solve.  Times on the tape are the engine's ascending
    s = sign * t.  Returns (t_bar or None, y0_bar, [param_bar])."""
    dev, T, sign = p.device, p.dtype, p.t_sign

    def F(s_, y_):                                        # reference-sense dynamics in ascending time (misc.py:158-165)
        out = p.fn(s_ * sign, y_)
        if isinstance(out, tuple):
            out = p.layout.flatten(list(out))
        out = out.reshape(-1)
        return out * sign if sign != 1.0 else out
    sa = StepAdjoint(F, params, need_t)
    S = tab.S
    n_out = grad_sol.shape[0]
    sbar = torch.zeros(n_out, dtype=torch.float64, device=dev) if need_t else None   # w.r.t. the ascending output times
    s_out = p.t_cpu.to(torch.float64)                     # ascending engine time of every output row (host copy)
    gy = None            # adjoint of the state the NEXT step starts from
    gk = None            # adjoint of the k_0 the next step starts from
    shift = None         # adjoint of a common shift of all step times (= d/d s[0])
    for st in reversed(tape):
        s0, dt = st["t0"], st["dt"]
        s1 = s0 + dt
        dtT, t0T, t1T = _T(dt, T, dev), _T(s0, T, dev), _T(s1, T, dev)
        y0, k0 = st["y0"], st["k0"]
        if sign != 1.0:
            k0 = k0 * sign                               # the engine keeps RAW func outputs; F is reference-sense
        times = [_dev(_prev(t1T) if a == 1.0 else t0T + _T(a, T) * dtT, dev) for a in tab.alpha]
        coefs = [[float(_T(b, T) * dtT) for b in row] for row in tab.beta]
        Ys, ks = sa.stages(times, y0, k0, coefs)
        if tab.fsal:
            y1 = Ys[-1]
        else:
            csol = [float(dtT * _T(c, T)) for c in tab.c_sol]
            y1 = y0 + sum(k * c for k, c in zip(ks, csol) if c != 0.0)
        kbar = [None] * (S + 1)
        ybar0, ybar1 = None, gy
        kbar[S] = gk
        # ---- dense output rows produced by this step (interp.py:1-48) ---------------------------------------------
        lo, hi = st["out_lo"], st["out_hi"]
        if hi > lo:
            cmid = [float(dtT * _T(c, T)) for c in tab.c_mid]
            f0, f1 = ks[0], ks[S]
            ymid = y0 + sum(k * c for k, c in zip(ks, cmid) if c != 0.0)
            dtf = float(dtT)
            a = 2 * dtf * (f1 - f0) - 8 * (y1 + y0) + 16 * ymid
            b = dtf * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * ymid
            c = dtf * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * ymid
            d = dtf * f0
            ab = bb = cb = db = eb = None
            for j in range(lo, hi):
                G = grad_sol[j]
                x = float(((s_out[j] - s0) / (s1 - s0)).to(T))          # interp.py:39-40
                eb = _acc(eb, G)
                db = _acc(db, G, x)
                cb = _acc(cb, G, x * x)
                bb = _acc(bb, G, x ** 3)
                ab = _acc(ab, G, x ** 4)
                if need_t:
                    dp = d + (2 * x) * c + (3 * x * x) * b + (4 * x ** 3) * a
                    xbar = torch.dot(G.double(), dp.double())
                    sbar[j] += xbar / (s1 - s0)
                    shift = _acc(shift, -xbar / (s1 - s0))
            ybar0 = _acc(_acc(_acc(_acc(ybar0, eb), bb, 18.0), ab, -8.0), cb, -11.0)
            ybar1 = _acc(_acc(_acc(ybar1, ab, -8.0), bb, 14.0), cb, -5.0)
            kbar[0] = _acc(_acc(_acc(_acc(kbar[0], ab, -2 * dtf), bb, 5 * dtf), cb, -4 * dtf), db, dtf)
            kbar[S] = _acc(_acc(_acc(kbar[S], ab, 2 * dtf), bb, -3 * dtf), cb, dtf)
            ymb = _acc(_acc(_acc(None, ab, 16.0), bb, -32.0), cb, 16.0)
            ybar0 = _acc(ybar0, ymb)
            for j, cm in enumerate(cmid):
                if cm != 0.0:
                    kbar[j] = _acc(kbar[j], ymb, cm)
        # ---- y1 (rk_common.py:83-87) and the stages ---------------------------------------------------------------
        Ybar_last = None
        if tab.fsal:
            Ybar_last = ybar1
        elif ybar1 is not None:
            ybar0 = _acc(ybar0, ybar1)
            for j, cs in enumerate(csol):
                if cs != 0.0:
                    kbar[j] = _acc(kbar[j], ybar1, cs)
        ybar0, kbar0, tsum, _ = sa.sweep(times, Ys, coefs, kbar, ybar0, Ybar_last)
        shift = _acc(shift, tsum.double() if tsum is not None else None)
        # ---- where this step's k_0 came from ----------------------------------------------------------------------
        if st["first"]:                                   # k_0 = f(t[0], y0) (rk_common.py:214)
            if kbar0 is not None:
                gyk, tb = sa.vjp(_dev(t0T, dev), y0, kbar0)
                ybar0 = _acc(ybar0, gyk)
                shift = _acc(shift, tb.double() if tb is not None else None)
            gk = None
        elif st["jumped_into"] is not None:               # re-evaluated after a discontinuity at next(t0) (rk_common.py:346-351)
            if kbar0 is not None:
                gyk, tb = sa.vjp(_dev(_next(t0T), dev), y0, kbar0)
                ybar0 = _acc(ybar0, gyk)
                shift = _acc(shift, tb.double() if tb is not None else None)
            gk = None
        else:
            gk = kbar0
        gy = ybar0
    y0bar = _acc(gy, grad_sol[0])                                     # solution[0] = y0 (solvers.py:30)
    tbar = None
    if need_t:
        sbar[0] += shift if shift is not None else 0.0
        tbar = (sbar * sign).to(t.dtype).to(t.device)
    return tbar, y0bar, sa.pbar


def fixed_backward(p, method, tape, grid, t_cpu, grad_sol, params, need_t):
    """Reverse sweep over the steps of a fixed-grid solve (solvers.py:102-128, linear interpolation :175-181).
    grid / t_cpu: ascending CPU tensors.  Returns (grid_bar, t_out_bar) as float64 device tensors (or None), y0_bar,
    [param_bar]."""
    dev, T, sign = p.device, p.dtype, p.t_sign
    alpha, beta, wts = FIXED_TABLEAUS[method]

    def F(s_, y_):
        out = p.fn(s_ * sign, y_)
        if isinstance(out, tuple):
            out = p.layout.flatten(list(out))
        out = out.reshape(-1)
        return out * sign if sign != 1.0 else out
    sa = StepAdjoint(F, params, need_t)
    gbar = torch.zeros(grid.numel(), dtype=torch.float64, device=dev) if need_t else None
    obar = torch.zeros(t_cpu.numel(), dtype=torch.float64, device=dev) if need_t else None
    gy = None
    for st in reversed(tape):
        k = st["k"]
        g0, g1 = grid[k], grid[k + 1]
        dt = g1 - g0                                                   # t's dtype, like the reference (solvers.py:112)
        dtT, t0T = dt.to(T), g0.to(T)
        dtf = float(dtT)
        y0 = st["y0"]
        # stage times by the reference's dtype rules: t0 + dt*alpha in t's dtype, cast to the state dtype by _PerturbFunc
        tc = [((g0 + dt * a) if a != 1.0 else (g0 + dt * 1.0 if method == "heun2" else g1)).to(T) for a in alpha]
        t0_eval = t0T
        if st["perturb"]:
            t0_eval = _next(t0T)
            tc = [(_prev(tt) if a == 1.0 else tt) for tt, a in zip(tc, alpha)]
        times = [_dev(tt, dev) for tt in tc]
        t0_dev = _dev(t0_eval, dev)
        with torch.no_grad():
            k1 = F(t0_dev, y0)
        coefs = [[b * dtf for b in row] for row in beta]
        Ys, ks = sa.stages(times, y0, k1, coefs)
        incr = None
        for kk, w in zip(ks, wts):
            if w != 0.0:
                incr = _acc(incr, kk, w)                               # dy / dt
        y1 = y0 + incr * dtf
        ybar1, ybar0 = gy, None
        h = float(g1 - g0)
        for (j, mode, slope) in st["outs"]:                            # solvers.py:175-181
            G = grad_sol[j]
            if mode == 0:
                ybar0 = _acc(ybar0, G)
            elif mode == 1:
                ybar1 = _acc(ybar1, G)
            else:
                sl = float(torch.as_tensor(slope).to(T))
                ybar0 = _acc(ybar0, G, 1.0 - sl)
                ybar1 = _acc(ybar1, G, sl)
                if need_t:
                    sb = torch.dot(G.double(), (y1 - y0).double())
                    frac = float(t_cpu[j] - g0) / h
                    obar[j] += sb / h
                    gbar[k] += sb * (-1.0 / h + frac / h)
                    gbar[k + 1] += sb * (-frac / h)
        kbar = [None] * (len(alpha) + 1)
        dtbar = None
        if ybar1 is not None:
            ybar0 = _acc(ybar0, ybar1)
            for j, w in enumerate(wts):
                if w != 0.0:
                    kbar[j] = _acc(kbar[j], ybar1, w * dtf)
            if need_t:
                dtbar = _acc(dtbar, torch.dot(ybar1.double(), incr.double()))
        ybar0, kbar0, tsum, per_stage = sa.sweep(times, Ys, coefs, kbar, ybar0, None)
        if need_t:
            for i, (Yb, tb) in enumerate(per_stage):
                if Yb is not None and dtf != 0.0:
                    dtbar = _acc(dtbar, torch.dot(Yb.double(), ((Ys[i] - y0) / dtf).double()))
                if tb is not None:
                    gbar[k] += tb.double()
                    dtbar = _acc(dtbar, tb.double(), alpha[i])
        if kbar0 is not None:                                          # k_1 = f(t0, y0), a fresh evaluation every step
            gyk, tb = sa.vjp(t0_dev, y0, kbar0)
            ybar0 = _acc(ybar0, gyk)
            if need_t and tb is not None:
                gbar[k] += tb.double()
        if need_t and dtbar is not None:
            gbar[k + 1] += dtbar
            gbar[k] -= dtbar
        gy = ybar0
    y0bar = _acc(gy, grad_sol[0])
    return gbar, obar, y0bar, sa.pbar


class _BackpropFunction(torch.autograd.Function):
    """odeint with gradients of the discrete solve (see the module docstring)."""

    @staticmethod
    def forward(ctx, p, run, t, y0_flat, *params):
        ctx.p, ctx.n_params = p, len(params)
        with torch.no_grad():
            sol, ctx.aux = run()
        ctx.save_for_backward(t, *params)
        ctx.need_t = t.requires_grad
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        p = ctx.p
        t, *params = ctx.saved_tensors
        grad_sol = grad_sol.contiguous()
        with on_solver_stream(p.device) as ss:
            if ctx.aux["kind"] == "adaptive":
                with torch.no_grad():
                    tbar, y0bar, pbar = adaptive_backward(p, ctx.aux["tab"], ctx.aux["tape"], t, grad_sol, params,
                                                          ctx.need_t)
            else:
                grid_req, grid, t_req = ctx.aux["grid_req"], ctx.aux["grid"], ctx.aux["t_req"]
                with torch.no_grad():
                    gbar, obar, y0bar, pbar = fixed_backward(p, p.method, ctx.aux["tape"], grid, p.t_cpu, grad_sol, params,
                                                             ctx.need_t)
                tbar = None
                if ctx.need_t:
                    # the grid as a differentiable function of the (ascending) output times, whatever constructor made it
                    tb = obar.to("cpu")
                    if grid_req.requires_grad:
                        (gt,) = torch.autograd.grad(grid_req, t_req, gbar.to("cpu").to(grid_req.dtype), allow_unused=True)
                        if gt is not None:
                            tb = tb + gt.double()
                    tbar = (tb * p.t_sign).to(t.dtype).to(t.device)
            if y0bar is None:
                y0bar = torch.zeros(p.n, dtype=p.dtype, device=p.device)
            pbar = [g if g is not None else torch.zeros_like(q) for g, q in zip(pbar, params)]
            ss.publish(y0bar, *pbar)
        return (None, None, tbar, y0bar, *pbar)
