"""Host side of the fixed-grid RK4 path (solvers.py:52-128 FixedGridODESolver.integrate,
fixed_grid.py:24-29 RK4, rk_common.py:110-118 rk4_alt_step_func).

The grid is known before the first step, so everything the reference decides per step on the host
(step sizes, stage times, which outputs fall into which step, interpolation slopes) is tabulated
once with the reference's own dtype rules and uploaded; one captured step graph then serves every
grid interval, indexed by a device step counter."""
import torch

from . import _lib
from ._engine import _DTYPES, _RetryWithCopies, _stream, pack_pieces, solver_stream

_ONE_THIRD = 1 / 3      # rk_common.py:94-96
_TWO_THIRDS = 2 / 3


def grid_from_step_size(step_size):
    """solvers.py:85-96 _grid_constructor_from_step_size."""
    def _grid_constructor(func, y0, t):
        start_time = t[0]
        end_time = t[-1]
        niters = torch.ceil((end_time - start_time) / step_size + 1).item()
        t_infer = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start_time
        t_infer[-1] = t[-1]
        return t_infer
    return _grid_constructor


FIXED_METHODS = ("euler", "midpoint", "heun2", "heun3", "rk4")


class FixedGridEngine:
    """Explicit fixed-step methods of fixed_grid.py:6-60 on one captured step graph."""

    def __init__(self, fn, n, dtype, device, *, method="rk4", t_sign=1.0, perturb=False, graph="auto",
                 callbacks=None, pieces=None):
        if method not in FIXED_METHODS:
            raise ValueError("unknown fixed-grid method %r" % method)
        self.method = method
        if device.type != "cuda":
            raise _lib.TdqError("torchdiffeq_b200 runs on CUDA devices only (got %s); there is no CPU path" % device)
        if dtype not in _DTYPES:
            raise _lib.TdqError("unsupported state dtype %s (float32 and float64 are implemented)" % dtype)
        self.lib = _lib.load()
        self.fn, self.n, self.dtype, self.device = fn, int(n), dtype, device
        self.dc = _DTYPES[dtype]
        self.t_sign = float(t_sign)
        self.perturb = bool(perturb)
        self.callbacks = callbacks or {}
        self.graph_opt = False if self.callbacks else graph
        self._always_copy = False
        self.pieces = pieces            # fn returns a tuple of pieces (tuple states, the adjoint's augmented state)
        self.nfe = 0
        self.launches = 0

    # ---- tables ---------------------------------------------------------------------------
    def _tabulate(self, grid, t):
        """grid, t: ascending CPU tensors of t's dtype.  Returns per-step and per-output tables."""
        T = self.dtype
        if grid.dtype != t.dtype:                  # a grid_constructor may return another float dtype: compare in
            common = torch.promote_types(grid.dtype, t.dtype)      # the promoted one, like the reference's mixed ops
            t = t.to(common)
        t0, t1 = grid[:-1], grid[1:]
        dt = t1 - t0                                                   # solvers.py:112
        # func times of the four evaluations, then _PerturbFunc's cast to the state dtype (misc.py:187)
        z = torch.zeros_like(t0)
        m = self.method
        if m == "rk4":                                                 # rk_common.py:110-118
            cols, prev_col = [t0, t0 + dt * _ONE_THIRD, t0 + dt * _TWO_THIRDS, t1], 3
        elif m == "euler":                                             # fixed_grid.py:9-11
            cols, prev_col = [t0, z, z, z], None
        elif m == "midpoint":                                          # fixed_grid.py:17-21
            cols, prev_col = [t0, t0 + 0.5 * dt, z, z], None
        elif m == "heun2":                                             # fixed_grid.py:51-60, rk_common.py:141-158
            cols, prev_col = [t0, t0 + dt * 1.0, z, z], 1
        else:                                                          # heun3: fixed_grid.py:35-45, rk_common.py:121-139
            cols, prev_col = [t0, t0 + dt * (1 / 3), t0 + dt * (2 / 3), z], None
        ts = torch.stack(cols, dim=1).to(T)
        if self.perturb:                                               # Perturb.NEXT / PREV, misc.py:188-193
            ts[:, 0] = torch.nextafter(ts[:, 0], ts[:, 0] + 1)
            if prev_col is not None:
                ts[:, prev_col] = torch.nextafter(ts[:, prev_col], ts[:, prev_col] - 1)
        ts = ts * self.t_sign
        dtT = dt.to(T) * self.t_sign           # sign of _ReverseFunc folded into dt (exact)
        # outputs: step s emits every t[j] with t1_s >= t[j] not emitted before (solvers.py:117)
        n_steps = grid.numel() - 1
        step_of = torch.searchsorted(t1.to(t.dtype).contiguous(), t[1:].contiguous(), right=False)
        if step_of.numel() and int(step_of.max()) >= n_steps:
            raise AssertionError("output time beyond the end of the grid")
        g0, g1, tj = t0[step_of], t1[step_of], t[1:]
        mode = torch.full_like(step_of, 2, dtype=torch.int32)
        mode[tj == g1] = 1
        mode[tj == g0] = 0                                             # solvers.py:176-179
        slope = ((tj - g0) / (g1 - g0)).to(T)                          # :180
        counts = torch.bincount(step_of, minlength=n_steps)
        rec_begin = torch.zeros(n_steps + 1, dtype=torch.int32)
        rec_begin[1:] = torch.cumsum(counts, 0).to(torch.int32)
        out_idx = torch.arange(1, t.numel(), dtype=torch.int32)
        return ts.contiguous(), dtT.contiguous(), rec_begin, out_idx, mode.contiguous(), slope.contiguous(), n_steps

    # ---- one step -----------------------------------------------------------------------------
    def _call_fn(self, t, y, own):
        self.nfe += 1
        f = self.fn(t, y)
        if not isinstance(f, torch.Tensor):
            buf = torch.zeros(self.n, dtype=self.dtype, device=self.device)
            self.launches += pack_pieces(self.lib, self.dc, self.dtype, buf, f, self.pieces)
            return buf
        if f.dtype != self.dtype:
            f = f.to(self.dtype)
        f = f.reshape(-1)
        if f.numel() != self.n:
            raise ValueError("func returned %d elements for a state of %d" % (f.numel(), self.n))
        if f.data_ptr() in self._taken and not self._always_copy:
            self._always_copy = True                       # func reuses one output buffer: redo the step with copies
            raise _RetryWithCopies()
        if (self._always_copy or (not f.is_contiguous()) or f.untyped_storage().data_ptr() in self._own
                or f.data_ptr() in self._taken):
            f = f.clone(memory_format=torch.contiguous_format)
        self._taken.add(f.data_ptr())
        return f

    def _step(self):
        try:
            return self._step_once()
        except _RetryWithCopies:                           # nothing of the step has been committed yet
            return self._step_once()

    def _step_once(self):
        lib, dc, n, st = self.lib, self.dc, self.n, _stream()
        y0, ya, y1 = self.y0w.data_ptr(), self.ytmp.data_ptr(), self.y1.data_ptr()
        dtp, stp = self.dt_dev.data_ptr(), self.step_dev.data_ptr()

        def stage(which, out, k1=None, k2=None, k3=None, k4=None):
            p = lambda k: k.data_ptr() if k is not None else None
            _lib.check(lib.tdq_rk4_stage(dc, which, out, y0, p(k1), p(k2), p(k3), p(k4), dtp, stp, n, st))
            self.launches += 1
        m = self.method
        self._taken = set()                                # stage outputs of this step (a func may reuse one buffer)
        k1 = self._call_fn(self.tcur[0], self.y0w, None)
        keep = [k1]
        if m == "rk4":
            stage(1, ya, k1)
            k2 = self._call_fn(self.tcur[1], self.ytmp, None)
            stage(2, y1, k1, k2)
            k3 = self._call_fn(self.tcur[2], self.y1, None)
            stage(3, ya, k1, k2, k3)
            k4 = self._call_fn(self.tcur[3], self.ytmp, None)
            stage(4, y1, k1, k2, k3, k4)
            keep += [k2, k3, k4]
        elif m == "euler":
            stage(5, y1, k1)                                   # dt * f0
        elif m == "midpoint":
            stage(6, ya, k1)                                   # y_mid = y0 + f0 * half_dt
            k2 = self._call_fn(self.tcur[1], self.ytmp, None)
            stage(5, y1, k2)                                   # dt * func(t0 + half_dt, y_mid)
            keep.append(k2)
        elif m == "heun2":
            stage(5, ya, k1)                                   # y0 + dt * k1 * 1.0
            k2 = self._call_fn(self.tcur[1], self.ytmp, None)
            stage(7, y1, k1, k2)
            keep.append(k2)
        else:                                                  # heun3
            stage(1, ya, k1)                                   # y0 + dt * k1 * (1/3)
            k2 = self._call_fn(self.tcur[1], self.ytmp, None)
            stage(8, y1, None, k2)                             # y0 + dt * (k1*0 + k2*(2/3))
            k3 = self._call_fn(self.tcur[2], self.y1, None)
            stage(9, y1, k1, None, k3)                         # y0 + dt * (k1/4 + k2*0 + 3*k3/4)
            keep += [k2, k3]
        _lib.check(lib.tdq_fixed_emit(dc, y0, y1, self.solution.data_ptr(), self.rec_begin.data_ptr(),
                                      self.out_idx.data_ptr(), self.mode.data_ptr(), self.slope.data_ptr(), stp,
                                      self.ts_all.data_ptr(), self.tcur.data_ptr(), self.n_steps, n, st))
        self.launches += 2
        return keep

    def solve(self, y0_flat, grid_cpu, t_cpu):
        dev, T = self.device, self.dtype
        ts, dtT, rec_begin, out_idx, mode, slope, n_steps = self._tabulate(grid_cpu, t_cpu)
        self.n_steps = n_steps
        self.ts_all = ts.to(dev)
        self.dt_dev = dtT.to(dev)
        self.rec_begin, self.out_idx = rec_begin.to(dev), out_idx.to(dev)
        self.mode = mode.to(dev)
        self.slope = slope.to(dev) if slope.numel() else torch.zeros(1, dtype=T, device=dev)
        if self.out_idx.numel() == 0:
            self.out_idx = torch.zeros(1, dtype=torch.int32, device=dev)
            self.mode = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.tcur = self.ts_all[0].clone() if n_steps > 0 else torch.zeros(4, dtype=T, device=dev)
        kw = dict(dtype=T, device=dev)
        self.solution = torch.empty(t_cpu.numel(), self.n, **kw)
        self.solution[0].copy_(y0_flat)
        self.y0w = y0_flat.detach().clone()
        self.ytmp = torch.empty(self.n, **kw)
        self.y1 = torch.empty(self.n, **kw)
        self._own = {x.untyped_storage().data_ptr() for x in (self.y0w, self.ytmp, self.y1, self.solution)}
        if n_steps == 0:
            return self.solution
        cb = self.callbacks.get("callback_step")
        if cb is not None:                                            # solvers.py:113, host in the loop
            t0s, dts = grid_cpu[:-1], grid_cpu[1:] - grid_cpu[:-1]
            for s in range(n_steps):
                cb(t0s[s].to(dev), self.y0w, dts[s].to(dev))
                self._step()
            torch.cuda.current_stream().synchronize()
            return self.solution
        done = 0
        self._step()                                                  # eager first step = warm-up for capture
        done += 1
        graph = None
        if self.graph_opt in (True, "auto") and n_steps > 2:
            try:
                graph = torch.cuda.CUDAGraph()
                nfe, launches = self.nfe, self.launches
                with torch.cuda.graph(graph, stream=solver_stream(self.device)):
                    keep = self._step()
                self._evals, self._graph_launches = self.nfe - nfe, self.launches - launches
                self.nfe, self.launches = nfe, launches
            except Exception as e:
                graph = None
                if self.graph_opt is True:
                    raise
                import warnings
                warnings.warn("torchdiffeq_b200: CUDA graph capture of the RK4 step failed (%s: %s); "
                              "continuing with eager launches" % (type(e).__name__, e))
        while done < n_steps:
            if graph is not None:
                graph.replay()
                self.nfe += self._evals
                self.launches += self._graph_launches
            else:
                self._step()
            done += 1
        torch.cuda.current_stream().synchronize()
        del graph
        return self.solution


FixedRK4Engine = FixedGridEngine      # r1 name
