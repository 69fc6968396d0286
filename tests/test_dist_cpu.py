"""world_size-2 gloo tests (CPU) of the host side of the sharded solve:
  * dist.make_reduce sums the norm partials and the segment counts across ranks;
  * a batch-sharded oracle solve whose RMS norm is all-reduced takes the accept/reject decisions and the
    dt sequence of the unsharded solve and reproduces its rows (the property the CUDA path relies on)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import problems as P
from oracle import ode_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torchdiffeq_b200.dist import make_reduce
        segs = [(0, 1), (4, 100 + rank)]
        reduce_fn, n_global, counts = make_reduce(True, segs, torch.device("cpu"))
        buf = torch.tensor([1.0 + rank, 2.0, 0.0], dtype=torch.float64)
        reduce_fn(buf)
        ok_reduce = (counts == [2, 201] and n_global == 203 and buf.tolist() == [3.0, 4.0, 0.0])

        torch.set_num_threads(1)
        f = P.BatchedLinear(16, torch.float64)
        y0 = torch.randn(8, 16, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
        t = torch.linspace(0., 2., 4, dtype=torch.float64)
        rows = slice(rank * 4, (rank + 1) * 4)
        n_total = float(y0.numel())

        def global_rms(x):                       # misc.py:22-23 over the WHOLE batch
            s = (x.abs() ** 2).sum().reshape(1)
            dist.all_reduce(s)
            return (s[0] / n_total).sqrt()
        rec = {}
        with torch.no_grad():
            y = O.odeint_adaptive(f, y0[rows], t, "dopri5", rtol=1e-6, atol=1e-8, norm=global_rms, record=rec)
        out[rank] = (ok_reduce, y, rec["dts"], rec["accepted"])
    finally:
        dist.destroy_process_group()


def test_sharded_norm_reproduces_unsharded_solve():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    f = P.BatchedLinear(16, torch.float64)
    y0 = torch.randn(8, 16, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    t = torch.linspace(0., 2., 4, dtype=torch.float64)
    rec = {}
    with torch.no_grad():
        want = O.odeint_adaptive(f, y0, t, "dopri5", rtol=1e-6, atol=1e-8, record=rec)
    for r in range(world):
        ok_reduce, y, dts, acc = out[r]
        assert ok_reduce
        assert acc == rec["accepted"]
        assert torch.allclose(torch.tensor(dts), torch.tensor(rec["dts"]), rtol=1e-12, atol=0)
        assert torch.allclose(y, want[:, r * 4:(r + 1) * 4], rtol=1e-12, atol=1e-14)


def _adjoint_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        f = P.MLPField(dim=4, hidden=8, seed=0, dtype=torch.float64)
        y0 = torch.randn(8, 4, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
        t = torch.tensor([0., 0.4, 1.0], dtype=torch.float64)
        rows = slice(rank * 4, (rank + 1) * 4)
        gy = torch.randn(3, 8, 4, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
        n_total = float(y0.numel())

        def global_rms(x):                       # a mean over the rows of ALL ranks
            s = (x.abs() ** 2).sum().reshape(1)
            dist.all_reduce(s)
            return (s[0] / n_total).sqrt()

        def reduce_partial(v):
            dist.all_reduce(v)
        ys, gy0, gp = O.adjoint_gradients(f, list(f.parameters()), y0[rows], t, gy[:, rows], "dopri5", rtol=1e-6, atol=1e-8,
                                          state_rms=global_rms, reduce_partial=reduce_partial)
        out[rank] = (ys, gy0, gp)
    finally:
        dist.destroy_process_group()


def test_sharded_adjoint_reproduces_unsharded_gradients():
    """SURVEY.md section 8(e), adjoint: y / adj_y rows stay local, vjp_t and the parameter gradients of every
    evaluation are all-reduced (so every rank integrates the GLOBAL adj_theta and the default adjoint norm sees the
    global gradient tensors), the state segments of the norm are global means.  Every rank then takes the unsharded
    solve's decisions and ends with the complete parameter gradient -- the scheme adjoint.py:_BackwardSolver implements
    on the GPU with NCCL (scripts/dist_check.py --adjoint checks that half)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_adjoint_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    f = P.MLPField(dim=4, hidden=8, seed=0, dtype=torch.float64)
    y0 = torch.randn(8, 4, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    t = torch.tensor([0., 0.4, 1.0], dtype=torch.float64)
    gy = torch.randn(3, 8, 4, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    ys, gy0, gp = O.adjoint_gradients(f, list(f.parameters()), y0, t, gy, "dopri5", rtol=1e-6, atol=1e-8)
    for r in range(world):
        ys_r, gy0_r, gp_r = out[r]
        assert torch.allclose(ys_r, ys[:, r * 4:(r + 1) * 4], rtol=1e-12, atol=1e-14)
        assert torch.allclose(gy0_r, gy0[r * 4:(r + 1) * 4], rtol=1e-10, atol=1e-13)
        for a, b in zip(gp_r, gp):
            assert torch.allclose(a, b, rtol=1e-10, atol=1e-13)          # complete on EVERY rank
