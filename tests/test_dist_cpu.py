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
