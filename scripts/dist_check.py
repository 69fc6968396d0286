"""Sharded-solve check, run under torchrun (NCCL, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
Each rank integrates its slice of ONE seeded batch with options['process_group']; rank 0 also solves the whole
batch alone.  The sharded rows must reproduce the unsharded solve (same dt sequence: the RMS norm is global)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P            # noqa: E402
import torchdiffeq_b200 as tdq  # noqa: E402


def log(*a):
    print("[rank %s]" % os.environ.get("RANK"), *a, file=sys.stderr, flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DIST_CHECK_DUMP_AFTER", "300")), exit=True)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    log("process group up")
    per = int(os.environ.get("DIST_CHECK_ROWS_PER_RANK", "0")) or (4096 // world)
    B, D = per * world, 128
    f_generic = P.BatchedLinear(D, torch.float32).to(dev)
    # the same field as a LinearField: the whole attempt (stages, norm, controller, peer exchange) is one tcgen05 launch
    f_linear = tdq.LinearField(P.skew_matrix(D, torch.float32).to(dev))
    y0 = torch.randn(B, D, generator=torch.Generator().manual_seed(1)).to(dev)      # ONE seeded batch, sharded by rows
    t = torch.linspace(0., 3., 5).to(dev)
    rows = slice(rank * per, (rank + 1) * per)
    ok = True
    modes = ({"graph": False, "run_ahead": 0}, {"graph": True, "run_ahead": 2},
             {"graph": False, "run_ahead": 0, "exchange": "nccl"}, {"graph": True, "run_ahead": 2, "exchange": "nccl"})
    if per > 8192:          # the big configuration (65,536 rows per rank): the two production modes only
        modes = ({"graph": True, "run_ahead": 2}, {"graph": True, "run_ahead": 2, "exchange": "nccl"})
    fields = [("generic", f_generic)] + ([("linear", f_linear)] if os.environ.get("DIST_CHECK_LINEAR", "1") == "1" else [])
    for fname, f, mode in [(n_, f_, m_) for n_, f_ in fields for m_ in modes]:
        st = {}
        log("field", fname, "mode", mode, "sharded solve ...")
        with torch.no_grad():
            y = tdq.odeint(f, y0[rows].contiguous(), t, method="dopri5", rtol=1e-5, atol=1e-7,
                           options=dict(mode, process_group=True), _stats=st)
        log("sharded solve done", st.get("n_accept"), st.get("n_reject"))
        gathered = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(gathered, y)
        if rank == 0:
            full = torch.cat(gathered, dim=1)
            st1 = {}
            with torch.no_grad():
                want = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-7,
                                  options={k: v for k, v in mode.items() if k != "exchange"}, _stats=st1)
            err = (full - want).abs().max().item()
            same_steps = (st["n_accept"], st["n_reject"]) == (st1["n_accept"], st1["n_reject"])
            print("field", fname, "fused_attempt", st.get("fused_attempt"), "mode", mode, "max|sharded - unsharded| =", err, "steps",
                  (st["n_accept"], st["n_reject"]), (st1["n_accept"], st1["n_reject"]), flush=True)
            ok = ok and err < 1e-5 and same_steps
    if os.environ.get("DIST_CHECK_ADJOINT", "1") == "1":
        ok = adjoint_check(rank, world, dev) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    # captured step graphs hold NCCL kernels: drop them before tearing the communicator down
    tdq.clear_cache()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag) == 1 else 1)


def adjoint_check(rank, world, dev):
    """Sharded odeint_adjoint: every rank integrates its rows forward and backward; vjp_t and
    the parameter gradients of every evaluation are all-reduced, so each rank ends with the COMPLETE dL/dtheta and its
    rows of dL/dy0.  Compared with the unsharded solve of the whole batch on rank 0 to 1e-6 relative."""
    B, D = 64 * world, 8
    ok = True
    for norm in ("default", "seminorm"):
        for mode in ({"graph": False, "run_ahead": 0}, {"graph": True, "run_ahead": 2}):
            torch.manual_seed(0)
            fm = P.MLPField(dim=D, hidden=16, seed=0, dtype=torch.float64).to(dev)
            y0 = torch.randn(B, D, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(dev)
            w = torch.randn(3, B, D, generator=torch.Generator().manual_seed(2), dtype=torch.float64).to(dev)
            t = torch.tensor([0., 0.4, 1.0], dtype=torch.float64, device=dev)
            per = B // world
            rows = slice(rank * per, (rank + 1) * per)
            ao = dict(mode, process_group=True)
            if norm == "seminorm":
                ao["norm"] = "seminorm"
            yy = y0[rows].clone().requires_grad_(True)
            fm.zero_grad()
            out = tdq.odeint_adjoint(fm, yy, t, method="dopri5", rtol=1e-7, atol=1e-9, options=dict(mode, process_group=True),
                                     adjoint_options=ao)
            (out * w[:, rows]).sum().backward()
            gy_all = [torch.empty_like(yy.grad) for _ in range(world)]
            dist.all_gather(gy_all, yy.grad)
            gp = [q.grad.clone() for q in fm.parameters()]
            # every rank must hold the same complete parameter gradient
            for g in gp:
                ref = g.clone()
                dist.broadcast(ref, 0)
                ok = ok and bool(torch.equal(ref, g))
            if rank == 0:
                y1 = y0.clone().requires_grad_(True)
                fm.zero_grad()
                ao1 = dict(mode)
                if norm == "seminorm":
                    ao1["norm"] = "seminorm"
                o1 = tdq.odeint_adjoint(fm, y1, t, method="dopri5", rtol=1e-7, atol=1e-9, options=dict(mode),
                                        adjoint_options=ao1)
                (o1 * w).sum().backward()
                rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
                e_y = rel(torch.cat(gy_all, 0), y1.grad)
                e_p = max(rel(a, q.grad) for a, q in zip(gp, fm.parameters()))
                print("adjoint", norm, mode, "rel err dL/dy0 %.2e dL/dtheta %.2e" % (e_y, e_p), flush=True)
                ok = ok and e_y < 1e-6 and e_p < 1e-6
    return ok


if __name__ == "__main__":
    main()
