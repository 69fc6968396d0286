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
    B, D = 4096, 128
    f = P.BatchedLinear(D, torch.float32).to(dev)
    y0 = torch.randn(B, D, generator=torch.Generator().manual_seed(1)).to(dev)
    t = torch.linspace(0., 3., 5).to(dev)
    per = B // world
    rows = slice(rank * per, (rank + 1) * per)
    ok = True
    modes = ({"graph": False, "run_ahead": 0}, {"graph": True, "run_ahead": 2},
             {"graph": False, "run_ahead": 0, "exchange": "nccl"}, {"graph": True, "run_ahead": 2, "exchange": "nccl"})
    for mode in modes:
        st = {}
        log("mode", mode, "sharded solve ...")
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
            print("mode", mode, "max|sharded - unsharded| =", err, "steps", (st["n_accept"], st["n_reject"]),
                  (st1["n_accept"], st1["n_reject"]), flush=True)
            ok = ok and err < 1e-5 and same_steps
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    # captured step graphs hold NCCL kernels: drop them before tearing the communicator down
    tdq.clear_cache()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
